"""Multi-process sample partition on CPU (gloo, world size 2): the split arithmetic, the single reduce per pass, and the
rank-invariance of the keyed RNG — checked with the oracle standing in for the device renderer."""
import os
import socket

import numpy as np
import pytest

from ptsharp_b200 import distributed as D


def test_split_arithmetic():
    for total in (1, 7, 16, 512, 1024):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                rs = D.interleaved_split(total, r, world)
                assert rs.total_spp == total and rs.sample_stride == world
                seen += D.global_sample_indices(rs)
            assert sorted(seen) == list(range(total))  # a partition of the job's samples
    for spp in (1, 16, 512):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                rs = D.blocked_split(spp, r, world)
                assert rs.total_spp == spp * world
                seen += D.global_sample_indices(rs)
            assert sorted(seen) == list(range(spp * world))
    with pytest.raises(ValueError):
        D.interleaved_split(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import orc
    from ptsharp_b200 import scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, total = 40, 30, 6
    w = orc.OracleWorld()
    scenes.build_c2(w)
    rs = D.interleaved_split(total, rank, world)
    mean, _, cnt = w.render(W, H, rs.spp, passes=1, threads=1, rng_mode=orc.RNG_KEYED, sample_base=rs.sample_base,
                            sample_stride=rs.sample_stride)
    part = torch.from_numpy(mean * rs.spp)  # the rank's per-pixel SUM, as the device accumulates it
    D.reduce_sum_to_root(part, 0)
    if rank == 0:
        q.put((part.numpy() / rs.total_spp, cnt["cameraSamples"]))
    dist.destroy_process_group()


def test_two_rank_reduce_matches_single_rank(orc):
    import torch.multiprocessing as mp
    from ptsharp_b200 import scenes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    img, n0 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w = orc.OracleWorld()
    scenes.build_c2(w)
    ref, _, _ = w.render(40, 30, 6, passes=1, threads=2, rng_mode=orc.RNG_KEYED)
    np.testing.assert_allclose(img, ref, rtol=1e-12, atol=1e-12)
    assert n0 == 40 * 30 * 3
