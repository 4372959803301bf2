"""Sample-space data parallelism: the scene is replicated, the samples of every pixel are split over
ranks, and the per-pass float sum buffers are summed onto rank 0 with ONE collective.

Only plumbing lives here (torch.distributed over NCCL on GPUs, gloo in the CPU tests); who renders what is pure
arithmetic so it can be tested without a GPU:

* interleaved split — rank r of R draws global samples r, r+R, r+2R, ... (a fixed job split R ways: strong scaling);
* blocked split     — rank r draws [r*spp, (r+1)*spp) (every rank renders a full-spp pass: weak scaling, what bench.py
  reports).

Because the Philox stream is keyed on the GLOBAL sample index, the multiset of samples — hence the reduced image up to
float-add order — does not depend on R.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional


@dataclass(frozen=True)
class RankSamples:
    spp: int            # samples per pixel this rank renders in one pass
    sample_base: int    # global index of its first sample
    sample_stride: int  # global index stride between its samples
    total_spp: int      # samples per pixel of the whole job (the divisor of the reduced sum)


def interleaved_split(total_spp: int, rank: int, world: int) -> RankSamples:
    """Strong scaling: `total_spp` samples per pixel split over `world` ranks, remainder to the lowest ranks."""
    if not (0 <= rank < world) or total_spp < 0:
        raise ValueError("bad rank/world/spp")
    n = total_spp // world + (1 if rank < total_spp % world else 0)
    return RankSamples(n, rank, world, total_spp)


def blocked_split(spp_per_rank: int, rank: int, world: int) -> RankSamples:
    """Weak scaling: every rank renders `spp_per_rank` samples per pixel with disjoint global indices."""
    if not (0 <= rank < world) or spp_per_rank < 0:
        raise ValueError("bad rank/world/spp")
    return RankSamples(spp_per_rank, rank * spp_per_rank, 1, spp_per_rank * world)


def global_sample_indices(rs: RankSamples):
    return [rs.sample_base + k * rs.sample_stride for k in range(rs.spp)]


def reduce_sum_to_root(tensor, dst: int = 0):
    """The one collective of a pass: element-wise sum of every rank's accumulation buffer onto `dst`."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM)
    return tensor


def render_pass_distributed(device, host_world, width: int, height: int, rs: RankSamples, d_sum, stream: Optional[int] = None,
                            pass_index: int = 0, seed: int = 0x50545348, rank: Optional[int] = None):
    """One multi-process multi-GPU pass (one rank per GPU): accumulate this rank's samples into `d_sum` (a torch CUDA tensor of
    width*height*3 floats), reduce, and on rank 0 apply Buffer.AddSample with the job's total spp.  Everything is ordered on
    `stream` (a cudaStream_t handle); None = torch's current stream, the one `d_sum.zero_()` and the NCCL reduce are issued on.
    (A single-process host uses ptgpu_params.devices instead: bindings.Device(devices=[...]).)"""
    import torch
    import torch.distributed as dist
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream if d_sum.is_cuda else 0
    d_sum.zero_()
    if rs.spp > 0:
        p = host_world.make_pass(width, height, rs.spp, pass_index=pass_index, seed=seed, sample_base=rs.sample_base,
                                 sample_stride=rs.sample_stride)
        device.accumulate_device(p, d_sum.data_ptr(), stream)
    reduce_sum_to_root(d_sum, 0)
    if rank is None:
        rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0
    if rank == 0:
        device.add_sample_device(width, height, d_sum.data_ptr(), float(rs.total_spp), stream)
