#!/usr/bin/env python
"""bench.py — headline benchmark: Msamples/s (and Gpaths·bounce/s) of one RenderParallel pass on a BASELINE config.

Default workload (config.workload = "c3_icospheres_1m", the configuration the north-star target is quoted on): two
displaced-icosphere meshes, 1 000 000 triangles in the reference's kd-trees, Glossy + Clear materials, 1920x1080,
SamplesPerPixel = 512, DefaultSampler.NewSampler(1, 4).  `--workload c1|c2|c4|c5` times the other BASELINE configs the same way.
Synthetic, generated from closed-form formulas (ptsharp_b200/scenes.py).

A step = one pass (Renderer.RenderParallel, Renderer.cs:199-338) over the whole frame at the config's SamplesPerPixel.
  value  device-timed throughput, scene resident in HBM, CUDA events, max over ranks;
  e2e    the same pass through the C ABI with HOST buffers: every step re-uploads the flat scene (host -> device) and reads
         the image back (device -> host);
  roofline       the dominant kernel of the profiled pass (k_mesh / k_march, k_shade for scenes of analytic shapes): algorithmic bytes per launch / launch
                 duration measured live with CUDA events on the launching stream, vs the measured HBM peak;
  cpu_baseline   the CPU restatement in oracle/ (kind "port": the C# reference cannot be built here) on all host cores, on a
                 bounded sample of the same workload: every k-th 32x32 task of the WHOLE frame (the reference's own task list,
                 Renderer.cs:257-281) at a reduced spp, so the sample has the frame's mix of paths.
Multi-GPU (torchrun, one rank per GPU): scene replicated, the config's spp are SPLIT over the ranks (rank r draws global samples
r, r+N, ...: strong scaling, a fixed job), per-pass float sum buffers are reduced to rank 0 with one NCCL reduce, rank 0 does the
Buffer.AddSample.  `weak` (every rank renders the full spp with disjoint global indices) is measured beside it and reported as an
extra key; `--scaling weak` makes it the headline instead.
`--impl reference` times the oracle alone (rank 0 only) on the same bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (builder key, kwargs)
    "c1": ("c1", {}),
    "c2": ("c2", {}),
    "c3": ("c3", {}),
    "c4": ("c4", {}),
    "c5": ("c5", {}),
    "c3_small": ("c3", dict(freq_a=60, freq_b=30, width=480, height=270, spp=16)),
}
SAMPLERS = {"c1": "DefaultSampler.NewSampler(16,4)", "c2": "NewSampler(1,8) LightModeAll", "c3": "DefaultSampler.NewSampler(1,4)",
            "c4": "DefaultSampler.NewSampler(1,4)", "c5": "NewSampler(4,4) LightModeAll SpecularModeAll", "c3_small": "DefaultSampler.NewSampler(1,4)"}
B_PER_BOUNCE_NEE, B_PER_BOUNCE, B_PER_SAMPLE = 208, 128, 24
TRACE_BYTES_PER_RAY = 32 + 24  # k_scene_trace: reads (o,pixel),(d,meta) = 2 x float4, writes the 24-byte hit record
MESH_BYTES_PER_ITEM = 48 + 12  # k_mesh / k_march: reads the 48-byte work item (co, ray | cd, root | tmin, tmax), writes T (8) + triangle (4)
TRAFFIC_FILES = ("r02_k_mesh_traffic.json", "r01_k_mesh_traffic.json")  # ncu --set full captures of k_mesh, newest first


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 6:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except ValueError:
                continue
            for k, n in enumerate(names):
                if r[2 + k].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_scene(world, workload: str, spp_override=None):
    from ptsharp_b200 import scenes
    key, kw = WORKLOADS[workload]
    cfg = scenes.BUILDERS[key](world, **kw)
    if spp_override:
        cfg.spp = spp_override
    return cfg


def config_dict(cfg, workload: str, scaling: str) -> dict:
    """The `config` object of the JSON line: the same keys and values for both arms (what is rendered, not how)."""
    return {"workload": cfg.name, "width": cfg.width, "height": cfg.height, "spp": cfg.spp, "triangles": cfg.triangles,
            "sampler": SAMPLERS[workload], "scaling": scaling}


def cpu_baseline(workload: str, seconds_target=12.0, steps=1, warmup=0, spp_override=None):
    """Time the oracle (CPU restatement) with every host core on a bounded sample of the workload: every k-th non-empty 32x32 task
    of the whole frame (a strided sample of the reference's own task queue, so the sample has the frame's mix of paths) at a
    reduced spp.  One step = one pass over that sample."""
    from oracle import orc
    ow = orc.OracleWorld()
    cfg = build_scene(ow, workload, spp_override)
    ow.compile()
    cores = os.cpu_count() or 1
    W, H = cfg.width, cfg.height
    tiles_x, tiles_y = (W + 255) // 256, (H + 255) // 256
    n_tasks = sum(1 for ty in range(tiles_y) for tx in range(tiles_x) for sy in range(8) for sx in range(8)
                  if tx * 256 + sx * 32 < W and ty * 256 + sy * 32 < H)
    # probe: one sample per pixel over a sparse task sample, to size the timed sample
    probe_stride = max(1, n_tasks // max(4 * cores, 32))
    ow.set_task_sample(probe_stride, 0)
    t0 = time.perf_counter()
    _, _, cnt = ow.render(W, H, 1, passes=1, threads=cores, seed=7)
    probe = max(time.perf_counter() - t0, 1e-3)
    rate = cnt["cameraSamples"] / probe
    spp = int(max(1, min(cfg.spp, 4)))
    want_tasks = max(cores, int(rate * seconds_target / (1024.0 * spp)))
    stride = max(1, n_tasks // want_tasks)
    ow.set_task_sample(stride, stride // 2)
    times, cnts = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, cnts = ow.render(W, H, spp, passes=1, threads=cores, seed=11 + i)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    ow.set_task_sample(1, 0)
    dt = sum(times) / len(times)
    return {
        "value": cnts["cameraSamples"] / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
        "gpaths_bounce_per_s": cnts["segments"] / dt / 1e9, "rays_per_s": (cnts["segments"] + cnts["shadowRays"]) / dt,
        "segments_per_sample": cnts["segments"] / max(cnts["cameraSamples"], 1),
        "sample": f"{workload}: every {stride}-th of the {n_tasks} 32x32 tasks of the whole {W}x{H} frame (Renderer.cs:257-281 queue order), "
                  f"{spp} spp, 1 pass = {cnts['cameraSamples']} camera samples in {dt:.2f} s; oracle/ C++ restatement of RenderParallel "
                  f"(not the .NET binary), {cores} threads on a shared FIFO of 32x32 tasks",
        "seconds": dt,
    }, cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, cfg = cpu_baseline(args.workload, seconds_target=args.cpu_seconds, steps=max(1, args.steps), warmup=min(args.warmup, 1),
                             spp_override=args.spp or None)
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": base["value"], "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 vectors + f64 scalars (the reference's Vector/double model)", "data": "synthetic",
        "config": config_dict(cfg, args.workload, args.scaling),
        "parallelism": "host CPU threads (rank 0 only)",
        "gpaths_bounce_per_s": base["gpaths_bounce_per_s"], "segments_per_sample": base["segments_per_sample"],
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = the config's spp split over the ranks (default); weak = every rank renders the full spp")
    ap.add_argument("--spp", type=int, default=0, help="override SamplesPerPixel (development only; invalid as a headline number)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end leg (default max(5, --steps); heavy configs: 1-2)")
    ap.add_argument("--no-other-scaling", action="store_true", help="N > 1: skip the extra measurement of the other scaling mode")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from ptsharp_b200.bindings import Device, HostWorld

    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libptgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    hw = HostWorld()
    t0 = time.perf_counter()
    cfg = build_scene(hw, args.workload, args.spp or None)
    flat = hw.flatten()
    build_s = time.perf_counter() - t0
    dev = Device(local_rank)
    dev.upload_flat(flat)
    W, H, spp = cfg.width, cfg.height, cfg.spp
    npix = W * H
    # all device work of the bench goes on one explicit (non-default) stream: libptgpu launches on the stream handle it is
    # given, torch ops / NCCL / the timing events are issued under the same stream context
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    d_sum = torch.zeros(npix * 3, dtype=torch.float32, device="cuda")

    from ptsharp_b200 import distributed as D

    def split(mode):
        # strong: rank r draws global samples r, r + N, ... of the config's spp; weak: rank r draws [r*spp, (r+1)*spp)
        return D.interleaved_split(spp, rank, world_size) if mode == "strong" else D.blocked_split(spp, rank, world_size)

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(mode, warmup, steps, first_pass):
        """`steps` passes, device-timed with CUDA events between barriers; max over ranks, counters summed over ranks."""
        rs = split(mode)
        for s in range(warmup):
            D.render_pass_distributed(dev, hw, W, H, rs, d_sum, stream, pass_index=first_pass + s, rank=rank)
        barrier()
        dev.reset_counters()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for s in range(steps):
            D.render_pass_distributed(dev, hw, W, H, rs, d_sum, stream, pass_index=first_pass + warmup + s, rank=rank)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        cnt = dev.counters()
        t = torch.tensor([ms, float(cnt["cameraSamples"]), float(cnt["segments"]), float(cnt["shadowRays"]), float(cnt["kernelLaunches"]),
                          float(cnt["queueOverflows"])], dtype=torch.float64, device="cuda")
        if world_size > 1:
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            ms = float(tmax[0]); rest = [float(tsum[i]) for i in range(1, 6)]
        else:
            rest = [float(t[i]) for i in range(1, 6)]
        return ms, rest

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms, (samples, segs, shadows, launches, overflows) = timed(args.scaling, args.warmup, args.steps, 0)
    clk = clocks.stop()
    if overflows:
        raise SystemExit("a ray queue overflowed inside the timed region: the number would be invalid")
    sec = ms / 1e3
    value = samples / sec / 1e6
    gpb = segs / sec / 1e9
    other = None
    if world_size > 1 and not args.no_other_scaling:  # the other scaling mode, a short measurement beside the headline
        mode2 = "weak" if args.scaling == "strong" else "strong"
        ms2, (samples2, segs2, _, _, _) = timed(mode2, 1, max(1, min(args.steps, 2)), 1000)
        other = {"scaling": mode2, "value": samples2 / (ms2 / 1e3) / 1e6, "unit": "Msamples/s", "gpaths_bounce_per_s": segs2 / (ms2 / 1e3) / 1e9,
                 "ms_per_step": ms2 / max(1, min(args.steps, 2)), "spp_per_gpu": split(mode2).spp}

    # ---- roofline of the dominant kernel: one profiled pass with CUDA events around every launch of it ------------------
    roofline = None
    stage = None
    if rank == 0:
        peak, peak_kind = measured_peaks()
        dev.set_profiling(True)
        dev.reset_counters()
        prof_spp = max(1, min(spp, 64))
        dev.render_pass(hw.make_pass(W, H, prof_spp, pass_index=10_000), want_mean=False)
        pc = dev.counters()
        dev.set_profiling(False)
        dev.reset_buffer()
        stage = {k: pc[k] for k in ("raygenMs", "traceMs", "shadeMs", "shadowMs", "meshMs")}
        total_stage = (pc["raygenMs"] + pc["traceMs"] + pc["shadeMs"] + pc["shadowMs"]) or 1.0
        pipeline_bytes = segs * B_PER_BOUNCE_NEE + samples * B_PER_SAMPLE
        # dominant kernel: the consumer of the deferred work items that takes the most time - k_mesh (Mesh.Intersect of every ray
        # that enters a mesh box), k_march<SDF> / k_march<VOLUME> (the marching loops) - for trace AND shadow rays.  Algorithmic HBM
        # bytes per work item: the 48-byte item in, the 12-byte Hit (T, triangle) out; the kd nodes, triangles, SDF programs and
        # voxels it reads are scene data, reported as measured traffic, not counted as algorithmic.  Scenes without deferred shapes
        # (only analytic primitives): k_scene_trace<START>, a streaming kernel reading a 32-byte ray and writing a 24-byte hit.
        kinds = [("k_mesh", pc["meshMs"], pc["meshItems"], pc["meshLaunches"]), ("k_march<SDF>", pc["sdfMs"], pc["sdfItems"], pc["sdfLaunches"]),
                 ("k_march<VOLUME>", pc["volumeMs"], pc["volumeItems"], pc["volumeLaunches"])]
        kname, kms, units, klaunches = max(kinds, key=lambda k: k[1])
        unit_bytes = MESH_BYTES_PER_ITEM
        # scenes of analytic primitives (C1, C2) spend most of the pass in k_shade (Hit.Info, Ray.Bounce, sampleLight ray generation; its time
        # here includes the three shade-order kernels): per hit record it reads the 52-byte ray record and the 24-byte hit, and writes a
        # 52-byte record per child and a 48-byte record per shadow ray (averages of the profiled pass)
        scene_ms = pc["traceMs"] + pc["shadowMs"] - pc["meshMs"] - pc["sdfMs"] - pc["volumeMs"]
        if pc["shadeMs"] > kms and pc["shadeMs"] >= scene_ms and pc["segments"]:
            kname, kms, units, klaunches = "k_shade", pc["shadeMs"], pc["segments"], pc["traceLaunches"]
            unit_bytes = 52 + 24 + 52 * max(0, pc["segments"] - pc["cameraSamples"]) / pc["segments"] + 48 * pc["shadowRays"] / pc["segments"]
        elif kms <= 0:
            kname, unit_bytes, units, kms, klaunches = "k_scene_trace<START>", TRACE_BYTES_PER_RAY, pc["segments"], pc["traceMs"], pc["traceLaunches"]
        elif scene_ms > kms:
            # instanced scenes (C4): the scene level - every k_scene_trace / k_scene_shadow launch (START, RESUME, FINISH) - outweighs
            # the consumer kernels.  Unit = a ray traced (path segment or shadow ray): 32-byte ray in, 24-byte hit out; its per-round
            # state and work items are traffic of the split, not algorithmic.  Launches: one START per depth and ray kind (the RESUME
            # rounds are folded into that launch's time).
            kname, unit_bytes, units, kms, klaunches = ("k_scene_trace+k_scene_shadow (scene level)", TRACE_BYTES_PER_RAY,
                                                        pc["segments"] + pc["shadowRays"], scene_ms, 2 * pc["traceLaunches"])
        stage.update({"sdfMs": pc["sdfMs"], "volumeMs": pc["volumeMs"]})
        achieved = units * unit_bytes / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        traffic, traffic_src = None, None
        if kname == "k_mesh" and args.workload == "c3":
            for name in TRAFFIC_FILES:
                tpath = os.path.join(ROOT, "profiles", name)
                if os.path.exists(tpath):
                    with open(tpath) as f:
                        tj = json.load(f)  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
                    # per launch like `achieved`: the captured launch's DRAM bytes per item x the items of an average launch here
                    traffic = tj["bytes_per_item"] * units / max(klaunches, 1)
                    traffic_src = (f"profiles/{name}: ncu --set full capture of one k_mesh launch ({tj.get('captured_on', 'round 1 code')}), "
                                   "dram read+write per item x items per launch of this run; not re-captured by this run")
                    break
        roofline = {
            "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": units * unit_bytes / max(klaunches, 1), "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
            "algorithmic_bytes_per_unit": unit_bytes, "units_in_profiled_pass": units, "launches_in_profiled_pass": klaunches,
            "avg_launch_ms": kms / max(klaunches, 1), "kernel_ms_in_profiled_pass": kms, "share_of_step": kms / total_stage,
            "profiled_pass_spp": prof_spp,
            "pipeline": {"bytes_per_bounce": B_PER_BOUNCE_NEE, "bytes_per_sample": B_PER_SAMPLE,
                         "achieved": pipeline_bytes / sec / 1e9, "frac": pipeline_bytes / sec / 1e9 / peak},
            "note": "latency/divergence-bound kd-tree walk: HBM traffic is the streamed queues plus the part of the scene working "
                    "set that misses the 126 MB L2; see profiles/ for SIMT efficiency, issue utilisation and hit rates",
        }

    # ---- e2e through the C ABI with host buffers ------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        out = np.empty((H, W, 3), np.float32)
        host_out = torch.from_numpy(out).pin_memory().numpy()
        scene_bytes = dev.scene_bytes()  # everything ptgpu_upload_scene copies host -> device: the flat scene and the records derived from it
        n_e2e = args.e2e_steps if args.e2e_steps > 0 else max(5, args.steps)
        rs = split(args.scaling)
        dev.reset_buffer()
        barrier()
        s_before = dev.counters()["cameraSamples"]
        t0 = time.perf_counter()
        for s in range(n_e2e):
            dev.upload_flat(flat)                                   # host -> device: the whole flat scene, on every rank
            if world_size == 1:
                p = hw.make_pass(W, H, spp, pass_index=20_000 + s)
                dev.render_pass(p, out=host_out)                     # device -> host: this pass's mean image
            else:
                D.render_pass_distributed(dev, hw, W, H, rs, d_sum, stream, pass_index=20_000 + s, rank=rank)
                if rank == 0:
                    torch.cuda.synchronize()
                    host_out[...] = dev.read_buffer(W, H, 0)         # device -> host: the Buffer's colour channel
        barrier()
        dt = time.perf_counter() - t0
        e2e_samples = dev.counters()["cameraSamples"] - s_before
        tt = torch.tensor([dt, float(e2e_samples)], dtype=torch.float64, device="cuda")
        if world_size > 1:
            a = tt.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX); dt = float(a[0])
            b = tt.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM); e2e_samples = float(b[1])
        e2e = {"value": e2e_samples / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(scene_bytes + 168) * world_size,
               "d2h_bytes_per_step": int(npix * 3 * 4), "steps": n_e2e, "flat_scene_bytes": int(hw.flat_bytes()),
               "api": "ptgpu_upload_scene + ptgpu_render_pass (what Renderer.RenderParallel calls), host buffers, wall clock" if world_size == 1 else
                      "per rank: ptgpu_upload_scene + ptgpu_accumulate_device, one NCCL reduce, rank 0: ptgpu_add_sample_device + ptgpu_read_buffer; wall clock"}

    base = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline(args.workload, seconds_target=args.cpu_seconds, spp_override=args.spp or None)
        base = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "gpaths_bounce_per_s", "segments_per_sample")}

    if rank == 0:
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32 vectors + f64 scalars (the reference's Vector/double model)", "data": "synthetic",
            "config": config_dict(cfg, args.workload, args.scaling),
            "parallelism": f"spp-split x{world_size} ({args.scaling}: {split(args.scaling).spp} spp per GPU per step), scene replicated, one NCCL reduce per pass",
            "l2": "inputs larger than L2 (scene + streamed queues)", "scene_build_s": round(build_s, 2),
            "gpaths_bounce_per_s": gpb, "shadow_rays_per_s": shadows / sec, "rays_per_s": (segs + shadows) / sec,
            "segments_per_sample": segs / max(samples, 1.0),
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "stage_ms_profiled_pass": stage,
            "e2e": e2e, "cpu_baseline": base, "other_scaling": other,
        }
        if base:
            line["vs_cpu_baseline"] = {"msamples_ratio": value / base["value"], "path_bounce_ratio": gpb / base["gpaths_bounce_per_s"],
                                       "note": "GPU value / CPU restatement on this box's host cores; tracks the core count"}
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
