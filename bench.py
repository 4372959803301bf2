#!/usr/bin/env python
"""bench.py — headline benchmark: Msamples/s (and Gpaths·bounce/s) of one RenderParallel pass on BASELINE config C3.

Workload (config.workload = "c3_icospheres_1m"): two displaced-icosphere meshes, 1 000 000 triangles in the reference's
kd-trees, Glossy + Clear materials, 1920x1080, SamplesPerPixel = 512, DefaultSampler.NewSampler(1, 4).  Synthetic,
generated from closed-form formulas (ptsharp_b200/scenes.py).

A step = one pass (Renderer.RenderParallel, Renderer.cs:199-338) over the whole frame.
  value  device-timed throughput, scene resident in HBM, CUDA events, max over ranks;
  e2e    the same pass through the host-side Renderer API with HOST buffers: every step re-uploads the flat scene
         (host -> device) and reads the pass's mean image back (device -> host);
  roofline       the dominant kernel (k_trace): algorithmic bytes per launch / measured launch duration vs measured HBM;
  cpu_baseline   the CPU restatement in oracle/ (kind "port": the C# reference cannot be built here) on all host cores,
                 on a bounded sample of the same workload.
Multi-GPU (torchrun): scene replicated, each rank renders its own 512 spp with disjoint global sample indices (weak
scaling), per-pass float sum buffers are reduced to rank 0 over NCCL, rank 0 does the Buffer.AddSample.
`--impl reference` times the oracle alone (rank 0 only).
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
    "c3": ("c3", {}),
    "c3_small": ("c3", dict(freq_a=60, freq_b=30, width=480, height=270, spp=16)),
    "c2": ("c2", {}),
    "c1": ("c1", {}),
}
B_PER_BOUNCE_NEE, B_PER_BOUNCE, B_PER_SAMPLE = 208, 128, 24
TRACE_BYTES_PER_RAY = 32 + 24  # k_trace: reads (o,pixel),(d,meta) = 2 x float4, writes the 24-byte hit record
MESH_BYTES_PER_ITEM = 48 + 12  # k_mesh: reads the 48-byte work item (co, ray | cd, root | tmin, tmax), writes T (8) + triangle (4)


def prof_launch_guess(pc):
    return 5


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


def cpu_baseline(workload: str, cfg_hint, seconds_target=15.0, steps=1, warmup=0):
    """Time the oracle (CPU restatement) on a bounded sample of the workload with every host core."""
    from oracle import orc
    ow = orc.OracleWorld()
    cfg = build_scene(ow, workload)
    ow.compile()
    cores = os.cpu_count() or 1
    W, H = cfg.width, cfg.height
    # bounded sample: the central quarter-size window of the frame, spp chosen from a short probe
    win = (W // 2 - W // 8, H // 2 - H // 8, W // 2 + W // 8, H // 2 + H // 8)
    t0 = time.perf_counter()
    _, _, cnt = ow.render(W, H, 1, passes=1, threads=cores, window=win, seed=7)
    probe = time.perf_counter() - t0
    spp = int(max(1, min(cfg.spp, seconds_target / max(probe, 1e-3))))
    times, cnts = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, cnts = ow.render(W, H, spp, passes=1, threads=cores, window=win, seed=11 + i)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    return {
        "value": cnts["cameraSamples"] / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
        "gpaths_bounce_per_s": cnts["segments"] / dt / 1e9, "rays_per_s": (cnts["segments"] + cnts["shadowRays"]) / dt,
        "sample": f"{workload}: central {win[2]-win[0]}x{win[3]-win[1]} window of the {W}x{H} frame, {spp} spp, 1 pass, "
                  f"{cnts['cameraSamples']} camera samples, {dt:.2f} s; oracle/ C++ restatement of RenderParallel "
                  f"(not the .NET binary), {cores} threads on a shared FIFO of 32x32 tasks",
        "seconds": dt,
    }, cfg


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, cfg = cpu_baseline(args.workload, None, seconds_target=args.cpu_seconds, steps=max(1, args.steps), warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": base["value"], "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 vectors + f64 scalars", "data": "synthetic",
        "config": {"workload": cfg.name, "width": cfg.width, "height": cfg.height, "spp_per_gpu": cfg.spp, "triangles": cfg.triangles,
                   "sampler": "DefaultSampler.NewSampler(1,4)" if args.workload.startswith("c3") else "see scenes.py",
                   "parallelism": "host CPU threads (rank 0 only)", "l2": "n/a (CPU)"},
        "gpaths_bounce_per_s": base["gpaths_bounce_per_s"],
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
    ap.add_argument("--spp", type=int, default=0, help="override SamplesPerPixel (development only; invalid as a headline number)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
    # weak scaling: rank r draws global samples [r*spp, (r+1)*spp) of every pixel; one NCCL reduce per pass
    rs = D.blocked_split(spp, rank, world_size)

    def one_step(step):
        D.render_pass_distributed(dev, hw, W, H, rs, d_sum, stream, pass_index=step, rank=rank)

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        one_step(s)
    barrier()
    dev.reset_counters()
    clocks = ClockSampler(local_rank)
    clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for s in range(args.steps):
        one_step(args.warmup + s)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    cnt = dev.counters()
    t = torch.tensor([ms, float(cnt["cameraSamples"]), float(cnt["segments"]), float(cnt["shadowRays"]), float(cnt["kernelLaunches"])],
                     dtype=torch.float64, device="cuda")
    if world_size > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms = float(tmax[0]); samples, segs, shadows, launches = (float(tsum[i]) for i in (1, 2, 3, 4))
    else:
        samples, segs, shadows, launches = (float(t[i]) for i in (1, 2, 3, 4))
    sec = ms / 1e3
    value = samples / sec / 1e6
    gpb = segs / sec / 1e9

    # ---- roofline of the dominant kernel (k_trace): one profiled pass with per-stage CUDA events ---------------------
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
        stage = {k: pc[k] for k in ("raygenMs", "traceMs", "shadeMs", "shadowMs", "meshMs")}
        total_stage = (pc["raygenMs"] + pc["traceMs"] + pc["shadeMs"] + pc["shadowMs"]) or 1.0
        pipeline_bytes = segs * B_PER_BOUNCE_NEE + samples * B_PER_SAMPLE
        if pc["meshLaunches"] > 0:
            # dominant kernel: k_mesh (Mesh.Intersect of every ray that enters a mesh box; trace AND shadow rays).
            # Algorithmic HBM bytes per work item: the 48-byte item in, the 12-byte Hit (T, triangle) out; the kd
            # nodes and triangles it walks are scene data, reported as measured traffic, not counted as algorithmic.
            kname, unit_bytes, units, kms, klaunches = "k_mesh", MESH_BYTES_PER_ITEM, pc["meshItems"], pc["meshMs"], pc["meshLaunches"]
        else:
            kname, unit_bytes, units, kms, klaunches = "k_trace", TRACE_BYTES_PER_RAY, pc["segments"], pc["traceMs"], prof_launch_guess(pc)
        achieved = units * unit_bytes / (kms / 1e3) / 1e9 if kms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_k_mesh_traffic.json")
        if kname == "k_mesh" and os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)        # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel
            # per launch like `achieved`: the captured launch's DRAM bytes per item x the items of an average launch here
            traffic = tj["bytes_per_item"] * units / max(klaunches, 1)
        roofline = {
            "bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": "profiles/r01_k_mesh_traffic.json (ncu --set full, dram read+write per item x items per launch)" if traffic else None,
            "algorithmic_bytes_per_launch": units * unit_bytes / max(klaunches, 1), "peak_source": f"MEASURED_PEAKS.json ({peak_kind})",
            "algorithmic_bytes_per_unit": unit_bytes, "units_in_profiled_pass": units, "launches_in_profiled_pass": klaunches,
            "avg_launch_ms": kms / max(klaunches, 1), "kernel_ms_in_profiled_pass": kms, "share_of_step": kms / total_stage,
            "pipeline": {"bytes_per_bounce": B_PER_BOUNCE_NEE, "bytes_per_sample": B_PER_SAMPLE,
                         "achieved": pipeline_bytes / sec / 1e9, "frac": pipeline_bytes / sec / 1e9 / peak},
            "note": "latency/divergence-bound kd-tree walk: HBM traffic is the streamed queues plus the part of the 221 MB "
                    "scene working set that misses the 126 MB L2; see profiles/ for SIMT efficiency, issue utilisation and hit rates",
        }

    # ---- e2e through the host Renderer API with host buffers ----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        out = np.empty((H, W, 3), np.float32)
        pinned = torch.from_numpy(out).pin_memory() if hasattr(torch.Tensor, "pin_memory") else None
        host_out = pinned.numpy() if pinned is not None else out
        scene_bytes = dev.scene_bytes()  # everything ptgpu_upload_scene copies host -> device: the flat scene and the records derived from it
        n_e2e = max(1, min(args.steps, 2))
        barrier()
        t0 = time.perf_counter()
        s_before = dev.counters()["cameraSamples"]
        for s in range(n_e2e):
            dev.upload_flat(flat)                                   # host -> device: the whole flat scene
            p = hw.make_pass(W, H, spp, pass_index=20_000 + s, sample_base=rank * spp)
            dev.render_pass(p, out=host_out)                         # device -> host: this pass's mean image
        barrier()
        dt = time.perf_counter() - t0
        e2e_samples = dev.counters()["cameraSamples"] - s_before
        tt = torch.tensor([dt, float(e2e_samples)], dtype=torch.float64, device="cuda")
        if world_size > 1:
            a = tt.clone(); dist.all_reduce(a, op=dist.ReduceOp.MAX); dt = float(a[0])
            b = tt.clone(); dist.all_reduce(b, op=dist.ReduceOp.SUM); e2e_samples = float(b[1])
        e2e = {"value": e2e_samples / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(scene_bytes + 128),
               "d2h_bytes_per_step": int(npix * 3 * 4), "steps": n_e2e,
               "flat_scene_bytes": int(hw.flat_bytes()),
               "api": "ptgpu_upload_scene + ptgpu_render_pass (what Renderer.RenderParallel calls), host buffers, wall clock"}

    base = None
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline(args.workload, cfg, seconds_target=args.cpu_seconds)
        base = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "gpaths_bounce_per_s")}

    if rank == 0:
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 vectors + f64 scalars (the reference's Vector/double model)", "data": "synthetic",
            "config": {"workload": cfg.name, "width": W, "height": H, "spp_per_gpu": spp, "triangles": cfg.triangles,
                       "sampler": "DefaultSampler.NewSampler(1,4)" if args.workload.startswith("c3") else "see scenes.py",
                       "parallelism": f"spp-split x{world_size}, scene replicated", "l2": "inputs larger than L2 (scene + streamed queues)",
                       "scene_build_s": round(build_s, 2)},
            "gpaths_bounce_per_s": gpb, "shadow_rays_per_s": shadows / sec, "rays_per_s": (segs + shadows) / sec,
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "stage_ms_profiled_pass": stage,
            "e2e": e2e, "cpu_baseline": base,
        }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
