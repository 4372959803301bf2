"""One pass of a BASELINE config at full size with per-stage timing.  Usage: python tools/profile_cfg.py c4 [spp]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
name = sys.argv[1]; spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hw = HostWorld(); cfg = scenes.BUILDERS[name](hw); flat = hw.flatten()
dev = Device(0); dev.upload_flat(flat)
W, H = cfg.width, cfg.height
for i in range(2):
    dev.reset_counters()
    dev.render_pass(hw.make_pass(W, H, spp, pass_index=i), want_mean=False)
    c = dev.counters()
    print(f"pass {i}: {c['lastPassMs']:.2f} ms, {c['cameraSamples']/c['lastPassMs']/1e3:.2f} Msamples/s, {c['segments']/c['lastPassMs']/1e6:.4f} Gseg/s, launches {c['kernelLaunches']}", flush=True)
dev.set_profiling(True)
dev.render_pass(hw.make_pass(W, H, spp, pass_index=9), want_mean=False)
c = dev.counters()
print("stage ms:", {k: round(c[k], 2) for k in ('raygenMs', 'traceMs', 'shadeMs', 'shadowMs', 'meshMs')}, "mesh items", c['meshItems'], "mesh launches", c['meshLaunches'])
