"""Short single-GPU run for ncu: one pass of a reduced C3 (same generator, fewer triangles/pixels) so each kernel
launches a handful of times.  Usage: python tools/profile_run.py [freq_a freq_b width height spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device

a = [int(x) for x in sys.argv[1:]]
fa, fb, W, H, spp = (a + [100, 50, 960, 540, 4][len(a):])[:5]
hw = HostWorld()
cfg = scenes.build_c3(hw, freq_a=fa, freq_b=fb, width=W, height=H, spp=spp)
dev = Device(0)
dev.upload(hw)
for i in range(2):
    dev.reset_counters()
    dev.render_pass(hw.make_pass(W, H, spp, pass_index=i), want_mean=False)
    c = dev.counters()
    print(f"pass {i}: {c['lastPassMs']:.2f} ms, {c['cameraSamples']/c['lastPassMs']/1e3:.2f} Msamples/s, "
          f"{c['segments']/c['lastPassMs']/1e6:.4f} Gseg/s, shadow {c['shadowRays']/c['lastPassMs']/1e6:.4f} G/s, launches {c['kernelLaunches']}")
dev.set_profiling(True)
dev.render_pass(hw.make_pass(W, H, spp, pass_index=9), want_mean=False)
c = dev.counters()
print("stage ms:", {k: round(c[k], 2) for k in ('raygenMs', 'traceMs', 'shadeMs', 'shadowMs')}, "triangles", cfg.triangles)
