"""Development: per-launch k_mesh times of one C3 pass (PTGPU_TRACE_DETAIL=1 prints them to stderr).  Usage: python tools/r02_detail.py [spp]"""
import os, sys
os.environ["PTGPU_TRACE_DETAIL"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
hw = HostWorld()
cfg = scenes.build_c3(hw)
dev = Device(0)
dev.upload(hw)
W, H = cfg.width, cfg.height
dev.render_pass(hw.make_pass(W, H, spp, pass_index=0), want_mean=False)
print("---- measured pass", file=sys.stderr, flush=True)
dev.set_profiling(True)
dev.render_pass(hw.make_pass(W, H, spp, pass_index=1), want_mean=False)
c = dev.counters()
print({k: round(c[k], 2) for k in ("raygenMs", "traceMs", "shadeMs", "shadowMs", "meshMs")}, c["meshItems"], c["meshLaunches"])
