"""Per-round k_mesh times of one C4 pass (development): PTGPU_TRACE_DETAIL=1 python tools/c4_detail.py [spp]"""
import sys, time
sys.path.insert(0, '/root/repo')
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
hw = HostWorld(); cfg = scenes.build_c4(hw); flat = hw.flatten()
dev = Device(0); dev.upload_flat(flat)
dev.render_pass(hw.make_pass(cfg.width, cfg.height, spp, pass_index=0), want_mean=False)
dev.reset_counters()
print("== measured pass", file=sys.stderr, flush=True)
t0 = time.perf_counter(); dev.render_pass(hw.make_pass(cfg.width, cfg.height, spp, pass_index=1), want_mean=False); t1 = time.perf_counter()
c = dev.counters()
print(f"pass {c['lastPassMs']:.1f} ms wall {1e3*(t1-t0):.1f}  segments {c['segments']} shadow {c['shadowRays']} launches {c['kernelLaunches']}  Gseg/s {c['segments']/c['lastPassMs']/1e6:.3f}")
