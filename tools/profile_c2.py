"""Short C2 (Cornell box) run for ncu: python tools/profile_c2.py [spp]"""
import sys
sys.path.insert(0, '/root/repo')
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
hw = HostWorld(); cfg = scenes.build_c2(hw); dev = Device(0); dev.upload(hw)
for i in range(2):
    dev.reset_counters(); dev.render_pass(hw.make_pass(cfg.width, cfg.height, spp, pass_index=i), want_mean=False); c = dev.counters()
    print(f"pass {i}: {c['lastPassMs']:.2f} ms, {c['segments']/c['lastPassMs']/1e6:.3f} Gseg/s")
