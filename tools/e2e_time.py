import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
hw = HostWorld(); cfg = scenes.build_c3(hw); flat = hw.flatten()
dev = Device(0); dev.upload_flat(flat)
W, H = cfg.width, cfg.height
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 128
out = np.empty((H, W, 3), np.float32)
for i in range(3):
    t0 = time.perf_counter(); dev.render_pass(hw.make_pass(W, H, spp, pass_index=i), out=out); t1 = time.perf_counter()
    c = dev.counters()
    print(f"render_pass {i}: wall {t1 - t0:.3f} s, device {c['lastPassMs']/1e3:.3f} s")
for i in range(2):
    t0 = time.perf_counter(); dev.upload_flat(flat); t1 = time.perf_counter(); dev.render_pass(hw.make_pass(W, H, spp, pass_index=10 + i), out=out); t2 = time.perf_counter()
    c = dev.counters()
    print(f"upload {t1 - t0:.3f} s + render_pass wall {t2 - t1:.3f} s, device {c['lastPassMs']/1e3:.3f} s")
