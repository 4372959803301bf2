import sys, time
sys.path.insert(0, '/root/repo')
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
hw = HostWorld(); cfg = scenes.build_c3(hw); flat = hw.flatten()
dev = Device(0)
for i in range(3):
    t0 = time.perf_counter(); dev.upload_flat(flat); t1 = time.perf_counter()
    print(f"upload {i}: {t1 - t0:.3f} s")
