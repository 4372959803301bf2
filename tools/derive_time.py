"""Time the host-side derivation of ptgpu_upload_scene (mesh_derive.hpp) without a device: python tools/derive_time.py [c3|c4] [old.so]"""
import ctypes as C, sys, time
sys.path.insert(0, '/root/repo')
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, gpu_lib
name = sys.argv[1] if len(sys.argv) > 1 else 'c3'
hw = HostWorld(); t0 = time.perf_counter(); getattr(scenes, 'build_' + name)(hw); flat = hw.flatten(); print(f"build+flatten {time.perf_counter() - t0:.2f} s")
lib = gpu_lib()
lib.ptgpu_debug_derive.restype = C.c_double
lib.ptgpu_debug_derive.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
h, nr, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
for i in range(3):
    ms = lib.ptgpu_debug_derive(flat, C.byref(h), C.byref(nr), C.byref(nt))
    print(f"derive {ms:.1f} ms  hash {h.value:016x}  records {nr.value}  leaf triangles {nt.value}")
if len(sys.argv) > 2:
    old = C.CDLL(sys.argv[2]); old.derive_old.restype = C.c_double; old.derive_old.argtypes = lib.ptgpu_debug_derive.argtypes
    h2 = C.c_uint64()
    ms = old.derive_old(flat, C.byref(h2), C.byref(nr), C.byref(nt))
    print(f"old    {ms:.1f} ms  hash {h2.value:016x}  records {nr.value}  leaf triangles {nt.value}  {'IDENTICAL' if h2.value == h.value else 'DIFFERENT'}")
