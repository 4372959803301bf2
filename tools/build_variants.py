"""Build tuning variants of libptgpu.so: python tools/build_variants.py name=-DFLAG=1,-DOTHER=2 name2=...
Output: ptsharp_b200/_lib/variants/libptgpu_<name>.so (select with PTGPU_LIB=<path>)."""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ptsharp_b200 import build as b
out_dir = os.path.join(b.LIBDIR, "variants")
os.makedirs(out_dir, exist_ok=True)
def one(spec):
    name, _, flags = spec.partition("=")
    out = os.path.join(out_dir, f"libptgpu_{name}.so")
    cmd = [b.NVCC] + b.NVCC_FLAGS + [f for f in flags.split(",") if f] + ["-o", out, os.path.join(b.PKG, "csrc", "ptgpu.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, r.stderr[-2000:]
with ThreadPoolExecutor(8) as ex:
    for name, rc, err in ex.map(one, sys.argv[1:]):
        print(name, "ok" if rc == 0 else "FAILED\n" + err)
