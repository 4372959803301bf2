"""Development: one ray through the product library, build variants and the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C
import importlib.util
spec = importlib.util.spec_from_file_location("tg", os.path.join(ROOT, "tests", "test_gpu_parity.py")); tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
from oracle import orc
from ptsharp_b200 import bindings
name = sys.argv[1]
o = np.array([[float(x) for x in sys.argv[2:5]]], np.float32); d = np.array([[float(x) for x in sys.argv[5:8]]], np.float32)
hw, ow, _ = tg._worlds(orc, bindings, name)
c = ow.intersect_batch(o, d)
print("oracle  ", c["shape"][0], c["prim"][0], repr(c["t"][0]))
libs = {"product": None}
vdir = os.path.join(ROOT, "ptsharp_b200", "_lib", "variants")
for f in sorted(os.listdir(vdir)):
    libs[f] = bindings._bind_gpu(C.CDLL(os.path.join(vdir, f), mode=C.RTLD_LOCAL))
for k, lib in libs.items():
    dev = bindings.Device(0, lib=lib); dev.upload(hw)
    # the ray alone, and inside a batch of copies (the SCENE_FINISH path vs the k_mesh path)
    for reps in (1, 100000):
        g = dev.intersect_batch(np.repeat(o, reps, 0), np.repeat(d, reps, 0), full=False)
        print("%-28s x%-6d" % (k, reps), g["shape"][0], g["prim"][0], repr(g["t"][0]), "all equal" if (g["t"] == g["t"][0]).all() else "VARIES")
    dev.close()
