"""Development: print the rays on which the product library and the no-cull checker build disagree (see test_cull_matches_no_cull)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib.util
spec = importlib.util.spec_from_file_location("tg", os.path.join(ROOT, "tests", "test_gpu_parity.py")); tg = importlib.util.module_from_spec(spec); spec.loader.exec_module(tg)
from oracle import orc
from ptsharp_b200 import bindings
np.set_printoptions(precision=9, linewidth=200)
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
total = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
hw, ow, _ = tg._worlds(orc, bindings, name)
dev = bindings.Device(0); dev.upload(hw)
arb = bindings.Device(0, lib=bindings.checker_lib("nocull")); arb.upload(hw)
o0, d0 = tg._ray_batch(ow, W=160, H=120, n_secondary=1, seed=1)
h0 = ow.intersect_batch(o0, d0); ok = h0["shape"] >= 0
hits = (h0["position"][ok], h0["normal"][ok])
scale = 3.0 if name == "c3" else 25.0
rng = np.random.default_rng(2024)
done = 0
while done < total:
    n = min(5_000_000, total - done)
    o, d = tg._fuzz_rays(rng, n, hits, scale)
    g, a = dev.intersect_batch(o, d, full=False), arb.intersect_batch(o, d, full=False)
    bad = (g["shape"] != a["shape"]) | (g["prim"] != a["prim"]) | ((g["t"].view(np.int64) != a["t"].view(np.int64)) & (a["shape"] >= 0))
    for i in np.flatnonzero(bad)[:10]:
        c = ow.intersect_batch(o[i:i + 1], d[i:i + 1])
        print("ray", done + i, "family", i // (n // 6), "o", o[i], "d", d[i])
        print("   product ", g["shape"][i], g["prim"][i], repr(g["t"][i]))
        print("   no-cull ", a["shape"][i], a["prim"][i], repr(a["t"][i]))
        print("   oracle  ", c["shape"][0], c["prim"][0], repr(c["t"][0]))
    print("chunk at", done, "mismatches", int(bad.sum()), flush=True)
    done += n
