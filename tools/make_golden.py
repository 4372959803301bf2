"""Generate tests/golden/*.npz from the CPU oracle (oracle/ — the C++ restatement of PTSharp's render path).

The reference ships no tests, golden vectors or assets and cannot be built here, so these vectors do NOT pin
the oracle to the .NET binary ("parity unpinned").  What they do: freeze the oracle's answers so that
(a) a later change to the oracle that alters any hit shows up in the CPU suite, and (b) the GPU path is compared with
answers that were committed before it ran.  Usage: python tools/make_golden.py   (deterministic; overwrites the files)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc
from ptsharp_b200 import scenes

OUT = os.path.join(ROOT, "tests", "golden")
CASES = {
    "c1": (scenes.build_c1, {}),
    "c2": (scenes.build_c2, {}),
    "c3": (scenes.build_c3, dict(freq_a=30, freq_b=16)),
    "c4": (scenes.build_c4, dict(freq=10, nx=5, nz=3, tex=64)),
    "c5": (scenes.build_c5, dict(volume_n=24)),
}


def ray_batch(ow, W=48, H=36, n_secondary=1500, seed=11):
    """Camera rays through pixel centres + random rays leaving the surfaces they hit, exactly on the surface."""
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    xs, ys = xs.ravel(), ys.ravel()
    o, d = ow.cast_rays(W, H, xs, ys, np.full(xs.shape, 0.5), np.full(xs.shape, 0.5), np.zeros(xs.shape, int))
    hit = ow.intersect_batch(o, d)
    ok = np.flatnonzero(hit["shape"] >= 0)
    rng = np.random.default_rng(seed)
    idx = rng.choice(ok, size=n_secondary, replace=True)
    d2 = rng.normal(size=(n_secondary, 3))
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    return (np.concatenate([o, hit["position"][idx]]).astype(np.float32), np.concatenate([d, d2.astype(np.float32)]).astype(np.float32))


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (builder, kw) in CASES.items():
        ow = orc.OracleWorld()
        builder(ow, **kw)
        o, d = ray_batch(ow)
        h = ow.intersect_batch(o, d)
        W, H = 40, 30
        img, _, cnt = ow.render(W, H, 1, passes=1, threads=1, rng_mode=orc.RNG_KEYED)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), o=o, d=d, shape=h["shape"].astype(np.int32), prim=h["prim"].astype(np.int32),
                            t=h["t"].astype(np.float64), normal=h["normal"].astype(np.float32), position=h["position"].astype(np.float32),
                            inside=h["inside"].astype(np.int32), replay=np.asarray(img, np.float64), replay_wh=np.array([W, H]),
                            segments=np.int64(cnt["segments"]), shadow_rays=np.int64(cnt["shadowRays"]))
        print(name, o.shape[0], "rays,", int((h["shape"] >= 0).sum()), "hits, replay", W, "x", H, "segments", cnt["segments"])


if __name__ == "__main__":
    main()
