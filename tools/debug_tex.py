"""Development: which texture slot makes the keyed replay differ from the oracle (tests/test_gpu_parity.py::test_texture_paths...)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device


def build(w, feat):
    n = 48
    v, u = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    height = 0.5 + 0.5 * np.sin(9 * u) * np.cos(7 * v)
    bump = w.texture(np.stack([height, height, height], axis=-1))
    g = 0.05 + 0.4 * (0.5 + 0.5 * np.sin(5 * u + 3 * v))
    gloss = w.texture(np.stack([g, 0.5 * g, 1.5 * g], axis=-1))
    albedo = w.texture(scenes.procedural_albedo(n))
    normal = w.texture(scenes.procedural_normal_map(n))
    kw = {}
    if "albedo" in feat: kw["texture"] = albedo
    if "bump" in feat: kw.update(bump_texture=bump, bump_multiplier=2.5)
    if "gloss" in feat: kw["gloss_texture"] = gloss
    if "normal" in feat: kw["normal_texture"] = normal
    V = scenes.displaced_icosphere(8, 1.0, (0, 1, 0), amplitude=0.03)
    w.add(w.mesh(V, w.GlossyMaterial((0.9, 0.9, 0.9), 1.5, 0.1, **kw), T=scenes.spherical_uv(V, (0, 1, 0))))
    skw = dict(texture=albedo, gloss_texture=gloss) if "sphere" in feat else {}
    w.add(w.sphere((2.0, 0.7, 0.3), 0.7, w.GlossyMaterial((1, 1, 1), 1.6, 0.2, **skw)))
    ckw = dict(texture=albedo, gloss_texture=gloss) if "cube" in feat else {}
    w.add(w.cube((0.5, 0.0, -2.2), (1.5, 0.8, -1.2), w.GlossyMaterial((1, 1, 1), 1.3, 0.0, **ckw)))
    w.add(w.plane((0, 0, 0), (0, 1, 0), w.DiffuseMaterial((0.7, 0.7, 0.7))))
    w.add(w.sphere((0, 6, -1), 0.8, w.LightMaterial((1, 1, 1), 30)))
    if "env" in feat:
        ev, eu = np.meshgrid((np.arange(32) + 0.5) / 32, (np.arange(64) + 0.5) / 64, indexing="ij")
        w.env(color=(0.1, 0.1, 0.1), texture=w.texture(np.stack([0.2 + 0.6 * eu, 0.3 + 0.5 * ev, 0.9 - 0.4 * eu * ev], axis=-1)), angle=0.7)
    else:
        w.env(color=(0.25, 0.3, 0.45))
    w.look_at((0.5, 2.2, -6.0), (0, 0.8, 0), (0, 1, 0), 45)
    w.sampler(1, int(os.environ.get("BOUNCES", "4")))


dev = Device(0)
W, H = 160, 120
FEATS = [f.split("+") if f else [] for f in os.environ.get("FEATS", ",albedo,bump,gloss,normal,sphere,cube,env").split(",")]
for feat in FEATS:
    hw, ow = HostWorld(), orc.OracleWorld()
    build(hw, feat); build(ow, feat)
    dev.upload(hw)
    dev.reset_counters()
    img = dev.render_pass(hw.make_pass(W, H, 1, pass_index=0)).astype(np.float64)
    cnt = dev.counters()
    ref, _, ocnt = ow.render(W, H, 1, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED, seed=0x50545348)
    rel = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
    bad = rel.max(axis=2) > 1e-4
    print(feat, "bad frac %.5f" % bad.mean(), "segments", cnt["segments"], ocnt["segments"], "shadow", cnt["shadowRays"], ocnt["shadowRays"],
          "median rel of bad %.2e" % (np.median(rel.max(axis=2)[bad]) if bad.any() else 0), flush=True)
    if bad.any() and os.environ.get("SHOW"):
        ys, xs = np.nonzero(bad)
        for y, x in list(zip(ys, xs))[:8]:
            print("   pixel", x, y, "gpu", img[y, x], "oracle", ref[y, x])
