"""Developer check run on a GPU box: GPU-vs-oracle parity on small scenes + a rough timing.  Not a test, not a bench."""
import sys, time, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from oracle import orc
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device

def rays_for(ow, W, H, n_secondary=20000, seed=1):
    xs, ys = np.meshgrid(np.arange(W), np.arange(H)); xs = xs.ravel(); ys = ys.ravel()
    o, d = ow.cast_rays(W, H, xs, ys, np.full(xs.shape, 0.5), np.full(xs.shape, 0.5), np.zeros(xs.shape, int))
    hit = ow.intersect_batch(o, d)
    ok = hit['shape'] >= 0
    rng = np.random.default_rng(seed)
    idx = rng.choice(np.flatnonzero(ok), size=min(n_secondary, ok.sum()), replace=True)
    d2 = rng.normal(size=(len(idx), 3)); d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    return np.concatenate([o, hit['position'][idx]]), np.concatenate([d, d2.astype(np.float32)])

def compare_hits(a, b, name):
    same_shape = (a['shape'] == b['shape']); same_prim = (a['prim'] == b['prim'])
    hitm = a['shape'] >= 0
    tbits = (a['t'].view(np.int64) == b['t'].view(np.int64))
    nb = (a['normal'].view(np.int32) == b['normal'].view(np.int32)).all(axis=1)
    pb = (a['position'].view(np.int32) == b['position'].view(np.int32)).all(axis=1)
    print(f"[{name}] rays {len(hitm)} hits {hitm.sum()} | shape mismatches {(~same_shape).sum()} prim mismatches {(~same_prim).sum()} "
          f"| t bit-exact {(tbits|~hitm).mean():.6f} normal bit-exact {(nb|~hitm).mean():.6f} position bit-exact {(pb|~hitm).mean():.6f} inside mism {(a['inside']!=b['inside']).sum()}")
    bad = np.flatnonzero(~same_shape | ~same_prim)
    for i in bad[:5]:
        print('   mismatch ray', i, 'gpu', a['shape'][i], a['prim'][i], a['t'][i], 'orc', b['shape'][i], b['prim'][i], b['t'][i])

def check(name, builder, W, H, spp, **kw):
    hw, ow = HostWorld(), orc.OracleWorld()
    cfg = builder(hw, **kw); builder(ow, **kw)
    dev = Device()
    t = time.time(); dev.upload(hw); print(f"[{name}] flatten+upload {time.time()-t:.2f}s scene bytes {dev.scene_bytes()}")
    o, d = rays_for(ow, 160, 120)
    g = dev.intersect_batch(o, d); c = ow.intersect_batch(o, d)
    compare_hits(g, c, name)
    # replay parity: one camera sample per pixel, keyed RNG on both sides
    p = hw.make_pass(W, H, 1)
    img = dev.render_pass(p)
    cnt = dev.counters()
    ref, _, ocnt = ow.render(W, H, 1, passes=1, threads=os.cpu_count(), rng_mode=orc.RNG_KEYED)
    diff = np.abs(img.astype(np.float64) - ref); scale = np.maximum(np.abs(ref), 1e-3)
    rel = (diff / scale).max(axis=2)
    print(f"[{name}] replay spp=1: pixels rel>1e-4: {(rel>1e-4).mean():.5f}  rel>1e-2: {(rel>1e-2).mean():.5f}  mean img gpu {img.mean():.6f} orc {ref.mean():.6f}")
    print(f"[{name}] counters gpu seg {cnt['segments']} shadow {cnt['shadowRays']} | orc seg {ocnt['segments']} shadow {ocnt['shadowRays']} nan {cnt['nanSamples']}")
    # timing
    dev.reset_counters()
    p = hw.make_pass(W, H, spp, pass_index=1)
    t = time.time(); dev.render_pass(p, want_mean=False); dt = time.time() - t
    cnt = dev.counters()
    print(f"[{name}] {W}x{H} {spp}spp: {dt*1e3:.1f} ms host, {cnt['lastPassMs']:.1f} ms device | {cnt['cameraSamples']/cnt['lastPassMs']/1e3:.2f} Msamples/s "
          f"{cnt['segments']/cnt['lastPassMs']/1e6:.4f} Gseg/s shadow {cnt['shadowRays']/cnt['lastPassMs']/1e6:.4f} G/s launches {cnt['kernelLaunches']}")
    dev.set_profiling(True); dev.reset_counters()
    dev.render_pass(hw.make_pass(W, H, max(1, spp // 4), pass_index=2), want_mean=False)
    cnt = dev.counters()
    print(f"[{name}] stage ms: raygen {cnt['raygenMs']:.1f} trace {cnt['traceMs']:.1f} shade {cnt['shadeMs']:.1f} shadow {cnt['shadowMs']:.1f}")
    dev.close()

if __name__ == '__main__':
    which = sys.argv[1:] or ['c1', 'c2', 'c3s']
    if 'c1' in which: check('c1', scenes.build_c1, 256, 256, 16)
    if 'c2' in which: check('c2', scenes.build_c2, 256, 256, 64)
    if 'c3s' in which: check('c3s', scenes.build_c3, 480, 270, 16, freq_a=60, freq_b=30)
    if 'c3' in which: check('c3', scenes.build_c3, 1920, 1080, 8)
