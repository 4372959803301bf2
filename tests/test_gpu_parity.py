"""GPU-vs-oracle parity, through the C ABI (ptsharp_b200.bindings.Device -> libptgpu.so).

Bars (BASELINE.json north_star, tightened where the design allows):
  * closest hit on fixed ray batches: shape / triangle IDs bit-exact; t, normal and position bit-exact too (the device
    mirrors the reference's FP32-storage / FP64-scalar arithmetic), which is stricter than the 1e-5 relative asked for;
  * one camera sample per pixel with the keyed Philox stream on both sides: per-pixel radiance within 1e-4 relative for
    >= 99.9 % of pixels (libm ulp differences may flip a stochastic branch in a handful), identical ray counts +-0.1 %;
  * converged images with independent RNGs: mean relative error < 1 %, no pixel outside 4 sigma + a small floor.
"""
import os

import numpy as np
import pytest

from ptsharp_b200 import scenes

pytestmark = pytest.mark.gpu

SMALL = {
    "c1": (scenes.build_c1, {}),
    "c2": (scenes.build_c2, {}),
    "c3": (scenes.build_c3, dict(freq_a=30, freq_b=16)),
    "c4": (scenes.build_c4, dict(freq=10, nx=5, nz=3, tex=64)),
    "c5": (scenes.build_c5, dict(volume_n=24)),
    "c5_nosdf": (scenes.build_c5, dict(volume_n=24, with_sdf=False)),
}


@pytest.fixture(scope="module")
def device(bindings):
    dev = bindings.Device(0)
    yield dev
    dev.close()


def _worlds(orc, bindings, name):
    builder, kw = SMALL[name]
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    cfg = builder(hw, **kw)
    builder(ow, **kw)
    return hw, ow, cfg


def _ray_batch(ow, W=128, H=96, n_secondary=12000, seed=3):
    """Camera rays through pixel centres plus random rays leaving the surfaces they hit (exactly on the surface: the
    self-intersection regime of SURVEY F4)."""
    xs, ys = np.meshgrid(np.arange(W), np.arange(H))
    xs, ys = xs.ravel(), ys.ravel()
    o, d = ow.cast_rays(W, H, xs, ys, np.full(xs.shape, 0.5), np.full(xs.shape, 0.5), np.zeros(xs.shape, int))
    hit = ow.intersect_batch(o, d)
    ok = np.flatnonzero(hit["shape"] >= 0)
    rng = np.random.default_rng(seed)
    idx = rng.choice(ok, size=n_secondary, replace=True)
    d2 = rng.normal(size=(n_secondary, 3))
    d2 /= np.linalg.norm(d2, axis=1, keepdims=True)
    return np.concatenate([o, hit["position"][idx]]), np.concatenate([d, d2.astype(np.float32)])


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
def test_closest_hit_bit_exact(orc, bindings, device, name):
    hw, ow, _ = _worlds(orc, bindings, name)
    device.upload(hw)
    o, d = _ray_batch(ow)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert hit.sum() > 1000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["inside"], c["inside"])
    # normal-mapped (c4) included: texels are held as doubles like the reference's Colour, so the perturbed normal is the same bits
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))


def test_closest_hit_large_random_batch(orc, bindings, device):
    """150k random rays against the mesh scene — near, inside, grazing and very far origins — to exercise the padded
    subtree bounds (bounds_hit) and the division-free triangle filter against the reference sequence."""
    hw, ow, _ = _worlds(orc, bindings, "c3")
    device.upload(hw)
    rng = np.random.default_rng(11)
    n = 50000
    def dirs(k):
        v = rng.normal(size=(k, 3)); return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    # (a) origins in a box around the meshes, random directions
    oa = (rng.random((n, 3)) * [5.0, 3.0, 4.0] + [-1.5, -0.2, -2.0]).astype(np.float32); da = dirs(n)
    # (b) far origins aimed at points near the big mesh (|o| ~ 1e3: the origin-dependent padding)
    tgt = (rng.normal(size=(n, 3)) * 0.6 + [0, 1, 0])
    ob = (dirs(n).astype(np.float64) * 1000.0 + [0, 1, 0]).astype(np.float32)
    db = tgt - ob; db = (db / np.linalg.norm(db, axis=1, keepdims=True)).astype(np.float32)
    # (c) rays starting on the surface, grazing along it
    hit0 = ow.intersect_batch(oa, da)
    ok = np.flatnonzero(hit0["shape"] >= 0)
    idx = rng.choice(ok, n, replace=True)
    nrm = hit0["normal"][idx].astype(np.float64)
    tang = np.cross(nrm, dirs(n)); tang /= np.maximum(np.linalg.norm(tang, axis=1, keepdims=True), 1e-9)
    oc = hit0["position"][idx]; dc = (tang + 0.02 * nrm * rng.normal(size=(n, 1))).astype(np.float32)
    dc /= np.linalg.norm(dc, axis=1, keepdims=True)
    o = np.concatenate([oa, ob, oc]); d = np.concatenate([da, db, dc.astype(np.float32)])
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert hit.sum() > 50000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))


def test_closest_hit_degenerate_rays(orc, bindings, device):
    """Axis-parallel rays (0/0 and x/0 in the slab and split tests), rays starting on a split plane, zero direction."""
    hw, ow, _ = _worlds(orc, bindings, "c3")
    device.upload(hw)
    o = np.array([[0, 1, -5], [0, 5, 0], [-5, 1, 0], [0, 1, 0], [0.3, 0.7, 0], [0, 1, -5], [1.6, 0.45, -3]], np.float32)
    d = np.array([[0, 0, 1], [0, -1, 0], [1, 0, 0], [0, 0, 1], [0, 1, 0], [0, 0, 0], [0, 0, 1]], np.float32)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    hit = c["shape"] >= 0
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))


def test_empty_batch_and_errors(bindings, device, orc):
    hw, ow, _ = _worlds(orc, bindings, "c1")
    device.upload(hw)
    out = device.intersect_batch(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert out["shape"].shape == (0,)
    fresh = bindings.Device(0)
    with pytest.raises(bindings.PtgpuError):
        fresh.render_pass(hw.make_pass(8, 8, 1))  # no scene uploaded
    fresh.close()
    bad = hw.make_pass(1, 8, 1)
    with pytest.raises(bindings.PtgpuError):
        device.render_pass(bad)


@pytest.mark.parametrize("aperture", [0.0, 0.1])
def test_cast_rays(orc, bindings, device, aperture):
    """Camera.CastRay incl. the thin-lens branch (Camera.cs:98-119)."""
    hw, ow, _ = _worlds(orc, bindings, "c1")
    if aperture > 0:
        hw.set_focus((0, 0, 0.5), aperture)
        ow.set_focus((0, 0, 0.5), aperture)
    device.upload(hw)
    rng = np.random.default_rng(5)
    n, W, H = 5000, 640, 360
    x, y = rng.integers(0, W, n), rng.integers(0, H, n)
    fu, fv = rng.random(n), rng.random(n)
    smp = rng.integers(0, 64, n)
    go, gd = device.cast_rays(hw.make_pass(W, H, 1), x, y, fu, fv, smp)
    co, cd = ow.cast_rays(W, H, x, y, fu, fv, smp)
    if aperture == 0:
        np.testing.assert_array_equal(go.view(np.int32), co.view(np.int32))
        np.testing.assert_array_equal(gd.view(np.int32), cd.view(np.int32))
    else:  # sin/cos of the lens angle come from two different libms: a few ulps
        np.testing.assert_allclose(go, co, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(gd, cd, rtol=2e-6, atol=1e-7)


def test_keyed_stream_addressing(orc, device):
    lib = orc.lib()
    rng = np.random.default_rng(9)
    for _ in range(40):
        seed, ps, pix, smp = (int(v) for v in rng.integers(0, 2 ** 31, 4))
        bits = int(rng.integers(0, 2 ** 16)); first = int(rng.integers(0, 4000)); depth = int(rng.integers(0, 17))
        sub = int(rng.integers(0, 200)); k = int(rng.integers(0, 12))
        want = lib.orc_keyed_draw(seed, ps, pix, smp, bits, first, depth, sub, k)
        assert device.keyed_draw(seed, ps, pix, smp, bits, first, depth, sub, k) == want
        assert 0.0 <= want < 1.0


@pytest.mark.parametrize("name,res", [("c1", (160, 120)), ("c2", (128, 128)), ("c3", (160, 90)), ("c4", (160, 90)), ("c5", (96, 54)),
                                      ("c5_nosdf", (128, 72))])
def test_replay_one_sample_per_pixel(orc, bindings, device, name, res):
    """Same Philox stream on both sides: every camera sample's radiance must agree (SURVEY A.2 linearity argument)."""
    hw, ow, _ = _worlds(orc, bindings, name)
    device.upload(hw)
    W, H = res
    img = device.render_pass(hw.make_pass(W, H, 1, pass_index=3)).astype(np.float64)
    ref, _, ocnt = ow.render(W, H, 1, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED, seed=0x50545348)
    # oracle pass index is 0 for its first pass: render the GPU with the same key
    device.reset_counters()
    img0 = device.render_pass(hw.make_pass(W, H, 1, pass_index=0)).astype(np.float64)
    cnt = device.counters()
    rel = np.abs(img0 - ref) / np.maximum(np.abs(ref), 1e-3)
    frac_bad = (rel.max(axis=2) > 1e-4).mean()
    assert frac_bad < 1e-3, f"{name}: {frac_bad:.5f} of pixels differ"
    assert abs(img0.mean() - ref.mean()) <= 2e-3 * abs(ref.mean()) + 1e-9
    assert cnt["cameraSamples"] == W * H
    # same stream => same path trees, up to the odd stochastic branch flipped by a libm ulp
    assert abs(cnt["segments"] - ocnt["segments"]) <= 5e-4 * ocnt["segments"] + 4
    assert abs(cnt["shadowRays"] - ocnt["shadowRays"]) <= 5e-4 * ocnt["shadowRays"] + 4
    assert not np.array_equal(img, img0)  # pass index is part of the key
    assert cnt["nanSamples"] == 0


def test_replay_counts_identical(orc, bindings, device):
    hw, ow, _ = _worlds(orc, bindings, "c2")
    device.upload(hw)
    device.reset_counters()
    device.render_pass(hw.make_pass(96, 96, 2, pass_index=0), want_mean=False)
    cnt = device.counters()
    _, _, ocnt = ow.render(96, 96, 2, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED)
    assert cnt["segments"] == ocnt["segments"]
    assert cnt["shadowRays"] == ocnt["shadowRays"]


_ORDER_SNIPPET = """
import sys, numpy as np
sys.path.insert(0, {root!r})
from ptsharp_b200 import scenes
from ptsharp_b200.bindings import HostWorld, Device
hw = HostWorld()
scenes.build_c3(hw, freq_a=12, freq_b=6)
dev = Device(0)
dev.upload(hw)
img = dev.render_pass(hw.make_pass(160, 90, 4, pass_index=3))
c = dev.counters()
np.save({out!r}, img)
print(c["segments"], c["shadowRays"], c["cameraSamples"])
"""


def test_shade_order_does_not_change_the_pass(tmp_path):
    """k_shade visits the hit records of a launch bin by bin (k_bin_count / k_bin_scan / k_bin_scatter, csrc/ptgpu.cu).  Every
    draw is keyed by its place in the path tree, so the pass must not depend on that order: the same pass with the order switched
    off (PTGPU_SHADE_ORDER=0, read when the library first renders, hence the two processes) gives the same ray counts and the same
    image up to the rounding of the unordered float additions into the pass accumulator."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs, counts = [], []
    for flag in ("1", "0"):
        out = str(tmp_path / f"order{flag}.npy")
        env = dict(os.environ, PTGPU_SHADE_ORDER=flag)
        r = subprocess.run([sys.executable, "-c", _ORDER_SNIPPET.format(root=root, out=out)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(np.load(out).astype(np.float64))
        counts.append(r.stdout.split())
    assert counts[0] == counts[1]
    rel = np.abs(outs[0] - outs[1]) / np.maximum(np.abs(outs[1]), 1e-3)
    assert rel.max() < 1e-4, rel.max()


def test_stratified_branch(orc, bindings, device):
    """Renderer.cs:231-246: sppRoot^2 samples at strata centres, each its own Buffer.AddSample."""
    hw, ow, _ = _worlds(orc, bindings, "c1")
    device.upload(hw)
    W, H, spp = 64, 48, 5  # floor(sqrt(5)) = 2 -> 4 samples
    device.reset_buffer()
    device.render_pass(hw.make_pass(W, H, spp, stratified=True, pass_index=0), want_mean=False)
    mean = device.read_buffer(W, H, 0).astype(np.float64)
    var = device.read_buffer(W, H, 1).astype(np.float64)
    ns = device.read_buffer(W, H, 3)
    assert (ns == 4).all()
    ref, rvar, _ = ow.render(W, H, spp, passes=1, stratified=True, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED)
    rel = np.abs(mean - ref) / np.maximum(np.abs(ref), 1e-3)
    assert (rel.max(axis=2) > 1e-4).mean() < 2e-3
    relv = np.abs(var - rvar) / np.maximum(np.abs(rvar), 1e-3)
    assert (relv.max(axis=2) > 1e-3).mean() < 5e-3
    device.reset_buffer()


def test_adaptive_and_firefly_passes(orc, bindings, device):
    """Renderer.cs:340-468 after the main pass: AdaptiveSamples extra samples for every pixel, then up to FireflySamples
    more for pixels above the standard-deviation threshold, stopping at the first sample IsFirefly() rejects."""
    hw, ow, _ = _worlds(orc, bindings, "c1")
    device.upload(hw)
    W, H, spp, A, F, thr = 64, 48, 2, 3, 5, 0.05
    device.reset_buffer()
    device.reset_counters()
    device.render_pass(hw.make_pass(W, H, spp, pass_index=0, adaptive_samples=A, firefly_samples=F, firefly_threshold=thr), want_mean=False)
    cnt = device.counters()
    mean = device.read_buffer(W, H, 0).astype(np.float64)
    var = device.read_buffer(W, H, 1).astype(np.float64)
    ns = device.read_buffer(W, H, 3)[..., 0].astype(np.int64)
    device.reset_buffer()
    ow.set_extra(A, F, thr)
    ref, rvar, ocnt = ow.render(W, H, spp, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED)
    rns = ow.last_samples(W, H)
    ow.set_extra(0, 0, 1.0)
    assert ns.min() == 1 + A and ns.max() <= 1 + A + F and ns.max() > 1 + A  # some pixels took firefly samples
    assert (ns != rns).mean() < 0.01                                          # same pixels, same stopping points
    same = ns == rns
    rel = np.abs(mean - ref) / np.maximum(np.abs(ref), 1e-3)
    assert (rel.max(axis=2)[same] > 1e-4).mean() < 5e-3
    relv = np.abs(var - rvar) / np.maximum(np.abs(rvar), 1e-3)
    assert (relv.max(axis=2)[same] > 1e-3).mean() < 1e-2
    assert abs(cnt["cameraSamples"] - ocnt["cameraSamples"]) <= 0.01 * ocnt["cameraSamples"]


def test_serial_render_rules(orc, bindings, device):
    """The serial Render() (Renderer.cs:150-191, what IterativeRender runs when NumCPU == 1): AdaptiveSamples more samples only for
    pixels whose deviation reaches AdaptiveThreshold (`samples = AdaptiveSamples * (int)v`), then FireflySamples more for pixels above
    FireflyThreshold with fu = (x + xi) * (1.0f / w) and no IsFirefly test.  Two passes: after the first every deviation is 0."""
    hw, ow, _ = _worlds(orc, bindings, "c1")
    device.upload(hw)
    W, H, spp, A, F, athr, fthr = 64, 48, 2, 2, 3, 0.08, 0.15
    device.reset_buffer()
    device.reset_counters()
    ns_pass = []
    for i in range(2):
        device.render_pass(hw.make_pass(W, H, spp, pass_index=i, adaptive_samples=A, firefly_samples=F, firefly_threshold=fthr, serial_rules=True,
                                        adaptive_threshold=athr, adaptive_exponent=1.0), want_mean=False)
        ns_pass.append(device.read_buffer(W, H, 3)[..., 0].astype(np.int64))
    cnt = device.counters()
    mean = device.read_buffer(W, H, 0).astype(np.float64)
    var = device.read_buffer(W, H, 1).astype(np.float64)
    ns = ns_pass[1]
    device.reset_buffer()
    ow.set_extra(A, F, fthr)
    ow.set_serial(True, athr, 1.0)
    ref, rvar, ocnt = ow.render(W, H, spp, passes=2, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED)
    rns = ow.last_samples(W, H)
    ow.set_extra(0, 0, 1.0)
    ow.set_serial(False)
    assert (ns_pass[0] == 1).all()                                   # one sample in the buffer: Variance() is 0, nothing is picked
    assert set(np.unique(ns)) <= {2, 2 + A, 2 + F, 2 + A + F} and (ns == 2).any() and (ns > 2).any()
    assert (ns != rns).mean() < 0.01
    same = ns == rns
    rel = np.abs(mean - ref) / np.maximum(np.abs(ref), 1e-3)
    assert (rel.max(axis=2)[same] > 1e-4).mean() < 5e-3
    relv = np.abs(var - rvar) / np.maximum(np.abs(rvar), 1e-3)
    assert (relv.max(axis=2)[same] > 1e-3).mean() < 1e-2
    assert abs(cnt["cameraSamples"] - ocnt["cameraSamples"]) <= 0.01 * ocnt["cameraSamples"]
    with pytest.raises(bindings.PtgpuError):                         # (int)pow(v < 1, negative) is unbounded: rejected
        device.render_pass(hw.make_pass(W, H, spp, adaptive_samples=1, serial_rules=True, adaptive_exponent=-1.0), want_mean=False)
    device.reset_buffer()


def test_buffer_welford_matches_reference_formula(orc, bindings, device):
    """Buffer.AddSample over several passes (Buffer.cs:33-57) against numpy on the per-pass means."""
    hw, _, _ = _worlds(orc, bindings, "c1")
    device.upload(hw)
    device.reset_buffer()
    W, H = 48, 32
    passes = [device.render_pass(hw.make_pass(W, H, 2, pass_index=i)).astype(np.float64) for i in range(5)]
    stack = np.stack(passes)
    np.testing.assert_allclose(device.read_buffer(W, H, 0), stack.mean(axis=0), rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(device.read_buffer(W, H, 1), stack.var(axis=0, ddof=1), rtol=2e-5, atol=1e-7)
    np.testing.assert_allclose(device.read_buffer(W, H, 2), np.sqrt(stack.var(axis=0, ddof=1)), rtol=2e-5, atol=1e-6)
    assert (device.read_buffer(W, H, 3) == 5).all()
    device.reset_buffer()


def test_sample_partition_is_rank_invariant(orc, bindings, device):
    """Two 'ranks' drawing global samples {0,2,4,6} and {1,3,5,7} sum to what one rank drawing {0..7} gets (float-add
    order aside): the multi-GPU split changes nothing but where a sample is computed."""
    import torch
    hw, _, _ = _worlds(orc, bindings, "c2")
    device.upload(hw)
    W, H = 80, 60
    whole = torch.zeros(W * H * 3, device="cuda")
    parts = torch.zeros(W * H * 3, device="cuda")
    device.accumulate_device(hw.make_pass(W, H, 8, pass_index=1), whole.data_ptr())
    for r in range(2):
        device.accumulate_device(hw.make_pass(W, H, 4, pass_index=1, sample_base=r, sample_stride=2), parts.data_ptr())
    device.counters()  # synchronises the library's stream
    torch.cuda.synchronize()
    a, b = whole.cpu().numpy().astype(np.float64), parts.cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5)
    assert a.sum() > 0


def _outlier_budget(npix, passes, z=4.0):
    """How many of npix pixels may lie outside z sigma when the reference and the device agree: sigma is ESTIMATED from `passes`
    per-pass means per side, so (GPU - oracle) / sigma follows a Student t with about 2 (passes - 1) degrees of freedom, not a
    normal.  Budget = the expected count + 4 standard deviations of a Poisson count, rounded up (0.11 expected for 1728 Gaussian
    pixels; 0.2-0.4 with the t tails at 48-64 passes)."""
    from scipy import stats
    expected = 2.0 * stats.t.sf(z, 2 * (passes - 1)) * npix
    return int(np.ceil(expected + 4.0 * np.sqrt(expected))), expected


@pytest.mark.parametrize("name,res,spp,passes", [("c1", (48, 36), 8, 64), ("c2", (40, 40), 8, 64), ("c3", (48, 27), 16, 64),
                                                 ("c4", (48, 27), 8, 48), ("c5", (32, 18), 4, 48)])
def test_converged_image_statistical(orc, bindings, device, name, res, spp, passes):
    """Independent RNGs (GPU Philox vs the oracle's sequential xoshiro), BASELINE's second check on every config: the image means agree
    to 1 % (plus four standard errors of the Monte-Carlo mean at this sample count) and no pixel lies outside 4 sigma of the two
    renders' own per-pixel standard errors beyond what the t-distribution of an estimated sigma predicts (_outlier_budget: at most
    2-3 of ~1700 pixels; a systematic difference would put hundreds there).  The only floor on sigma is 1e-4 of the pixel value: the
    FP32 rounding of the device's pass accumulator, for pixels both sides render without noise."""
    hw, ow, _ = _worlds(orc, bindings, name)
    device.upload(hw)
    W, H = res
    ref, var, _ = ow.render(W, H, spp, passes=passes, threads=os.cpu_count() or 1, rng_mode=orc.RNG_SEQUENTIAL, seed=123)
    device.reset_buffer()
    for i in range(passes):
        device.render_pass(hw.make_pass(W, H, spp, pass_index=100 + i), want_mean=False)
    img = device.read_buffer(W, H, 0).astype(np.float64)
    gvar = device.read_buffer(W, H, 1).astype(np.float64)
    device.reset_buffer()
    lum_ref, lum = ref.mean(axis=2), img.mean(axis=2)
    sigma = np.sqrt((var.mean(axis=2) + gvar.mean(axis=2)) / passes)       # per pixel: standard error of (GPU - oracle), channels fully correlated (the conservative reading)
    se_mean = np.sqrt((sigma ** 2).sum()) / sigma.size / lum_ref.mean()    # of the relative difference of the image means
    mean_rel = np.abs(lum.mean() - lum_ref.mean()) / lum_ref.mean()
    assert mean_rel < 0.01 + 4 * se_mean, (mean_rel, se_mean)
    z = np.abs(lum - lum_ref) / (sigma + 1e-4 * lum_ref + 1e-12)
    budget, expected = _outlier_budget(z.size, passes)
    assert (z > 4).sum() <= budget, ((z > 4).sum(), budget, expected, np.sort(z.ravel())[-5:])
    assert (z > 8).sum() == 0, np.sort(z.ravel())[-5:]                    # and nothing far out
    per_pixel_rel = np.abs(lum - lum_ref).mean() / lum_ref.mean()
    assert per_pixel_rel < 0.01 + 2.0 * sigma.mean() / lum_ref.mean(), (per_pixel_rel, sigma.mean() / lum_ref.mean())  # the Monte-Carlo noise floor at this spp


def test_renderer_api_roundtrip(bindings, tmp_path):
    """Renderer.NewRenderer / SamplesPerPixel / IterativeRender through the C++ host mirror (Renderer.cs:35-56, 702-765)."""
    hw = bindings.HostWorld()
    scenes.build_c1(hw)
    hw.new_renderer(64, 48)
    hw.renderer_set(4)
    path = str(tmp_path / "out_{0}.ppm")
    hw.iterative_render(path, 2)
    assert os.path.getsize(str(tmp_path / "out_2.ppm")) > 64 * 48 * 3
    img = hw.renderer_image(64, 48, 0)
    assert np.isfinite(img).all() and img.mean() > 0.01
    assert (hw.renderer_image(64, 48, 3) == 2).all()
    c = hw.renderer_counters()
    assert c["cameraSamples"] == 2 * 64 * 48 * 4 and c["kernelLaunches"] > 0


def test_albedo_and_normal_channels(orc, bindings, device):
    """Channel.AlbedoChannel / NormalChannel of Buffer.Image (Buffer.cs:99-124, 222-282) against the same formulas in numpy on
    the exported FP64 means."""
    hw, _, _ = _worlds(orc, bindings, "c1")
    device.upload(hw)
    device.reset_buffer()
    W, H = 48, 40
    for i in range(2):
        device.render_pass(hw.make_pass(W, H, 2, pass_index=i), want_mean=False)
    M, _, n = device.export_buffer()
    assert M.shape == (H, W, 3) and (n == 2).all()
    albedo = device.read_buffer(W, H, 4).astype(np.float64)
    mx = M.max(axis=2, keepdims=True)
    ref = np.where(mx != 0, np.clip(M / np.where(mx != 0, mx, 1), 0, 1), 0.0)
    np.testing.assert_allclose(albedo, ref, rtol=2e-7, atol=1e-12)
    normal = device.read_buffer(W, H, 5).astype(np.float64)
    P = np.zeros((H + 2, W + 2, 3)); P[1:-1, 1:-1] = M  # pixels outside the frame are `new Pixel()`
    r, g = P[..., 0], P[..., 1]
    nx = (r[:-2, :-2] + 2 * r[1:-1, :-2] + r[2:, :-2] - r[:-2, 2:] - 2 * r[1:-1, 2:] - r[2:, 2:]) / 8.0
    ny = (g[:-2, :-2] + 2 * g[:-2, 1:-1] + g[:-2, 2:] - g[2:, :-2] - 2 * g[2:, 1:-1] - g[2:, 2:]) / 8.0
    ln = np.sqrt(nx * nx + ny * ny + 1.0)
    refn = np.stack([(nx / ln + 1) * 0.5, (ny / ln + 1) * 0.5, 1.0 / ln], axis=2)
    np.testing.assert_allclose(normal, refn, rtol=3e-7, atol=1e-7)


def test_checkpoint_round_trip_and_resume(orc, bindings, device):
    """ptgpu_export_buffer / ptgpu_import_buffer: the Welford state survives a round trip bit for bit and a resumed loop
    continues counting from it."""
    hw, _, _ = _worlds(orc, bindings, "c2")
    device.upload(hw)
    device.reset_buffer()
    W, H = 40, 32
    for i in range(3):
        device.render_pass(hw.make_pass(W, H, 2, pass_index=i), want_mean=False)
    M, V, n = device.export_buffer()
    assert (n == 3).all() and np.isfinite(M).all() and (V >= 0).all()
    other = bindings.Device(0)
    try:
        other.upload(hw)
        other.import_buffer(M, V, n)
        M2, V2, n2 = other.export_buffer()
        assert np.array_equal(M2.view(np.int64), M.view(np.int64)) and np.array_equal(V2.view(np.int64), V.view(np.int64)) and np.array_equal(n2, n)
        other.render_pass(hw.make_pass(W, H, 2, pass_index=3), want_mean=False)
        device.render_pass(hw.make_pass(W, H, 2, pass_index=3), want_mean=False)
        Ma, _, na = other.export_buffer()
        Mb, _, nb = device.export_buffer()
        assert (na == 4).all() and (nb == 4).all()
        np.testing.assert_allclose(Ma, Mb, rtol=1e-4, atol=1e-6)  # same samples; only the float atomics' order differs
    finally:
        other.close()


# ------------------------------------------------------------------------------------------------ BASELINE.json full sizes
@pytest.fixture(scope="module")
def full_c3(orc, bindings):
    """configs[2] at its real size: the 1 000 000-triangle displaced icospheres (the scene bench.py times)."""
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    cfg = scenes.build_c3(hw)
    scenes.build_c3(ow)
    assert cfg.triangles == 1_000_000 and (cfg.width, cfg.height) == (1920, 1080)
    return hw, ow, cfg


def test_full_size_c3_closest_hit_bit_exact(full_c3, device):
    """Tree.Intersect on the full 1 M-triangle kd-trees: camera rays over the whole frame plus rays leaving the surfaces in
    random directions, hit shape / triangle / T / position / normal bit for bit against the oracle."""
    hw, ow, cfg = full_c3
    device.upload(hw)
    o, d = _ray_batch(ow, W=192, H=108, n_secondary=30000, seed=17)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert (c["prim"] >= 0).sum() > 10000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))
    np.testing.assert_array_equal(g["inside"], c["inside"])


def test_full_size_c3_replay_window(orc, full_c3, device):
    """One keyed camera sample per pixel of a 240x135 frame of the full scene: per-pixel radiance and ray counts against
    the oracle on the same Philox stream."""
    hw, ow, cfg = full_c3
    device.upload(hw)
    W, H = 240, 135
    device.reset_counters()
    img = device.render_pass(hw.make_pass(W, H, 1, pass_index=0)).astype(np.float64)
    cnt = device.counters()
    ref, _, ocnt = ow.render(W, H, 1, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED, seed=0x50545348)
    rel = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
    assert (rel.max(axis=2) > 1e-4).mean() < 1e-3
    assert abs(cnt["segments"] - ocnt["segments"]) <= 5e-4 * ocnt["segments"] + 4
    assert abs(cnt["shadowRays"] - ocnt["shadowRays"]) <= 5e-4 * ocnt["shadowRays"] + 4
    assert cnt["nanSamples"] == 0


def test_full_size_c3_frame_properties(full_c3, device):
    """Size-independent properties on the full 1920x1080 frame (too large for the CPU oracle): the pass is a pure function
    of (pixel, global sample index, pass) — a re-run is bit-identical, batch size does not matter, and a 2-way sample
    partition sums to the unpartitioned pass — and the counters are consistent with the sampler's bounds."""
    import torch
    hw, _, cfg = full_c3
    device.upload(hw)
    W, H, spp = cfg.width, cfg.height, 2
    bufs = [torch.zeros(W * H * 3, device="cuda") for _ in range(3)]
    device.reset_counters()
    device.accumulate_device(hw.make_pass(W, H, spp, pass_index=5), bufs[0].data_ptr())
    cnt = device.counters()
    device.accumulate_device(hw.make_pass(W, H, spp, pass_index=5), bufs[1].data_ptr())
    for r in range(2):
        device.accumulate_device(hw.make_pass(W, H, 1, pass_index=5, sample_base=r, sample_stride=2), bufs[2].data_ptr())
    device.counters()
    torch.cuda.synchronize()
    a, b, c = (t.cpu().numpy() for t in bufs)
    assert np.isfinite(a).all() and a.min() >= 0 and a.sum() > 0
    # identical up to the order of the atomic float adds into a pixel (2 samples per pixel and a few path vertices each)
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(a, c, rtol=1e-5, atol=1e-5)
    n = W * H * spp
    assert cnt["cameraSamples"] == n
    assert n <= cnt["segments"] <= 5 * n            # NewSampler(1, 4): one camera segment + at most 4 bounces
    assert 0 < cnt["shadowRays"] <= cnt["segments"]  # LightModeRandom: at most one shadow ray per path vertex
    assert cnt["nanSamples"] == 0


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_full_size_replay_c1_c2(orc, bindings, device, name):
    """configs[0] (512x512, NewSampler(16, 4)) and configs[1] (Cornell 1024x1024, LightModeAll, 8 bounces) at their full
    frame size, one keyed camera sample per pixel: every pixel against the oracle, ray counts identical or within the odd
    libm-ulp branch flip."""
    builder = {"c1": scenes.build_c1, "c2": scenes.build_c2}[name]
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    cfg = builder(hw)
    builder(ow)
    assert (cfg.width, cfg.height) == {"c1": (512, 512), "c2": (1024, 1024)}[name]
    device.upload(hw)
    W, H = cfg.width, cfg.height
    device.reset_counters()
    img = device.render_pass(hw.make_pass(W, H, 1, pass_index=0)).astype(np.float64)
    cnt = device.counters()
    ref, _, ocnt = ow.render(W, H, 1, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED, seed=0x50545348)
    rel = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
    assert (rel.max(axis=2) > 1e-4).mean() < 1e-3
    assert cnt["cameraSamples"] == W * H
    assert abs(cnt["segments"] - ocnt["segments"]) <= 5e-4 * ocnt["segments"] + 4
    assert abs(cnt["shadowRays"] - ocnt["shadowRays"]) <= 5e-4 * ocnt["shadowRays"] + 4


@pytest.mark.parametrize("name", ["c4", "c5"])
def test_full_size_scene_closest_hit_c4_c5(orc, bindings, device, name):
    """configs[3] (200 TransformedShape instances of a 50 k-triangle mesh = 10 M triangles, 2 x 1024^2 textures + normal map)
    and configs[4] (SDF + Cylinder + 64^3 Volume) with their full scene data: closest hits bit for bit."""
    builder = {"c4": scenes.build_c4, "c5": scenes.build_c5}[name]
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    cfg = builder(hw)
    builder(ow)
    if name == "c4":
        assert cfg.triangles >= 9_500_000 and (cfg.width, cfg.height) == (3840, 2160)
    device.upload(hw)
    o, d = _ray_batch(ow, W=128, H=72, n_secondary=8000 if name == "c5" else 20000, seed=29)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert hit.sum() > 5000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["inside"], c["inside"])
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))


def test_loaded_model_renders_like_the_oracle(orc, bindings, device, tmp_path):
    """a model that came through STL.Load + FitInside + SmoothNormals (host/loaders.cpp) on the device vs the
    oracle given the same triangles and vertex normals: closest hits and interpolated normals bit for bit."""
    import struct
    tris = scenes.displaced_icosphere(12, 1.0, (0.3, 0.2, 0.1))
    path = str(tmp_path / "model.stl")
    with open(path, "wb") as f:
        f.write(b"x".ljust(80, b" ")); f.write(struct.pack("<i", len(tris)))
        for t in tris:
            f.write(struct.pack("<3f", 0, 0, 0)); f.write(t.astype("<f4").tobytes()); f.write(struct.pack("<H", 0))
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    m = hw.load_stl(path, hw.GlossyMaterial((0.8, 0.6, 0.3), 1.5, 0.2))
    hw.mesh_fit_inside(m, (-1, 0, -1), (1, 2, 1), (0.5, 0, 0.5))
    hw.mesh_smooth_normals(m)
    V, N, T = hw.mesh_triangles(m)
    om = ow.mesh(V, ow.GlossyMaterial((0.8, 0.6, 0.3), 1.5, 0.2), N=N, T=T)
    for w, s in ((hw, m), (ow, om)):
        w.add(s)
        w.add(w.plane((0, 0, 0), (0, 1, 0), w.DiffuseMaterial((0.9, 0.9, 0.9))))
        w.add(w.sphere((0, 6, 0), 1.0, w.LightMaterial((1, 1, 1), 40)))
        w.look_at((0, 2.5, -5), (0, 1, 0), (0, 1, 0), 35)
        w.sampler(1, 4)
    device.upload(hw)
    o, d = _ray_batch(ow, W=128, H=96, n_secondary=12000, seed=5)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert (c["prim"] >= 0).sum() > 3000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))


def _random_scene(w, rng, n_mesh, n_inst, n_analytic):
    """A random scene through the shared authoring verbs: small displaced icospheres (some instanced under random rotate /
    non-uniform scale / translate matrices), spheres, cubes, cylinders, transformed cylinders, a floor plane and a light."""
    from ptsharp_b200 import hostmath as hm
    mats = [w.DiffuseMaterial(tuple(rng.random(3))), w.GlossyMaterial(tuple(rng.random(3)), 1.3 + rng.random(), 0.1 * rng.random()),
            w.ClearMaterial(1.5, 0.0), w.SpecularMaterial(tuple(rng.random(3)), 2.0)]
    pick = lambda: mats[int(rng.integers(len(mats)))]
    w.add(w.plane((0, -1.0, 0), (0, 1, 0), mats[0]))
    meshes = []
    for _ in range(n_mesh):
        c = rng.uniform(-2.5, 2.5, 3); c[1] = rng.uniform(-0.5, 1.5)
        V = scenes.displaced_icosphere(int(rng.integers(2, 7)), float(rng.uniform(0.3, 0.9)), tuple(c), amplitude=float(rng.uniform(0, 0.15)), k=float(rng.uniform(3, 12)))
        m = w.mesh(V, pick())
        meshes.append(m)
        w.add(m)
    for _ in range(n_inst):
        base = scenes.displaced_icosphere(int(rng.integers(2, 5)), 0.5, (0, 0, 0), amplitude=0.1)
        m = w.mesh(base, pick())
        axis = rng.normal(size=3)
        M = hm.mul(hm.translate(hm.vec(rng.uniform(-3, 3, 3))), hm.mul(hm.rotate(hm.vec(axis), float(rng.uniform(0, 6.28))), hm.scale(hm.vec(rng.uniform(0.4, 1.8, 3)))))
        w.add(w.transformed(m, M))
    for _ in range(n_analytic):
        kind = int(rng.integers(4))
        p = rng.uniform(-3, 3, 3)
        if kind == 0: w.add(w.sphere(tuple(p), float(rng.uniform(0.2, 0.8)), pick()))
        elif kind == 1: w.add(w.cube(tuple(p - rng.uniform(0.1, 0.6, 3)), tuple(p + rng.uniform(0.1, 0.6, 3)), pick()))
        elif kind == 2: w.add(w.cylinder(float(rng.uniform(0.1, 0.5)), float(rng.uniform(-1, 0)), float(rng.uniform(0.1, 1)), pick()))
        else: w.add(w.transformed_cylinder(tuple(p), tuple(p + rng.uniform(-1, 1, 3)), float(rng.uniform(0.05, 0.3)), pick()))
    w.add(w.sphere((0, 6, 0), 0.8, w.LightMaterial((1, 1, 1), 30)))
    w.look_at((0, 2.0, -7.0), (0, 0.3, 0), (0, 1, 0), 45)
    w.sampler(1, 3)


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_scenes_closest_hit_bit_exact(orc, bindings, device, seed):
    """Randomised scenes (meshes of different sizes, instances under general affine matrices, every analytic shape, many shapes per
    Scene.tree leaf): closest hits of camera rays and of rays leaving the surfaces, bit for bit against the oracle."""
    n_mesh, n_inst, n_analytic = [(2, 0, 3), (1, 6, 4), (0, 12, 0), (5, 5, 10), (3, 0, 20), (0, 0, 12)][seed - 1]
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    _random_scene(hw, np.random.default_rng(1000 + seed), n_mesh, n_inst, n_analytic)
    _random_scene(ow, np.random.default_rng(1000 + seed), n_mesh, n_inst, n_analytic)
    device.upload(hw)
    o, d = _ray_batch(ow, W=96, H=72, n_secondary=9000, seed=seed)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert hit.sum() > 4000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))
    np.testing.assert_array_equal(g["inside"], c["inside"])
    # and one keyed camera sample per pixel through the whole sampler
    W, H = 80, 60
    device.reset_counters()
    img = device.render_pass(hw.make_pass(W, H, 1, pass_index=0)).astype(np.float64)
    cnt = device.counters()
    ref, _, ocnt = ow.render(W, H, 1, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED, seed=0x50545348)
    rel = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
    assert (rel.max(axis=2) > 1e-4).mean() < 2e-3
    assert abs(cnt["segments"] - ocnt["segments"]) <= 1e-3 * ocnt["segments"] + 4


# ------------------------------------------------------------------------------------------------ sub-paths of the shading code
def _replay_check(orc, device, hw, ow, W, H, frac=1e-3, count_tol=5e-4):
    """One keyed camera sample per pixel on both sides: per-pixel radiance and ray counts."""
    device.upload(hw)
    device.reset_counters()
    img = device.render_pass(hw.make_pass(W, H, 1, pass_index=0)).astype(np.float64)
    cnt = device.counters()
    ref, _, ocnt = ow.render(W, H, 1, passes=1, threads=os.cpu_count() or 1, rng_mode=orc.RNG_KEYED, seed=0x50545348)
    rel = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
    bad = (rel.max(axis=2) > 1e-4).mean()
    assert bad < frac, f"{bad:.5f} of pixels differ"
    assert abs(img.mean() - ref.mean()) <= 3e-3 * abs(ref.mean()) + 1e-9
    assert abs(cnt["segments"] - ocnt["segments"]) <= count_tol * ocnt["segments"] + 4
    assert abs(cnt["shadowRays"] - ocnt["shadowRays"]) <= count_tol * ocnt["shadowRays"] + 4
    assert cnt["nanSamples"] == 0 and cnt["queueOverflows"] == 0
    return img, ref


def _textured_scene(w, env_texture=True):
    """Every texture slot of Material (Material.cs:11-17, 124-138): albedo, normal map, bump map (Triangle.cs:173-186) and gloss map on
    a Mesh with texture coordinates; albedo + gloss map on a Sphere and a Cube (their own UVector formulas); an environment texture
    (Sampler.cs:177-189) seen by the rays that leave the scene."""
    n = 48
    v, u = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    height = 0.5 + 0.5 * np.sin(9 * u) * np.cos(7 * v)
    bump = w.texture(np.stack([height, height, height], axis=-1))
    g = 0.05 + 0.4 * (0.5 + 0.5 * np.sin(5 * u + 3 * v))          # Gloss is an angle in radians (Util.Cone)
    gloss = w.texture(np.stack([g, 0.5 * g, 1.5 * g], axis=-1))     # MaterialAt takes the mean of r, g, b
    albedo = w.texture(scenes.procedural_albedo(n))
    normal = w.texture(scenes.procedural_normal_map(n))
    V = scenes.displaced_icosphere(8, 1.0, (0, 1, 0), amplitude=0.03)
    T = scenes.spherical_uv(V, (0, 1, 0))
    w.add(w.mesh(V, w.GlossyMaterial((0.9, 0.9, 0.9), 1.5, 0.1, texture=albedo, bump_texture=bump, bump_multiplier=2.5, gloss_texture=gloss), T=T))
    V2 = scenes.displaced_icosphere(6, 0.6, (-2.0, 0.6, 0.5), amplitude=0.0)
    w.add(w.mesh(V2, w.GlossyMaterial((0.8, 0.8, 0.8), 1.4, 0.05, normal_texture=normal, bump_texture=bump, bump_multiplier=1.0), T=scenes.spherical_uv(V2, (-2.0, 0.6, 0.5))))
    w.add(w.sphere((2.0, 0.7, 0.3), 0.7, w.GlossyMaterial((1, 1, 1), 1.6, 0.2, texture=albedo, gloss_texture=gloss)))
    w.add(w.cube((0.5, 0.0, -2.2), (1.5, 0.8, -1.2), w.GlossyMaterial((1, 1, 1), 1.3, 0.0, texture=albedo, gloss_texture=gloss)))
    w.add(w.plane((0, 0, 0), (0, 1, 0), w.DiffuseMaterial((0.7, 0.7, 0.7))))
    w.add(w.sphere((0, 6, -1), 0.8, w.LightMaterial((1, 1, 1), 30)))
    if env_texture:
        ev, eu = np.meshgrid((np.arange(32) + 0.5) / 32, (np.arange(64) + 0.5) / 64, indexing="ij")
        sky = np.stack([0.2 + 0.6 * eu, 0.3 + 0.5 * ev, 0.9 - 0.4 * eu * ev], axis=-1)
        w.env(color=(0.1, 0.1, 0.1), texture=w.texture(sky), angle=0.7)
    else:
        w.env(color=(0.25, 0.3, 0.45))
    w.look_at((0.5, 2.2, -6.0), (0, 0.8, 0), (0, 1, 0), 45)
    w.sampler(1, 4)


def test_texture_paths_bump_gloss_environment(orc, bindings, device):
    """Bump texture (Triangle.cs:173-186, Texture.BumpSample), gloss texture (Material.cs:131-135), normal + bump together and the
    environment texture (Sampler.cs:177-189): perturbed normals of the closest hits against the oracle bit for bit, then every camera
    sample of a keyed replay."""
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    _textured_scene(hw); _textured_scene(ow)
    device.upload(hw)
    o, d = _ray_batch(ow, W=128, H=96, n_secondary=12000, seed=21)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert (c["prim"] >= 0).sum() > 3000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))  # texels are doubles on both sides
    np.testing.assert_array_equal(g["material"], c["material"])
    img, ref = _replay_check(orc, device, hw, ow, 160, 120, frac=2e-3)
    xs, ys = np.meshgrid(np.arange(160), np.arange(120))
    oc, dc = ow.cast_rays(160, 120, xs.ravel(), ys.ravel(), np.full(xs.size, 0.5), np.full(xs.size, 0.5), np.zeros(xs.size, int))
    miss = ow.intersect_batch(oc, dc)["shape"] < 0
    assert miss.sum() > 500                       # pixels that see the environment texture directly
    sky = ref.reshape(-1, 3)[miss]
    assert sky.std(axis=0).max() > 0.02           # ... and it is a texture, not Scene.Color
    # the same scene with a constant environment: the Scene.Color branch
    hw2, ow2 = bindings.HostWorld(), orc.OracleWorld()
    _textured_scene(hw2, env_texture=False); _textured_scene(ow2, env_texture=False)
    _replay_check(orc, device, hw2, ow2, 96, 72, frac=2e-3)


def test_replay_specular_mode_first(orc, bindings, device):
    """SpecularModeFirst (Sampler.cs:83-87; Example.cs:208, 353, 959, 1098, 1360): the camera vertex takes BOTH a diffuse and a
    specular bounce, later vertices one BounceTypeAny bounce."""
    from ptsharp_b200.authoring import SpecularModeFirst
    def build(w):
        w.add(w.plane((0, 0, 0), (0, 0, 1), w.GlossyMaterial((0.8, 0.7, 0.6), 1.3, 0.15)))
        w.add(w.sphere((0, 0, 1), 1.0, w.GlossyMaterial((0.3, 0.6, 0.9), 1.5, 0.1)))
        w.add(w.sphere((-2.2, 0.4, 0.7), 0.7, w.ClearMaterial(1.5, 0.0)))
        w.add(w.cube((1.4, -1.0, 0.0), (2.4, 0.0, 1.2), w.MetallicMaterial((0.9, 0.8, 0.3), 0.05, 0.8)))
        w.add(w.sphere((0, 0, 5.0), 1.0, w.LightMaterial((1, 1, 1), 8)))
        w.look_at((3, 3, 3), (0, 0, 0.5), (0, 0, 1), 50)
        w.sampler(4, 4, specular_mode=SpecularModeFirst)
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    build(hw); build(ow)
    device.upload(hw)
    device.reset_counters()
    _replay_check(orc, device, hw, ow, 128, 96)
    cnt = device.counters()
    assert cnt["segments"] > 128 * 96 * (1 + 0.8 * 8)  # n^2 x 2 modes = 8 children of (nearly) every camera vertex


def test_replay_with_aperture(orc, bindings, device):
    """A full keyed replay through the thin-lens branch of Camera.CastRay (Camera.cs:107-117): the lens draws come from the sample's
    stream on both sides; sin / cos of the lens angle come from two libms, so a ray may differ in its last bit."""
    hw, ow, _ = _worlds(orc, bindings, "c1")
    for w in (hw, ow):
        w.set_focus((0, 0, 1.0), 0.08)
    _replay_check(orc, device, hw, ow, 128, 96, frac=5e-3, count_tol=2e-3)
    hw3, ow3, _ = _worlds(orc, bindings, "c3")
    for w in (hw3, ow3):
        w.set_focus((0, 1, 0), 0.05)
    _replay_check(orc, device, hw3, ow3, 128, 72, frac=5e-3, count_tol=2e-3)


def test_replay_light_mode_all_two_lights_on_meshes(orc, bindings, device):
    """LightModeAll with several lights on a mesh scene (Sampler.cs:197-205: the mean over lights; one shadow ray per light and
    diffuse vertex, each with its own Philox sub-stream, through the split tracer's any-hit cut-off)."""
    from ptsharp_b200.authoring import LightModeAll
    def build(w):
        scenes.build_c3(w, freq_a=20, freq_b=10)
        w.add(w.sphere((4, 6, -3), 0.7, w.LightMaterial((1.0, 0.8, 0.6), 40)))
        w.add(w.cube((-3.5, 4.0, 1.0), (-2.5, 4.2, 2.0), w.LightMaterial((0.6, 0.8, 1.0), 60)))
        w.sampler(1, 4, light_mode=LightModeAll)
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    build(hw); build(ow)
    assert ow.num_lights() == 3
    device.upload(hw)
    device.reset_counters()
    _replay_check(orc, device, hw, ow, 160, 90)
    cnt = device.counters()
    assert cnt["shadowRays"] > 2 * 160 * 90  # up to three shadow rays per diffuse vertex (the lights it faces)


def test_any_hit_cut_off_matches_the_full_closest_hit_walk(orc, bindings, device):
    """The exact any-hit cut-off of the shadow rays (scene_advance<SHADOW>, k_mesh<ANYHIT>) against the checker build without it
    (libptgpu_nocull.so: every shadow ray takes the full closest-hit walk, then `hit.Shape == light`): same ray counts, same image
    up to the order of the float additions - on the mesh scene, on the instanced scene and with three lights."""
    from ptsharp_b200.authoring import LightModeAll
    arbiter = bindings.Device(0, lib=bindings.checker_lib("nocull"))
    try:
        for name, res, lights_all in (("c3", (160, 90), False), ("c3", (128, 72), True), ("c4", (160, 90), False)):
            hw, _, _ = _worlds(orc, bindings, name)
            if lights_all:
                hw.add(hw.sphere((4, 6, -3), 0.7, hw.LightMaterial((1.0, 0.8, 0.6), 40)))
                hw.add(hw.cube((-3.5, 4.0, 1.0), (-2.5, 4.2, 2.0), hw.LightMaterial((0.6, 0.8, 1.0), 60)))
                hw.sampler(1, 4, light_mode=LightModeAll)
            W, H = res
            outs, counts = [], []
            for dev in (device, arbiter):
                dev.upload(hw)
                dev.reset_counters()
                outs.append(dev.render_pass(hw.make_pass(W, H, 4, pass_index=3)).astype(np.float64))
                c = dev.counters()
                counts.append((c["segments"], c["shadowRays"], c["cameraSamples"]))
            assert counts[0] == counts[1]
            rel = np.abs(outs[0] - outs[1]) / np.maximum(np.abs(outs[1]), 1e-3)
            assert rel.max() < 1e-4, (name, rel.max())
    finally:
        arbiter.close()


def test_russian_roulette_is_unbiased(orc, bindings, device):
    """Opt-in Russian roulette (ptgpu_pass.flags bit 0; dead code in the reference, SURVEY F6, so NOT part of parity mode): fewer
    path segments, same expectation - the converged image agrees with the roulette-free one within 4 sigma of the image mean and per
    pixel within the budget of test_converged_image_statistical."""
    hw, _, _ = _worlds(orc, bindings, "c2")
    device.upload(hw)
    W, H, spp, passes = 40, 40, 16, 48
    res = []
    for rr in (False, True):
        device.reset_buffer(); device.reset_counters()
        for i in range(passes):
            device.render_pass(hw.make_pass(W, H, spp, pass_index=300 + i + (1000 if rr else 0), russian_roulette=rr), want_mean=False)
        res.append((device.read_buffer(W, H, 0).astype(np.float64).mean(axis=2), device.read_buffer(W, H, 1).astype(np.float64).mean(axis=2), device.counters()["segments"]))
    device.reset_buffer()
    (a, va, sa), (b, vb, sb) = res
    assert sb < 0.9 * sa                                  # it does terminate paths (max 8 bounces in this scene)
    sigma = np.sqrt((va + vb) / passes)
    se_mean = np.sqrt((sigma ** 2).sum()) / sigma.size
    assert abs(a.mean() - b.mean()) < 4 * se_mean + 1e-3 * a.mean(), (a.mean(), b.mean(), se_mean)
    z = np.abs(a - b) / (sigma + 1e-4 * a + 1e-12)
    budget, _ = _outlier_budget(z.size, passes)
    assert (z > 4).sum() <= budget and (z > 8).sum() == 0


def test_multi_device_handle(orc, bindings):
    """ptgpu_params.devices: one handle, N GPUs, one process (what a C# host gets).  The pass split over two devices gives the image
    and the counters of the single-device pass (same global sample indices; float-add order aside), and the Buffer on devices[0]
    counts one sample."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    hw, ow, _ = _worlds(orc, bindings, "c3")
    W, H, spp = 160, 90, 6
    one = bindings.Device(0)
    one.upload(hw)
    a = one.render_pass(hw.make_pass(W, H, spp, pass_index=2)).astype(np.float64)
    ca = one.counters()
    one.close()
    for devs in ([0, 1], [1, 0]):
        two = bindings.Device(devices=devs)
        two.upload(hw)
        b = two.render_pass(hw.make_pass(W, H, spp, pass_index=2)).astype(np.float64)
        cb = two.counters()
        assert cb["devices"] == 2
        assert (cb["cameraSamples"], cb["segments"], cb["shadowRays"]) == (ca["cameraSamples"], ca["segments"], ca["shadowRays"])
        np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5)
        assert (two.read_buffer(W, H, 3) == 1).all()
        two.close()
    with pytest.raises(bindings.PtgpuError):
        bindings.Device(devices=[0, 0])


def _fuzz_rays(rng, n, ow_hits, scale):
    """n rays in six families: (a) random origins around the scene, (b) far origins (|o| = 10 .. 1e6 x scale) aimed at the scene,
    (c) rays leaving surface points (the self-intersection regime), (d) rays grazing along the surface they start on,
    (e) axis-parallel and near-axis-parallel directions, (f) origins on the kd split / box planes of round coordinates."""
    k = n // 6
    def dirs(m):
        v = rng.normal(size=(m, 3)).astype(np.float32)
        return v / np.linalg.norm(v, axis=1, keepdims=True)
    pos, nrm = ow_hits
    oa = (rng.random((k, 3), dtype=np.float32) * 2 - 1) * scale * 1.5 + np.float32([0, 0.5 * scale, 0]); da = dirs(k)
    dist = (10.0 ** rng.uniform(1, 6, size=(k, 1))).astype(np.float32) * scale
    tgt = pos[rng.integers(0, len(pos), k)] + rng.normal(size=(k, 3)).astype(np.float32) * 0.05 * scale
    ob = tgt + dirs(k) * dist
    db = tgt - ob; db /= np.linalg.norm(db, axis=1, keepdims=True)
    ic = rng.integers(0, len(pos), k); oc = pos[ic]; dc = dirs(k)
    idd = rng.integers(0, len(pos), k); od = pos[idd]
    tang = np.cross(nrm[idd], dirs(k)); tang /= np.maximum(np.linalg.norm(tang, axis=1, keepdims=True), 1e-9)
    dd = tang + nrm[idd] * (rng.normal(size=(k, 1)) * 0.01).astype(np.float32); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    oe = (rng.random((k, 3), dtype=np.float32) * 2 - 1) * scale * 2; de = np.zeros((k, 3), np.float32)
    ax = rng.integers(0, 3, k); de[np.arange(k), ax] = rng.choice(np.float32([-1, 1]), k)
    de += (rng.random((k, 3), dtype=np.float32) < 0.5) * rng.normal(size=(k, 3)).astype(np.float32) * 1e-6
    m = n - 5 * k
    of = np.round((rng.random((m, 3), dtype=np.float32) * 2 - 1) * scale * 4) / 4; df = dirs(m)
    return np.concatenate([oa, ob, oc, od, oe, of]).astype(np.float32), np.concatenate([da, db, dc, dd.astype(np.float32), de, df]).astype(np.float32)


@pytest.mark.parametrize("name", ["c3", "c4"])
def test_cull_matches_no_cull(orc, bindings, device, name):
    """Every shortcut of the tracer (padded subtree / mesh / instance bounds, the bounds-only hierarchy under the reference leaves,
    the division-free triangle filter, walks clipped to the running best) against the SAME source compiled without them
    (libptgpu_nocull.so, -DPT_NO_CULL=1: Tree.Intersect / IntersectShapes as written, every reference leaf tested in full): closest
    hits - shape, triangle, T - bit for bit on 1e8 rays (PTGPU_FUZZ_RAYS to change) in six families including far origins up to
    1e6 scene sizes and grazing rays.  The no-cull build is itself checked against the oracle on the first chunk."""
    total = int(float(os.environ.get("PTGPU_FUZZ_RAYS", "1e8"))) // 2  # per scene
    hw, ow, _ = _worlds(orc, bindings, name)
    device.upload(hw)
    arbiter = bindings.Device(0, lib=bindings.checker_lib("nocull"))
    try:
        arbiter.upload(hw)
        o0, d0 = _ray_batch(ow, W=160, H=120, n_secondary=1, seed=1)
        h0 = ow.intersect_batch(o0, d0)
        ok = h0["shape"] >= 0
        hits = (h0["position"][ok], h0["normal"][ok])
        scale = 3.0 if name == "c3" else 25.0
        rng = np.random.default_rng(2024)
        chunk, done, nhit = 5_000_000, 0, 0
        while done < total:
            n = min(chunk, total - done)
            o, d = _fuzz_rays(rng, n, hits, scale)
            g, a = device.intersect_batch(o, d, full=False), arbiter.intersect_batch(o, d, full=False)
            bad = (g["shape"] != a["shape"]) | (g["prim"] != a["prim"]) | ((g["t"].view(np.int64) != a["t"].view(np.int64)) & (a["shape"] >= 0))
            assert not bad.any(), (name, done, int(bad.sum()), o[bad][:4], d[bad][:4], g["t"][bad][:4], a["t"][bad][:4])
            if done == 0:  # the arbiter itself against the oracle (CPU: a slice)
                c = ow.intersect_batch(o[::50], d[::50])
                np.testing.assert_array_equal(a["shape"][::50], c["shape"])
                np.testing.assert_array_equal(a["prim"][::50], c["prim"])
                hit = c["shape"] >= 0
                np.testing.assert_array_equal(a["t"][::50][hit].view(np.int64), c["t"][hit].view(np.int64))
            nhit += int((a["shape"] >= 0).sum())
            done += n
        assert nhit > 0.2 * total
    finally:
        arbiter.close()


def _visible_volume_scene(w, n=24):
    """A Volume the rays actually hit.  Volume.Sample reads the grid at (x, z, z) of the point (the `y <- z` and `(z + 2) / 2` quirks of
    Volume.cs:75-78), so voxels are only reached for object z in [-1, 0): the box sits there.  Thin windows give hits, refinements
    (Volume.cs:181-190) and long empty stretches (vol_skip), directly and under a TransformedShape."""
    from ptsharp_b200 import hostmath as hm
    F = lambda x: float(np.float32(x))
    colors = [0x004358, 0x1F8A70, 0xBEDB39, 0xFFE11A, 0xFD7400]
    windows = [(F(0.2) + F(0.1) * i, F(0.2) + F(0.1) * i + F(0.02), w.GlossyMaterial(hm.hex_color(c), F(1.3), hm.radians(10))) for i, c in enumerate(colors)]
    c = (np.arange(n) + 0.5) / n * 2 - 1
    _, qy, qx = np.meshgrid(c, c, c, indexing="ij")               # index [z, y, x]; Sample() reads (x, y(z), z(z)): a blob around object z = -0.5
    data = np.exp(-3 * (qx ** 2 + (2 * qy + 1) ** 2)) * (0.85 + 0.15 * np.sin(9 * qx) * np.sin(7 * qy))
    w.add(w.volume((-1, -1, -1), (1, 1, F(-0.001)), data, 1.0, windows))
    w.add(w.transformed(w.volume((-1, -1, -1), (1, 1, F(-0.001)), data, 1.0, windows[:2]),
                        hm.mul(hm.translate(hm.vec((2.6, 0.3, 0.2))), hm.rotate((0, 0, 1), 0.4))))
    w.add(w.plane((0, 0, F(-1.2)), (0, 0, 1), w.DiffuseMaterial((0.8, 0.8, 0.8))))
    w.add(w.sphere((1, -3, 4), 0.7, w.LightMaterial((1, 1, 1), 60)))
    w.look_at((1.2, -5.0, 2.2), (1.0, 0, -0.5), (0, 0, 1), 45)
    w.sampler(1, 3)


def test_volume_hits_refinement_and_empty_space_skipping(orc, bindings, device):
    """Volume.Intersect (Volume.cs:169-197) on a volume that is actually hit: closest hits bit for bit against the oracle (T, the
    window material, the gradient normal), the product library (vol_skip: steps that can only repeat `sign == 1` are skipped with t
    advanced in closed form) against the checker build that takes every step, and a keyed replay."""
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    _visible_volume_scene(hw); _visible_volume_scene(ow)
    device.upload(hw)
    o, d = _ray_batch(ow, W=160, H=120, n_secondary=15000, seed=31)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert ((c["shape"] == 0) | (c["shape"] == 1)).sum() > 3000          # rays that end on one of the two volumes
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))
    np.testing.assert_array_equal(g["material"], c["material"])
    arbiter = bindings.Device(0, lib=bindings.checker_lib("nocull"))
    try:
        arbiter.upload(hw)
        rng = np.random.default_rng(8)
        n = 400_000
        oo = (rng.random((n, 3), dtype=np.float32) * 2 - 1) * np.float32([4, 4, 2.5]) + np.float32([1, 0, 0])
        dd = rng.normal(size=(n, 3)).astype(np.float32); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        a, b = device.intersect_batch(oo, dd, full=False), arbiter.intersect_batch(oo, dd, full=False)
        np.testing.assert_array_equal(a["shape"], b["shape"])
        np.testing.assert_array_equal(a["t"].view(np.int64), b["t"].view(np.int64))
        assert (b["shape"] <= 1).sum() > 20000
    finally:
        arbiter.close()
    _replay_check(orc, device, hw, ow, 128, 96, frac=2e-3)


def test_marching_cubes_mesh_and_spherical_harmonic_render_like_the_oracle(orc, bindings, device):
    """SURVEY 8f rank 4, second half.  (1) MC.NewSDFMesh of a carved SDF (host/mc.cpp) on the device vs the oracle given the same
    triangles as a plain Mesh: closest hits bit for bit.  (2) SphericalHarmonic (SH.cs): Intersect walks the marching-cubes mesh but
    the Hit names the solid - NormalAt is the gradient of |p| - |Y(p/|p|)|, MaterialAt the sign of Y, `inside` is forced false
    (Hit.cs:41-47): closest hits incl. normals and materials bit for bit, directly and under a TransformedShape, then a keyed replay."""
    from ptsharp_b200 import hostmath as hm
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    sdf = hw.sdf_difference([hw.sdf_intersection([hw.sdf_sphere(0.8), hw.sdf_cube((1.2, 1.2, 1.2))]), hw.sdf_cylinder(0.3, 2.0)])
    gm = lambda w: w.GlossyMaterial((0.7, 0.8, 0.3), 1.4, 0.1)
    m = hw.mc_mesh(sdf, (-1, -1, -1), (1, 1, 1), 0.04, gm(hw))
    V, N, T = hw.mesh_triangles(m)
    assert V.shape[0] > 5000
    om = ow.mesh(V, gm(ow), N=N, T=T)
    sh_args = []
    for w in (hw, ow):
        pm, nm = w.GlossyMaterial((0.9, 0.3, 0.2), 1.4, 0.1), w.DiffuseMaterial((0.2, 0.4, 0.9))
        if w is hw:
            s1 = w.spherical_harmonic(3, 2, pm, nm, 0.03); s2 = w.spherical_harmonic(4, -1, pm, nm, 0.04)
            sh_args = [w.mesh_triangles(s1)[0], w.mesh_triangles(s2)[0]]
        else:
            s1 = w.spherical_harmonic(3, 2, pm, nm, sh_args[0]); s2 = w.spherical_harmonic(4, -1, pm, nm, sh_args[1])
        w.add(w.transformed(m if w is hw else om, hm.translate(hm.vec((-2.2, 0, 0)))))
        w.add(s1)
        w.add(w.transformed(s2, hm.mul(hm.translate(hm.vec((2.3, 0.2, 0.1))), hm.mul(hm.rotate((0, 1, 0), 0.5), hm.scale(hm.vec((1.2, 0.9, 1.1)))))))
        w.add(w.plane((0, 0, -1.1), (0, 0, 1), w.DiffuseMaterial((0.8, 0.8, 0.8))))
        w.add(w.sphere((1, -3, 5), 0.8, w.LightMaterial((1, 1, 1), 60)))
        w.look_at((0.3, -6.5, 2.0), (0, 0, 0), (0, 0, 1), 42)
        w.sampler(1, 3)
    device.upload(hw)
    o, d = _ray_batch(ow, W=192, H=108, n_secondary=20000, seed=41)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert (c["shape"] == 0).sum() > 1000 and (c["shape"] == 1).sum() > 300 and (c["shape"] == 2).sum() > 300
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])      # the triangle for the mesh, -1 for the harmonic solids
    assert (c["prim"][c["shape"] == 1] == -1).all() and (c["prim"][c["shape"] == 0] >= 0).all()
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))
    np.testing.assert_array_equal(g["inside"], c["inside"])
    np.testing.assert_array_equal(g["material"], c["material"])
    assert len(np.unique(c["material"][c["shape"] == 1])) == 2  # both lobes' materials
    _replay_check(orc, device, hw, ow, 160, 90, frac=2e-3)


def test_nested_transformed_shapes_match_the_oracle(orc, bindings, device):
    """TransformedShape.NewTransformedShape accepts any IShape (TransformedShape.cs:29-32), another TransformedShape included.  The nested
    Intersect re-measures Hit.T at every level and leaves hit.Shape = the innermost shape, hit.HitInfo = the outermost level's - so
    NormalAt / MaterialAt of the innermost shape are evaluated at a point of the FIRST shape space (TransformedShape.cs:52-58), which the
    device reproduces (nested_fold, hit_info).  Two and three levels, around a Mesh, a Sphere and a Cylinder; closest hits bit for bit
    (T, position, normal, inside, material, triangle), then a keyed replay."""
    from ptsharp_b200 import hostmath as hm, scenes
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    V = scenes.spatial_order(scenes.displaced_icosphere(8, 1.0, (0, 0, 0)), "friendly")
    for w in (hw, ow):
        gm = w.GlossyMaterial((0.8, 0.5, 0.3), 1.4, 0.15)
        dm = w.DiffuseMaterial((0.3, 0.6, 0.9))
        mesh = w.mesh(V, gm)
        inner = w.transformed(mesh, hm.mul(hm.rotate((0, 0, 1), 0.4), hm.scale(hm.vec((1.0, 0.7, 1.2)))))
        w.add(w.transformed(inner, hm.translate(hm.vec((-2.4, 0.3, 0.1)))))                                   # two levels around a Mesh
        s2 = w.transformed(w.transformed(w.sphere((0, 0, 0), 0.8, dm), hm.scale(hm.vec((1.3, 0.8, 1.0)))), hm.rotate((1, 0, 0), 0.6))
        w.add(w.transformed(s2, hm.translate(hm.vec((0.2, 0.0, 0.2)))))                                       # three levels around a Sphere
        cyl = w.transformed(w.cylinder(0.5, -0.7, 0.7, gm), hm.rotate((0, 1, 0), 0.9))
        w.add(w.transformed(cyl, hm.mul(hm.translate(hm.vec((2.5, -0.2, 0.0))), hm.scale(hm.vec((1.0, 1.4, 0.8))))))   # two levels around a Cylinder
        w.add(w.plane((0, 0, -1.3), (0, 0, 1), w.DiffuseMaterial((0.8, 0.8, 0.8))))
        w.add(w.sphere((1, -4, 6), 1.0, w.LightMaterial((1, 1, 1), 50)))
        w.look_at((0.2, -7.0, 2.2), (0, 0, 0), (0, 0, 1), 40)
        w.sampler(1, 4)
    device.upload(hw)
    o, d = _ray_batch(ow, W=160, H=90, n_secondary=20000, seed=77)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    for k in range(3):
        assert (c["shape"] == k).sum() > 300, k
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))
    np.testing.assert_array_equal(g["inside"], c["inside"])
    np.testing.assert_array_equal(g["material"], c["material"])
    _replay_check(orc, device, hw, ow, 160, 90, frac=2e-3)


def _many_shapes_world(w):
    """340 scene shapes - instanced meshes, spheres and cubes on both sides of index 255, every one of them in several Scene.tree leaves."""
    from ptsharp_b200 import hostmath as hm, scenes
    V = scenes.spatial_order(scenes.displaced_icosphere(3, 1.0, (0, 0, 0)), "friendly")
    rng = np.random.default_rng(5)
    pos = rng.uniform(-9, 9, size=(338, 3)) * np.array([1, 1, 0.25])
    kind = rng.integers(0, 3, size=338)
    gm = w.GlossyMaterial((0.7, 0.6, 0.4), 1.4, 0.2)
    dm = w.DiffuseMaterial((0.4, 0.7, 0.8))
    mesh = w.mesh(V, gm)
    for i in range(338):
        p = tuple(float(v) for v in pos[i])
        if kind[i] == 0:
            m = hm.mul(hm.translate(hm.vec(p)), hm.mul(hm.rotate((0, 0, 1), 0.37 * i), hm.scale(hm.vec((0.5 + 0.001 * i, 0.45, 0.6)))))
            w.add(w.transformed(mesh, m))
        elif kind[i] == 1:
            w.add(w.sphere(p, 0.45, dm))
        else:
            w.add(w.cube((p[0] - 0.4, p[1] - 0.4, p[2] - 0.4), (p[0] + 0.4, p[1] + 0.4, p[2] + 0.4), gm))
    w.add(w.cube((-12, -12, -3.2), (12, 12, -3.0), w.DiffuseMaterial((0.8, 0.8, 0.8))))
    w.add(w.sphere((2, -3, 9), 1.5, w.LightMaterial((1, 1, 1), 40)))
    w.look_at((0.5, -22.0, 7.0), (0, 0, 0), (0, 0, 1), 45)
    w.sampler(1, 3)
    assert (kind[255:] == 0).sum() > 10 and (kind[:255] == 0).sum() > 50


def test_candidate_mask_with_more_than_255_scene_shapes(orc, bindings, device):
    """The scene level keeps one bit per scene shape for the first 255 shapes (candidate mask + mailboxing) and one shared bit
    for all the others, which keep the per-item path.  340 scene shapes against the oracle: closest hits bit for bit and a keyed replay;
    then the product library against the no-cull arbiter build (no mask, every item of every leaf evaluated) on 2e7 fuzz rays."""
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    for w in (hw, ow):
        _many_shapes_world(w)
    device.upload(hw)
    o, d = _ray_batch(ow, W=192, H=108, n_secondary=30000, seed=11)
    g, c = device.intersect_batch(o, d), ow.intersect_batch(o, d)
    hit = c["shape"] >= 0
    assert (c["shape"][hit] >= 255).sum() > 500 and (c["shape"][hit] < 255).sum() > 2000
    np.testing.assert_array_equal(g["shape"], c["shape"])
    np.testing.assert_array_equal(g["prim"], c["prim"])
    np.testing.assert_array_equal(g["t"][hit].view(np.int64), c["t"][hit].view(np.int64))
    np.testing.assert_array_equal(g["position"][hit].view(np.int32), c["position"][hit].view(np.int32))
    np.testing.assert_array_equal(g["normal"][hit].view(np.int32), c["normal"][hit].view(np.int32))
    np.testing.assert_array_equal(g["inside"], c["inside"])
    arbiter = bindings.Device(0, lib=bindings.checker_lib("nocull"))
    try:
        arbiter.upload(hw)
        hits = (c["position"][hit], c["normal"][hit])
        rng = np.random.default_rng(77)
        total, done, nhit = int(float(os.environ.get("PTGPU_FUZZ_RAYS", "1e8"))) // 5, 0, 0
        while done < total:
            n = min(5_000_000, total - done)
            fo, fd = _fuzz_rays(rng, n, hits, 12.0)
            gg, aa = device.intersect_batch(fo, fd, full=False), arbiter.intersect_batch(fo, fd, full=False)
            bad = (gg["shape"] != aa["shape"]) | (gg["prim"] != aa["prim"]) | ((gg["t"].view(np.int64) != aa["t"].view(np.int64)) & (aa["shape"] >= 0))
            assert not bad.any(), (done, int(bad.sum()), fo[bad][:4], fd[bad][:4], gg["t"][bad][:4], aa["t"][bad][:4])
            nhit += int((aa["shape"] >= 0).sum())
            done += n
        assert nhit > 0.2 * total
    finally:
        arbiter.close()
    _replay_check(orc, device, hw, ow, 160, 90, frac=2e-3)
