"""Committed golden vectors (tests/golden/*.npz, written by tools/make_golden.py from the CPU oracle).

The reference ships no golden vectors, so these freeze the ORACLE's answers: the CPU suite checks that the
oracle still reproduces them bit for bit (hits) / to 1e-12 (the keyed-stream replay image), the GPU suite checks the
CUDA path against the same committed numbers through the C ABI."""
import os

import numpy as np
import pytest

from ptsharp_b200 import scenes

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {
    "c1": (scenes.build_c1, {}),
    "c2": (scenes.build_c2, {}),
    "c3": (scenes.build_c3, dict(freq_a=30, freq_b=16)),
    "c4": (scenes.build_c4, dict(freq=10, nx=5, nz=3, tex=64)),
    "c5": (scenes.build_c5, dict(volume_n=24)),
}


def _load(name):
    return np.load(os.path.join(GOLDEN, f"{name}.npz"))


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_hits(orc, name):
    g = _load(name)
    builder, kw = CASES[name]
    ow = orc.OracleWorld()
    builder(ow, **kw)
    h = ow.intersect_batch(g["o"], g["d"])
    hit = g["shape"] >= 0
    assert hit.sum() > 1000
    np.testing.assert_array_equal(h["shape"], g["shape"])
    np.testing.assert_array_equal(h["prim"], g["prim"])
    np.testing.assert_array_equal(h["t"][hit].view(np.int64), g["t"][hit].view(np.int64))
    np.testing.assert_array_equal(h["normal"][hit].view(np.int32), g["normal"][hit].view(np.int32))
    np.testing.assert_array_equal(h["position"][hit].view(np.int32), g["position"][hit].view(np.int32))


@pytest.mark.parametrize("name", ["c1", "c2", "c3"])
def test_oracle_reproduces_golden_replay(orc, name):
    g = _load(name)
    builder, kw = CASES[name]
    ow = orc.OracleWorld()
    builder(ow, **kw)
    W, H = (int(v) for v in g["replay_wh"])
    img, _, cnt = ow.render(W, H, 1, passes=1, threads=2, rng_mode=orc.RNG_KEYED)  # the keyed stream does not depend on threading
    np.testing.assert_allclose(np.asarray(img, np.float64), g["replay"], rtol=1e-12, atol=1e-14)
    assert cnt["segments"] == int(g["segments"]) and cnt["shadowRays"] == int(g["shadow_rays"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_matches_golden_hits(bindings, name):
    g = _load(name)
    builder, kw = CASES[name]
    hw = bindings.HostWorld()
    builder(hw, **kw)
    dev = bindings.Device(0)
    try:
        dev.upload(hw)
        h = dev.intersect_batch(g["o"], g["d"])
    finally:
        dev.close()
    hit = g["shape"] >= 0
    np.testing.assert_array_equal(h["shape"], g["shape"])
    np.testing.assert_array_equal(h["prim"], g["prim"])
    np.testing.assert_array_equal(h["t"][hit].view(np.int64), g["t"][hit].view(np.int64))
    np.testing.assert_array_equal(h["position"][hit].view(np.int32), g["position"][hit].view(np.int32))
    if name == "c4":  # FP32 texels on the device
        np.testing.assert_allclose(h["normal"][hit], g["normal"][hit], rtol=1e-5, atol=2e-6)
    else:
        np.testing.assert_array_equal(h["normal"][hit].view(np.int32), g["normal"][hit].view(np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "c2", "c3"])
def test_device_matches_golden_replay(bindings, name):
    g = _load(name)
    builder, kw = CASES[name]
    hw = bindings.HostWorld()
    builder(hw, **kw)
    W, H = (int(v) for v in g["replay_wh"])
    dev = bindings.Device(0)
    try:
        dev.upload(hw)
        dev.reset_counters()
        img = dev.render_pass(hw.make_pass(W, H, 1)).astype(np.float64)
        cnt = dev.counters()
    finally:
        dev.close()
    ref = g["replay"].reshape(img.shape)
    rel = np.abs(img - ref) / np.maximum(np.abs(ref), 1e-3)
    assert (rel > 1e-4).mean() < 2e-3, f"{(rel > 1e-4).mean():.4f} of the pixels differ"
    assert abs(cnt["segments"] - int(g["segments"])) <= max(2, int(g["segments"]) // 1000)


@pytest.mark.gpu
def test_device_matches_golden_through_flat_file(bindings, tmp_path):
    """Scene -> flat-scene file -> another process' world -> device: same hits as the committed vectors (the route a .NET
    box dumping PTSharp scenes for this library would take)."""
    g = _load("c3")
    builder, kw = CASES["c3"]
    hw = bindings.HostWorld()
    builder(hw, **kw)
    hw.flatten()
    path = str(tmp_path / "c3.ptfs")
    hw.save_flat(path)
    other = bindings.HostWorld()
    flat = other.load_flat(path)
    dev = bindings.Device(0)
    try:
        dev.upload_flat(flat)
        h = dev.intersect_batch(g["o"], g["d"])
    finally:
        dev.close()
    hit = g["shape"] >= 0
    np.testing.assert_array_equal(h["shape"], g["shape"])
    np.testing.assert_array_equal(h["prim"], g["prim"])
    np.testing.assert_array_equal(h["t"][hit].view(np.int64), g["t"][hit].view(np.int64))
