"""kd-tree builder: hand-worked example of the reference's pseudo-median rule, host builder vs oracle builder, and the
leaf statistics of the builder-friendly triangle order."""
import numpy as np
import pytest

from ptsharp_b200 import scenes


def _same(a, b):
    for k in ("axis", "point", "a", "b", "items"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    np.testing.assert_array_equal(a["box"], b["box"])


def _eight_spheres(w):
    m = w.DiffuseMaterial((1, 1, 1))
    xs = [0, 10, 2, 8, 4, 6, 12, 14]  # insertion order matters
    for x in xs:
        w.add(w.sphere((x, 0, 0), 0.5, m))
    return xs


def test_pseudo_median_hand_worked(orc, bindings):
    """N = 8: the bag holds m0,M0,...,m7,M7 and enumerates in reverse, so Median takes reversed elements 7 and 8 =
    m4 and M3: point = (min_x(shape 4) + max_x(shape 3)) / 2 (Tree.cs:130-148, 208-226)."""
    for W in (orc.OracleWorld, bindings.HostWorld):
        w = W()
        xs = _eight_spheres(w)
        t = w.tree_dump(-1)
        assert t["axis"][0] == 1
        assert t["point"][0] == ((xs[4] - 0.5) + (xs[3] + 0.5)) / 2 == 6.0
        # children hold the reversed sub-sequences (ConcurrentBag.ToArray): left = x-0.5 <= 6, right = x+0.5 >= 6
        left, right = t["a"][0], t["b"][0]
        assert t["axis"][left] == 0 and t["axis"][right] == 0
        li = t["items"][t["a"][left]: t["a"][left] + t["b"][left]].tolist()
        ri = t["items"][t["a"][right]: t["a"][right] + t["b"][right]].tolist()
        assert li == [5, 4, 2, 0] and ri == [7, 6, 5, 3, 1]


def test_odd_count_uses_middle_box_centre(orc):
    w = orc.OracleWorld()
    m = w.DiffuseMaterial((1, 1, 1))
    for x in [0, 10, 2, 8, 5, 6, 12, 14, 3]:
        w.add(w.sphere((x, 0, 0), 0.5, m))
    t = w.tree_dump(-1)
    assert t["axis"][0] == 1 and t["point"][0] == 5.0  # centre of shape (N-1)/2 = 4


@pytest.mark.parametrize("name,kw", [("c1", {}), ("c2", {}), ("c3", dict(freq_a=24, freq_b=12))])
def test_host_builder_matches_oracle(orc, bindings, name, kw):
    hw, ow = bindings.HostWorld(), orc.OracleWorld()
    scenes.BUILDERS[name](hw, **kw)
    scenes.BUILDERS[name](ow, **kw)
    _same(hw.tree_dump(-1), ow.tree_dump(-1))
    if name == "c3":
        for mesh in (0, 1):
            _same(hw.tree_dump(mesh), ow.tree_dump(mesh))


def test_plane_lives_in_every_leaf(orc):
    """A Plane's box is +-1e9 (Plane.cs:33-36), so it is copied into both children of every split."""
    w = orc.OracleWorld()
    m = w.DiffuseMaterial((1, 1, 1))
    w.add(w.plane((0, 0, 0), (0, 0, 1), m))
    for i in range(12):
        w.add(w.sphere((i * 3.0, 0, 1), 1, m))
    t = w.tree_dump(-1)
    leaves = np.flatnonzero(t["axis"] == 0)
    assert len(leaves) > 1
    for leaf in leaves:
        assert 0 in t["items"][t["a"][leaf]: t["a"][leaf] + t["b"][leaf]]


def test_builder_friendly_order_quality(bindings):
    """The reference builder on a Morton-ordered mesh leaves leaves of thousands of triangles; on the
    builder-friendly order (host/host.cpp) the same, unmodified builder stays below a few hundred."""
    V = scenes.displaced_icosphere(40, 1.0, (0, 1, 0))
    stats = {}
    for mode in ("morton", "friendly"):
        w = bindings.HostWorld()
        mesh = w.mesh(scenes.spatial_order(V, mode), w.DiffuseMaterial((1, 1, 1)))
        w.add(mesh)
        t = w.tree_dump(mesh)
        sizes = t["b"][t["axis"] == 0].astype(np.float64)
        stats[mode] = dict(max=sizes.max(), wmean=(sizes ** 2).sum() / sizes.sum(), depth=w.tree_stats(mesh)["maxDepth"])
    assert stats["morton"]["max"] > 2000
    assert stats["friendly"]["max"] < 600 and stats["friendly"]["wmean"] < 40 and stats["friendly"]["depth"] < 40
    perm = bindings.builder_friendly_order(scenes.spatial_order(V, "morton"))
    assert sorted(perm.tolist()) == list(range(len(V)))
