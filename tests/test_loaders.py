"""OBJ / STL loaders and Mesh utilities of the C++ host mirror (host/loaders.cpp; OBJ.cs, STL.cs, Mesh.cs:141-289), SURVEY 8f
rank 4.  The expectations below restate the reference's rules independently in numpy, quirks included."""
import os
import struct

import numpy as np
import pytest

from ptsharp_b200 import scenes


def _flat_normal(V):
    e1, e2 = V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]
    n = np.cross(e1, e2).astype(np.float32)
    return n / np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)


def _write_binary_stl(path, V):
    with open(path, "wb") as f:
        f.write(b"binary stl written by tests".ljust(80, b" "))
        f.write(struct.pack("<i", len(V)))
        for t in V:
            f.write(struct.pack("<3f", 0, 0, 0))
            f.write(t.astype("<f4").tobytes())
            f.write(struct.pack("<H", 0))


def _write_text_stl(path, V):
    with open(path, "w") as f:
        f.write("solid test\n")
        for t in V:
            f.write("  facet normal 0 0 0\n    outer loop\n")
            for p in t:
                f.write("      vertex %.9g %.9g %.9g\n" % tuple(float(x) for x in p))
            f.write("    endloop\n  endfacet\n")
        f.write("endsolid test\n")


@pytest.fixture()
def tris():
    return scenes.displaced_icosphere(4, 1.0, (0.25, 1.0, -0.5))  # 320 triangles, float32


@pytest.mark.parametrize("kind", ["binary", "text"])
def test_stl_load_matches_source_triangles(bindings, tmp_path, tris, kind):
    path = str(tmp_path / f"m_{kind}.stl")
    (_write_binary_stl if kind == "binary" else _write_text_stl)(path, tris)
    hw = bindings.HostWorld()
    s = hw.load_stl(path, hw.DiffuseMaterial((0.7, 0.7, 0.7)))
    V, N, T = hw.mesh_triangles(s)
    np.testing.assert_array_equal(V.view(np.int32), tris.view(np.int32))    # %.9g round-trips a float32
    np.testing.assert_allclose(N[:, 0], _flat_normal(tris), rtol=0, atol=2e-6)  # file normals dropped, FixNormals (STL.cs:94)
    assert (N[:, 0] == N[:, 1]).all() and (N[:, 0] == N[:, 2]).all() and (T == 0).all()


def test_stl_edge_cases(bindings, tmp_path, tris):
    hw = bindings.HostWorld()
    mat = hw.DiffuseMaterial((1, 1, 1))
    # a binary file cut in the middle of a facet keeps the facets read so far (STL.cs:218-221 catches and returns)
    path = str(tmp_path / "cut.stl")
    _write_binary_stl(path, tris)
    data = open(path, "rb").read()
    open(path, "wb").write(data[: 84 + 50 * 10 + 17])
    assert hw.mesh_triangles(hw.load_stl(path, mat))[0].shape[0] == 10
    # a binary file whose header starts with "solid" but has no "facet" in its first 256 bytes is still binary (STL.cs:60-68)
    path2 = str(tmp_path / "solid_header.stl")
    open(path2, "wb").write(b"solid but binary".ljust(80, b" ") + data[80:])
    assert hw.mesh_triangles(hw.load_stl(path2, mat))[0].shape[0] == len(tris)
    # text without the 'solid' line, or with a malformed vertex: empty mesh (STL.cs:134-138)
    p3 = str(tmp_path / "bad.stl")
    open(p3, "w").write("facet normal 0 0 1\nvertex 0 0 0\n" + " " * 100)
    assert hw.mesh_triangles(hw.load_stl(p3, mat))[0].shape[0] == 0
    p4 = str(tmp_path / "bad2.stl")
    open(p4, "w").write("solid x\nfacet normal 0 0 1\nvertex 0 zero 0\nendfacet\nendsolid\n" + " " * 100)
    assert hw.mesh_triangles(hw.load_stl(p4, mat))[0].shape[0] == 0
    with pytest.raises(RuntimeError):
        hw.load_stl(str(tmp_path / "missing.stl"), mat)


def test_obj_load_rules(bindings, tmp_path):
    """v / vt / vn / f with fan triangulation, 1-based indices, upper-case tolerated (the line is lower-cased), and the two
    index quirks of OBJ.cs: the normal list starts with a dummy zero vector (`vn` index k names the (k-1)-th normal, k = 1
    gives the face normal through FixNormals) and "a//c" puts c into the TEXTURE slot."""
    path = str(tmp_path / "quad.obj")
    open(path, "w").write(
        "# a unit quad, a triangle with // indices and one with bare indices\n"
        "mtllib none.mtl\nusemtl whatever\n"
        "V 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 1\n"
        "vt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\n"
        "vn 0 0 1\nvn 0 1 0\nvn 1 0 0\n"
        "f 1/1/2 2/2/2 3/3/3 4/4/1\n"
        "f 1//2 2//3   5//4\n"
        "f 1 5 4\n")
    hw = bindings.HostWorld()
    V, N, T = hw.mesh_triangles(hw.load_obj(path, hw.DiffuseMaterial((0.5, 0.5, 0.5))))
    vs = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    vts = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    vns = np.array([[0, 0, 0], [0, 0, 1], [0, 1, 0], [1, 0, 0]], np.float32)  # dummy first
    assert V.shape[0] == 4
    np.testing.assert_array_equal(V[0], vs[[0, 1, 2]]); np.testing.assert_array_equal(V[1], vs[[0, 2, 3]])   # fan (0, i, i+1)
    np.testing.assert_array_equal(T[0], vts[[0, 1, 2]]); np.testing.assert_array_equal(T[1], vts[[0, 2, 3]])
    np.testing.assert_array_equal(N[0], vns[[1, 1, 2]])                      # vn indices 2,2,3 -> list entries 1,1,2
    np.testing.assert_array_equal(N[1][:2], vns[[1, 2]]); np.testing.assert_array_equal(N[1][2], [0, 0, 1])  # index 1 -> dummy -> face normal
    np.testing.assert_array_equal(V[2], vs[[0, 1, 4]])
    np.testing.assert_array_equal(T[2], vts[[1, 2, 3]])                      # "1//2": the 2 is read as a texture index
    face = _flat_normal(V[2:3])[0]
    np.testing.assert_allclose(N[2], np.tile(face, (3, 1)), atol=1e-6)       # no third field -> normal index 0 -> dummy
    np.testing.assert_array_equal(T[3], vts[[0, 0, 0]])                      # bare indices: texture index defaults to the first vt
    with pytest.raises(RuntimeError):
        hw.load_obj(str(tmp_path / "missing.obj"), 0)


def test_mesh_utilities(bindings, tmp_path, tris):
    path = str(tmp_path / "m.stl")
    _write_binary_stl(path, tris)
    hw = bindings.HostWorld()
    s = hw.load_stl(path, hw.DiffuseMaterial((0.7, 0.7, 0.7)))
    # FitInside (Mesh.cs:243-252): uniform scale, the box is touched along the tightest axis, anchored at the centre
    hw.mesh_fit_inside(s, (-1, 0, -1), (1, 3, 1), (0.5, 0.5, 0.5))
    V, N, _ = hw.mesh_triangles(s)
    lo, hi = V.reshape(-1, 3).min(0), V.reshape(-1, 3).max(0)
    assert (lo >= np.array([-1, 0, -1]) - 1e-5).all() and (hi <= np.array([1, 3, 1]) + 1e-5).all()
    assert np.isclose((hi - lo).max(), 2.0, atol=1e-4)
    np.testing.assert_allclose((lo + hi) / 2, [0, 1.5, 0], atol=1e-4)
    np.testing.assert_allclose(np.linalg.norm(N, axis=2), 1.0, atol=1e-5)
    # MoveTo (Mesh.cs:237-241): the anchor point of the bounding box lands on `position`
    hw.mesh_move_to(s, (5, 6, 7), (0, 0, 0))
    V2, _, _ = hw.mesh_triangles(s)
    np.testing.assert_allclose(V2.reshape(-1, 3).min(0), [5, 6, 7], atol=1e-5)
    # SmoothNormals (Mesh.cs:191-229): one normal per position = normalised sum of the corner normals meeting there
    before = hw.mesh_triangles(s)[1]
    hw.mesh_smooth_normals(s)
    V3, N3, _ = hw.mesh_triangles(s)
    keys = {}
    for p, n in zip(V3.reshape(-1, 3), before.reshape(-1, 3)):
        keys.setdefault(p.tobytes(), []).append(n.astype(np.float64))
    for p, n in zip(V3.reshape(-1, 3), N3.reshape(-1, 3)):
        e = np.sum(keys[p.tobytes()], axis=0)
        np.testing.assert_allclose(n, e / np.linalg.norm(e), atol=2e-6)
    # Transform (Mesh.cs:254-274) with a rotation about z: positions rotate, normals rotate and stay unit
    c, sn = np.cos(0.5), np.sin(0.5)
    R = np.array([[c, -sn, 0, 0], [sn, c, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])
    hw.mesh_transform(s, R)
    V4, N4, _ = hw.mesh_triangles(s)
    np.testing.assert_allclose(V4.reshape(-1, 3), V3.reshape(-1, 3).astype(np.float64) @ R[:3, :3].T, rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(N4.reshape(-1, 3), N3.reshape(-1, 3).astype(np.float64) @ R[:3, :3].T, atol=1e-5)


def test_smooth_normals_threshold_quirk(bindings, tmp_path):
    """Mesh.cs:155-189 snapshots the running list of ALL N1 / N2 / N3 seen so far for each vertex (not the normals meeting at
    the vertex): restated literally here."""
    V = scenes.displaced_icosphere(2, 1.0, (0, 0, 0))
    path = str(tmp_path / "m.stl")
    _write_binary_stl(path, V)
    hw = bindings.HostWorld()
    s = hw.load_stl(path, hw.DiffuseMaterial((0.7, 0.7, 0.7)))
    _, N0, _ = hw.mesh_triangles(s)
    hw.mesh_smooth_normals_threshold(s, 0.9)
    _, N1, _ = hw.mesh_triangles(s)
    thr = np.cos(0.9)
    lookup = {}
    for i, t in enumerate(V):
        for k in range(3):
            lookup[(t[k] + 0.0).tobytes()] = (k, i + 1)
    for i, t in enumerate(V):
        for k in range(3):
            lst, cnt = lookup[(t[k] + 0.0).tobytes()]
            cand = N0[:cnt, lst].astype(np.float64)
            keep = cand[(cand @ N0[i, k].astype(np.float64)) >= thr - 1e-7]
            e = keep.sum(axis=0)
            np.testing.assert_allclose(N1[i, k], e / np.linalg.norm(e), atol=1e-4)


def test_loaded_mesh_flattens_like_an_authored_mesh(bindings, tmp_path, tris):
    """A mesh that came through STL.Load is the same flat scene, byte for byte, as the same triangles handed to Mesh.NewMesh."""
    path = str(tmp_path / "m.stl")
    _write_binary_stl(path, tris)
    files = []
    for loaded in (True, False):
        hw = bindings.HostWorld()
        mat = hw.GlossyMaterial((0.9, 0.5, 0.2), 1.4, 0.1)
        s = hw.load_stl(path, mat) if loaded else hw.mesh(tris, mat)
        hw.add(s)
        hw.add(hw.sphere((0, 5, 0), 1.0, hw.LightMaterial((1, 1, 1), 30)))
        hw.look_at((0, 2, -5), (0, 1, 0), (0, 1, 0), 40)
        hw.sampler(1, 4)
        hw.flatten()
        out = str(tmp_path / f"flat_{int(loaded)}.ptfs")
        hw.save_flat(out)
        files.append(open(out, "rb").read())
    assert files[0] == files[1] and len(files[0]) > 320 * 48
