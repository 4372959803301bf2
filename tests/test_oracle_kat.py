"""Known-answer tests that pin the CPU oracle (oracle/).

The reference ships no tests, golden vectors or runnable binary here, so these answers are derived
by hand from the formulas in the reference source (cited per test) or from published vectors (Philox).
"""
import ctypes as C
import math

import numpy as np
import pytest


def _d3(v):
    return (C.c_double * 3)(*v)


def test_philox_known_answers(orc):
    # Random123 kat_vectors, philox4x32 10 rounds
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    lib = orc.lib()
    for ctr, key, want in kats:
        out = (C.c_uint * 4)()
        lib.orc_philox((C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), out)
        assert tuple(out) == want


def _one_shape_world(orc, kind):
    w = orc.OracleWorld()
    m = w.DiffuseMaterial((1, 1, 1))
    if kind == "sphere":
        w.add(w.sphere((0, 0, 0), 1.0, m))
    elif kind == "plane":
        w.add(w.plane((0, 0, 0), (0, 0, 2), m))  # normal is normalised by NewPlane (Plane.cs:28)
    elif kind == "cube":
        w.add(w.cube((-1, -1, -1), (1, 1, 1), m))
    elif kind == "cylinder":
        w.add(w.transformed(w.cylinder(1.0, 0.0, 2.0, m), np.eye(4)))
    elif kind == "triangle":
        V = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
        w.add(w.mesh(V, m))
    w.look_at((0, 0, 5), (0, 0, 0), (0, 1, 0), 40)
    return w


def test_sphere_roots(orc):  # Sphere.cs:40-60
    w = _one_shape_world(orc, "sphere")
    o = np.array([[0, 0, -3], [0, 0, 0], [0, 2, -3], [0, 0, 3]], np.float32)
    d = np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, 1]], np.float32)
    h = w.intersect_batch(o, d)
    assert h["shape"].tolist() == [0, 0, -1, -1]
    assert h["t"][0] == 2.0 and h["t"][1] == 1.0          # near root outside, far root from inside
    assert h["inside"].tolist()[:2] == [0, 1]
    np.testing.assert_array_equal(h["normal"][0], [0, 0, -1])
    np.testing.assert_array_equal(h["normal"][1], [0, 0, -1])  # flipped to face the ray (Hit.cs:37-40)


def test_plane_and_cube(orc):  # Plane.cs:38-52, Cube.cs:35-69
    w = _one_shape_world(orc, "plane")
    h = w.intersect_batch(np.array([[0, 0, 2], [0, 0, 2], [0, 0, 2]], np.float32),
                          np.array([[0, 0, -1], [1, 0, 0], [0, 0, 1]], np.float32))
    assert h["shape"].tolist() == [0, -1, -1] and h["t"][0] == 2.0
    w = _one_shape_world(orc, "cube")
    h = w.intersect_batch(np.array([[0, 0, -3], [0, 0, 0], [-3, 0, 0]], np.float32),
                          np.array([[0, 0, 1], [0, 0, 1], [1, 0, 0]], np.float32))
    assert h["shape"].tolist() == [0, -1, 0]               # never hit from inside (Cube.cs:41)
    assert h["t"][0] == 2.0
    np.testing.assert_array_equal(h["normal"][0], [0, 0, -1])
    np.testing.assert_array_equal(h["normal"][2], [-1, 0, 0])


def test_triangle_moller_trumbore(orc):  # Triangle.cs:95-124, 208-223
    w = _one_shape_world(orc, "triangle")
    o = np.array([[0.25, 0.25, 1], [0.75, 0.75, 1], [0.25, 0.25, -1], [0, 0, 1]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, -1], [0, 0, 1], [0, 0, -1]], np.float32)
    h = w.intersect_batch(o, d)
    assert h["shape"].tolist() == [0, -1, 0, 0]            # u+v>1 misses; no back-face culling; the vertex itself hits
    assert h["prim"].tolist() == [0, -1, 0, 0]
    assert h["t"][0] == 1.0 and h["t"][2] == 1.0
    np.testing.assert_array_equal(h["normal"][0], [0, 0, 1])
    np.testing.assert_array_equal(h["normal"][2], [0, 0, -1])


def test_cylinder_order_quirk(orc):  # Cylinder.cs:60-106: caps first, then the FAR lateral root
    w = _one_shape_world(orc, "cylinder")
    # sideways ray through the middle of the bar: the reference returns the far root (t = 4), not the near one (t = 2)
    h = w.intersect_batch(np.array([[-3, 0, 1]], np.float32), np.array([[1, 0, 0]], np.float32))
    assert h["shape"][0] == 0 and h["t"][0] == 4.0
    # a ray that passes the top cap's plane inside the radius hits the cap even though the side is nearer
    h = w.intersect_batch(np.array([[0, 0, 5]], np.float32), np.array([[0, 0, -1]], np.float32))
    assert h["t"][0] == 3.0
    np.testing.assert_array_equal(h["normal"][0], [0, 0, 1])


def test_fresnel_and_refract(orc):  # Vector.cs:500-536
    lib = orc.lib()
    n, i = _d3((0, 0, 1)), _d3((0, 0, -1))
    r = lib.orc_reflectance(n, i, 1.0, 1.5)
    assert r == pytest.approx(((1 - 1.5) / (1 + 1.5)) ** 2, rel=1e-15)
    assert lib.orc_reflectance(n, i, 1.0, 1.0) == 0.0
    # total internal reflection
    s = math.sin(math.radians(60))
    i2 = _d3((s, 0, -math.cos(math.radians(60))))
    assert lib.orc_reflectance(n, i2, 1.5, 1.0) == 1.0
    # Snell: sin(theta_t) = n1/n2 sin(theta_i)
    th = math.radians(30)
    i3 = _d3((math.sin(th), 0, -math.cos(th)))
    out = (C.c_float * 3)()
    lib.orc_refract(n, i3, 1.0, 1.5, out)
    assert out[0] == pytest.approx(math.sin(th) / 1.5, rel=1e-6)
    assert math.hypot(out[0], out[2]) == pytest.approx(1.0, rel=1e-6)


def test_matrix_inverse_and_quirks(orc):  # Matrix.cs:33-54, 196-217
    lib = orc.lib()
    m = np.array([[2, 0, 0, 1], [0, 4, 0, -2], [0, 0, 0.5, 3], [0, 0, 0, 1]], np.float64)
    out = np.empty((4, 4))
    lib.orc_matrix_inverse(m.ctypes.data_as(orc.c_double_p), out.ctypes.data_as(orc.c_double_p))
    np.testing.assert_allclose(out @ m, np.eye(4), atol=1e-15)
    rot = np.empty((4, 4))
    lib.orc_matrix_rotate(_d3((0, 0, 2)), math.pi / 2, rot.ctypes.data_as(orc.c_double_p))
    # the reference's Rotate is the transpose of the usual right-handed rotation about +z (Matrix.cs:50-53)
    np.testing.assert_allclose(rot[:3, :3], [[0, 1, 0], [-1, 0, 0], [0, 0, 1]], atol=1e-15)


def test_hexcolor_and_welford(orc):  # Colour.cs:125-132, Buffer.cs:33-55
    lib = orc.lib()
    out = (C.c_double * 3)()
    lib.orc_hexcolor(0xFF8000, out)
    g = float(np.float32(128) / np.float32(255))
    assert out[0] == 1.0 and out[2] == 0.0 and out[1] == math.pow(g, float(np.float32(2.2)))
    s = np.array([[1, 2, 3], [3, 2, 1], [5, 5, 5], [0, 1, 0]], np.float64)
    mean = np.empty(3); var = np.empty(3)
    lib.orc_welford(4, s.ctypes.data_as(orc.c_double_p), mean.ctypes.data_as(orc.c_double_p), var.ctypes.data_as(orc.c_double_p))
    np.testing.assert_allclose(mean, s.mean(axis=0), rtol=1e-15)
    np.testing.assert_allclose(var, s.var(axis=0, ddof=1), rtol=1e-14)


def test_camera_lookat(orc):  # Camera.cs:23-35
    w = orc.OracleWorld()
    w.look_at((0, 0, 5), (0, 0, 0), (0, 1, 0), 90)
    puvw, mfa = w.camera()
    np.testing.assert_array_equal(puvw[0], [0, 0, 5])
    np.testing.assert_array_equal(puvw[3], [0, 0, -1])     # w = norm(center - eye)
    np.testing.assert_array_equal(puvw[1], [-1, 0, 0])     # u = norm(up x w)
    np.testing.assert_array_equal(puvw[2], [0, 1, 0])      # v = norm(w x u)
    assert mfa[0] == pytest.approx(1.0, rel=1e-15)
    # centre-of-image ray through CastRay (Camera.cs:98-105): odd resolution so the centre pixel maps to px = py = 0
    o, d = w.cast_rays(3, 3, [1], [1], [0.5], [0.5], [0])
    np.testing.assert_array_equal(o[0], [0, 0, 5])
    np.testing.assert_allclose(d[0], [0, 0, -1], atol=1e-7)


def test_light_registration_and_struct_identity(orc):  # Scene.cs:29-38, SURVEY F7
    w = orc.OracleWorld()
    light = w.LightMaterial((1, 1, 1), 10)
    w.add(w.sphere((0, 0, 5), 1, light))                                  # class: can light
    w.add(w.transformed(w.sphere((3, 0, 5), 1, light), np.eye(4)))        # struct wrapper: registered, never matches
    V = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], np.float32)
    w.add(w.mesh(V, light))                                               # Mesh.MaterialAt is `new Material()`: not a light
    assert w.num_lights() == 2


def test_serial_render_adaptive_rule(orc):  # Renderer.cs:153-158: samples = AdaptiveSamples * (int)pow(clamp(sd / threshold, 0, 1), exponent)
    from ptsharp_b200 import scenes
    ow = orc.OracleWorld()
    scenes.build_c1(ow)
    W, H = 24, 16
    ow.set_extra(3, 0, 1.0)
    ow.set_serial(True, 0.05, 1.0)
    ow.render(W, H, 1, passes=1, threads=2, rng_mode=orc.RNG_KEYED)
    assert (ow.last_samples(W, H) == 1).all()        # first pass: one sample per pixel, deviation 0, (int)0 = 0 extra samples
    ow.set_serial(True, 0.05, 0.0)
    ow.render(W, H, 1, passes=1, threads=2, rng_mode=orc.RNG_KEYED)
    assert (ow.last_samples(W, H) == 1 + 3).all()    # exponent 0: pow(v, 0) = 1 for every pixel
    ow.set_serial(True, 0.05, 1.0)
    ow.render(W, H, 1, passes=2, threads=2, rng_mode=orc.RNG_KEYED)
    ns = ow.last_samples(W, H)
    assert set(np.unique(ns)) == {2, 5}              # second pass: only the pixels whose deviation reached the threshold
    ow.set_extra(0, 0, 1.0)
    ow.set_serial(False)
