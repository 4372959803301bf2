"""ORACLE — TEST INFRASTRUCTURE ONLY (parity unpinned; see oracle/orc_math.hpp and DESIGN.md).

ctypes loader for oracle/liborc.so, the CPU restatement of PTSharp's render path.  Only tests/, the smoke check in
__graft_entry__.py and bench.py's CPU-baseline / `--impl reference` legs may import this module; nothing in
ptsharp_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from ptsharp_b200.authoring import World, bind, c_double_p, c_float_p, c_int_p  # noqa: E402  (marshalling only)

RNG_SEQUENTIAL, RNG_KEYED = 0, 1
c_ll_p = C.POINTER(C.c_longlong)
c_uint_p = C.POINTER(C.c_uint)

_EXTRA = {
    "num_lights": (C.c_int, [C.c_void_p]),
    "set_extra": (None, [C.c_void_p, C.c_int, C.c_int, C.c_double]),
    "set_serial": (None, [C.c_void_p, C.c_int, C.c_double, C.c_double]),
    "set_task_sample": (None, [C.c_void_p, C.c_int, C.c_int]),
    "spherical_harmonic": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p]),
    "last_samples": (C.c_int, [C.c_void_p, C.c_int, c_int_p]),
    "camera_get": (None, [C.c_void_p, c_float_p, c_double_p]),
    "intersect_batch": (None, [C.c_void_p, C.c_int, c_float_p, c_float_p, c_int_p, c_int_p, c_double_p, c_float_p,
                               c_float_p, c_int_p, c_int_p]),
    "cast_rays": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int_p, c_int_p, c_double_p, c_double_p, c_int_p,
                         C.c_uint, C.c_uint, c_float_p, c_float_p]),
    "render": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_int, C.c_int,
                      c_int_p, c_double_p, c_double_p, c_ll_p]),
    "tree_stats": (C.c_int, [C.c_void_p, C.c_int, c_ll_p, c_float_p]),
    "tree_dump": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_double_p, c_int_p, c_int_p, c_int_p]),
    "traversal_cost": (None, [C.c_void_p, C.c_int, c_float_p, c_float_p, c_double_p]),
    "philox": (None, [c_uint_p, c_uint_p, c_uint_p]),
    "reflectance": (C.c_double, [c_double_p, c_double_p, C.c_double, C.c_double]),
    "refract": (None, [c_double_p, c_double_p, C.c_double, C.c_double, c_float_p]),
    "matrix_inverse": (None, [c_double_p, c_double_p]),
    "matrix_rotate": (None, [c_double_p, C.c_double, c_double_p]),
    "hexcolor": (None, [C.c_int, c_double_p]),
    "welford": (None, [C.c_int, c_double_p, c_double_p, c_double_p]),
    "keyed_draw": (C.c_double, [C.c_uint] * 9),
}

_lib = None


def build(force: bool = False) -> str:
    so = os.path.join(HERE, "liborc.so")
    srcs = [os.path.join(HERE, f) for f in ("orc_capi.cpp", "orc_math.hpp", "orc_scene.hpp", "orc_render.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return so


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        bind(_lib, "orc_", _EXTRA)
    return _lib


class OracleWorld(World):
    def __init__(self):
        super().__init__(lib(), "orc_")

    # -- queries -------------------------------------------------------------------------------------------
    def intersect_batch(self, o: np.ndarray, d: np.ndarray) -> dict:
        o = np.ascontiguousarray(o, dtype=np.float32)
        d = np.ascontiguousarray(d, dtype=np.float32)
        n = o.shape[0]
        out = dict(shape=np.empty(n, np.int32), prim=np.empty(n, np.int32), t=np.empty(n, np.float64),
                   normal=np.empty((n, 3), np.float32), position=np.empty((n, 3), np.float32),
                   inside=np.empty(n, np.int32), material=np.empty(n, np.int32))
        self.lib.orc_intersect_batch(self.h, n, o.ctypes.data_as(c_float_p), d.ctypes.data_as(c_float_p),
                                     out["shape"].ctypes.data_as(c_int_p), out["prim"].ctypes.data_as(c_int_p),
                                     out["t"].ctypes.data_as(c_double_p), out["normal"].ctypes.data_as(c_float_p),
                                     out["position"].ctypes.data_as(c_float_p), out["inside"].ctypes.data_as(c_int_p),
                                     out["material"].ctypes.data_as(c_int_p))
        return out

    def cast_rays(self, W, H, x, y, fu, fv, sample, seed=0x50545348, pass_index=0):
        x = np.ascontiguousarray(x, np.int32); y = np.ascontiguousarray(y, np.int32)
        fu = np.ascontiguousarray(fu, np.float64); fv = np.ascontiguousarray(fv, np.float64)
        sample = np.ascontiguousarray(sample, np.int32)
        n = x.shape[0]
        o = np.empty((n, 3), np.float32); d = np.empty((n, 3), np.float32)
        self.lib.orc_cast_rays(self.h, W, H, n, x.ctypes.data_as(c_int_p), y.ctypes.data_as(c_int_p),
                               fu.ctypes.data_as(c_double_p), fv.ctypes.data_as(c_double_p),
                               sample.ctypes.data_as(c_int_p), seed, pass_index, o.ctypes.data_as(c_float_p),
                               d.ctypes.data_as(c_float_p))
        return o, d

    def render(self, W, H, spp, passes=1, stratified=False, threads=1, rng_mode=RNG_SEQUENTIAL, seed=0x50545348,
               sample_base=0, sample_stride=1, window=None):
        mean = np.zeros((H, W, 3), np.float64)
        var = np.zeros((H, W, 3), np.float64)
        cnt = (C.c_longlong * 3)()
        win = None if window is None else (C.c_int * 4)(*window)
        self.lib.orc_render(self.h, W, H, spp, passes, int(stratified), threads, rng_mode, seed, sample_base, sample_stride, win,
                            mean.ctypes.data_as(c_double_p), var.ctypes.data_as(c_double_p), cnt)
        return mean, var, dict(cameraSamples=cnt[0], segments=cnt[1], shadowRays=cnt[2])

    def set_extra(self, adaptive_samples=0, firefly_samples=0, firefly_threshold=1.0):
        self.lib.orc_set_extra(self.h, adaptive_samples, firefly_samples, float(firefly_threshold))

    def set_serial(self, serial=True, adaptive_threshold=1.0, adaptive_exponent=1.0):
        """The extra samples follow the serial Render() (Renderer.cs:150-191)."""
        self.lib.orc_set_serial(self.h, int(serial), float(adaptive_threshold), float(adaptive_exponent))

    def spherical_harmonic(self, l, m, pm, nm, V):
        """SphericalHarmonic.NewSphericalHarmonic(l, m, pm, nm) over the given marching-cubes triangles V (ntri, 3, 3)."""
        V = np.ascontiguousarray(V, dtype=np.float32)
        self._keep.append(V)
        return self.lib.orc_spherical_harmonic(self.h, l, m, pm, nm, V.shape[0], V.ctypes.data_as(c_float_p))

    def set_task_sample(self, stride=1, offset=0):
        """render() only renders every `stride`-th non-empty 32x32 task of the frame (a bounded, evenly spread timing sample)."""
        self.lib.orc_set_task_sample(self.h, int(stride), int(offset))

    def last_samples(self, W, H):
        """Pixel.Samples of the Buffer the last render() filled."""
        out = np.empty(W * H, np.int32)
        if self.lib.orc_last_samples(self.h, W * H, out.ctypes.data_as(c_int_p)) != 0:
            raise ValueError("size mismatch")
        return out.reshape(H, W)

    def tree_stats(self, which=-1):
        out = (C.c_longlong * 4)()
        box = (C.c_float * 6)()
        rc = self.lib.orc_tree_stats(self.h, which, out, box)
        if rc != 0:
            raise ValueError("not a mesh")
        return dict(nodes=out[0], leafItems=out[1], maxLeaf=out[2], maxDepth=out[3], box=np.array(list(box), np.float32))

    def tree_dump(self, which=-1):
        st = self.tree_stats(which)
        n, m = st["nodes"], st["leafItems"]
        axis = np.empty(n, np.int32); point = np.empty(n, np.float64)
        a = np.empty(n, np.int32); b = np.empty(n, np.int32); items = np.empty(max(m, 1), np.int32)
        self.lib.orc_tree_dump(self.h, which, axis.ctypes.data_as(c_int_p), point.ctypes.data_as(c_double_p),
                               a.ctypes.data_as(c_int_p), b.ctypes.data_as(c_int_p), items.ctypes.data_as(c_int_p))
        return dict(axis=axis, point=point, a=a, b=b, items=items[:m], box=st["box"])

    def traversal_cost(self, o, d):
        o = np.ascontiguousarray(o, dtype=np.float32); d = np.ascontiguousarray(d, dtype=np.float32)
        out = (C.c_double * 2)()
        self.lib.orc_traversal_cost(self.h, o.shape[0], o.ctypes.data_as(c_float_p), d.ctypes.data_as(c_float_p), out)
        return out[0], out[1]

    def num_lights(self):
        return self.lib.orc_num_lights(self.h)

    def camera(self):
        puvw = np.empty(12, np.float32); mfa = np.empty(3, np.float64)
        self.lib.orc_camera_get(self.h, puvw.ctypes.data_as(c_float_p), mfa.ctypes.data_as(c_double_p))
        return puvw.reshape(4, 3), mfa
