"""Scene authoring verbs shared by the two back ends.

The reference builds scenes with factory calls (``Sphere.NewSphere``, ``Material.GlossyMaterial``,
``TransformedShape.NewTransformedShape`` ... see ``Example.cs``).  ``World`` exposes the same verbs with the
same argument meaning over a C library that implements them:

* prefix ``pth_`` — ``libpthost.so``, the product's C++ host side (builds the reference's kd-tree, flattens the
  scene into SoA buffers, drives the CUDA library through the C ABI in ``include/ptgpu.h``);
* prefix ``orc_`` — ``oracle/liborc.so``, the CPU restatement used only as a checker by tests and bench.py.

Nothing here computes anything: it marshals arguments.  Host-side helpers that the reference evaluates in C#
before a shape exists (``Colour.HexColor``, ``Matrix.Rotate`` ...) live in :mod:`ptsharp_b200.hostmath`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)


def _d3(v) -> "C.Array":
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def _d16(m) -> "C.Array":
    a = np.ascontiguousarray(np.asarray(m, dtype=np.float64).reshape(16))
    return (C.c_double * 16)(*a.tolist())


def _fp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(c_float_p)


def _ip(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(c_int_p)


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(c_double_p)


_SIGS = {
    "world_new": (C.c_void_p, []),
    "world_free": (None, [C.c_void_p]),
    "texture": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "material": (C.c_int, [C.c_void_p, c_double_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                           C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "sphere": (C.c_int, [C.c_void_p, c_double_p, C.c_double, C.c_int]),
    "cube": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int]),
    "plane": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int]),
    "cylinder": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]),
    "mesh": (C.c_int, [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p, c_int_p, C.c_int]),
    "transformed": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sdf_sphere": (C.c_int, [C.c_void_p, C.c_double]),
    "sdf_cube": (C.c_int, [C.c_void_p, c_double_p]),
    "sdf_cylinder": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "sdf_capsule": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_double]),
    "sdf_torus": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "sdf_transform": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sdf_scale": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "sdf_repeat": (C.c_int, [C.c_void_p, C.c_int, c_double_p]),
    "sdf_combine": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_int_p]),
    "sdf_shape": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "volume": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int, C.c_int, C.c_int, C.c_double, c_double_p,
                         C.c_int, c_double_p, c_double_p, c_int_p]),
    "scene_add": (None, [C.c_void_p, C.c_int]),
    "scene_env": (None, [C.c_void_p, c_double_p, C.c_int, C.c_double]),
    "camera_lookat": (None, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_double]),
    "camera_focus": (None, [C.c_void_p, c_double_p, C.c_double]),
    "sampler": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "compile": (None, [C.c_void_p]),
}


def bind(lib: C.CDLL, prefix: str, extra: Optional[dict] = None) -> None:
    """Attach argtypes/restype for the authoring verbs (and `extra`) on `lib`."""
    table = dict(_SIGS)
    if extra:
        table.update(extra)
    for name, (res, args) in table.items():
        fn = getattr(lib, prefix + name)
        fn.restype = res
        fn.argtypes = args


# enums (LightMode.cs, SpecularMode.cs, BounceType.cs): the integer codes are part of the ABI
LightModeRandom, LightModeAll = 0, 1
SpecularModeNaive, SpecularModeFirst, SpecularModeAll = 0, 1, 2


class World:
    """One scene + camera + sampler being authored on a back end."""

    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib = lib
        self.prefix = prefix
        self._f = lambda n: getattr(lib, prefix + n)
        self.h = C.c_void_p(self._f("world_new")())
        self._keep = []  # numpy buffers the C side may borrow until flatten/compile
        self.sampler_settings = dict(firstHit=1, maxBounces=4, directLighting=1, softShadows=1,
                                     lightMode=LightModeRandom, specularMode=SpecularModeNaive)

    def close(self):
        if self.h:
            self._f("world_free")(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # --- textures / materials (Material.cs:48-100) -------------------------------------------------
    def texture(self, rgb: np.ndarray) -> int:
        """rgb: (H, W, 3) float64 linear colours (the loader's Pow(2.2) already applied, Texture.cs:163)."""
        a = np.ascontiguousarray(rgb, dtype=np.float64)
        h, w = a.shape[0], a.shape[1]
        return self._f("texture")(self.h, w, h, _dp(a))

    def material(self, color, texture=-1, normal_texture=-1, bump_texture=-1, gloss_texture=-1, bump_multiplier=1.0,
                 emittance=0.0, index=1.0, gloss=0.0, tint=0.0, reflectivity=-1.0, transparent=False) -> int:
        return self._f("material")(self.h, _d3(color), texture, normal_texture, bump_texture, gloss_texture,
                                   bump_multiplier, emittance, index, gloss, tint, reflectivity, int(transparent))

    def DiffuseMaterial(self, color):
        return self.material(color, index=1, gloss=0, tint=0, reflectivity=-1)

    def SpecularMaterial(self, color, index):
        return self.material(color, index=index)

    def GlossyMaterial(self, color, index, gloss, **kw):
        return self.material(color, index=index, gloss=gloss, **kw)

    def ClearMaterial(self, index, gloss):
        return self.material((0, 0, 0), index=index, gloss=gloss, transparent=True)

    def TransparentMaterial(self, color, index, gloss, tint):
        return self.material(color, index=index, gloss=gloss, tint=tint, transparent=True)

    def MetallicMaterial(self, color, gloss, tint):
        return self.material(color, index=1, gloss=gloss, tint=tint, reflectivity=1)

    def LightMaterial(self, color, emittance):
        return self.material(color, emittance=emittance)

    # --- shapes ---------------------------------------------------------------------------------------
    def sphere(self, center, radius, mat) -> int:
        return self._f("sphere")(self.h, _d3(center), float(radius), mat)

    def cube(self, mn, mx, mat) -> int:
        return self._f("cube")(self.h, _d3(mn), _d3(mx), mat)

    def plane(self, point, normal, mat) -> int:
        return self._f("plane")(self.h, _d3(point), _d3(normal), mat)

    def cylinder(self, radius, z0, z1, mat) -> int:
        return self._f("cylinder")(self.h, float(radius), float(z0), float(z1), mat)

    def mesh(self, V: np.ndarray, mat: int, N: Optional[np.ndarray] = None, T: Optional[np.ndarray] = None,
             mats: Optional[np.ndarray] = None) -> int:
        """V, N, T: (ntri, 3, 3) float32.  N=None -> flat normals (Triangle.FixNormals)."""
        V = np.ascontiguousarray(V, dtype=np.float32)
        ntri = V.shape[0]
        N = None if N is None else np.ascontiguousarray(N, dtype=np.float32)
        T = None if T is None else np.ascontiguousarray(T, dtype=np.float32)
        mats = None if mats is None else np.ascontiguousarray(mats, dtype=np.int32)
        self._keep += [V, N, T, mats]
        return self._f("mesh")(self.h, ntri, _fp(V), _fp(N), _fp(T), _ip(mats), mat)

    def transformed(self, shape: int, matrix) -> int:
        return self._f("transformed")(self.h, shape, _d16(matrix))

    def transformed_cylinder(self, v0, v1, radius, mat) -> int:
        """Cylinder.NewTransformedCylinder (Cylinder.cs:22-35).  `new Matrix().Rotate(u, a).Translate(v0)` is
        just Translate(v0) because Matrix.Translate ignores `this` (Matrix.cs:33-36), so the bar stays z-aligned."""
        from . import hostmath as hm
        d = hm.vsub(hm.vec(v1), hm.vec(v0))
        z = hm.vlength(d)
        c = self.cylinder(radius, 0.0, z, mat)
        return self.transformed(c, hm.translate(hm.vec(v0)))

    # --- SDF ------------------------------------------------------------------------------------------
    def sdf_sphere(self, r): return self._f("sdf_sphere")(self.h, float(r))
    def sdf_cube(self, size): return self._f("sdf_cube")(self.h, _d3(size))
    def sdf_cylinder(self, r, h): return self._f("sdf_cylinder")(self.h, float(r), float(h))
    def sdf_capsule(self, a, b, r): return self._f("sdf_capsule")(self.h, _d3(a), _d3(b), float(r))
    def sdf_torus(self, major, minor): return self._f("sdf_torus")(self.h, float(major), float(minor))
    def sdf_transform(self, sdf, m): return self._f("sdf_transform")(self.h, sdf, _d16(m))
    def sdf_scale(self, sdf, f): return self._f("sdf_scale")(self.h, sdf, float(f))
    def sdf_repeat(self, sdf, step): return self._f("sdf_repeat")(self.h, sdf, _d3(step))

    def _combine(self, op, items):
        a = (C.c_int * len(items))(*items)
        return self._f("sdf_combine")(self.h, op, len(items), a)

    def sdf_union(self, items): return self._combine(0, items)
    def sdf_difference(self, items): return self._combine(1, items)
    def sdf_intersection(self, items): return self._combine(2, items)
    def sdf_shape(self, sdf, mat): return self._f("sdf_shape")(self.h, sdf, mat)

    def volume(self, bmin, bmax, data: np.ndarray, zscale: float, windows: Sequence[tuple]) -> int:
        """data: (D, H, W) float64 — index x + y*W + z*W*H (Volume.cs:40-46).  windows: [(lo, hi, mat)]."""
        a = np.ascontiguousarray(data, dtype=np.float64)
        d, h, w = a.shape
        n = len(windows)
        lo = (C.c_double * n)(*[float(x[0]) for x in windows])
        hi = (C.c_double * n)(*[float(x[1]) for x in windows])
        ms = (C.c_int * n)(*[int(x[2]) for x in windows])
        return self._f("volume")(self.h, _d3(bmin), _d3(bmax), w, h, d, float(zscale), _dp(a), n, lo, hi, ms)

    # --- scene / camera / sampler -------------------------------------------------------------------
    def add(self, shape: int) -> None:
        self._f("scene_add")(self.h, shape)

    def env(self, color=(0, 0, 0), texture=-1, angle=0.0) -> None:
        self._f("scene_env")(self.h, _d3(color), texture, float(angle))

    def look_at(self, eye, center, up, fovy) -> None:
        self._f("camera_lookat")(self.h, _d3(eye), _d3(center), _d3(up), float(fovy))

    def set_focus(self, focal_point, aperture) -> None:
        self._f("camera_focus")(self.h, _d3(focal_point), float(aperture))

    def sampler(self, first_hit, max_bounces, direct_lighting=True, soft_shadows=True,
                light_mode=LightModeRandom, specular_mode=SpecularModeNaive) -> None:
        """DefaultSampler.NewSampler(firstHitSamples, maxBounces) + its public fields (Sampler.cs:30-53)."""
        self.sampler_settings = dict(firstHit=first_hit, maxBounces=max_bounces, directLighting=int(direct_lighting),
                                     softShadows=int(soft_shadows), lightMode=light_mode, specularMode=specular_mode)
        self._f("sampler")(self.h, first_hit, max_bounces, int(direct_lighting), int(soft_shadows), light_mode,
                           specular_mode)

    def compile(self) -> None:
        self._f("compile")(self.h)
