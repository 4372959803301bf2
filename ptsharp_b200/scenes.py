"""The five BASELINE.json configurations as deterministic scene generators.

Each `build_*` function authors one scene on a `World` (either back end) with the same calls an `Example.cs`
scene function makes, and returns a `Config` with resolution / spp.  No RNG and no asset files: meshes, textures
and volumes come from closed-form formulas so every back end sees identical floats.

Triangle order matters: the reference's kd-tree builder takes its split position from whichever shapes sit in the
*middle of the array* (Tree.cs:130-148,208-226 — the "median" of an unsorted bag), so tree quality is a function
of input order.  `spatial_order` arranges triangles along a Morton curve and then applies the host
library's `BuilderFriendlyOrder`, so that the reference builder, unmodified, produces a balanced tree;
tests/test_tree.py records the resulting leaf statistics.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from . import hostmath as hm
from .authoring import (LightModeAll, LightModeRandom, SpecularModeAll, SpecularModeNaive, World)


@dataclass
class Config:
    name: str
    width: int
    height: int
    spp: int
    description: str
    triangles: int = 0
    extra: dict = field(default_factory=dict)


# --------------------------------------------------------------------------------------------------------------------
# mesh generation
# --------------------------------------------------------------------------------------------------------------------
def _icosahedron():
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    return v, f


def displaced_icosphere(freq: int, radius: float, center, amplitude: float = 0.04, k: float = 9.0) -> np.ndarray:
    """20*freq^2 triangles, (ntri,3,3) float32: p = c + R * p_hat * (1 + A sin(k x) sin(k y) sin(k z)), p_hat unit."""
    v, faces = _icosahedron()
    f = freq
    ii, jj = np.meshgrid(np.arange(f + 1), np.arange(f + 1), indexing="ij")
    # upward triangles (i,j),(i+1,j),(i,j+1) for i+j < f ; downward (i+1,j),(i+1,j+1),(i,j+1) for i+j < f-1
    up = np.argwhere(ii + jj < f)
    dn = np.argwhere(ii + jj < f - 1)
    tri_ij = np.concatenate([
        np.stack([up, up + [1, 0], up + [0, 1]], axis=1),
        np.stack([dn + [1, 0], dn + [1, 1], dn + [0, 1]], axis=1)], axis=0).astype(np.float64)  # (f*f, 3, 2)
    out = []
    c = np.asarray(center, dtype=np.float64)
    for fa in faces:
        A, B, Cc = v[fa[0]], v[fa[1]], v[fa[2]]
        p = A[None, None, :] + (B - A)[None, None, :] * (tri_ij[..., 0:1] / f) + (Cc - A)[None, None, :] * (tri_ij[..., 1:2] / f)
        p /= np.linalg.norm(p, axis=-1, keepdims=True)
        disp = 1.0 + amplitude * np.sin(k * p[..., 0:1]) * np.sin(k * p[..., 1:2]) * np.sin(k * p[..., 2:3])
        out.append(c + radius * p * disp)
    return np.concatenate(out, axis=0).astype(np.float32)


def _part1by2(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_order(V: np.ndarray) -> np.ndarray:
    """Permutation sorting triangles (ntri,3,3) by the 63-bit Morton code of their centroid."""
    cen = V.astype(np.float64).mean(axis=1)
    lo, hi = cen.min(axis=0), cen.max(axis=0)
    q = ((cen - lo) / np.maximum(hi - lo, 1e-30) * (2 ** 21 - 1)).astype(np.uint64)
    code = (_part1by2(q[:, 0]) << np.uint64(2)) | (_part1by2(q[:, 1]) << np.uint64(1)) | _part1by2(q[:, 2])
    return np.argsort(code, kind="stable")


def kd_order(V: np.ndarray, leaf: int = 1) -> np.ndarray:
    """Permutation from a balanced median split of centroids on cycling axes (level-synchronous)."""
    cen = V.astype(np.float64).mean(axis=1)
    n = cen.shape[0]
    perm = np.arange(n)
    seg = np.zeros(n, dtype=np.int64)
    level = 0
    size = n
    while size > leaf:
        axis = level % 3
        key = cen[perm, axis]
        order = np.lexsort((key, seg))
        perm = perm[order]
        seg = seg[order]
        # split every segment at its midpoint
        starts = np.flatnonzero(np.r_[True, seg[1:] != seg[:-1]])
        lens = np.diff(np.r_[starts, n])
        pos = np.arange(n) - np.repeat(starts, lens)
        half = np.repeat((lens + 1) // 2, lens)
        seg = seg * 2 + (pos >= half)
        size = (size + 1) // 2
        level += 1
    return perm


def spatial_order(V: np.ndarray, mode: str = "friendly") -> np.ndarray:
    if mode == "friendly":
        # Morton order for memory locality, then the host library's BuilderFriendlyOrder (host/host.cpp): the order for
        # which the UNMODIFIED reference builder yields a balanced tree.  Both back ends receive the same triangles.
        from .bindings import builder_friendly_order
        V = V[morton_order(V)]
        return V[builder_friendly_order(V)]
    if mode == "morton":
        return V[morton_order(V)]
    if mode == "kd":
        return V[kd_order(V)]
    if mode == "none":
        return V
    raise ValueError(mode)


def spherical_uv(V: np.ndarray, center) -> np.ndarray:
    """(ntri,3,3) texture coordinates (u, v, 0) from the direction of each vertex about `center`."""
    d = V.astype(np.float64) - np.asarray(center, dtype=np.float64)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    u = (np.arctan2(d[..., 2], d[..., 0]) + math.pi) / (2 * math.pi)
    v = (np.arcsin(np.clip(d[..., 1], -1, 1)) + math.pi / 2) / math.pi
    return np.stack([u, v, np.zeros_like(u)], axis=-1).astype(np.float32)


# --------------------------------------------------------------------------------------------------------------------
# C1 — Example.simplesphere (Example.cs:1670-1697)
# --------------------------------------------------------------------------------------------------------------------
def build_c1(w: World, width=512, height=512, spp=16, first_hit=16) -> Config:
    material = w.DiffuseMaterial(hm.WHITE)
    w.add(w.plane((0, 0, 0), (0, 0, 1), material))
    w.add(w.sphere((0, 0, 1), 1.0, material))
    w.add(w.sphere((0, 0, 5.0), 1.0, w.LightMaterial(hm.WHITE, 8)))
    w.look_at((3, 3, 3), (0, 0, 0.5), (0, 0, 1), 50)
    w.sampler(first_hit, 4)
    return Config("c1_simplesphere", width, height, spp, "Example.cs simplesphere: plane + sphere + sphere light")


# --------------------------------------------------------------------------------------------------------------------
# C2 — Cornell box from Cube/Plane primitives with a Cube area light
# --------------------------------------------------------------------------------------------------------------------
def build_c2(w: World, width=1024, height=1024, spp=256) -> Config:
    white = w.DiffuseMaterial((0.73, 0.73, 0.73))
    red = w.DiffuseMaterial((0.65, 0.05, 0.05))
    green = w.DiffuseMaterial((0.12, 0.45, 0.15))
    w.add(w.plane((0, 0, 0), (0, 1, 0), white))    # floor
    w.add(w.plane((0, 2, 0), (0, -1, 0), white))   # ceiling
    w.add(w.plane((0, 0, 1), (0, 0, -1), white))   # back
    w.add(w.plane((-1, 0, 0), (1, 0, 0), red))     # left
    w.add(w.plane((1, 0, 0), (-1, 0, 0), green))   # right
    w.add(w.cube((-0.65, 0, 0.05), (-0.05, 1.2, 0.65), white))
    w.add(w.cube((0.1, 0, -0.6), (0.7, 0.6, 0.0), white))
    w.add(w.cube((-0.25, 1.98, -0.25), (0.25, 1.999, 0.25), w.LightMaterial(hm.WHITE, 15)))
    w.look_at((0, 1, -3.4), (0, 1, 0), (0, 1, 0), 40)
    w.sampler(1, 8, light_mode=LightModeAll)
    return Config("c2_cornell", width, height, spp, "Cornell box from Cube/Plane primitives, Cube area light")


# --------------------------------------------------------------------------------------------------------------------
# C3 — 1 000 000-triangle displaced icospheres in kd-trees, Glossy + Clear
# --------------------------------------------------------------------------------------------------------------------
def build_c3(w: World, width=1920, height=1080, spp=512, freq_a=200, freq_b=100, order="friendly") -> Config:
    glossy = w.GlossyMaterial(hm.hex_color(0xB7CA79), 1.5, hm.radians(20))
    clear = w.ClearMaterial(1.5, 0)
    floor = w.GlossyMaterial(hm.hex_color(0xD8CAA8), 1.2, hm.radians(5))
    va = spatial_order(displaced_icosphere(freq_a, 1.0, (0, 1, 0)), order)
    vb = spatial_order(displaced_icosphere(freq_b, 0.45, (1.6, 0.45, -0.4)), order)
    w.add(w.mesh(va, glossy))
    w.add(w.mesh(vb, clear))
    w.add(w.cube((-50, -1, -50), (50, 0, 50), floor))
    w.add(w.sphere((-1, 10, 4), 1, w.LightMaterial(hm.WHITE, 75)))
    w.look_at((-3, 2, -1.5), (0.3, 0.7, 0), (0, 1, 0), 35)
    w.sampler(1, 4)
    return Config("c3_icospheres_1m", width, height, spp,
                  "two displaced icosphere meshes (Glossy + Clear) on a glossy floor, sphere light",
                  triangles=va.shape[0] + vb.shape[0], extra={"order": order})


# --------------------------------------------------------------------------------------------------------------------
# C4 — 200 TransformedShape instances of one 50 000-triangle mesh, albedo + normal textures
# --------------------------------------------------------------------------------------------------------------------
def procedural_albedo(n: int = 1024) -> np.ndarray:
    """(n, n, 3) linear colours: 0.2 + 0.8*checker(16), tinted by (u, v)."""
    v, u = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    checker = ((np.floor(u * 16) + np.floor(v * 16)) % 2)
    base = 0.2 + 0.8 * checker
    return np.stack([base * (0.5 + 0.5 * u), base * (0.5 + 0.5 * v), base * (1.0 - 0.5 * u)], axis=-1)


def procedural_normal_map(n: int = 1024, strength: float = 0.15) -> np.ndarray:
    """(n, n, 3) tangent-space normal map of h = 0.5 + 0.5 sin(40u) sin(40v), encoded (n + 1) / 2."""
    v, u = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    dhdu = 0.5 * 40 * np.cos(40 * u) * np.sin(40 * v)
    dhdv = 0.5 * 40 * np.sin(40 * u) * np.cos(40 * v)
    nrm = np.stack([-strength * dhdu / 40, -strength * dhdv / 40, np.ones_like(u)], axis=-1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    return (nrm + 1.0) / 2.0


def build_c4(w: World, width=3840, height=2160, spp=1024, freq=50, nx=20, nz=10, tex=1024, order="friendly") -> Config:
    albedo = w.texture(procedural_albedo(tex))
    normal = w.texture(procedural_normal_map(tex))
    mat = w.GlossyMaterial(hm.WHITE, 1.3, hm.radians(15), texture=albedo, normal_texture=normal)
    V = spatial_order(displaced_icosphere(freq, 1.0, (0, 0, 0)), order)
    T = spherical_uv(V, (0, 0, 0))
    base = w.mesh(V, mat, T=T)
    for i in range(nx * nz):
        x = (i % nx - (nx - 1) / 2) * 2.2
        z = (i // nx - (nz - 1) / 2) * 2.2
        sy = 0.6 + 0.002 * i
        # composed with .Mul as in Example.cs:1115-1117 (Translate/Rotate/Scale themselves ignore their receiver)
        m = hm.mul(hm.mul(hm.translate(hm.vec((x, sy, z))), hm.rotate((0, 1, 0), 0.3 * i)), hm.scale(hm.vec((1, sy, 1))))
        w.add(w.transformed(base, m))
    w.add(w.cube((-60, -1, -60), (60, 0, 60), w.GlossyMaterial(hm.hex_color(0xD8CAA8), 1.2, hm.radians(5))))
    w.add(w.sphere((-10, 25, -10), 3, w.LightMaterial(hm.WHITE, 60)))
    w.add(w.sphere((15, 20, 10), 3, w.LightMaterial(hm.WHITE, 40)))
    w.look_at((0, 18, -30), (0, 0, 0), (0, 1, 0), 45)
    w.sampler(1, 4)
    return Config("c4_instanced_10m", width, height, spp, "TransformedShape instances of one textured mesh, floor cube, two sphere lights",
                  triangles=V.shape[0] * nx * nz, extra={"base_triangles": V.shape[0], "instances": nx * nz})


# --------------------------------------------------------------------------------------------------------------------
# C5 — SDF solid (Example.sdf, Example.cs:1402-1418) + transformed cylinders + procedural Volume
# --------------------------------------------------------------------------------------------------------------------
def procedural_volume(n: int = 64) -> np.ndarray:
    """(n, n, n) density rho(q) = exp(-4|q|^2) (0.75 + 0.25 sin(12qx) sin(12qy) sin(12qz)), q in [-1, 1]^3, index [z, y, x]."""
    c = (np.arange(n) + 0.5) / n * 2 - 1
    qz, qy, qx = np.meshgrid(c, c, c, indexing="ij")
    return np.exp(-4 * (qx ** 2 + qy ** 2 + qz ** 2)) * (0.75 + 0.25 * np.sin(12 * qx) * np.sin(12 * qy) * np.sin(12 * qz))


def build_c5(w: World, width=1920, height=1080, spp=256, volume_n=64, bars=8, with_sdf=True, with_volume=True) -> Config:
    F = lambda x: float(np.float32(x))  # a C# `F` literal widened to double
    light = w.LightMaterial(hm.WHITE, 180)
    d = F(4.0)
    for v in ((-1, -1, F(0.5)), (0, -1, F(0.25)), (-1, 1, 0)):
        w.add(w.sphere(hm.vmuls(hm.vnormalize(hm.vec(v)), d), F(0.25), light))
    if with_sdf:
        material = w.GlossyMaterial(hm.hex_color(0x468966), F(1.2), hm.radians(20))
        sphere = w.sdf_sphere(F(0.65))
        cube = w.sdf_cube((1, 1, 1))
        rounded = w.sdf_intersection([sphere, cube])
        a = w.sdf_cylinder(F(0.25), F(1.1))
        b = w.sdf_transform(a, hm.rotate((1, 0, 0), hm.radians(90)))
        c = w.sdf_transform(a, hm.rotate((0, 0, 1), hm.radians(90)))
        diff = w.sdf_difference([rounded, a, b, c])
        sdf = w.sdf_transform(diff, hm.rotate((0, 0, 1), hm.radians(30)))
        w.add(w.sdf_shape(sdf, material))
    w.add(w.plane((0, 0, F(-0.5)), (0, 0, 1), w.GlossyMaterial(hm.hex_color(0xFFF0A5), F(1.2), hm.radians(20))))
    colors = [0x730046, 0xBFBB11, 0xFFC200, 0xE88801, 0xC93C00]  # Example.cs:1288-1294
    mats = [w.GlossyMaterial(hm.hex_color(c), F(1.6), hm.radians(45)) for c in colors]
    for i in range(bars):
        v0 = (-1.5 + 0.45 * i, 2.0 + 0.25 * i, -0.5)
        v1 = (v0[0], v0[1], v0[2] + 1.0 + 0.1 * i)
        w.add(w.transformed_cylinder(v0, v1, 0.2, mats[i % len(mats)]))
    if with_volume:
        vcolors = [0x004358, 0x1F8A70, 0xBEDB39, 0xFFE11A, 0xFD7400]  # Example.cs:1446-1453
        windows = []
        for i, c in enumerate(vcolors):
            lo = F(0.2) + F(0.1) * i
            windows.append((lo, lo + F(0.01), w.GlossyMaterial(hm.hex_color(c), F(1.3), hm.radians(0))))
        vol = w.volume((-1, -1, F(-0.2)), (1, 1, 1), procedural_volume(volume_n), 1.0, windows)
        w.add(w.transformed(vol, hm.translate(hm.vec((0, -2.5, 0)))))
    w.look_at((-5, 0, 2), (0, 0, 0), (0, 0, 1), 45)
    w.sampler(4, 4, light_mode=LightModeAll, specular_mode=SpecularModeAll)
    return Config("c5_sdf_cylinder_volume", width, height, spp, "SDF solid + transformed cylinders + procedural Volume, three sphere lights")


BUILDERS = {"c1": build_c1, "c2": build_c2, "c3": build_c3, "c4": build_c4, "c5": build_c5}
