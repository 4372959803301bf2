"""Host-side value helpers that the reference evaluates in C# while a scene is being authored.

These run once per scene on the CPU (never per ray) and mirror the reference's quirks so that the Python scene
generators hand both back ends exactly the floats a C# `Example.cs` scene would:

* `Vector` stores float32 (Vector.cs:201-234): `vec()` rounds to float32, `vnormalize`/`vlength`/`vcross` use
  float32 arithmetic like System.Numerics.Vector3;
* `Matrix.Translate/Scale/Rotate` ignore `this` and return a fresh matrix (Matrix.cs:33-54);
* `Colour.HexColor` is sRGB/255f raised to 2.2f (Colour.cs:125-132).
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def vec(v) -> np.ndarray:
    return np.asarray(v, dtype=np.float64).astype(np.float32)


def vsub(a, b) -> np.ndarray:
    return (a.astype(np.float64) - b.astype(np.float64)).astype(np.float32)


def vadd(a, b) -> np.ndarray:
    return (a.astype(np.float64) + b.astype(np.float64)).astype(np.float32)


def vmuls(a, s: float) -> np.ndarray:
    return (a.astype(np.float64) * float(s)).astype(np.float32)


def vdot(a, b) -> float:
    xx, yy, zz = f32(a[0]) * f32(b[0]), f32(a[1]) * f32(b[1]), f32(a[2]) * f32(b[2])
    return float(f32(f32(xx + yy) + zz))


def vlength(a) -> float:
    return float(np.sqrt(f32(vdot(a, a))))


def vnormalize(a) -> np.ndarray:
    ln = np.sqrt(f32(vdot(a, a)))
    return np.array([f32(a[0]) / ln, f32(a[1]) / ln, f32(a[2]) / ln], dtype=np.float32)


def vcross(a, b) -> np.ndarray:
    a = a.astype(np.float32)
    b = b.astype(np.float32)
    return np.array([f32(a[1] * b[2]) - f32(a[2] * b[1]), f32(a[2] * b[0]) - f32(a[0] * b[2]),
                     f32(a[0] * b[1]) - f32(a[1] * b[0])], dtype=np.float32)


def radians(deg: float) -> float:  # Util.cs:13
    return deg * math.pi / 180


def identity() -> np.ndarray:
    return np.eye(4, dtype=np.float64)


def translate(v) -> np.ndarray:  # Matrix.cs:33-36
    m = np.eye(4, dtype=np.float64)
    m[0, 3], m[1, 3], m[2, 3] = float(v[0]), float(v[1]), float(v[2])
    return m


def scale(v) -> np.ndarray:  # Matrix.cs:38-41
    m = np.eye(4, dtype=np.float64)
    m[0, 0], m[1, 1], m[2, 2] = float(v[0]), float(v[1]), float(v[2])
    return m


def rotate(v, a: float) -> np.ndarray:  # Matrix.cs:44-54
    v = vnormalize(vec(v))
    x, y, z = float(v[0]), float(v[1]), float(v[2])
    s, c = math.sin(a), math.cos(a)
    k = 1 - c
    return np.array([
        [k * x * x + c, k * x * y + z * s, k * z * x - y * s, 0],
        [k * x * y - z * s, k * y * y + c, k * y * z + x * s, 0],
        [k * z * x + y * s, k * y * z - x * s, k * z * z + c, 0],
        [0, 0, 0, 1]], dtype=np.float64)


def mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Matrix.Mul (Matrix.cs:111-131): each entry is a left-to-right sum of four products."""
    r = np.zeros((4, 4), dtype=np.float64)
    for i in range(4):
        for j in range(4):
            r[i, j] = a[i, 0] * b[0, j] + a[i, 1] * b[1, j] + a[i, 2] * b[2, j] + a[i, 3] * b[3, j]
    return r


def hex_color(x: int):  # Colour.cs:125-132
    comps = [f32((x >> 16) & 0xFF) / f32(255.0), f32((x >> 8) & 0xFF) / f32(255.0), f32(x & 0xFF) / f32(255.0)]
    e = float(f32(2.2))
    return tuple(math.pow(float(c), e) for c in comps)


WHITE = (1.0, 1.0, 1.0)
BLACK = (0.0, 0.0, 0.0)
