"""ctypes bindings for the two product libraries.

* ``libptgpu.so``  — the C ABI of include/ptgpu.h (CUDA, sm_100a).  `Device` wraps one ``ptgpu_ctx``.
* ``libpthost.so`` — the C++ host side (scene classes, reference kd-tree builder, flattener, Renderer), reached
  through its ``pth_*`` C entry points.  `HostWorld` is a scene being authored on it.

There is no CPU fallback anywhere in this module: if the CUDA library cannot be built/loaded, or no device is present,
the calls raise `PtgpuError`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import build as _build
from .authoring import World, bind, c_double_p, c_float_p, c_int_p

c_ll_p = C.POINTER(C.c_longlong)


class PtgpuError(RuntimeError):
    pass


class Camera(C.Structure):
    _fields_ = [("p", C.c_float * 3), ("u", C.c_float * 3), ("v", C.c_float * 3), ("w", C.c_float * 3),
                ("m", C.c_double), ("focalDistance", C.c_double), ("apertureRadius", C.c_double)]


class Pass(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("stratified", C.c_int32),
                ("sampleBase", C.c_int32), ("sampleStride", C.c_int32), ("firstHitSamples", C.c_int32),
                ("maxBounces", C.c_int32), ("directLighting", C.c_int32), ("softShadows", C.c_int32),
                ("lightMode", C.c_int32), ("specularMode", C.c_int32), ("seed", C.c_uint32), ("passIndex", C.c_uint32),
                ("camera", Camera), ("adaptiveSamples", C.c_int32), ("fireflySamples", C.c_int32),
                ("fireflyThreshold", C.c_double), ("serialRules", C.c_int32), ("flags", C.c_int32),
                ("adaptiveThreshold", C.c_double), ("adaptiveExponent", C.c_double)]


MAX_DEVICES = 8
PASS_RUSSIAN_ROULETTE = 1


class Params(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_int32), ("queueCapacity", C.c_uint64), ("numDevices", C.c_int32),
                ("devices", C.c_int32 * MAX_DEVICES), ("reserved", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("cameraSamples", C.c_uint64), ("segments", C.c_uint64), ("shadowRays", C.c_uint64),
                ("nanSamples", C.c_uint64), ("kernelLaunches", C.c_uint64), ("lastPassMs", C.c_double),
                ("traceMs", C.c_double), ("shadeMs", C.c_double), ("shadowMs", C.c_double), ("raygenMs", C.c_double),
                ("meshMs", C.c_double), ("meshItems", C.c_uint64), ("meshLaunches", C.c_uint64), ("traceLaunches", C.c_uint64),
                ("queueOverflows", C.c_uint64), ("devices", C.c_uint64), ("sdfMs", C.c_double), ("sdfItems", C.c_uint64),
                ("sdfLaunches", C.c_uint64), ("volumeMs", C.c_double), ("volumeItems", C.c_uint64), ("volumeLaunches", C.c_uint64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/ptgpu.h declares (tests check the library exports all of them)
PTGPU_SYMBOLS = [
    "ptgpu_abi_version", "ptgpu_abi_sizeof", "ptgpu_create", "ptgpu_destroy", "ptgpu_last_error", "ptgpu_upload_scene", "ptgpu_scene_bytes",
    "ptgpu_render_pass", "ptgpu_accumulate_device", "ptgpu_add_sample_device", "ptgpu_read_buffer", "ptgpu_reset_buffer",
    "ptgpu_intersect_batch", "ptgpu_cast_rays", "ptgpu_keyed_draw", "ptgpu_get_counters", "ptgpu_reset_counters",
    "ptgpu_set_profiling", "ptgpu_export_buffer", "ptgpu_import_buffer",
]

_gpu = None
_host = None


def _bind_gpu(lib: C.CDLL) -> C.CDLL:
    lib.ptgpu_abi_version.restype = C.c_int
    lib.ptgpu_abi_sizeof.restype = C.c_int
    lib.ptgpu_abi_sizeof.argtypes = [C.c_int]
    lib.ptgpu_create.restype = C.c_int
    lib.ptgpu_create.argtypes = [C.POINTER(Params), C.POINTER(C.c_void_p)]
    lib.ptgpu_destroy.restype = None
    lib.ptgpu_destroy.argtypes = [C.c_void_p]
    lib.ptgpu_last_error.restype = C.c_char_p
    lib.ptgpu_last_error.argtypes = [C.c_void_p]
    lib.ptgpu_upload_scene.argtypes = [C.c_void_p, C.c_void_p]
    lib.ptgpu_scene_bytes.restype = C.c_uint64
    lib.ptgpu_scene_bytes.argtypes = [C.c_void_p]
    lib.ptgpu_render_pass.argtypes = [C.c_void_p, C.POINTER(Pass), c_float_p]
    lib.ptgpu_accumulate_device.argtypes = [C.c_void_p, C.POINTER(Pass), C.c_void_p, C.c_void_p]
    lib.ptgpu_add_sample_device.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_double, C.c_void_p]
    lib.ptgpu_read_buffer.argtypes = [C.c_void_p, C.c_int32, c_float_p]
    lib.ptgpu_reset_buffer.argtypes = [C.c_void_p]
    lib.ptgpu_export_buffer.argtypes = [C.c_void_p, c_int_p, c_int_p, c_double_p, c_double_p, c_int_p]
    lib.ptgpu_import_buffer.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, c_int_p]
    lib.ptgpu_intersect_batch.argtypes = [C.c_void_p, C.c_int32, c_float_p, c_float_p, c_int_p, c_int_p, c_double_p,
                                          c_float_p, c_float_p, c_int_p, c_int_p]
    lib.ptgpu_cast_rays.argtypes = [C.c_void_p, C.POINTER(Pass), C.c_int32, c_int_p, c_int_p, c_double_p, c_double_p,
                                    c_int_p, c_float_p, c_float_p]
    lib.ptgpu_keyed_draw.argtypes = [C.c_void_p] + [C.c_uint32] * 9 + [c_double_p]
    lib.ptgpu_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.ptgpu_reset_counters.argtypes = [C.c_void_p]
    lib.ptgpu_set_profiling.argtypes = [C.c_void_p, C.c_int32]
    return lib


def gpu_lib() -> C.CDLL:
    global _gpu
    if _gpu is None:
        path = os.environ.get("PTGPU_LIB") or _build.build_gpu()  # PTGPU_LIB: development override (tuning variants)
        _gpu = _bind_gpu(C.CDLL(path, mode=C.RTLD_GLOBAL))
    return _gpu


def checker_lib(name: str = "nocull") -> C.CDLL:
    """A checker build of libptgpu (ptsharp_b200/build.py VARIANTS) loaded BESIDE the product library: tests only."""
    return _bind_gpu(C.CDLL(_build.build_variants()[name], mode=C.RTLD_LOCAL))


_HOST_EXTRA = {
    "last_error": (C.c_char_p, [C.c_void_p]),
    "flatten": (C.c_void_p, [C.c_void_p]),
    "flat_bytes": (C.c_uint64, [C.c_void_p]),
    "save_flat": (C.c_int, [C.c_void_p, C.c_char_p]),
    "load_flat": (C.c_void_p, [C.c_void_p, C.c_char_p]),
    "make_pass": (None, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.c_uint, C.c_int, C.c_int, C.POINTER(Pass)]),
    "tree_stats": (C.c_int, [C.c_void_p, C.c_int, c_ll_p, c_float_p]),
    "tree_dump": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_double_p, c_int_p, c_int_p, c_int_p]),
    "builder_friendly_order": (None, [C.c_int, c_float_p, C.c_double, C.c_int, C.c_int, c_int_p]),
    "renderer_new": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "renderer_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_uint]),
    "renderer_devices": (C.c_int, [C.c_void_p, C.c_int, c_int_p]),
    "renderer_set_extra": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "renderer_render": (C.c_int, [C.c_void_p, c_float_p]),
    "renderer_iterative": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "renderer_image": (C.c_int, [C.c_void_p, C.c_int, c_float_p]),
    "renderer_counters": (C.c_int, [C.c_void_p, C.POINTER(Counters)]),
    "renderer_ctx": (C.c_void_p, [C.c_void_p]),
    "load_model": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_int]),
    "mc_mesh": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, C.c_double, C.c_int]),
    "spherical_harmonic": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "mesh_op": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_double_p]),
    "mesh_get": (C.c_int, [C.c_void_p, C.c_int, c_float_p, c_float_p, c_float_p]),
}


def host_lib() -> C.CDLL:
    global _host
    if _host is None:
        gpu_lib()
        lib = C.CDLL(_build.build_host())
        bind(lib, "pth_", _HOST_EXTRA)
        lib.pth_mc_case.restype = C.c_int; lib.pth_mc_case.argtypes = [C.c_int, c_int_p]
        lib.pth_mc_edges.restype = C.c_int; lib.pth_mc_edges.argtypes = [C.c_int]
        _host = lib
    return _host


def builder_friendly_order(V: np.ndarray, balance: float = 0.70, min_repair: int = 16, verbose: bool = False) -> np.ndarray:
    """Permutation of triangles (ntri,3,3) for which the unmodified reference kd builder yields a balanced tree."""
    V = np.ascontiguousarray(V, dtype=np.float32)
    perm = np.empty(V.shape[0], np.int32)
    host_lib().pth_builder_friendly_order(V.shape[0], V.ctypes.data_as(c_float_p), float(balance), min_repair, int(verbose),
                                          perm.ctypes.data_as(c_int_p))
    return perm


class HostWorld(World):
    """A scene authored on the C++ host library (the product path)."""

    def __init__(self):
        super().__init__(host_lib(), "pth_")
        self._flat = None

    def _err(self) -> str:
        e = self.lib.pth_last_error(self.h)
        return e.decode() if e else ""

    def flatten(self) -> int:
        """Scene.Compile() + flatten; returns the address of the ptgpu_flat_scene view."""
        p = self.lib.pth_flatten(self.h)
        if not p:
            raise PtgpuError("flatten: " + self._err())
        self._flat = p
        return p

    def save_flat(self, path: str):
        """Write the flattened scene (flatten() first) to a flat-scene file."""
        if self.lib.pth_save_flat(self.h, os.fsencode(path)) != 0:
            raise RuntimeError(self._err())

    def load_flat(self, path: str) -> int:
        """Replace this world's flat scene by the one in `path`; returns the ptgpu_flat_scene pointer for Device.upload_flat."""
        p = self.lib.pth_load_flat(self.h, os.fsencode(path))
        if not p:
            raise RuntimeError(self._err())
        return p

    def flat_bytes(self) -> int:
        return int(self.lib.pth_flat_bytes(self.h))

    # ---- model loaders and Mesh utilities (host/loaders.cpp; OBJ.cs, STL.cs, Mesh.cs:141-289) ----
    def load_obj(self, path: str, mat: int) -> int:
        return self._model(0, path, mat)

    def load_stl(self, path: str, mat: int) -> int:
        return self._model(1, path, mat)

    def _model(self, kind: int, path: str, mat: int) -> int:
        s = self.lib.pth_load_model(self.h, kind, os.fsencode(path), mat)
        if s < 0:
            raise RuntimeError(self._err())
        return s

    # ---- marching cubes / spherical harmonics (host/mc.cpp; MC.cs, SH.cs) ----
    def mc_mesh(self, sdf: int, bmin, bmax, step: float, mat: int = -1) -> int:
        """MC.NewSDFMesh(sdf, box, step): the marching-cubes Mesh of an SDF (triangles in the reference's order)."""
        from .authoring import _d3
        s = self.lib.pth_mc_mesh(self.h, sdf, _d3(bmin), _d3(bmax), float(step), mat)
        if s < 0:
            raise RuntimeError(self._err())
        return s

    def spherical_harmonic(self, l: int, m: int, pm: int, nm: int, step: float = 0.0) -> int:
        """SphericalHarmonic.NewSphericalHarmonic(l, m, pm, nm); step = 0: the reference's 0.01F grid."""
        s = self.lib.pth_spherical_harmonic(self.h, l, m, pm, nm, float(step))
        if s < 0:
            raise RuntimeError(self._err())
        return s

    @staticmethod
    def mc_case(index: int):
        """(triangleTable[index] as a list of cube edges, edgetable[index]) of the marching-cubes tables (MC.cs:135-429)."""
        lib = host_lib()
        out = (C.c_int * 15)()
        n = lib.pth_mc_case(index, out)
        return [out[i] for i in range(3 * n)], lib.pth_mc_edges(index)

    def _mesh_op(self, shape: int, op: int, args) -> None:
        a = np.ascontiguousarray(np.asarray(list(args) + [0.0], dtype=np.float64))
        if self.lib.pth_mesh_op(self.h, shape, op, a.ctypes.data_as(c_double_p)) != 0:
            raise RuntimeError(self._err())

    def mesh_smooth_normals(self, shape): self._mesh_op(shape, 0, [])
    def mesh_smooth_normals_threshold(self, shape, radians): self._mesh_op(shape, 1, [radians])
    def mesh_move_to(self, shape, position, anchor): self._mesh_op(shape, 2, list(position) + list(anchor))
    def mesh_fit_inside(self, shape, bmin, bmax, anchor): self._mesh_op(shape, 3, list(bmin) + list(bmax) + list(anchor))
    def mesh_transform(self, shape, m16): self._mesh_op(shape, 4, list(np.asarray(m16, dtype=np.float64).ravel()))
    def mesh_set_material(self, shape, mat): self._mesh_op(shape, 5, [mat])

    def mesh_triangles(self, shape: int):
        """(V, N, T) of a Mesh, each (ntri, 3, 3) float32."""
        n = self.lib.pth_mesh_get(self.h, shape, None, None, None)
        if n < 0:
            raise RuntimeError(self._err())
        V, N, T = (np.empty((n, 3, 3), np.float32) for _ in range(3))
        self.lib.pth_mesh_get(self.h, shape, V.ctypes.data_as(c_float_p), N.ctypes.data_as(c_float_p), T.ctypes.data_as(c_float_p))
        return V, N, T

    def make_pass(self, width, height, spp, stratified=False, seed=0x50545348, pass_index=0, sample_base=0,
                  sample_stride=1, adaptive_samples=0, firefly_samples=0, firefly_threshold=1.0, serial_rules=False,
                  adaptive_threshold=1.0, adaptive_exponent=1.0, russian_roulette=False) -> Pass:
        p = Pass()
        self.lib.pth_make_pass(self.h, width, height, spp, int(stratified), seed, pass_index, sample_base, sample_stride,
                               C.byref(p))
        p.adaptiveSamples, p.fireflySamples, p.fireflyThreshold = adaptive_samples, firefly_samples, firefly_threshold
        p.serialRules, p.adaptiveThreshold, p.adaptiveExponent = int(serial_rules), adaptive_threshold, adaptive_exponent
        if russian_roulette:
            p.flags |= PASS_RUSSIAN_ROULETTE  # opt-in, not part of parity mode (ptgpu.h)
        return p

    def tree_stats(self, which=-1):
        out = (C.c_longlong * 4)()
        box = (C.c_float * 6)()
        if self.lib.pth_tree_stats(self.h, which, out, box) != 0:
            raise ValueError("not a mesh")
        return dict(nodes=out[0], leafItems=out[1], maxLeaf=out[2], maxDepth=out[3], box=np.array(list(box), np.float32))

    def tree_dump(self, which=-1):
        st = self.tree_stats(which)
        n, m = st["nodes"], st["leafItems"]
        axis = np.empty(n, np.int32); point = np.empty(n, np.float64)
        a = np.empty(n, np.int32); b = np.empty(n, np.int32); items = np.empty(max(m, 1), np.int32)
        self.lib.pth_tree_dump(self.h, which, axis.ctypes.data_as(c_int_p), point.ctypes.data_as(c_double_p),
                               a.ctypes.data_as(c_int_p), b.ctypes.data_as(c_int_p), items.ctypes.data_as(c_int_p))
        return dict(axis=axis, point=point, a=a, b=b, items=items[:m], box=st["box"])

    # Renderer.cs mirror ------------------------------------------------------------------------------------------
    def new_renderer(self, width, height, device=0):
        if self.lib.pth_renderer_new(self.h, width, height, device) != 0:
            raise PtgpuError(self._err())

    def renderer_devices(self, devices):
        """Renderer.Devices: split every pass over these GPUs inside the handle (single process)."""
        a = (C.c_int * len(devices))(*devices)
        if self.lib.pth_renderer_devices(self.h, len(devices), a) != 0:
            raise PtgpuError(self._err())

    def renderer_set(self, samples_per_pixel, stratified=False, seed=0x50545348):
        if self.lib.pth_renderer_set(self.h, samples_per_pixel, int(stratified), seed) != 0:
            raise PtgpuError(self._err())

    def renderer_set_extra(self, adaptive_samples=0, firefly_samples=0):
        if self.lib.pth_renderer_set_extra(self.h, adaptive_samples, firefly_samples) != 0:
            raise PtgpuError(self._err())

    def render_parallel(self, width, height) -> np.ndarray:
        out = np.empty((height, width, 3), np.float32)
        if self.lib.pth_renderer_render(self.h, out.ctypes.data_as(c_float_p)) != 0:
            raise PtgpuError(self._err())
        return out

    def iterative_render(self, path_template: str, iterations: int):
        if self.lib.pth_renderer_iterative(self.h, path_template.encode(), iterations) != 0:
            raise PtgpuError(self._err())

    def renderer_image(self, width, height, channel=0) -> np.ndarray:
        out = np.empty((height, width, 3), np.float32)
        if self.lib.pth_renderer_image(self.h, channel, out.ctypes.data_as(c_float_p)) != 0:
            raise PtgpuError(self._err())
        return out

    def renderer_counters(self) -> dict:
        c = Counters()
        if self.lib.pth_renderer_counters(self.h, C.byref(c)) != 0:
            raise PtgpuError(self._err())
        return c.as_dict()


class Device:
    """One ptgpu_ctx: a CUDA device with an uploaded scene, its wavefront queues and its image Buffer."""

    def __init__(self, device: int = 0, queue_capacity: int = 0, devices=None, lib: Optional[C.CDLL] = None):
        """devices: a list of CUDA ordinals -> one handle that splits every pass over them (ptgpu_params.devices).
        lib: a checker build (checker_lib()) instead of the product library - tests only."""
        self.lib = lib or gpu_lib()
        self.h = C.c_void_p()
        p = Params()
        p.device, p.flags, p.queueCapacity = device, 0, queue_capacity
        if devices:
            p.numDevices = len(devices)
            for k, d in enumerate(devices):
                p.devices[k] = d
        rc = self.lib.ptgpu_create(C.byref(p), C.byref(self.h))
        if rc != 0:
            e = self.lib.ptgpu_last_error(None)
            raise PtgpuError(f"ptgpu_create failed ({rc}): {e.decode() if e else ''}")

    def close(self):
        if self.h:
            self.lib.ptgpu_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            e = self.lib.ptgpu_last_error(self.h)
            raise PtgpuError(f"{what} failed ({rc}): {e.decode() if e else ''}")

    def upload(self, world: HostWorld):
        self._ck(self.lib.ptgpu_upload_scene(self.h, world.flatten()), "ptgpu_upload_scene")

    def upload_flat(self, flat_ptr: int):
        self._ck(self.lib.ptgpu_upload_scene(self.h, flat_ptr), "ptgpu_upload_scene")

    def scene_bytes(self) -> int:
        return int(self.lib.ptgpu_scene_bytes(self.h))

    def render_pass(self, p: Pass, want_mean=True, out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        if want_mean and out is None:
            out = np.empty((p.height, p.width, 3), np.float32)
        ptr = out.ctypes.data_as(c_float_p) if out is not None else None
        self._ck(self.lib.ptgpu_render_pass(self.h, C.byref(p), ptr), "ptgpu_render_pass")
        return out

    def accumulate_device(self, p: Pass, d_sum_ptr: int, stream: int = 0):
        self._ck(self.lib.ptgpu_accumulate_device(self.h, C.byref(p), C.c_void_p(d_sum_ptr), C.c_void_p(stream)),
                 "ptgpu_accumulate_device")

    def add_sample_device(self, width, height, d_sum_ptr: int, divisor: float, stream: int = 0):
        self._ck(self.lib.ptgpu_add_sample_device(self.h, width, height, C.c_void_p(d_sum_ptr), float(divisor),
                                                  C.c_void_p(stream)), "ptgpu_add_sample_device")

    def read_buffer(self, width, height, channel=0) -> np.ndarray:
        out = np.empty((height, width, 3), np.float32)
        self._ck(self.lib.ptgpu_read_buffer(self.h, channel, out.ctypes.data_as(c_float_p)), "ptgpu_read_buffer")
        return out

    def export_buffer(self):
        """Exact Welford state (M, V as float64 HxWx3, samples as int32 HxW): the checkpoint of an IterativeRender loop."""
        w, h = C.c_int32(0), C.c_int32(0)
        self._ck(self.lib.ptgpu_export_buffer(self.h, C.byref(w), C.byref(h), None, None, None), "ptgpu_export_buffer")
        M = np.empty((h.value, w.value, 3), np.float64); V = np.empty_like(M); n = np.empty((h.value, w.value), np.int32)
        self._ck(self.lib.ptgpu_export_buffer(self.h, C.byref(w), C.byref(h), M.ctypes.data_as(c_double_p), V.ctypes.data_as(c_double_p),
                                              n.ctypes.data_as(c_int_p)), "ptgpu_export_buffer")
        return M, V, n

    def import_buffer(self, M: np.ndarray, V: np.ndarray, samples: np.ndarray):
        M = np.ascontiguousarray(M, np.float64); V = np.ascontiguousarray(V, np.float64); samples = np.ascontiguousarray(samples, np.int32)
        h, w = samples.shape
        self._ck(self.lib.ptgpu_import_buffer(self.h, w, h, M.ctypes.data_as(c_double_p), V.ctypes.data_as(c_double_p),
                                              samples.ctypes.data_as(c_int_p)), "ptgpu_import_buffer")

    def reset_buffer(self):
        self._ck(self.lib.ptgpu_reset_buffer(self.h), "ptgpu_reset_buffer")

    def intersect_batch(self, o: np.ndarray, d: np.ndarray, full: bool = True) -> dict:
        """Scene.Intersect (+ Hit.Info when `full`) on caller-supplied rays."""
        o = np.ascontiguousarray(o, dtype=np.float32); d = np.ascontiguousarray(d, dtype=np.float32)
        n = o.shape[0]
        out = dict(shape=np.empty(n, np.int32), prim=np.empty(n, np.int32), t=np.empty(n, np.float64))
        if full:
            out.update(normal=np.empty((n, 3), np.float32), position=np.empty((n, 3), np.float32),
                       inside=np.empty(n, np.int32), material=np.empty(n, np.int32))
        fp = lambda k: out[k].ctypes.data_as(c_float_p) if k in out else None
        ip = lambda k: out[k].ctypes.data_as(c_int_p) if k in out else None
        self._ck(self.lib.ptgpu_intersect_batch(
            self.h, n, o.ctypes.data_as(c_float_p), d.ctypes.data_as(c_float_p), out["shape"].ctypes.data_as(c_int_p),
            out["prim"].ctypes.data_as(c_int_p), out["t"].ctypes.data_as(c_double_p),
            fp("normal"), fp("position"), ip("inside"), ip("material")), "ptgpu_intersect_batch")
        return out

    def cast_rays(self, p: Pass, x, y, fu, fv, sample):
        x = np.ascontiguousarray(x, np.int32); y = np.ascontiguousarray(y, np.int32)
        fu = np.ascontiguousarray(fu, np.float64); fv = np.ascontiguousarray(fv, np.float64)
        sample = np.ascontiguousarray(sample, np.int32)
        n = x.shape[0]
        o = np.empty((n, 3), np.float32); d = np.empty((n, 3), np.float32)
        self._ck(self.lib.ptgpu_cast_rays(self.h, C.byref(p), n, x.ctypes.data_as(c_int_p), y.ctypes.data_as(c_int_p),
                                          fu.ctypes.data_as(c_double_p), fv.ctypes.data_as(c_double_p),
                                          sample.ctypes.data_as(c_int_p), o.ctypes.data_as(c_float_p),
                                          d.ctypes.data_as(c_float_p)), "ptgpu_cast_rays")
        return o, d

    def keyed_draw(self, seed, pass_index, pixel, sample, bits, first, depth, sub, draw_index) -> float:
        out = C.c_double()
        self._ck(self.lib.ptgpu_keyed_draw(self.h, seed, pass_index, pixel, sample, bits, first, depth, sub, draw_index,
                                           C.byref(out)), "ptgpu_keyed_draw")
        return out.value

    def counters(self) -> dict:
        c = Counters()
        self._ck(self.lib.ptgpu_get_counters(self.h, C.byref(c)), "ptgpu_get_counters")
        return c.as_dict()

    def reset_counters(self):
        self._ck(self.lib.ptgpu_reset_counters(self.h), "ptgpu_reset_counters")

    def set_profiling(self, on: bool):
        self._ck(self.lib.ptgpu_set_profiling(self.h, int(on)), "ptgpu_set_profiling")
