"""Build the two in-tree shared libraries.

* ``ptsharp_b200/_lib/libptgpu.so``  — csrc/ptgpu.cu, nvcc, sm_100a only (``-gencode arch=compute_100a,code=sm_100a``),
  ``-fmad=false`` because the numeric model forbids a*b+c contraction (see csrc/pt_device.cuh), ``-lineinfo`` so ncu's
  source page maps to our code.  cudart is linked statically, so the library loads (and exports every symbol of
  include/ptgpu.h) on a box without a GPU; compute entry points then fail with PTGPU_E_CUDA.
* ``ptsharp_b200/_lib/libpthost.so`` — host/*.cpp, g++, ``-ffp-contract=off``; links libptgpu.so by $ORIGIN rpath.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIBDIR = os.path.join(PKG, "_lib")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "--expt-relaxed-constexpr", "--extended-lambda", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
              "-Xlinker", "-Bsymbolic"]  # -Bsymbolic: the library's own calls to ptgpu_* bind inside it (a checker build may be loaded beside it)
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wextra",
             "-Wno-unused-parameter"]


def _newer(srcs, out) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build_gpu(force=False, verbose=False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    out = os.path.join(LIBDIR, "libptgpu.so")
    srcs = [os.path.join(PKG, "csrc", "ptgpu.cu"), os.path.join(PKG, "csrc", "pt_device.cuh"),
            os.path.join(PKG, "csrc", "mesh_derive.hpp"), os.path.join(PKG, "csrc", "sh_funcs.hpp"), os.path.join(ROOT, "include", "ptgpu.h")]
    if force or _newer(srcs, out):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, srcs[0]]
        subprocess.check_call(cmd)
    return out


# Checker builds of the same source (tests only; never loaded by the product path).
VARIANTS = {"nocull": ["-DPT_NO_CULL=1"]}


def variant_path(name: str) -> str:
    return os.path.join(LIBDIR, "variants", f"libptgpu_{name}.so")


def build_variants(force=False) -> dict:
    """libptgpu_nocull.so: every culling shortcut of the tracer compiled out (PT_NO_CULL in csrc/pt_device.cuh) - the arbiter the
    GPU tests compare the product library with on 1e8 rays."""
    os.makedirs(os.path.join(LIBDIR, "variants"), exist_ok=True)
    srcs = [os.path.join(PKG, "csrc", "ptgpu.cu"), os.path.join(PKG, "csrc", "pt_device.cuh"),
            os.path.join(PKG, "csrc", "mesh_derive.hpp"), os.path.join(ROOT, "include", "ptgpu.h")]
    out = {}
    for name, flags in VARIANTS.items():
        path = variant_path(name)
        if force or _newer(srcs, path):
            subprocess.check_call([NVCC] + NVCC_FLAGS + flags + ["-o", path, srcs[0]])
        out[name] = path
    return out


def build_host(force=False) -> str:
    gpu = build_gpu()
    out = os.path.join(LIBDIR, "libpthost.so")
    cpps = [os.path.join(PKG, "host", f) for f in ("host.cpp", "capi.cpp", "loaders.cpp", "mc.cpp")]
    srcs = cpps + [os.path.join(PKG, "host", "ptsharp.hpp"), os.path.join(PKG, "host", "mc_table.inc"), os.path.join(PKG, "csrc", "sh_funcs.hpp"),
                   os.path.join(ROOT, "include", "ptgpu.h")]
    if force or _newer(srcs + [gpu], out):
        cmd = ["g++"] + CXX_FLAGS + ["-o", out] + cpps + ["-L" + LIBDIR, "-lptgpu", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return out


def build_all(force=False, verbose=False):
    out = build_gpu(force, verbose), build_host(force)
    build_variants(force)
    return out


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
