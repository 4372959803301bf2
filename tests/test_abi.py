"""The C-ABI library loads without a GPU, exports every symbol include/ptgpu.h declares, and fails loudly (no CPU
fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ptgpu.h")).read()
    return sorted(set(re.findall(r"\b(ptgpu_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported(bindings):
    lib = bindings.gpu_lib()
    declared = _declared_symbols()
    assert set(declared) == set(bindings.PTGPU_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ptgpu_abi_version() == 2


ABI_STRUCTS = ["ptgpu_params", "ptgpu_pass", "ptgpu_camera", "ptgpu_counters", "ptgpu_flat_scene", "ptgpu_node", "ptgpu_tree", "ptgpu_shape",
               "ptgpu_sphere", "ptgpu_cube", "ptgpu_plane", "ptgpu_cylinder", "ptgpu_mesh", "ptgpu_tri_geom", "ptgpu_tri_shade",
               "ptgpu_instance", "ptgpu_sdf_op", "ptgpu_sdf_shape", "ptgpu_volume_window", "ptgpu_volume", "ptgpu_material", "ptgpu_texture", "ptgpu_sh"]
# what a plain C compiler (hence a P/Invoke [StructLayout(LayoutKind.Sequential)] mirror) lays out for include/ptgpu.h
ABI_SIZES = [56, 168, 72, 176, 392, 16, 32, 16, 32, 24, 24, 24, 16, 48, 64, 272, 136, 32, 24, 64, 96, 16, 32]


def test_struct_layouts_match_header(bindings, tmp_path):
    """Every struct of the ABI: the size nvcc compiled into libptgpu (ptgpu_abi_sizeof) == the size gcc gives the same header as C
    == the documented constant, and the ctypes mirrors the tests marshal through agree."""
    import subprocess
    lib = bindings.gpu_lib()
    got = [lib.ptgpu_abi_sizeof(k) for k in range(len(ABI_STRUCTS))]
    assert got == ABI_SIZES, dict(zip(ABI_STRUCTS, got))
    assert lib.ptgpu_abi_sizeof(len(ABI_STRUCTS)) == -1
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "ptgpu.h"\nint main(void){' +
                   "".join(f'printf("%zu\\n", sizeof({n}));' for n in ABI_STRUCTS) + "return 0;}\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    assert [int(x) for x in subprocess.check_output([str(exe)]).split()] == ABI_SIZES
    for k, cls in ((0, bindings.Params), (1, bindings.Pass), (2, bindings.Camera), (3, bindings.Counters)):
        assert C.sizeof(cls) == ABI_SIZES[k], cls


def test_flatten_c1(bindings):
    from ptsharp_b200 import scenes
    w = bindings.HostWorld()
    scenes.build_c1(w)
    assert w.flatten() != 0
    assert w.flat_bytes() > 0
    p = w.make_pass(64, 48, 4)
    assert (p.width, p.height, p.spp, p.firstHitSamples, p.maxBounces) == (64, 48, 4, 16, 4)
    assert p.camera.m == pytest.approx(1 / __import__("math").tan(50 * __import__("math").pi / 360))


def test_nested_transform_flattens_as_a_chain(bindings):
    """TransformedShape of TransformedShape (TransformedShape.cs:29-32 takes any IShape): the flattener keeps the chain, inner shapes
    after the scene shapes; the device walks it level by level (tests/test_gpu_parity.py::test_nested_transformed_shapes_match_the_oracle)."""
    import numpy as np
    w = bindings.HostWorld()
    m = w.DiffuseMaterial((1, 1, 1))
    inner = w.transformed(w.sphere((0, 0, 0), 1, m), np.eye(4))
    w.add(w.transformed(inner, np.eye(4)))
    assert w.flatten()


def test_no_cpu_fallback(bindings):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(bindings.PtgpuError) as e:
        bindings.Device()
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    w = bindings.HostWorld()
    from ptsharp_b200 import scenes
    scenes.build_c1(w)
    w.new_renderer(32, 32)
    with pytest.raises(bindings.PtgpuError):
        w.render_parallel(32, 32)


def test_flat_scene_file_round_trip(bindings, tmp_path):
    """SaveFlatScene / LoadFlatScene: every array and header field of the flat scene survives the file bit for bit."""
    import ctypes as C
    from ptsharp_b200 import scenes
    hw = bindings.HostWorld()
    scenes.build_c4(hw, freq=6, nx=3, nz=2, tex=16)
    p0 = hw.flatten()
    n0 = hw.flat_bytes()
    path = str(tmp_path / "c4.ptfs")
    hw.save_flat(path)

    def snapshot(ptr):
        # header scalars + a checksum of the arrays through the known layout: count fields are the u32/u64 before each pointer
        raw = C.string_at(ptr, 12)
        return raw

    head0 = snapshot(p0)
    other = bindings.HostWorld()
    p1 = other.load_flat(path)
    assert other.flat_bytes() == n0 and snapshot(p1) == head0
    # a second save of the loaded scene is byte-identical to the first file
    path2 = str(tmp_path / "c4b.ptfs")
    other.save_flat(path2)
    with open(path, "rb") as a, open(path2, "rb") as b:
        assert a.read() == b.read()
    with open(path, "r+b") as f:  # corrupt the magic: loud failure, not garbage
        f.write(b"XXXX")
    import pytest
    with pytest.raises(RuntimeError):
        bindings.HostWorld().load_flat(path)


def _derive(bindings, flat):
    lib = bindings.gpu_lib()
    lib.ptgpu_debug_derive.restype = C.c_double
    lib.ptgpu_debug_derive.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    h, nr, nt = C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert lib.ptgpu_debug_derive(flat, C.byref(h), C.byref(nr), C.byref(nt)) >= 0
    return h.value, nr.value, nt.value


def test_mesh_derivation_is_thread_count_invariant(bindings, monkeypatch):
    """ptgpu_upload_scene derives the walk's node records / sorted leaf triangles on all host threads (csrc/mesh_derive.hpp):
    the bytes must not depend on the thread count and must equal the
    single-threaded derivation the earlier GPU parity runs were made with (hash frozen from that implementation)."""
    from ptsharp_b200 import scenes
    hw = bindings.HostWorld()
    scenes.build_c3(hw, freq_a=24, freq_b=12)
    flat = hw.flatten()
    monkeypatch.setenv("PTGPU_HOST_THREADS", "1")
    one = _derive(bindings, flat)
    monkeypatch.setenv("PTGPU_HOST_THREADS", "7")
    many = _derive(bindings, flat)
    assert one == many
    assert many[1:] == (14883, 46795)            # node records (reference + bounds-only), leaf triangles
    assert many[0] == 0xF03E2BC40DFF34E1, hex(many[0])
