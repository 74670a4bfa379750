"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol the public
header declares, and the ctypes mirrors of the parameter structs match the C layout."""
import ctypes as C

from vicasplat_b200 import _lib


def test_header_declares_expected_entry_points():
    syms = _lib.declared_symbols()
    for s in ["vs_rope_2d", "vs_gemm", "vs_layernorm", "vs_attention", "vs_raster_forward",
              "vs_raster_backward", "vs_raster_workspace_bytes", "vs_gaussian_adapter",
              "vs_last_error", "vs_struct_size"]:
        assert s in syms, s


def test_library_exports_every_declared_symbol(lib):
    for s in _lib.declared_symbols():
        assert hasattr(lib, s), s


def test_struct_layouts_match(lib):
    for name, cls in _lib.STRUCTS.items():
        assert lib.vs_struct_size(name.encode()) == C.sizeof(cls), name
    assert lib.vs_struct_size(b"nope") == -1


def test_argument_validation_without_gpu(lib):
    # NULL params are rejected before any CUDA call, with a message (mirrors TORCH_CHECK)
    assert lib.vs_gemm(None, None) == -1
    assert b"null" in lib.vs_last_error()
    assert lib.vs_raster_forward(None, None) == -1
    assert lib.vs_attention(None, None) == -1
    p = _lib.LayerNormParams()
    assert lib.vs_layernorm(C.byref(p), None) == -1
    assert lib.vs_raster_workspace_bytes(0, 0, 0, 0, 0) == 0
    assert lib.vs_raster_workspace_bytes(2, 1000, 64, 64, 5000) > 0


def test_rope_rejects_bad_dim(lib):
    buf = (C.c_float * 8)()
    pos = (C.c_int64 * 2)()
    rc = lib.vs_rope_2d(buf, 0, 1, 1, 1, 6, C.c_int64(6), C.c_int64(6), pos, C.c_float(100.0),
                        C.c_float(1.0), None)
    assert rc == -1 and b"multiple of 4" in lib.vs_last_error()
