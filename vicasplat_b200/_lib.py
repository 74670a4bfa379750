"""ctypes binding of the C-ABI in ``include/vicasplat_b200.h``.

There is no CPU fallback: if the shared library is missing or a symbol is absent, loading raises.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

_ROOT = Path(__file__).resolve().parent
import os as _os
# VICASPLAT_B200_LIB: load another build of the same C-ABI (A/B of kernel variants); same loud failure if absent
LIB_PATH = Path(_os.environ.get("VICASPLAT_B200_LIB") or _ROOT / "lib" / "libvicasplat_b200.so")
HEADER_PATH = _ROOT.parent / "include" / "vicasplat_b200.h"

VS_F32, VS_BF16, VS_F16, VS_F64 = 0, 1, 2, 3
VS_ACT_NONE, VS_ACT_GELU, VS_ACT_RELU = 0, 1, 2

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class GemmParams(C.Structure):
    _fields_ = [
        ("A", _vp), ("a_mode", _i32), ("a_rows", _i32), ("a_groups", _i32),
        ("a_row_stride", _i64), ("a_group_stride", _i64),
        ("cn", _i32), ("ch", _i32), ("cw", _i32), ("cin", _i32), ("kh", _i32), ("kw", _i32),
        ("pad", _i32), ("conv_in_h", _i32),
        ("conv_stride_x", _i64), ("conv_stride_y", _i64), ("conv_stride_n", _i64),
        ("W", _vp), ("w_row_stride", _i64), ("N", _i32), ("K", _i32),
        ("bias", _vp), ("act", _i32),
        ("gate", _vp), ("gate_ld", _i64), ("gate_rows", _i32), ("first_row_mode", _i32),
        ("res1", _vp), ("res2", _vp), ("res_dtype", _i32), ("res_up2", _i32), ("res_ld", _i64),
        ("C", _vp), ("c_dtype", _i32), ("ldc", _i64),
        ("C2", _vp), ("ldc2", _i64),
        ("out_gin", _i32), ("out_gout", _i32), ("out_off", _i32), ("block_n", _i32),
        ("rope_pos", _vp), ("rope_q_col", _i32), ("rope_k_col", _i32), ("rope_heads", _i32),
        ("rope_base", _f32), ("rope_cam_theta", _f32),
        ("mask_mode", _i32), ("c_accumulate", _i32), ("split_k", _i32), ("out_scale", _f32),
        ("operand_dtype", _i32),
    ]


class AdamWParams(C.Structure):
    _fields_ = [
        ("n_tensors", _i32), ("n_chunks", _i32),
        ("params", _vp), ("grads", _vp), ("exp_avg", _vp), ("exp_avg_sq", _vp),
        ("sizes", _vp), ("lrs", _vp), ("chunk_tensor", _vp), ("chunk_start", _vp),
        ("beta1", _f32), ("beta2", _f32), ("eps", _f32), ("weight_decay", _f32),
        ("step", _i32), ("max_grad_norm", _f32), ("skip_nonfinite", _i32),
        ("partials", _vp), ("counter", _vp), ("grad_norm_out", _vp), ("found_inf_out", _vp),
        ("step_counter", _vp),
    ]


class LayerNormParams(C.Structure):
    _fields_ = [
        ("x", _vp), ("ldx", _i64), ("rows", _i32), ("C", _i32),
        ("w", _vp), ("b", _vp), ("w0", _vp), ("b0", _vp),
        ("scale", _vp), ("shift", _vp), ("mod_ld", _i64), ("rows_per_frame", _i32),
        ("eps", _f32), ("normalize", _i32),
        ("y_bf16", _vp), ("ldy_bf16", _i64), ("y_f32", _vp), ("ldy_f32", _i64), ("y16_dtype", _i32),
    ]


class LayerNormBwdParams(C.Structure):
    _fields_ = [
        ("x", _vp), ("ldx", _i64), ("dy", _vp), ("dy_dtype", _i32), ("ldy", _i64),
        ("gamma", _vp), ("dres", _vp), ("ldres", _i64), ("dx", _vp), ("lddx", _i64),
        ("dgamma", _vp), ("dbeta", _vp), ("rows", _i32), ("C", _i32), ("eps", _f32),
    ]


class LnModBwdParams(C.Structure):
    _fields_ = [
        ("x", _vp), ("ldx", _i64), ("dh", _vp), ("dh_dtype", _i32), ("lddh", _i64),
        ("gamma", _vp), ("scale", _vp), ("mod_ld", _i64), ("dres", _vp), ("ldres", _i64),
        ("dx", _vp), ("lddx", _i64), ("frame_a", _vp), ("frame_b", _vp), ("frame_ld", _i64),
        ("frames", _i32), ("rows_per_frame", _i32), ("skip_first", _i32), ("C", _i32), ("eps", _f32),
    ]


class AttentionParams(C.Structure):
    _fields_ = [
        ("Q", _vp), ("K", _vp), ("V", _vp), ("O", _vp),
        ("ldq", _i64), ("ldk", _i64), ("ldv", _i64), ("ldo", _i64),
        ("q_rows", _i32), ("kv_rows", _i32), ("heads", _i32), ("items", _i32),
        ("q_start", _vp), ("q_len", _vp), ("kv_start0", _vp), ("kv_len0", _vp),
        ("kv_start1", _vp), ("kv_len1", _vp),
        ("max_q_len", _i32), ("max_kv_len", _i32), ("causal_block", _i32), ("scale", _f32),
        ("lse", _vp), ("dtype", _i32),
    ]


class AttentionBwdParams(C.Structure):
    _fields_ = [
        ("fwd", AttentionParams), ("dO", _vp), ("lddo", _i64),
        ("dQ", _vp), ("dK", _vp), ("dV", _vp), ("lddq", _i64), ("lddk", _i64), ("lddv", _i64),
        ("delta", _vp),
        ("dkv_items", _i32), ("dkv_max_kv_len", _i32),
        ("dkv_kv_start", _vp), ("dkv_kv_len", _vp), ("dkv_q_start0", _vp), ("dkv_q_len0", _vp),
        ("dkv_q_start1", _vp), ("dkv_q_len1", _vp),
    ]


class RasterParams(C.Structure):
    _fields_ = [
        ("V", _i32), ("G", _i32), ("H", _i32), ("W", _i32), ("gaussians_shared", _i32),
        ("means3D", _vp), ("cov3D", _vp), ("opacities", _vp), ("shs", _vp),
        ("sh_M", _i32), ("sh_degree", _i32), ("sh_stride_coef", _i32), ("sh_stride_chan", _i32),
        ("colors_precomp", _vp), ("viewmatrix", _vp), ("projmatrix", _vp), ("campos", _vp),
        ("tanfov", _vp), ("bg", _vp), ("scale_modifier", _f32),
        ("out_color", _vp), ("out_depth", _vp), ("out_alpha", _vp), ("radii", _vp),
        ("n_touched", _vp), ("final_T", _vp), ("n_contrib", _vp),
        ("workspace", _vp), ("workspace_bytes", _i64), ("max_pairs", _i64),
        ("num_pairs_out", _vp), ("max_tile_pairs", _i32),
    ]


class RasterBwdParams(C.Structure):
    _fields_ = [
        ("fwd", RasterParams),
        ("dL_dcolor", _vp), ("dL_ddepth", _vp), ("dL_dalpha", _vp),
        ("dL_dmeans3D", _vp), ("dL_dcov3D", _vp), ("dL_dopacity", _vp), ("dL_dshs", _vp),
        ("dL_dcolors", _vp), ("dL_dtau", _vp),
        ("bwd_workspace", _vp), ("bwd_workspace_bytes", _i64),
    ]


STRUCTS = {
    "vs_gemm_params": GemmParams,
    "vs_layernorm_params": LayerNormParams,
    "vs_attention_params": AttentionParams,
    "vs_raster_params": RasterParams,
    "vs_raster_bwd_params": RasterBwdParams,
    "vs_adamw_params": AdamWParams,
    "vs_layernorm_bwd_params": LayerNormBwdParams,
    "vs_attention_bwd_params": AttentionBwdParams,
    "vs_ln_mod_bwd_params": LnModBwdParams,
}

_DECL = re.compile(r"^\s*(?:const\s+char\s*\*|int64_t|int)\s+(vs_\w+)\s*\(", re.M)


def declared_symbols() -> list[str]:
    """Every function the public header declares."""
    return sorted(set(_DECL.findall(HEADER_PATH.read_text())))


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first. "
            "vicasplat_b200 has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"{LIB_PATH} does not export: {missing}")
    lib.vs_last_error.restype = C.c_char_p
    lib.vs_raster_workspace_bytes.restype = _i64
    lib.vs_raster_workspace_bytes.argtypes = [_i32, _i32, _i32, _i32, _i64]
    lib.vs_launch_count.restype = _i64
    lib.vs_mse_workspace_bytes.restype = _i64
    lib.vs_struct_size.restype = _i64
    lib.vs_struct_size.argtypes = [C.c_char_p]
    for name, cls in STRUCTS.items():
        n = lib.vs_struct_size(name.encode())
        if n != C.sizeof(cls):
            raise RuntimeError(f"ctypes layout of {name} ({C.sizeof(cls)} B) != C layout ({n} B)")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Turn a negative VS_ERR_* code into the RuntimeError the reference ops raise (TORCH_CHECK)."""
    if rc != 0:
        msg = load().vs_last_error().decode(errors="replace")
        raise RuntimeError(f"{what + ': ' if what else ''}{msg} (code {rc})")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
