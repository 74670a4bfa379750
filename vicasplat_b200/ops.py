"""Tensor-level wrappers of the C-ABI entry points (include/vicasplat_b200.h).

PyTorch is used for device memory and streams only; every function here launches hand-written
sm_100a kernels on ``torch.cuda.current_stream()`` and never synchronises.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (AttentionParams, GemmParams, LayerNormParams, VS_ACT_GELU, VS_ACT_NONE,
                   VS_ACT_RELU, VS_BF16, VS_F16, VS_F32, check, ptr, stream_ptr)

_DT = {torch.float32: VS_F32, torch.bfloat16: VS_BF16, torch.float16: VS_F16}
_HALF = (torch.bfloat16, torch.float16)   # the two 16-bit operand formats (speed / parity mode)

# Optional per-kernel-family timing (bench.py's roofline leg): when TIMERS is a dict, every wrapper
# brackets its launch with CUDA events on the launching stream; `family_ms()` sums them afterwards.
TIMERS = None


class _timed:
    def __init__(self, family, meta=None):
        self.family, self.meta = family, meta

    def __enter__(self):
        if TIMERS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if TIMERS is not None:
            self.e1.record()
            TIMERS.setdefault(self.family, []).append((self.e0, self.e1, self.meta))


def family_ms(timers) -> dict:
    torch.cuda.synchronize()
    return {k: sum(e[0].elapsed_time(e[1]) for e in v) for k, v in timers.items()}


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("vicasplat_b200 ops need CUDA tensors (there is no CPU fallback)")


def gemm(A, W, *, N=None, K=None, a_rows=None, a_groups=1, a_row_stride=None, a_group_stride=0,
         bias=None, act=VS_ACT_NONE, gate=None, gate_rows=0, first_row_mode=0, res1=None,
         res2=None, out=None, out_dtype=torch.bfloat16, ldc=None, out2=None, ldc2=None,
         out_gin=0, out_gout=0, out_off=0, out_rows=None, block_n=0, w_row_stride=None,
         rope=None, tn=False):
    """C = epilogue(A @ W^T): rows mode of vs_gemm.  A bf16 (rows, K) [or strided groups], W bf16 (N, K).
    rope = (pos_i32 (out_rows, 2), q_col, k_col, heads, base, cam_theta): rotary embedding of the
    q / k columns applied in the epilogue (see vs_gemm_params.rope_pos)."""
    _need_cuda(A, W)
    lib = _lib.load()
    p = GemmParams()
    if tn:
        # a_mode 2: A (K, rows), W (K, N) as stored -> C (rows, N) = A^T W  (the wgrad form)
        assert A.dim() == 2 and W.dim() == 2 and A.shape[0] == W.shape[0] and a_groups == 1
        K, rows, N = A.shape[0], A.shape[1], W.shape[1]
        p.A, p.a_mode, p.a_rows, p.a_groups = ptr(A), 2, rows, 1
        p.a_row_stride = A.stride(0)
        p.W, p.w_row_stride, p.N, p.K = ptr(W), W.stride(0), N, K
    else:
        K = K if K is not None else A.shape[-1]
        N = N if N is not None else W.shape[0]
        rows = a_rows if a_rows is not None else A.numel() // A.shape[-1]
        p.A, p.a_mode, p.a_rows, p.a_groups = ptr(A), 0, rows, a_groups
        p.a_row_stride = a_row_stride if a_row_stride is not None else A.stride(-2)
        p.a_group_stride = a_group_stride
        p.W, p.w_row_stride, p.N, p.K = ptr(W), (w_row_stride or W.stride(0)), N, K
    _fill_epilogue(p, bias, act, gate, gate_rows, first_row_mode, res1, res2)
    assert A.dtype in _HALF and W.dtype == A.dtype, "GEMM operands are bf16 (speed) or fp16 (parity mode)"
    p.operand_dtype = _DT[A.dtype]
    if out_dtype in _HALF:
        out_dtype = A.dtype
    total = rows * a_groups
    if out is None:
        n_out = out_rows if out_rows is not None else total
        out = torch.empty((n_out, N), dtype=out_dtype, device=A.device)
    p.C, p.c_dtype, p.ldc = ptr(out), _DT[out.dtype], (ldc if ldc is not None else out.stride(-2))
    if out2 is not None:
        p.C2, p.ldc2 = ptr(out2), (ldc2 if ldc2 is not None else out2.stride(-2))
    p.out_gin, p.out_gout, p.out_off, p.block_n = out_gin, out_gout, out_off, block_n
    if rope is not None:
        pos, p.rope_q_col, p.rope_k_col, p.rope_heads, p.rope_base, p.rope_cam_theta = rope
        _need_cuda(pos)
        assert pos.dtype == torch.int32 and pos.is_contiguous()
        p.rope_pos = ptr(pos)
    with _timed("gemm", ("lin", total, N, K, act, res1 is not None, out.dtype == torch.float32)):
        check(lib.vs_gemm(C.byref(p), C.c_void_p(stream_ptr())), "vs_gemm")
    return out


def conv_gemm(x_nhwc, Wp, *, kh, kw, pad, N, bias=None, act=VS_ACT_NONE, res1=None, res2=None,
              out=None, out_dtype=torch.bfloat16, out2=None, block_n=0, res_up2=False, view=None,
              out_gin=0, out_gout=0, out_off=0):
    """Stride-1 kh x kw convolution on an NHWC bf16 map as an implicit GEMM (conv mode of vs_gemm).
    Wp: packed weights (N, kh*kw*cin_pad) bf16, tap-major / channel-minor.
    view = (n, out_h, out_w, cin, in_h, stride_x, stride_y, stride_n): explicit (possibly
    overlapping) input view instead of the dense NHWC shape of x_nhwc.
    res_up2: res1 is a half-resolution NHWC map, bilinearly upsampled x2 in the epilogue."""
    _need_cuda(x_nhwc, Wp)
    lib = _lib.load()
    p = GemmParams()
    if view is None:
        n, h, w, cin = x_nhwc.shape
    else:
        n, h, w, cin, p.conv_in_h, p.conv_stride_x, p.conv_stride_y, p.conv_stride_n = view
    p.A, p.a_mode = ptr(x_nhwc), 1
    p.cn, p.ch, p.cw, p.cin, p.kh, p.kw, p.pad = n, h, w, cin, kh, kw, pad
    p.W, p.w_row_stride, p.N = ptr(Wp), Wp.stride(0), N
    _fill_epilogue(p, bias, act, None, 0, 0, res1, res2)
    assert x_nhwc.dtype in _HALF and Wp.dtype == x_nhwc.dtype, "conv operands are bf16 (speed) or fp16 (parity mode)"
    p.operand_dtype = _DT[x_nhwc.dtype]
    if out_dtype in _HALF:
        out_dtype = x_nhwc.dtype
    p.res_up2 = int(res_up2)
    if out is None:
        out = torch.empty((n, h, w, N), dtype=out_dtype, device=x_nhwc.device)
    p.C, p.c_dtype, p.ldc = ptr(out), _DT[out.dtype], out.stride(-2)
    if out2 is not None:
        p.C2, p.ldc2 = ptr(out2), out2.stride(-2)
    p.block_n = block_n
    p.out_gin, p.out_gout, p.out_off = out_gin, out_gout, out_off   # pixel rows -> grouped output rows
    with _timed("gemm", (f"conv{kh}x{kw}", n * h * w, N, kh * kw * ((cin + 63) // 64 * 64), act,
                         res1 is not None, out.dtype == torch.float32)):
        check(lib.vs_gemm(C.byref(p), C.c_void_p(stream_ptr())), "vs_gemm(conv)")
    return out


_EXTRA = None   # mask_mode / out_scale of the call being assembled (set by _masked)


def _fill_epilogue(p, bias, act, gate, gate_rows, first_row_mode, res1, res2):
    p.bias, p.act = ptr(bias), act
    if _EXTRA is not None:
        p.mask_mode, p.out_scale = _EXTRA["mask_mode"], _EXTRA["out_scale"]
    if gate is not None:
        p.gate, p.gate_ld = ptr(gate), gate.stride(-2)
    p.gate_rows, p.first_row_mode = gate_rows, first_row_mode
    if res1 is not None:
        p.res1, p.res_dtype, p.res_ld = ptr(res1), _DT[res1.dtype], res1.stride(-2)
        if res2 is not None:
            assert res2.dtype == res1.dtype and res2.stride(-2) == res1.stride(-2)
            p.res2 = ptr(res2)


def layernorm(x, w=None, b=None, *, eps=1e-6, w0=None, b0=None, scale=None, shift=None,
              rows_per_frame=0, normalize=True, out_bf16=None, out_f32=None, want_bf16=True,
              want_f32=False, half=torch.bfloat16):
    """Row LayerNorm (+ per-frame AdaLN modulate) on fp32 rows; see vs_layernorm in the header."""
    _need_cuda(x)
    lib = _lib.load()
    rows, Cc = x.shape
    p = LayerNormParams()
    p.x, p.ldx, p.rows, p.C = ptr(x), x.stride(0), rows, Cc
    p.w, p.b, p.w0, p.b0 = ptr(w), ptr(b), ptr(w0), ptr(b0)
    if scale is not None:
        p.scale, p.shift, p.mod_ld = ptr(scale), ptr(shift), scale.stride(-2)
    p.rows_per_frame, p.eps, p.normalize = rows_per_frame, eps, int(normalize)
    if out_bf16 is None and want_bf16:
        out_bf16 = torch.empty((rows, Cc), dtype=half, device=x.device)
    if out_f32 is None and want_f32:
        out_f32 = torch.empty((rows, Cc), dtype=torch.float32, device=x.device)
    if out_bf16 is not None:
        p.y_bf16, p.ldy_bf16, p.y16_dtype = ptr(out_bf16), out_bf16.stride(0), _DT[out_bf16.dtype]
    if out_f32 is not None:
        p.y_f32, p.ldy_f32 = ptr(out_f32), out_f32.stride(0)
    with _timed("layernorm"):
        check(lib.vs_layernorm(C.byref(p), C.c_void_p(stream_ptr())), "vs_layernorm")
    return out_bf16, out_f32


def attention(Q, K, V, O, *, heads, q_start, q_len, kv_start0, kv_len0, kv_start1=None,
              kv_len1=None, max_q_len, max_kv_len=0, causal_block=0, scale=0.125, lse=None):
    """softmax(Q K^T * scale) V per (item, head); Q/K/V/O are 2-D bf16 views (rows, >= heads*64).
    lse (rows, heads) f32, optional: log2-domain log-sum-exp per query row (kept for the backward pass)."""
    _need_cuda(Q, K, V, O)
    lib = _lib.load()
    p = AttentionParams()
    p.Q, p.K, p.V, p.O = ptr(Q), ptr(K), ptr(V), ptr(O)
    p.ldq, p.ldk, p.ldv, p.ldo = Q.stride(0), K.stride(0), V.stride(0), O.stride(0)
    p.q_rows, p.kv_rows = Q.shape[0], K.shape[0]
    p.heads, p.items = heads, q_start.numel()
    p.q_start, p.q_len = ptr(q_start), ptr(q_len)
    p.kv_start0, p.kv_len0 = ptr(kv_start0), ptr(kv_len0)
    p.kv_start1, p.kv_len1 = ptr(kv_start1), ptr(kv_len1)
    p.max_q_len, p.max_kv_len, p.causal_block, p.scale = max_q_len, max_kv_len, causal_block, scale
    assert Q.dtype in _HALF and K.dtype == V.dtype == O.dtype == Q.dtype
    p.dtype = _DT[Q.dtype]
    if lse is not None:
        assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.shape == (Q.shape[0], heads)
        p.lse = ptr(lse)
    with _timed("attention", ("attn", p.items, heads, max_q_len, max_kv_len, causal_block)):
        check(lib.vs_attention(C.byref(p), C.c_void_p(stream_ptr())), "vs_attention")
    return O


def rope_rows(qkv, pos_i32, *, heads, q_col, k_col, base=100.0, cam_theta=30.0):
    lib = _lib.load()
    _need_cuda(qkv, pos_i32)
    with _timed("rope_rows"):
        check(lib.vs_rope_rows(C.c_void_p(ptr(qkv)), C.c_int64(qkv.stride(0)), qkv.shape[0], heads,
                               q_col, k_col, C.c_void_p(ptr(pos_i32)), C.c_float(base),
                               C.c_float(cam_theta), C.c_void_p(stream_ptr())), "vs_rope_rows")
    return qkv


def patchify(img, P=16, half=torch.bfloat16):
    lib = _lib.load()
    _need_cuda(img)
    n, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous()
    out = torch.empty((n * (h // P) * (w // P), 3 * P * P), dtype=half, device=img.device)
    check(lib.vs_patchify(C.c_void_p(ptr(img)), C.c_void_p(ptr(out)), n, h, w, P, _DT[half],
                          C.c_void_p(stream_ptr())), "vs_patchify")
    return out


def im2col(src, *, nchw_f32, n, h, w, c, k, stride, pad, kpad, half=None):
    lib = _lib.load()
    _need_cuda(src)
    half = half or (src.dtype if src.dtype in _HALF else torch.bfloat16)
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = torch.empty((n * ho * wo, kpad), dtype=half, device=src.device)
    check(lib.vs_im2col(C.c_void_p(ptr(src)), int(nchw_f32), C.c_void_p(ptr(out)), n, h, w, c, k,
                        stride, pad, kpad, _DT[half], C.c_void_p(stream_ptr())), "vs_im2col")
    return out


def image_nhwc8(img, pad=3, half=torch.bfloat16):
    """fp32 NCHW (n,3,h,w) -> zero-bordered 16-bit (n, h+2*pad, w+8, 8) for the 7x7 stem."""
    lib = _lib.load()
    _need_cuda(img)
    n, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous()
    out = torch.empty((n, h + 2 * pad, w + 8, 8), dtype=half, device=img.device)
    check(lib.vs_image_nhwc8(C.c_void_p(ptr(img)), C.c_void_p(ptr(out)), n, h, w, pad, _DT[half],
                             C.c_void_p(stream_ptr())), "vs_image_nhwc8")
    return out


def upsample2x(x_nhwc, add=None):
    """bilinear x2 (align_corners=True) of an NHWC bf16 map (+ `add`, a full-resolution map)."""
    lib = _lib.load()
    _need_cuda(x_nhwc, add)
    n, h, w, c = x_nhwc.shape
    assert x_nhwc.dtype in _HALF
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=x_nhwc.dtype, device=x_nhwc.device)
    with _timed("upsample2x"):
        if add is None:
            check(lib.vs_upsample2x(C.c_void_p(ptr(x_nhwc)), C.c_void_p(ptr(out)), n, h, w, c, _DT[x_nhwc.dtype],
                                    C.c_void_p(stream_ptr())), "vs_upsample2x")
        else:
            assert add.shape == out.shape and add.dtype == x_nhwc.dtype and add.is_contiguous()
            check(lib.vs_upsample2x_add(C.c_void_p(ptr(x_nhwc)), C.c_void_p(ptr(add)), C.c_void_p(ptr(out)),
                                        n, h, w, c, _DT[x_nhwc.dtype], C.c_void_p(stream_ptr())),
                  "vs_upsample2x_add")
    return out


def pixel_shuffle(src, n, h, w, c, k):
    lib = _lib.load()
    _need_cuda(src)
    out = torch.empty((n, h * k, w * k, c), dtype=src.dtype, device=src.device)
    check(lib.vs_pixel_shuffle(C.c_void_p(ptr(src)), C.c_void_p(ptr(out)), n, h, w, c, k,
                               C.c_void_p(stream_ptr())), "vs_pixel_shuffle")
    return out


def intrinsic_token(K9, w, b, x, frames, E, rows_per_frame, row_off):
    lib = _lib.load()
    check(lib.vs_intrinsic_token(C.c_void_p(ptr(K9)), C.c_void_p(ptr(w)), C.c_void_p(ptr(b)),
                                 C.c_void_p(ptr(x)), frames, E, rows_per_frame, row_off,
                                 C.c_void_p(stream_ptr())), "vs_intrinsic_token")


def camera_tokens(intr_tok, extr_tok, x, frames, T, Cdim, rows_per_frame):
    lib = _lib.load()
    check(lib.vs_camera_tokens(C.c_void_p(ptr(intr_tok)), C.c_void_p(ptr(extr_tok)),
                               C.c_void_p(ptr(x)), frames, T, Cdim, rows_per_frame,
                               C.c_void_p(stream_ptr())), "vs_camera_tokens")


def silu_bf16(x, rows, Cdim, ldx=None, half=torch.bfloat16):
    lib = _lib.load()
    y = torch.empty((rows, Cdim), dtype=half, device=x.device)
    check(lib.vs_silu_bf16(C.c_void_p(ptr(x)), C.c_int64(ldx if ldx is not None else x.stride(0)),
                           C.c_void_p(ptr(y)), C.c_int64(Cdim), rows, Cdim, _DT[half],
                           C.c_void_p(stream_ptr())), "vs_silu_bf16")
    return y


def camera_head(cam_feat, ld, w, b, B, T, Cdim):
    lib = _lib.load()
    pred = torch.empty((B, T - 1, 8), dtype=torch.float32, device=cam_feat.device)
    c2w = torch.empty((B, T, 4, 4), dtype=torch.float32, device=cam_feat.device)
    check(lib.vs_camera_head(C.c_void_p(ptr(cam_feat)), C.c_int64(ld), C.c_void_p(ptr(w)),
                             C.c_void_p(ptr(b)), B, T, Cdim, C.c_void_p(ptr(pred)),
                             C.c_void_p(ptr(c2w)), C.c_void_p(stream_ptr())), "vs_camera_head")
    return pred, c2w


def pts_tail(feat, Cf, w, b, raw, px):
    lib = _lib.load()
    check(lib.vs_pts_tail(C.c_void_p(ptr(feat)), Cf, C.c_void_p(ptr(w)), C.c_void_p(ptr(b)),
                          C.c_void_p(ptr(raw)), C.c_int64(raw.stride(-2)), C.c_int64(px), _DT[feat.dtype],
                          C.c_void_p(stream_ptr())), "vs_pts_tail")


def gaussian_adapter(src, d_sh, sh_mask, *, center_col=0, param_col=3, want_cov=True, raw_out=None):
    """src (G, ld) fp32 head outputs -> dict of means/cov/cov6/sh/opac/scales/rot
    (gaussian_adapter.py:167-212); optionally also writes the reference-layout raw (G, 86)."""
    lib = _lib.load()
    _need_cuda(src)
    G = src.shape[0]
    dev, f32 = src.device, torch.float32
    out = dict(
        means=torch.empty((G, 3), dtype=f32, device=dev),
        cov=torch.empty((G, 3, 3), dtype=f32, device=dev) if want_cov else None,
        cov6=torch.empty((G, 6), dtype=f32, device=dev),
        sh=torch.empty((G, 3, d_sh), dtype=f32, device=dev),
        opac=torch.empty((G,), dtype=f32, device=dev),
        scales=torch.empty((G, 3), dtype=f32, device=dev),
        rot=torch.empty((G, 4), dtype=f32, device=dev),
    )
    check(lib.vs_gaussian_adapter(
        C.c_void_p(ptr(src)), C.c_int64(src.stride(0)), center_col, param_col, C.c_int64(G), d_sh,
        C.c_void_p(ptr(sh_mask)), C.c_void_p(ptr(raw_out)), C.c_void_p(ptr(out["means"])),
        C.c_void_p(ptr(out["cov"])), C.c_void_p(ptr(out["cov6"])), C.c_void_p(ptr(out["sh"])),
        C.c_void_p(ptr(out["opac"])), C.c_void_p(ptr(out["scales"])), C.c_void_p(ptr(out["rot"])),
        C.c_void_p(stream_ptr())), "vs_gaussian_adapter")
    return out


_mse_ws: dict = {}


def mse_loss(pred, target, weight=1.0, want_grad=True):
    """weight * mean((pred - target)^2) and (optionally) its gradient w.r.t. pred in one pass
    (vs_mse_loss; LossMse.forward, src/loss/loss_mse.py:23-31).  Returns (loss 0-d, grad | None)."""
    lib = _lib.load()
    _need_cuda(pred, target)
    assert pred.shape == target.shape and pred.dtype == target.dtype == torch.float32
    pred, target = pred.contiguous(), target.contiguous()
    key = (pred.device, torch.cuda.current_stream().cuda_stream)
    if key not in _mse_ws:   # zeroed once: the kernel leaves its completion counter at zero
        lib.vs_mse_workspace_bytes.restype = C.c_int64
        _mse_ws[key] = torch.zeros((lib.vs_mse_workspace_bytes(),), dtype=torch.uint8, device=pred.device)
    loss = torch.empty((), dtype=torch.float32, device=pred.device)
    grad = torch.empty_like(pred) if want_grad else None
    with _timed("mse"):
        check(lib.vs_mse_loss(C.c_void_p(ptr(pred)), C.c_void_p(ptr(target)), C.c_int64(pred.numel()),
                              C.c_float(weight), C.c_void_p(ptr(loss)), C.c_void_p(ptr(grad)),
                              C.c_void_p(ptr(_mse_ws[key])), C.c_void_p(stream_ptr())), "vs_mse_loss")
    return loss, grad


# ------------------------------------------------------------------ encoder training path (backward)
def grad_prep(src, *, z=None, want_copy=True, want_t=True, colsum=None):
    """One pass over a gradient matrix (rows, cols) f32 / bf16 (vs_grad_prep): optional gelu'(z)
    factor, bf16 row-major copy, bf16 transposed copy (returned as a (cols, rows) view of a buffer
    whose row stride is rows rounded up to 8) and column sums accumulated into `colsum` (f32)."""
    lib = _lib.load()
    _need_cuda(src, z, colsum)
    assert src.dim() == 2 and src.stride(1) == 1 and src.dtype in (torch.float32, torch.bfloat16)
    rows, cols = src.shape
    copy = torch.empty((rows, cols), dtype=torch.bfloat16, device=src.device) if want_copy else None
    tr = None
    ld_t = (rows + 7) // 8 * 8
    if want_t:
        tr = torch.empty((cols, ld_t), dtype=torch.bfloat16, device=src.device)
    if z is not None:
        assert z.shape == src.shape and z.dtype == torch.bfloat16 and z.stride(1) == 1
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.numel() == cols and colsum.is_contiguous()
    with _timed("grad_prep"):
        check(lib.vs_grad_prep(C.c_void_p(ptr(src)), _DT[src.dtype], C.c_int64(src.stride(0)),
                               C.c_void_p(ptr(z)), C.c_int64(z.stride(0) if z is not None else 0),
                               rows, cols, C.c_void_p(ptr(copy)), C.c_int64(cols),
                               C.c_void_p(ptr(tr)), C.c_int64(ld_t), C.c_void_p(ptr(colsum)),
                               C.c_void_p(stream_ptr())), "vs_grad_prep")
    return copy, (tr[:, :rows] if tr is not None else None)


def gelu_bf16(z, out=None):
    """a = gelu(z) (exact erf), bf16 -> bf16 (vs_gelu_bf16): the training-mode activation."""
    lib = _lib.load()
    _need_cuda(z)
    assert z.dim() == 2 and z.dtype == torch.bfloat16 and z.stride(1) == 1
    if out is None:
        out = torch.empty_like(z)
    check(lib.vs_gelu_bf16(C.c_void_p(ptr(z)), C.c_int64(z.stride(0)), C.c_void_p(ptr(out)),
                           C.c_int64(out.stride(0)), z.shape[0], z.shape[1],
                           C.c_void_p(stream_ptr())), "vs_gelu_bf16")
    return out


def layernorm_backward(x, dy, gamma, *, dres=None, dx=None, dgamma=None, dbeta=None, eps=1e-6):
    """dx = dres + d LayerNorm / dx; dgamma / dbeta accumulated (vs_layernorm_backward)."""
    lib = _lib.load()
    _need_cuda(x, dy, gamma)
    rows, Cc = x.shape
    assert dy.shape == x.shape and x.dtype == torch.float32
    if dx is None:
        dx = torch.empty((rows, Cc), dtype=torch.float32, device=x.device)
    p = _lib.LayerNormBwdParams()
    p.x, p.ldx = ptr(x), x.stride(0)
    p.dy, p.dy_dtype, p.ldy = ptr(dy), _DT[dy.dtype], dy.stride(0)
    p.gamma = ptr(gamma)
    if dres is not None:
        p.dres, p.ldres = ptr(dres), dres.stride(0)
    p.dx, p.lddx = ptr(dx), dx.stride(0)
    p.dgamma, p.dbeta = ptr(dgamma), ptr(dbeta)
    p.rows, p.C, p.eps = rows, Cc, eps
    with _timed("layernorm_bwd"):
        check(lib.vs_layernorm_backward(C.byref(p), C.c_void_p(stream_ptr())), "vs_layernorm_backward")
    return dx


def attention_backward(Q, K, V, O, dO, lse, dQ, dK, dV, *, heads, q_start, q_len, kv_start0, kv_len0,
                       kv_start1=None, kv_len1=None, max_q_len, max_kv_len, causal_block=0,
                       scale=0.125, dkv_tables=None):
    """dQ, dK, dV of ops.attention from dO, the forward output O and its lse (vs_attention_backward).
    dkv_tables: key-centric items of the dK / dV pass (dict of int32 device tensors kv_start, kv_len,
    q_start0, q_len0, q_start1, q_len1 + int max_kv_len) -- REQUIRED when forward items share key rows
    (neighbour attention); without them overlapping key rows would be overwritten, not accumulated."""
    _need_cuda(Q, K, V, O, dO, lse, dQ, dK, dV)
    if kv_start1 is not None and dkv_tables is None and q_start.numel() > 1:
        # dK / dV are written, not accumulated: items must not share key rows.  Single-segment tables of
        # the encoder are disjoint by construction; for ad-hoc two-segment tables it is checked here (a
        # host read of the small tables -- the training path passes dkv_tables and never comes here)
        segs = torch.stack([torch.stack([kv_start0, kv_len0]), torch.stack([kv_start1, kv_len1])]).cpu()
        spans = sorted((int(a), int(a) + int(n)) for a, n in segs.permute(0, 2, 1).reshape(-1, 2).tolist() if n > 0)
        if any(b0 < a1 for (_, a1), (b0, _) in zip(spans, spans[1:])):
            raise ValueError("attention_backward: items share key rows (dK / dV would be overwritten, not "
                             "accumulated); pass the key-centric dkv_tables "
                             "(encoder_grad.neighbour_backward_tables)")
    lib = _lib.load()
    p = _lib.AttentionBwdParams()
    f = p.fwd
    f.Q, f.K, f.V, f.O = ptr(Q), ptr(K), ptr(V), ptr(O)
    f.ldq, f.ldk, f.ldv, f.ldo = Q.stride(0), K.stride(0), V.stride(0), O.stride(0)
    f.q_rows, f.kv_rows = Q.shape[0], K.shape[0]
    f.heads, f.items = heads, q_start.numel()
    f.q_start, f.q_len = ptr(q_start), ptr(q_len)
    f.kv_start0, f.kv_len0 = ptr(kv_start0), ptr(kv_len0)
    f.kv_start1, f.kv_len1 = ptr(kv_start1), ptr(kv_len1)
    f.max_q_len, f.max_kv_len, f.causal_block, f.scale = max_q_len, max_kv_len, causal_block, scale
    assert lse.dtype == torch.float32 and lse.is_contiguous() and lse.shape == (Q.shape[0], heads)
    f.lse = ptr(lse)
    p.dO, p.lddo = ptr(dO), dO.stride(0)
    p.dQ, p.dK, p.dV = ptr(dQ), ptr(dK), ptr(dV)
    p.lddq, p.lddk, p.lddv = dQ.stride(0), dK.stride(0), dV.stride(0)
    delta = torch.empty((Q.shape[0], heads), dtype=torch.float32, device=Q.device)
    p.delta = ptr(delta)
    if dkv_tables is not None:
        t = dkv_tables
        p.dkv_items, p.dkv_max_kv_len = t["kv_start"].numel(), int(t["max_kv_len"])
        p.dkv_kv_start, p.dkv_kv_len = ptr(t["kv_start"]), ptr(t["kv_len"])
        p.dkv_q_start0, p.dkv_q_len0 = ptr(t["q_start0"]), ptr(t["q_len0"])
        p.dkv_q_start1, p.dkv_q_len1 = ptr(t["q_start1"]), ptr(t["q_len1"])
    with _timed("attention_bwd", ("attn_bwd", f.items, heads, max_q_len, max_kv_len, causal_block)):
        check(lib.vs_attention_backward(C.byref(p), C.c_void_p(stream_ptr())), "vs_attention_backward")
    return dQ, dK, dV


def rope_rows_backward(dqkv, pos_i32, *, heads, q_col, k_col, base=100.0, cam_theta=30.0):
    """Inverse rotation on the q / k columns of the packed bf16 gradient (in place)."""
    lib = _lib.load()
    _need_cuda(dqkv, pos_i32)
    check(lib.vs_rope_rows_backward(C.c_void_p(ptr(dqkv)), C.c_int64(dqkv.stride(0)), dqkv.shape[0],
                                    heads, q_col, k_col, C.c_void_p(ptr(pos_i32)), C.c_float(base),
                                    C.c_float(cam_theta), C.c_void_p(stream_ptr())),
          "vs_rope_rows_backward")
    return dqkv


# ------------------------------------------------------------------ decoder / head training path (backward)
def conv_wgrad(dy_nhwc, x_nhwc, dW, *, kh, kw, pad, view=None, split_k=0):
    """dW (Cout, kh*kw*cin_pad) fp32 += sum over pixels of dY[pixel, :] (x) X[pixel + tap, :]: the weight
    gradient of ops.conv_gemm in its packed layout, without an im2col buffer (a_mode 3 of vs_gemm;
    both operands are read as stored, pixels are the contraction dimension, split-K + atomic adds).
    `view` describes the conv's INPUT exactly as in conv_gemm."""
    _need_cuda(dy_nhwc, x_nhwc, dW)
    lib = _lib.load()
    p = GemmParams()
    if view is None:
        n, h, w, cin = x_nhwc.shape
    else:
        n, h, w, cin, p.conv_in_h, p.conv_stride_x, p.conv_stride_y, p.conv_stride_n = view
    cout = dy_nhwc.shape[-1]
    assert dy_nhwc.dtype == torch.bfloat16 and x_nhwc.dtype == torch.bfloat16 and dW.dtype == torch.float32
    assert dy_nhwc.numel() == n * h * w * cout and dy_nhwc.is_contiguous()
    cin_pad = (cin + 63) // 64 * 64
    assert dW.shape == (cout, kh * kw * cin_pad) and dW.is_contiguous()
    p.A, p.a_mode, p.a_rows, p.a_groups, p.a_row_stride = ptr(dy_nhwc), 3, cout, 1, cout
    p.cn, p.ch, p.cw, p.cin, p.kh, p.kw, p.pad = n, h, w, cin, kh, kw, pad
    p.W, p.w_row_stride, p.N = ptr(x_nhwc), 8, kh * kw * cin_pad
    p.C, p.c_dtype, p.ldc = ptr(dW), VS_F32, dW.stride(0)
    p.c_accumulate, p.split_k = 1, split_k
    with _timed("gemm", (f"wgrad{kh}x{kw}", cout, kh * kw * cin_pad, n * h * w, 0, False, True)):
        check(lib.vs_gemm(C.byref(p), C.c_void_p(stream_ptr())), "vs_gemm(conv wgrad)")
    return dW


def gemm_tn_acc(dy, x, dW, split_k=0):
    """dW (N, K) fp32 += dy^T x with dy (M, N), x (M, K) bf16 as stored (a_mode 2, split-K, atomic adds)."""
    _need_cuda(dy, x, dW)
    lib = _lib.load()
    assert dy.dim() == 2 and x.dim() == 2 and dy.shape[0] == x.shape[0]
    assert dW.dtype == torch.float32 and dW.shape == (dy.shape[1], x.shape[1])
    p = GemmParams()
    p.A, p.a_mode, p.a_rows, p.a_groups, p.a_row_stride = ptr(dy), 2, dy.shape[1], 1, dy.stride(0)
    p.W, p.w_row_stride, p.N, p.K = ptr(x), x.stride(0), x.shape[1], dy.shape[0]
    p.C, p.c_dtype, p.ldc = ptr(dW), VS_F32, dW.stride(0)
    p.c_accumulate, p.split_k = 1, split_k
    with _timed("gemm", ("wgrad", dy.shape[1], x.shape[1], dy.shape[0], 0, False, True)):
        check(lib.vs_gemm(C.byref(p), C.c_void_p(stream_ptr())), "vs_gemm(wgrad)")
    return dW


def gemm_masked(A, W, *, mask, res1=None, out_dtype=torch.bfloat16, out=None, out_scale=0.0, **kw):
    """ops.gemm with the ReLU-mask epilogue: C = (mask > 0 ? A W^T * out_scale : 0) (+ res1)."""
    return _masked(gemm, A, W, mask, res1, out_dtype, out, out_scale, kw)


def conv_gemm_masked(x, Wp, *, mask, res1=None, out_dtype=torch.bfloat16, out=None, out_scale=0.0, **kw):
    return _masked(conv_gemm, x, Wp, mask, res1, out_dtype, out, out_scale, kw)


def _masked(fn, A, W, mask, res1, out_dtype, out, out_scale, kw):
    assert mask.dtype == torch.bfloat16
    global _EXTRA
    if res1 is None:
        _EXTRA = dict(mask_mode=2, out_scale=out_scale)
        try:
            return fn(A, W, res1=mask, out=out, out_dtype=out_dtype, **kw)
        finally:
            _EXTRA = None
    _EXTRA = dict(mask_mode=1, out_scale=out_scale)
    try:
        return fn(A, W, res1=res1, res2=mask, out=out, out_dtype=out_dtype, **kw)
    finally:
        _EXTRA = None


def layernorm_mod_backward(x, dh, gamma, *, frame_a, frame_b, frames, rows_per_frame, skip_first=False,
                           scale=None, dres=None, dx=None, eps=1e-6):
    """vs_layernorm_mod_backward (see the header): dx and the per-frame sums A_f, B_f (accumulated)."""
    lib = _lib.load()
    _need_cuda(x, dh, gamma, frame_a, frame_b)
    rows, Cc = x.shape
    assert rows == frames * rows_per_frame and dh.shape == x.shape and x.dtype == torch.float32
    assert frame_a.shape == (frames, Cc) and frame_b.shape == (frames, Cc) and frame_a.stride(0) == frame_b.stride(0)
    if dx is None:
        dx = torch.empty((rows, Cc), dtype=torch.float32, device=x.device)
    p = _lib.LnModBwdParams()
    p.x, p.ldx = ptr(x), x.stride(0)
    p.dh, p.dh_dtype, p.lddh = ptr(dh), _DT[dh.dtype], dh.stride(0)
    p.gamma = ptr(gamma)
    if scale is not None:
        p.scale, p.mod_ld = ptr(scale), scale.stride(0)
    if dres is not None:
        p.dres, p.ldres = ptr(dres), dres.stride(0)
    p.dx, p.lddx = ptr(dx), dx.stride(0)
    p.frame_a, p.frame_b, p.frame_ld = ptr(frame_a), ptr(frame_b), frame_a.stride(0)
    p.frames, p.rows_per_frame, p.skip_first, p.C, p.eps = frames, rows_per_frame, int(skip_first), Cc, eps
    with _timed("layernorm_bwd"):
        check(lib.vs_layernorm_mod_backward(C.byref(p), C.c_void_p(stream_ptr())), "vs_layernorm_mod_backward")
    return dx


def adaln_reduce(frame_a, frame_b, gamma, beta, *, scale=None, dscale=None, dshift=None, dgamma=None, dbeta=None):
    lib = _lib.load()
    frames, Cc = frame_a.shape
    check(lib.vs_adaln_reduce(
        C.c_void_p(ptr(frame_a)), C.c_void_p(ptr(frame_b)), C.c_int64(frame_a.stride(0)), C.c_void_p(ptr(gamma)),
        C.c_void_p(ptr(beta)), C.c_void_p(ptr(scale)), C.c_int64(scale.stride(0) if scale is not None else 0),
        C.c_void_p(ptr(dscale)), C.c_void_p(ptr(dshift)), C.c_int64(dscale.stride(0) if dscale is not None else 0),
        C.c_void_p(ptr(dgamma)), C.c_void_p(ptr(dbeta)), frames, Cc, C.c_void_p(stream_ptr())), "vs_adaln_reduce")


def gate_residual(x, branch, *, gate=None, rows_per_frame=0, first_row_mode=0, out=None):
    """out = x + (1 + gate_f) * branch (out=None: in place): training forward of the gated residual."""
    lib = _lib.load()
    _need_cuda(x, branch)
    assert x.dtype == torch.float32 and branch.dtype == torch.bfloat16 and x.shape == branch.shape
    out = x if out is None else out
    with _timed("gate"):
        check(lib.vs_gate_residual(
            C.c_void_p(ptr(x)), C.c_int64(x.stride(0)), C.c_void_p(ptr(out)), C.c_int64(out.stride(0)),
            C.c_void_p(ptr(branch)), C.c_int64(branch.stride(0)),
            C.c_void_p(ptr(gate)), C.c_int64(gate.stride(0) if gate is not None else 0), C.c_int64(x.shape[0]),
            x.shape[1], rows_per_frame, first_row_mode, C.c_void_p(stream_ptr())), "vs_gate_residual")
    return out


def gate_backward(dout, *, frames, rows_per_frame, branch=None, gate=None, dgate=None, colsum=None,
                  first_row_mode=0):
    """-> dbranch bf16 (rows, C); dgate (frames, C) and colsum (C) accumulated."""
    lib = _lib.load()
    _need_cuda(dout)
    rows, Cc = dout.shape
    assert rows == frames * rows_per_frame and dout.dtype == torch.float32
    dbr = torch.empty((rows, Cc), dtype=torch.bfloat16, device=dout.device)
    with _timed("gate"):
        check(lib.vs_gate_backward(
            C.c_void_p(ptr(dout)), C.c_int64(dout.stride(0)), C.c_void_p(ptr(branch)),
            C.c_int64(branch.stride(0) if branch is not None else 0), C.c_void_p(ptr(gate)),
            C.c_int64(gate.stride(0) if gate is not None else 0), C.c_void_p(ptr(dbr)), C.c_int64(Cc),
            C.c_void_p(ptr(dgate)), C.c_int64(dgate.stride(0) if dgate is not None else 0),
            C.c_void_p(ptr(colsum)), frames, rows_per_frame, Cc, first_row_mode, C.c_void_p(stream_ptr())),
            "vs_gate_backward")
    return dbr


def silu_backward(x, dy, dx, accumulate=True):
    lib = _lib.load()
    rows, Cc = x.shape
    assert x.dtype == dy.dtype == dx.dtype == torch.float32
    check(lib.vs_silu_backward(C.c_void_p(ptr(x)), C.c_int64(x.stride(0)), C.c_void_p(ptr(dy)),
                               C.c_int64(dy.stride(0)), C.c_void_p(ptr(dx)), C.c_int64(dx.stride(0)), rows, Cc,
                               int(accumulate), C.c_void_p(stream_ptr())), "vs_silu_backward")
    return dx


def upsample2x_backward(dy_nhwc):
    lib = _lib.load()
    _need_cuda(dy_nhwc)
    n, h2, w2, c = dy_nhwc.shape
    assert dy_nhwc.is_contiguous() and dy_nhwc.dtype == torch.bfloat16
    out = torch.empty((n, h2 // 2, w2 // 2, c), dtype=torch.bfloat16, device=dy_nhwc.device)
    with _timed("upsample2x"):
        check(lib.vs_upsample2x_backward(C.c_void_p(ptr(dy_nhwc)), C.c_void_p(ptr(out)), n, h2 // 2, w2 // 2, c,
                                         C.c_void_p(stream_ptr())), "vs_upsample2x_backward")
    return out


def pixel_unshuffle(src_nhwc, k):
    lib = _lib.load()
    n, hk, wk, c = src_nhwc.shape
    h, w = hk // k, wk // k
    out = torch.empty((n * h * w, k * k * c), dtype=torch.bfloat16, device=src_nhwc.device)
    check(lib.vs_pixel_unshuffle(C.c_void_p(ptr(src_nhwc)), C.c_void_p(ptr(out)), n, h, w, c, k,
                                 C.c_void_p(stream_ptr())), "vs_pixel_unshuffle")
    return out


def col2im(dcols, *, n, h, w, c, k, stride, pad):
    lib = _lib.load()
    out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=dcols.device)
    check(lib.vs_col2im(C.c_void_p(ptr(dcols)), C.c_void_p(ptr(out)), n, h, w, c, k, stride, pad,
                        dcols.shape[1], C.c_void_p(stream_ptr())), "vs_col2im")
    return out


def relu_backward(dy, y, *, out=None, colsum=None, want_dx=True):
    """dx = dy * (y > 0) on bf16 (rows, C) matrices; colsum (C) f32 accumulated."""
    lib = _lib.load()
    _need_cuda(dy, y)
    rows, Cc = dy.shape
    if want_dx and out is None:
        out = torch.empty_like(dy)
    check(lib.vs_relu_backward(C.c_void_p(ptr(dy)), C.c_int64(dy.stride(0)), C.c_void_p(ptr(y)),
                               C.c_int64(y.stride(0)), C.c_void_p(ptr(out)),
                               C.c_int64(out.stride(0) if out is not None else 0), C.c_void_p(ptr(colsum)),
                               C.c_int64(rows), Cc, C.c_void_p(stream_ptr())), "vs_relu_backward")
    return out


def pts_tail_backward(feat, Cf, w, b, d_xyz, dw, db):
    """-> d_feat bf16 (px, Cf); dw (3, Cf) / db (3) accumulated.  d_xyz: fp32 rows, xyz at columns 0..2."""
    lib = _lib.load()
    px = d_xyz.shape[0]
    d_feat = torch.empty((px, Cf), dtype=torch.bfloat16, device=d_xyz.device)
    check(lib.vs_pts_tail_backward(C.c_void_p(ptr(feat)), Cf, C.c_void_p(ptr(w)), C.c_void_p(ptr(b)),
                                   C.c_void_p(ptr(d_xyz)), C.c_int64(d_xyz.stride(0)), C.c_void_p(ptr(d_feat)),
                                   C.c_void_p(ptr(dw)), C.c_void_p(ptr(db)), C.c_int64(px),
                                   C.c_void_p(stream_ptr())), "vs_pts_tail_backward")
    return d_feat


def gaussian_adapter_backward(src, d_sh, sh_mask, d_src, *, center_col=0, param_col=3, d_raw=None, d_means=None,
                              d_cov=None, d_cov6=None, d_shs=None, d_opac=None):
    lib = _lib.load()
    _need_cuda(src, d_src)
    G = src.shape[0]
    for t in (d_raw, d_means, d_cov, d_cov6, d_shs, d_opac):
        assert t is None or (t.is_contiguous() and t.dtype == torch.float32)
    check(lib.vs_gaussian_adapter_backward(
        C.c_void_p(ptr(src)), C.c_int64(src.stride(0)), center_col, param_col, C.c_int64(G), d_sh,
        C.c_void_p(ptr(sh_mask)), C.c_void_p(ptr(d_raw)), C.c_void_p(ptr(d_means)), C.c_void_p(ptr(d_cov)),
        C.c_void_p(ptr(d_cov6)), C.c_void_p(ptr(d_shs)), C.c_void_p(ptr(d_opac)), C.c_void_p(ptr(d_src)),
        C.c_int64(d_src.stride(0)), C.c_void_p(stream_ptr())), "vs_gaussian_adapter_backward")
    return d_src


def camera_head_backward(cam_feat, w, b, B, T, Cdim, d_pred, dw, db):
    lib = _lib.load()
    d_feat = torch.empty((B * T, Cdim), dtype=torch.float32, device=cam_feat.device)
    check(lib.vs_camera_head_backward(
        C.c_void_p(ptr(cam_feat)), C.c_int64(cam_feat.stride(0)), C.c_void_p(ptr(w)), C.c_void_p(ptr(b)), B, T,
        Cdim, C.c_void_p(ptr(d_pred)), C.c_void_p(ptr(d_feat)), C.c_int64(Cdim), C.c_void_p(ptr(dw)),
        C.c_void_p(ptr(db)), C.c_void_p(stream_ptr())), "vs_camera_head_backward")
    return d_feat


def dropout_(x, p: float, seed: int):
    """In-place training-mode dropout on a contiguous bf16 tensor (vs_dropout_bf16)."""
    lib = _lib.load()
    _need_cuda(x)
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    check(lib.vs_dropout_bf16(C.c_void_p(ptr(x)), C.c_int64(x.numel()), C.c_float(p), C.c_uint64(seed & (2 ** 64 - 1)),
                              C.c_void_p(stream_ptr())), "vs_dropout_bf16")
    return x


# ------------------------------------------------------------------ LPIPS consumer
def maxpool2(x_nhwc):
    lib = _lib.load()
    n, h, w, c = x_nhwc.shape
    assert x_nhwc.dtype == torch.bfloat16 and x_nhwc.is_contiguous()
    y = torch.empty((n, h // 2, w // 2, c), dtype=torch.bfloat16, device=x_nhwc.device)
    check(lib.vs_maxpool2(C.c_void_p(ptr(x_nhwc)), C.c_void_p(ptr(y)), n, h, w, c, C.c_void_p(stream_ptr())),
          "vs_maxpool2")
    return y


def maxpool2_backward(x, y, dy, add=None, relu_mask=False):
    lib = _lib.load()
    n, h, w, c = x.shape
    dx = torch.empty_like(x)
    check(lib.vs_maxpool2_backward(C.c_void_p(ptr(x)), C.c_void_p(ptr(y)), C.c_void_p(ptr(dy)), C.c_void_p(ptr(add)),
                                   C.c_void_p(ptr(dx)), n, h, w, c, int(relu_mask), C.c_void_p(stream_ptr())),
          "vs_maxpool2_backward")
    return dx


def lpips_layer(f0, f1, wlin, per_image, grad_scale, want_grad=True):
    """per_image (N,) += layer distance; returns d f0 (bf16, masked by f0 > 0) of grad_scale * sum over pixels."""
    lib = _lib.load()
    n, h, w, c = f0.shape
    df0 = torch.empty_like(f0) if want_grad else None
    check(lib.vs_lpips_layer(C.c_void_p(ptr(f0)), C.c_void_p(ptr(f1)), C.c_void_p(ptr(wlin)), C.c_int64(n * h * w), c,
                             h * w, C.c_float(grad_scale), C.c_void_p(ptr(per_image)), C.c_void_p(ptr(df0)),
                             C.c_void_p(stream_ptr())), "vs_lpips_layer")
    return df0
