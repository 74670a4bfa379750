"""Drop-in for the reference's ``curope`` extension and its autograd wrapper
(src/model/encoder/backbone/croco/curope/curope.cpp:49-69, curope2d.py:12-44).

``rope_2d(tokens, positions, base, fwd)`` rotates ``tokens`` (B, N, H, D) IN PLACE; the same
argument checks raise RuntimeError with the reference's messages.  The kernel runs on the current
PyTorch stream (the reference launches on the legacy default stream, kernels.cu:102).  CUDA only.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import VS_BF16, VS_F16, VS_F32, check, stream_ptr

_DT = {torch.float32: VS_F32, torch.float16: VS_F16, torch.bfloat16: VS_BF16}


def rope_2d(tokens: torch.Tensor, positions: torch.Tensor, base: float, fwd: float) -> None:
    if tokens.dim() != 4:
        raise RuntimeError("tokens must have 4 dimensions")
    if positions.dim() != 3:
        raise RuntimeError("positions must have 3 dimensions")
    if tokens.size(0) != positions.size(0):
        raise RuntimeError("batch size differs between tokens & positions")
    if tokens.size(1) != positions.size(1):
        raise RuntimeError("seq_length differs between tokens & positions")
    if positions.size(2) != 2:
        raise RuntimeError("positions.shape[2] must be equal to 2")
    if tokens.is_cuda != positions.is_cuda:
        raise RuntimeError("tokens and positions are not on the same device")
    if not tokens.is_cuda:
        raise RuntimeError("vicasplat_b200.curope runs on CUDA tensors only (no CPU fallback)")
    B, N, H, D = tokens.shape
    if tokens.stride(3) != 1 or tokens.stride(2) != D:
        raise RuntimeError("tokens must be contiguous in the last two dimensions")   # kernels.cu:91
    if D % 4 != 0:
        raise RuntimeError("token dim must be multiple of 4")
    if positions.dtype != torch.int64 or not positions.is_contiguous():
        raise RuntimeError("positions must be a contiguous int64 tensor")
    if tokens.dtype not in _DT:
        raise RuntimeError(f"rope_2d: unsupported dtype {tokens.dtype}")
    lib = _lib.load()
    check(lib.vs_rope_2d(C.c_void_p(tokens.data_ptr()), _DT[tokens.dtype], B, N, H, D,
                         C.c_int64(tokens.stride(0)), C.c_int64(tokens.stride(1)),
                         C.c_void_p(positions.data_ptr()), C.c_float(base), C.c_float(fwd),
                         C.c_void_p(stream_ptr())), "rope_2d")


class cuRoPE2D_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, positions, base, F0=1):
        ctx.save_for_backward(positions)
        ctx.saved_base, ctx.saved_F0 = base, F0
        rope_2d(tokens, positions, base, F0)
        ctx.mark_dirty(tokens)
        return tokens

    @staticmethod
    def backward(ctx, grad_res):
        positions, = ctx.saved_tensors
        rope_2d(grad_res, positions, ctx.saved_base, -ctx.saved_F0)   # inverse rotation
        ctx.mark_dirty(grad_res)
        return grad_res, None, None, None


class cuRoPE2D(torch.nn.Module):
    def __init__(self, freq: float = 100.0, F0: float = 1.0):
        super().__init__()
        self.base, self.F0 = freq, F0

    def forward(self, tokens, positions):
        """tokens (B, H, N, D), rotated in place through the (B, N, H, D) view."""
        cuRoPE2D_func.apply(tokens.transpose(1, 2), positions, self.base, self.F0)
        return tokens
