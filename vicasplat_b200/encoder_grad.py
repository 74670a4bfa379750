"""Training path of the ViT encoder block (croco/blocks.py:81-130), differentiated by
hand: the reference obtains these gradients from torch.autograd over ~20 eager kernels per block; here
every contraction of the backward pass runs on the same tcgen05 GEMM as the forward pass and the rest is
four fused bandwidth-bound kernels (``vs_grad_prep``, ``vs_layernorm_backward``, ``vs_gelu_bf16``,
``vs_rope_rows_backward``) plus the flash-style attention backward.

Numerics follow the forward path: bf16 GEMM operands (activations, weights AND gradients), fp32
accumulation, fp32 residual-stream gradient, fp32 weight gradients accumulated in place.

Per linear layer y = x W^T + b (W is (N, K)):
    dgrad   dx  = dy W          vs_gemm(A = dy (M, N), W = W^T (K, N))
    wgrad   dW += dy^T x        vs_gemm a_mode 2: dy (M, N) and x (M, K) as they are stored -- the tensor
                                core reads MN-major shared-memory tiles, there are no transposed copies
    bias    db += colsum(dy)    a by-product of the pass that makes the bf16 copy of dy
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

from . import ops
from ._lib import VS_ACT_NONE

_LIN = ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")


def pack_block(sd: Dict[str, torch.Tensor], key: str, device) -> Dict[str, torch.Tensor]:
    """bf16 operand copies of one block's weights, both orientations (W for the forward GEMM and
    wgrad's output layout, W^T for dgrad), fp32 biases / LayerNorm parameters.  `sd` uses the
    reference's state_dict names (``backbone.enc_blocks.N.*``)."""
    w = {}
    for name in _LIN:
        W = sd[f"{key}.{name}.weight"].to(device=device, dtype=torch.float32)
        w[name] = W.to(torch.bfloat16).contiguous()
        w[name + ".t"] = W.t().contiguous().to(torch.bfloat16)
        w[name + ".bias"] = sd[f"{key}.{name}.bias"].to(device=device, dtype=torch.float32).contiguous()
    for name in ("norm1", "norm2"):
        w[name + ".weight"] = sd[f"{key}.{name}.weight"].to(device=device, dtype=torch.float32).contiguous()
        w[name + ".bias"] = sd[f"{key}.{name}.bias"].to(device=device, dtype=torch.float32).contiguous()
    return w


def zero_grads(w: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """fp32 gradient accumulators named like the reference's parameters (``attn.qkv.weight`` ...)."""
    g = {}
    for name in _LIN:
        g[name + ".weight"] = torch.zeros(w[name].shape, dtype=torch.float32, device=w[name].device)
        g[name + ".bias"] = torch.zeros_like(w[name + ".bias"])
    for name in ("norm1", "norm2"):
        g[name + ".weight"] = torch.zeros_like(w[name + ".weight"])
        g[name + ".bias"] = torch.zeros_like(w[name + ".bias"])
    return g


@dataclass
class FrameLayout:
    """Token rows of the encoder stream: `frames` items of `n` consecutive rows (257 = 256 patches +
    the intrinsic token), their (y, x) rope positions and the item tables of vs_attention."""
    frames: int
    n: int
    heads: int
    pos: torch.Tensor            # (frames * n, 2) int32
    start: torch.Tensor          # (frames,) int32
    length: torch.Tensor         # (frames,) int32
    rope_base: float = 100.0

    @staticmethod
    def make(frames: int, gh: int, gw: int, heads: int, device, with_intrinsic: bool = True):
        ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
        pos = torch.stack([ys, xs], -1).reshape(-1, 2)
        if with_intrinsic:   # backbone_vica.py:455-459: the intrinsic token sits at (gh, 0)
            pos = torch.cat([pos, torch.tensor([[gh, 0]])], 0)
        n = pos.shape[0]
        pos = pos.repeat(frames, 1).to(device=device, dtype=torch.int32).contiguous()
        start = (torch.arange(frames, dtype=torch.int32) * n).to(device)
        length = torch.full((frames,), n, dtype=torch.int32, device=device)
        return FrameLayout(frames, n, heads, pos, start, length)


@dataclass
class Saved:
    t: Dict[str, torch.Tensor] = field(default_factory=dict)


# ------------------------------------------------------------------ linear layer pieces
def linear_backward(dy, x, w_t, dW, *, dx_dtype=torch.float32):
    """dgrad + wgrad of one linear layer (see module docstring).
    dy (M, N) bf16, x (M, K) bf16 (the layer's input as saved by the forward pass), w_t (K, N) bf16,
    dW (N, K) fp32 accumulated in place."""
    ops.gemm_tn_acc(dy, x, dW)                                 # dW += dy^T x (split-K over the tokens)
    return ops.gemm(dy, w_t, out_dtype=dx_dtype)               # dx = dy W


# ------------------------------------------------------------------ MLP half: x + fc2(gelu(fc1(LN2(x))))
def mlp_half_forward(x, w, saved: Optional[Saved] = None):
    """croco/blocks.py:58-79,128: returns x + mlp(norm2(x)); keeps what the backward pass needs."""
    h, _ = ops.layernorm(x, w["norm2.weight"], w["norm2.bias"], eps=w.get("ln_eps", 1e-6))
    z = ops.gemm(h, w["mlp.fc1"], bias=w["mlp.fc1.bias"], act=VS_ACT_NONE)
    a = ops.gelu_bf16(z)
    out = ops.gemm(a, w["mlp.fc2"], bias=w["mlp.fc2.bias"], res1=x, out_dtype=torch.float32)
    if saved is not None:
        saved.t.update(mlp_x=x, mlp_h=h, mlp_z=z, mlp_a=a)
    return out


def mlp_half_backward(dout, w, g, saved: Saved):
    """dout: fp32 gradient of the block output (M, E).  Returns the fp32 gradient of the half's input
    and accumulates the parameter gradients into `g`."""
    s = saved.t
    dy, _ = ops.grad_prep(dout, want_t=False, colsum=g["mlp.fc2.bias"])
    da = linear_backward(dy, s["mlp_a"], w["mlp.fc2.t"], g["mlp.fc2.weight"], dx_dtype=torch.bfloat16)
    dz, _ = ops.grad_prep(da, z=s["mlp_z"], want_t=False, colsum=g["mlp.fc1.bias"])   # da * gelu'(z)
    dh = linear_backward(dz, s["mlp_h"], w["mlp.fc1.t"], g["mlp.fc1.weight"])
    return ops.layernorm_backward(s["mlp_x"], dh, w["norm2.weight"], dres=dout,
                                  dgamma=g["norm2.weight"], dbeta=g["norm2.bias"], eps=w.get("ln_eps", 1e-6))


# ------------------------------------------------------------------ attention half: x + proj(attn(LN1(x)))
def attn_half_forward(x, w, lay: FrameLayout, saved: Optional[Saved] = None):
    """croco/blocks.py:81-127 with RoPE2D fused into the qkv epilogue."""
    E = x.shape[1]
    h, _ = ops.layernorm(x, w["norm1.weight"], w["norm1.bias"], eps=w.get("ln_eps", 1e-6))
    qkv = ops.gemm(h, w["attn.qkv"], bias=w["attn.qkv.bias"],
                   rope=(lay.pos, 0, E, lay.heads, lay.rope_base, 30.0))
    o = torch.empty((x.shape[0], E), dtype=torch.bfloat16, device=x.device)
    lse = torch.empty((x.shape[0], lay.heads), dtype=torch.float32, device=x.device)
    ops.attention(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], o, heads=lay.heads, q_start=lay.start,
                  q_len=lay.length, kv_start0=lay.start, kv_len0=lay.length, max_q_len=lay.n,
                  max_kv_len=lay.n, scale=0.125, lse=lse)
    out = ops.gemm(o, w["attn.proj"], bias=w["attn.proj.bias"], res1=x, out_dtype=torch.float32)
    if saved is not None:
        saved.t.update(att_x=x, att_h=h, att_qkv=qkv, att_o=o, att_lse=lse)
    return out


def attn_half_backward(dout, w, g, lay: FrameLayout, saved: Saved):
    s = saved.t
    E = dout.shape[1]
    dy, _ = ops.grad_prep(dout, want_t=False, colsum=g["attn.proj.bias"])
    do = linear_backward(dy, s["att_o"], w["attn.proj.t"], g["attn.proj.weight"], dx_dtype=torch.bfloat16)
    qkv = s["att_qkv"]
    dqkv = torch.empty_like(qkv)
    ops.attention_backward(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], s["att_o"], do, s["att_lse"],
                           dqkv[:, :E], dqkv[:, E:2 * E], dqkv[:, 2 * E:], heads=lay.heads,
                           q_start=lay.start, q_len=lay.length, kv_start0=lay.start,
                           kv_len0=lay.length, max_q_len=lay.n, max_kv_len=lay.n, scale=0.125)
    ops.rope_rows_backward(dqkv, lay.pos, heads=lay.heads, q_col=0, k_col=E, base=lay.rope_base)
    ops.grad_prep(dqkv, want_copy=False, want_t=False, colsum=g["attn.qkv.bias"])     # bias gradient only
    dh = linear_backward(dqkv, s["att_h"], w["attn.qkv.t"], g["attn.qkv.weight"])
    return ops.layernorm_backward(s["att_x"], dh, w["norm1.weight"], dres=dout,
                                  dgamma=g["norm1.weight"], dbeta=g["norm1.bias"], eps=w.get("ln_eps", 1e-6))


# ------------------------------------------------------------------ whole block
def block_forward(x, w, lay: FrameLayout, saved: Optional[Saved] = None):
    return mlp_half_forward(attn_half_forward(x, w, lay, saved), w, saved)


def block_backward(dout, w, g, lay: FrameLayout, saved: Saved):
    return attn_half_backward(mlp_half_backward(dout, w, g, saved), w, g, lay, saved)


# ------------------------------------------------------------------ neighbour attention: backward item tables
def neighbour_backward_tables(B: int, T: int, rows_per_frame: int, n_img: int, first_img_row: int = 1):
    """Item tables for the dK / dV pass of CrossNeighborAttention (backbone_vica.py:129-191).

    Forward: query frame t reads the image rows of frames prev(t) = t-1 (t+1 for t = 0) and nxt(t) = t+1
    (t-1 for t = T-1); when both coincide (end frames, T = 2) the frame is read ONCE -- softmax over a
    duplicated key set equals softmax over the set, and so does the gradient summed over the copies.
    A key frame j is therefore read by at most two query frames; dK / dV are computed per KEY frame
    with those query frames as (up to) two query segments, so every dK / dV row is written exactly once
    -- no atomics, no accumulation pass.  The softmax statistics (lse, delta) are indexed by absolute
    query row, each over the query frame's own key set.

    Rows: frame f of scene b starts at (b * T + f) * rows_per_frame; its n_img image rows start
    `first_img_row` rows later (row 0 is the camera token).  Returns int32 CPU tensors of B * T items:
    kv_start, kv_len, q_start0, q_len0, q_start1, q_len1."""
    def prev(t):
        return t - 1 if t > 0 else t + 1

    def nxt(t):
        return t + 1 if t < T - 1 else t - 1

    kv_start, q0, l0, q1, l1 = [], [], [], [], []
    for b in range(B):
        for j in range(T):
            readers = [t for t in range(T) if j in {prev(t), nxt(t)}] if T > 1 else []
            assert len(readers) <= 2
            row = lambda f: (b * T + f) * rows_per_frame + first_img_row
            kv_start.append(row(j))
            q0.append(row(readers[0]) if readers else 0)
            l0.append(n_img if readers else 0)
            q1.append(row(readers[1]) if len(readers) > 1 else 0)
            l1.append(n_img if len(readers) > 1 else 0)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32)
    return dict(kv_start=i32(kv_start), kv_len=torch.full((B * T,), n_img, dtype=torch.int32),
                q_start0=i32(q0), q_len0=i32(l0), q_start1=i32(q1), q_len1=i32(l1))
