"""TEST INFRASTRUCTURE (oracle): the backward formulas of the DPT heads' operators
(heads/dpt_block.py:79-229,264-459; forward restated in encoder_ref.dpt_trunk / pts_head / gs_head),
written the way the CUDA training path will run them on its implicit-GEMM kernel:

* stride-1 convolution  dgrad = the SAME stride-1 convolution of dY with the taps flipped and Cin / Cout
  swapped (padding k-1-p);  wgrad = dY^T . im2col(X), i.e. the wgrad-form GEMM (vs_gemm a_mode 2) over
  the pixel dimension with the A operand gathered per tap;
* stride-2 convolution (act_postprocess.3.1)  dgrad = zero-stuffed dY through the flipped stride-1 conv;
* ConvTranspose2d with kernel == stride (act_postprocess.{0,1}.1) = GEMM + pixel shuffle, so its
  backward is pixel-unshuffle + the two linear-layer GEMMs;
* bilinear x2, align_corners=True (Interpolate): y = U_h x U_w^T with two small interpolation matrices,
  backward dx = U_h^T dy U_w;
* ReLU masks from the saved outputs.

``tests/test_oracle_decoder_backward_cpu.py`` holds every formula to torch.autograd.  Nothing here is
imported by the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def conv_dgrad_s1(dy: Tensor, W: Tensor, pad: int) -> Tensor:
    """dX of y = conv2d(x, W, stride=1, padding=pad), W (O, C, k, k): conv2d(dy, W', padding=k-1-pad)
    with W'[c, o, u, v] = W[o, c, k-1-u, k-1-v]."""
    k = W.shape[-1]
    return F.conv2d(dy, W.flip(2, 3).transpose(0, 1), padding=k - 1 - pad)


def conv_wgrad(dy: Tensor, x: Tensor, k: int, pad: int, stride: int = 1) -> Tensor:
    """dW (O, C, k, k) = sum over pixels of dy[pixel, o] * im2col(x)[pixel, (c, u, v)]."""
    cols = F.unfold(x, k, padding=pad, stride=stride)            # (N, C k k, L)
    dW = torch.einsum("nol,nkl->ok", dy.flatten(2), cols)
    return dW.reshape(dy.shape[1], x.shape[1], k, k)


def conv_dgrad_strided(dy: Tensor, W: Tensor, pad: int, stride: int, in_hw) -> Tensor:
    """dX of a strided convolution: stuff stride-1 zeros between the elements of dy, then the flipped
    stride-1 convolution; rows / columns of the input that no window reaches get zero."""
    N, O, h, w = dy.shape
    k = W.shape[-1]
    H, Wd = in_hw
    stuffed = dy.new_zeros((N, O, (h - 1) * stride + 1, (w - 1) * stride + 1))
    stuffed[:, :, ::stride, ::stride] = dy
    full = F.conv2d(stuffed, W.flip(2, 3).transpose(0, 1), padding=k - 1)   # covers padded input rows
    out = dy.new_zeros((N, W.shape[1], H + 2 * pad, Wd + 2 * pad))
    out[:, :, : full.shape[2], : full.shape[3]] = full[:, :, : H + 2 * pad, : Wd + 2 * pad]
    return out[:, :, pad: pad + H, pad: pad + Wd]


def pixel_unshuffle_rows(dy: Tensor, s: int) -> Tensor:
    """(N, O, h s, w s) -> rows (N h w, O s s), column order (o, dy, dx) = ConvTranspose weight (C, O, s, s)
    flattened over its last three dimensions."""
    N, O, H, W = dy.shape
    t = dy.reshape(N, O, H // s, s, W // s, s).permute(0, 2, 4, 1, 3, 5)
    return t.reshape(N * (H // s) * (W // s), O * s * s)


def conv_transpose_ks_backward(dy: Tensor, x: Tensor, Wt: Tensor):
    """ConvTranspose2d(kernel = stride = s), Wt (C, O, s, s): forward rows(y) = rows(x) @ Wt.flatten(1).
    Returns (dx (N,C,h,w), dW (C,O,s,s), db (O))."""
    s = Wt.shape[-1]
    N, C, h, w = x.shape
    dyr = pixel_unshuffle_rows(dy, s)                              # (N h w, O s s)
    xr = x.permute(0, 2, 3, 1).reshape(-1, C)
    dx = (dyr @ Wt.flatten(1).t()).reshape(N, h, w, C).permute(0, 3, 1, 2)
    return dx, (xr.t() @ dyr).reshape(Wt.shape), dy.sum((0, 2, 3))


def interp_matrix(n_in: int, dtype=torch.float64) -> Tensor:
    """(2 n_in, n_in) matrix of bilinear x2 interpolation with align_corners=True."""
    n_out = 2 * n_in
    U = torch.zeros((n_out, n_in), dtype=dtype)
    for o in range(n_out):
        src = o * (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
        i0 = min(int(src), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        f = src - i0
        U[o, i0] += 1 - f
        U[o, i1] += f
    return U


def up2_forward(x: Tensor) -> Tensor:
    Uh, Uw = interp_matrix(x.shape[2], x.dtype), interp_matrix(x.shape[3], x.dtype)
    return torch.einsum("oh,nchw,pw->ncop", Uh, x, Uw)


def up2_backward(dy: Tensor) -> Tensor:
    Uh, Uw = interp_matrix(dy.shape[2] // 2, dy.dtype), interp_matrix(dy.shape[3] // 2, dy.dtype)
    return torch.einsum("oh,ncop,pw->nchw", Uh, dy, Uw)


def rcu_backward(dy: Tensor, x: Tensor, W1: Tensor, W2: Tensor, y1: Tensor):
    """ResidualConvUnit: out = conv2(relu(y1)) + x, y1 = conv1(relu(x)) (dpt_block.py:79-137).
    Returns dx, (dW1, db1), (dW2, db2)."""
    a1, a0 = F.relu(y1), F.relu(x)
    dW2, db2 = conv_wgrad(dy, a1, 3, 1), dy.sum((0, 2, 3))
    d_y1 = conv_dgrad_s1(dy, W2, 1) * (y1 > 0)
    dW1, db1 = conv_wgrad(d_y1, a0, 3, 1), d_y1.sum((0, 2, 3))
    dx = dy + conv_dgrad_s1(d_y1, W1, 1) * (x > 0)
    return dx, (dW1, db1), (dW2, db2)
