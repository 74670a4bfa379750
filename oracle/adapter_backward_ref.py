"""TEST INFRASTRUCTURE (oracle): HAND-DERIVED backward of the per-pixel tails of the encoder plugin --
``MyGaussianAdapter.forward`` (common/gaussian_adapter.py:167-212, common/gaussians.py:8-44), the
'exp' depth postprocess of the centre head (heads/postprocess.py:42-61) and the dual-quaternion
normalisation of the pose head (vicasplat.py:179-199) -- i.e. the chain rules that connect
``vs_raster_backward``'s outputs (d means, d covariances, d SH, d opacity) and the camera loss to the
gradient of the DPT heads' outputs.  Written without autograd, as the formulas the CUDA kernels will
implement; ``tests/test_oracle_decoder_backward_cpu.py`` holds them to torch.autograd over
oracle/encoder_ref.py.  Nothing here is imported by the product.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

from . import encoder_ref as er

Tensor = torch.Tensor


def _rot_matrix_and_grad(q: Tensor, dR: Optional[Tensor]):
    """R(q) of gaussians.py:8-31 (xyzw, two_s = 2 / (q.q + 1e-8)) and, if dR (...,3,3) is given,
    dL/dq (...,4)."""
    i, j, k, r = q.unbind(-1)
    u = (q * q).sum(-1) + 1e-8
    s = 2 / u
    # R = I + s * A(q) with A the quadratic form below
    A = torch.stack([-(j * j + k * k), i * j - k * r, i * k + j * r,
                     i * j + k * r, -(i * i + k * k), j * k - i * r,
                     i * k - j * r, j * k + i * r, -(i * i + j * j)], -1).reshape(q.shape[:-1] + (3, 3))
    R = torch.eye(3, dtype=q.dtype, device=q.device) + s[..., None, None] * A
    if dR is None:
        return R, None
    g = dR.reshape(q.shape[:-1] + (9,))
    g00, g01, g02, g10, g11, g12, g20, g21, g22 = g.unbind(-1)
    # dL/dA = s * dR; derivatives of A's entries w.r.t. (i, j, k, r)
    di = -2 * i * (g11 + g22) + j * (g01 + g10) + k * (g02 + g20) + r * (g21 - g12)
    dj = -2 * j * (g00 + g22) + i * (g01 + g10) + k * (g12 + g21) + r * (g02 - g20)
    dk = -2 * k * (g00 + g11) + i * (g02 + g20) + j * (g12 + g21) + r * (g10 - g01)
    dr = k * (g10 - g01) + j * (g02 - g20) + i * (g21 - g12)
    dq = s[..., None] * torch.stack([di, dj, dk, dr], -1)
    # through s = 2 / u: ds/dq = -s^2 q
    ds = (dR * A).sum((-1, -2))
    return R, dq - (ds * s * s)[..., None] * q


def adapter_backward(raw: Tensor, cfg: er.EncoderConfig, d_means: Tensor, d_cov: Tensor, d_sh: Tensor,
                     d_opac: Tensor, d_scales: Optional[Tensor] = None,
                     d_rot: Optional[Tensor] = None) -> Tensor:
    """dL/d raw (...,86) from the gradients of the adapter's outputs (vs_raster_backward delivers
    d means, d covariances, d harmonics, d opacities)."""
    xyz, o, s_raw, r_raw = raw[..., :11].split((3, 1, 3, 4), dim=-1)
    # opacity = sigmoid(o)
    op = torch.sigmoid(o)
    d_o = d_opac * op * (1 - op)
    # scales = min(0.001 * softplus(s), 0.3)
    sc_un = 0.001 * F.softplus(s_raw)
    sc = sc_un.clamp_max(0.3)
    # rotations = r / max(|r|, 1e-12)
    n = r_raw.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    q = r_raw / n
    # covariance = R diag(sc^2) R^T
    G = d_cov + d_cov.transpose(-1, -2)
    R, _ = _rot_matrix_and_grad(q, None)
    D = sc * sc
    dR = G @ (R * D[..., None, :])                       # (G + G^T) R D
    RtGR = R.transpose(-1, -2) @ d_cov @ R
    d_sc = 2 * sc * torch.diagonal(RtGR, dim1=-2, dim2=-1)
    if d_scales is not None:
        d_sc = d_sc + d_scales
    _, d_q = _rot_matrix_and_grad(q, dR)
    if d_rot is not None:
        d_q = d_q + d_rot
    d_s = d_sc * 0.001 * torch.sigmoid(s_raw) * (sc_un < 0.3)
    d_r = (d_q - q * (q * d_q).sum(-1, keepdim=True)) / n
    d_raw_sh = (d_sh * er.sh_mask(cfg, raw.device).to(raw.dtype)).flatten(-2)
    return torch.cat([d_means, d_o, d_s, d_r, d_raw_sh], dim=-1)


def exp_postprocess_backward(x: Tensor, d_xyz: Tensor) -> Tensor:
    """xyz = x / max(|x|, 1e-8) * expm1(|x|)  ->  dL/dx.  With f(d) = expm1(d) / d:
    dx = f g + x (x . g) f'(d) / d,  f'(d) = (e^d d - expm1(d)) / d^2."""
    d = x.norm(dim=-1, keepdim=True).clamp_min(1e-8)
    em = torch.expm1(d)
    f = em / d
    fp = (torch.exp(d) * d - em) / (d * d)
    return f * d_xyz + x * (x * d_xyz).sum(-1, keepdim=True) * fp / d


def dq_normalise_backward(v: Tensor, d_pred: Tensor) -> Tensor:
    """pred = v / |v[..., :4]| (all 8 components divided by the norm of the REAL part, after
    v[..., 3] += 1 -- vicasplat.py:183-190): dL/dv."""
    n = v[..., :4].norm(dim=-1, keepdim=True)
    y = v / n
    dot = (d_pred * y).sum(-1, keepdim=True)
    real = torch.cat([torch.ones_like(v[..., :4]), torch.zeros_like(v[..., 4:])], -1)
    return (d_pred - real * dot * y) / n
