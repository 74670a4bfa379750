"""Seeded synthetic workloads of the shapes BASELINE.json names (there is no dataset / checkpoint
in the image): video clips for the encoder and pixel-aligned Gaussian scenes for the rasterizer
(SURVEY.md §8d -- random-weight encoder output is degenerate as raster input, so the raster leg is
fed Gaussians laid out the way a trained encoder lays them out: one per context pixel, on the
pixel's ray, about one pixel wide)."""
from __future__ import annotations

from math import isqrt
from typing import Optional

import torch


def clip(B: int, T: int, size: int, seed: int = 250307):
    """image (B,T,3,size,size) in [-1,1] and normalised intrinsics (B,T,3,3), on the CPU."""
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((B, T, 3, size, size), generator=g) * 2 - 1
    K = torch.tensor([[0.86, 0, 0.5], [0, 0.86, 0.5], [0, 0, 1.0]]).expand(B, T, 3, 3).clone()
    return image, K


def gaussian_scene(n_ctx: int, h: int, w: int, n_tgt: int, seed: int = 250307, d_sh: int = 25,
                   depth_range=(1.0, 20.0), focal: float = 0.86, n_gauss: Optional[int] = None, device=None):
    """n_ctx*h*w Gaussians seen from n_ctx cameras on a unit baseline, n_tgt target cameras on the
    same line.  Returns fp32 tensors (on the CPU, or generated on `device` with that device's seeded
    generator -- a different but equally distributed scene): means (G,3), covariances (G,3,3), harmonics
    (G,3,d_sh), opacities (G,), extrinsics (n_tgt,4,4) c2w, intrinsics (n_tgt,3,3), near, far (n_tgt,)."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    g = torch.Generator(device=device).manual_seed(seed)
    f64 = torch.float64
    _rand, _randn, _arange, _eye, _ones = torch.rand, torch.randn, torch.arange, torch.eye, torch.ones

    class _T:   # the torch factory functions of the code below, on `device`
        rand = staticmethod(lambda *a, **k: _rand(*a, device=device, **k))
        randn = staticmethod(lambda *a, **k: _randn(*a, device=device, **k))
    K = torch.tensor([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], dtype=f64, device=device)

    def cams(xs):
        c = _eye(4, dtype=f64, device=device).repeat(len(xs), 1, 1)
        c[:, 0, 3] = torch.tensor(xs, dtype=f64, device=device)
        return c

    ctx = cams([i / max(n_ctx - 1, 1) for i in range(n_ctx)])
    tgt = cams([(i + 0.5) / n_tgt for i in range(n_tgt)])
    ys, xs = torch.meshgrid((_arange(h, dtype=f64, device=device) + 0.5) / h,
                            (_arange(w, dtype=f64, device=device) + 0.5) / w, indexing="ij")
    rays = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).reshape(-1, 3) @ torch.linalg.inv(K).T
    n = h * w
    lo, hi = depth_range
    z = lo * (hi / lo) ** _T.rand((n_ctx, n), generator=g, dtype=f64)
    pts = rays[None] * z[..., None] + ctx[:, None, :3, 3]
    sigma = z / (focal * w) * (0.5 + _T.rand((n_ctx, n), generator=g, dtype=f64))
    aniso = 1.0 + 2.0 * _T.rand((n_ctx, n, 3), generator=g, dtype=f64)
    scales = sigma[..., None] * aniso / aniso.mean(-1, keepdim=True)
    q = _T.randn((n_ctx, n, 4), generator=g, dtype=f64)
    q = q / q.norm(dim=-1, keepdim=True)
    i, j, k, r = q.unbind(-1)
    R = torch.stack([1 - 2 * (j * j + k * k), 2 * (i * j - k * r), 2 * (i * k + j * r),
                     2 * (i * j + k * r), 1 - 2 * (i * i + k * k), 2 * (j * k - i * r),
                     2 * (i * k - j * r), 2 * (j * k + i * r), 1 - 2 * (i * i + j * j)],
                    dim=-1).reshape(n_ctx, n, 3, 3)
    cov = R @ torch.diag_embed(scales ** 2) @ R.transpose(-1, -2)
    opac = 0.05 + 0.9 * _T.rand((n_ctx, n), generator=g, dtype=f64)
    sh = 0.5 * _T.randn((n_ctx, n, 3, d_sh), generator=g, dtype=f64)
    mask = _ones(d_sh, dtype=f64, device=device)
    for deg in range(1, isqrt(d_sh)):
        mask[deg * deg:(deg + 1) ** 2] = 0.1 * 0.25 ** deg
    out = dict(means=pts.reshape(-1, 3), covariances=cov.reshape(-1, 3, 3),
               harmonics=(sh * mask).reshape(-1, 3, d_sh), opacities=opac.reshape(-1))
    if n_gauss is not None:
        sel = torch.randperm(out["means"].shape[0], generator=g, device=device)[:n_gauss]
        out = {k_: v[sel] for k_, v in out.items()}
    out = {k_: v.float().contiguous() for k_, v in out.items()}
    out.update(extrinsics=tgt.float(), intrinsics=K.float()[None].repeat(n_tgt, 1, 1),
               near=torch.full((n_tgt,), 0.01, device=device), far=torch.full((n_tgt,), 100.0, device=device))
    return out


def vit_encoder_state_dict(depth: int = 24, embed: int = 1024, patch: int = 16, mlp_ratio: float = 4.0,
                           seed: int = 0):
    """Seeded random weights of the image encoder (patch embedding, intrinsic token, `depth` ViT blocks,
    final norm) under the reference's state_dict names (backbone_vica.py:380-399,431-448: xavier-uniform
    linears) -- with non-trivial biases / LayerNorm parameters so that every gradient path is exercised.
    fp32 CPU tensors."""
    g = torch.Generator().manual_seed(seed)

    def xavier(out_f, in_f, shape=None):
        a = (6.0 / (in_f + out_f)) ** 0.5
        return (torch.rand(shape or (out_f, in_f), generator=g) * 2 - 1) * a

    def small(n):
        return 0.02 * torch.randn((n,), generator=g)

    hidden = int(embed * mlp_ratio)
    sd = {
        "backbone.patch_embed.proj.weight": xavier(embed, 3 * patch * patch, (embed, 3, patch, patch)),
        "backbone.patch_embed.proj.bias": small(embed),
        "backbone.intrinsic_encoder.weight": xavier(embed, 9),
        "backbone.intrinsic_encoder.bias": small(embed),
        "backbone.enc_norm.weight": 1 + small(embed),
        "backbone.enc_norm.bias": small(embed),
    }
    for i in range(depth):
        k = f"backbone.enc_blocks.{i}."
        for name, (o, n) in (("attn.qkv", (3 * embed, embed)), ("attn.proj", (embed, embed)),
                             ("mlp.fc1", (hidden, embed)), ("mlp.fc2", (embed, hidden))):
            sd[k + name + ".weight"] = xavier(o, n)
            sd[k + name + ".bias"] = small(o)
        for name in ("norm1", "norm2"):
            sd[k + name + ".weight"] = 1 + small(embed)
            sd[k + name + ".bias"] = small(embed)
    return sd
