"""TEST INFRASTRUCTURE (oracle): the HAND-DERIVED backward pass of one MixDecoderBlock
(src/model/encoder/backbone/backbone_vica.py:280-335; forward restated in encoder_ref.dec_block).

The reference gets these gradients from torch.autograd.  The CUDA training path differentiates by
hand, kernel by kernel; this file is the CPU statement of exactly those formulas -- LayerNorm backward,
AdaLN modulate / gate backward with their per-frame reductions, SiLU / GELU derivatives, the softmax
attention backward (dP, dS, dQ, dK, dV), the rotary embeddings' transposes, the shared-weight
accumulations (qkv / proj see image AND camera rows) and the neighbour attention's scatter of dK / dV
into key frames that several query frames share -- written without autograd, so that every stage of the
decoder kernels to come has a stage-level oracle.  ``tests/test_oracle_decoder_backward_cpu.py`` holds
it to torch.autograd over ``encoder_ref.dec_block`` (which is pinned to the reference's own gradients
by tests/golden/model_grad_small.npz).  Nothing here is imported by the product.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import encoder_ref as er

Tensor = torch.Tensor


# ------------------------------------------------------------------ primitives (forward with cache, backward)
def ln_fwd(x: Tensor, w: Tensor, b: Tensor, eps: float):
    mu = x.mean(-1, keepdim=True)
    rstd = (x.var(-1, unbiased=False, keepdim=True) + eps).rsqrt()
    xhat = (x - mu) * rstd
    return xhat * w + b, (xhat, rstd)


def ln_bwd(dy: Tensor, cache, w: Tensor):
    """dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * w  (vs_layernorm_backward)."""
    xhat, rstd = cache
    g = dy * w
    dx = rstd * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    red = tuple(range(dy.dim() - 1))
    return dx, (dy * xhat).sum(red), dy.sum(red)


def linear_bwd(dy: Tensor, x: Tensor, W: Tensor):
    """y = x W^T + b: dgrad dy W, wgrad dy^T x, bias colsum(dy)."""
    dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
    return dy @ W, dy2.t() @ x2, dy2.sum(0)


def gelu_grad(z: Tensor) -> Tensor:
    return 0.5 * (1 + torch.erf(z / math.sqrt(2))) + z * torch.exp(-0.5 * z * z) / math.sqrt(2 * math.pi)


def silu_grad(x: Tensor) -> Tensor:
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


def sdpa_fwd(q, k, v, mask=None):
    s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    P = s.softmax(-1)
    return P @ v, P


def sdpa_bwd(do, q, k, v, P):
    """dV = P^T dO, dP = dO V^T, dS = P (dP - rowsum(dP P)) * scale, dQ = dS K, dK = dS^T Q
    (rowsum(dP P) = rowsum(dO O): the `delta` of vs_attention_backward)."""
    dv = P.transpose(-1, -2) @ do
    dP = do @ v.transpose(-1, -2)
    dS = P * (dP - (dP * P).sum(-1, keepdim=True)) * (q.shape[-1] ** -0.5)
    return dS @ k, dS.transpose(-1, -2) @ q, dv


def rope2d_bwd(dt: Tensor, pos: Tensor, base: float) -> Tensor:
    """The rotation is orthogonal: its transpose is the rotation by the negated angle, i.e. the same
    routine with the sine flipped = rope on positions of opposite sign (cuRoPE2D backward: fwd = -1)."""
    return er.rope2d(dt, -pos, base)


def rope1d_bwd(dt: Tensor, frame: Tensor, theta: float) -> Tensor:
    return er.rope1d_interleaved(dt, -frame, theta)


def _acc(g: Dict[str, Tensor], key: str, val: Tensor) -> None:
    g[key] = g[key] + val if key in g else val


def _lin_fwd(sd, key, x):
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def _lin_bwd(sd, g, key, dy, x):
    dx, dW, db = linear_bwd(dy, x, sd[key + ".weight"])
    _acc(g, key + ".weight", dW)
    if key + ".bias" in sd:
        _acc(g, key + ".bias", db)
    return dx


def _ln_bwd(sd, g, key, dy, cache):
    dx, dw, db = ln_bwd(dy, cache, sd[key + ".weight"])
    _acc(g, key + ".weight", dw)
    _acc(g, key + ".bias", db)
    return dx


# ------------------------------------------------------------------ AdaLN modulation (backbone_vica.py:194-212)
def modulation_fwd(sd, key, emb, n):
    a = F.silu(emb)
    return _lin_fwd(sd, key + ".proj", a).chunk(n, dim=-1), (emb, a)


def modulation_bwd(sd, g, key, d_chunks, cache):
    emb, a = cache
    d = torch.cat(d_chunks, dim=-1)
    return _lin_bwd(sd, g, key + ".proj", d, a) * silu_grad(emb)


def modulate_bwd(dy, ln_out, sc):
    """y = ln_out * (1 + sc) + sh with per-frame sc / sh (B,T,1,C) broadcast over the N rows of a frame:
    d ln_out = dy (1 + sc); d sc = sum_rows dy ln_out; d sh = sum_rows dy."""
    return dy * (1 + sc), (dy * ln_out).sum(2, keepdim=True), dy.sum(2, keepdim=True)


def gate_bwd(dout, branch, gt):
    """out = x + (1 + gt) * branch: d branch = dout (1 + gt); d gt = sum_rows dout branch."""
    return dout * (1 + gt), (dout * branch).sum(2, keepdim=True)


# ------------------------------------------------------------------ MLP (croco/blocks.py:58-79)
def mlp_fwd(sd, key, x):
    z = _lin_fwd(sd, key + ".fc1", x)
    a = F.gelu(z)
    return _lin_fwd(sd, key + ".fc2", a), (x, z, a)


def mlp_bwd(sd, g, key, dy, cache):
    x, z, a = cache
    da = _lin_bwd(sd, g, key + ".fc2", dy, a)
    return _lin_bwd(sd, g, key + ".fc1", da * gelu_grad(z), x)


# ------------------------------------------------------------------ VideoCameraAttention (backbone_vica.py:57-126)
def video_attn_fwd(sd, key, img, cam, pos, cfg):
    B, T, N, C = img.shape
    H = cfg.dec_num_heads
    hd = C // H
    heads = lambda t, L: t.reshape(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    qi0, ki0, vi = heads(_lin_fwd(sd, key + ".qkv", img), T * N)
    p = pos.reshape(B, T * N, 2)
    qi, ki = er.rope2d(qi0, p, cfg.rope_base), er.rope2d(ki0, p, cfg.rope_base)
    qc0, kc0, vc = heads(_lin_fwd(sd, key + ".qkv", cam), T)
    fr = torch.arange(T, device=img.device)
    qc = er.rope1d_interleaved(qc0, fr, cfg.temporal_rope_theta)
    kc = er.rope1d_interleaved(kc0, fr, cfg.temporal_rope_theta)
    kf = torch.cat([kc[:, :, :, None], ki.reshape(B, H, T, N, hd)], dim=3).reshape(B, H, -1, hd)
    vf = torch.cat([vc[:, :, :, None], vi.reshape(B, H, T, N, hd)], dim=3).reshape(B, H, -1, hd)
    mask = er.camera_mask(T, N, img.device)
    oi, Pi = sdpa_fwd(qi, kf, vf)
    oc, Pc = sdpa_fwd(qc, kf, vf, mask)
    oi2 = oi.transpose(1, 2).reshape(B, T, N, C)
    oc2 = oc.transpose(1, 2).reshape(B, T, C)
    cache = dict(img=img, cam=cam, p=p, fr=fr, qi=qi, qc=qc, kf=kf, vf=vf, Pi=Pi, Pc=Pc, oi2=oi2, oc2=oc2)
    return _lin_fwd(sd, key + ".proj", oi2), _lin_fwd(sd, key + ".proj", oc2), cache


def video_attn_bwd(sd, g, key, d_ai, d_ac, c, cfg):
    B, T, N, C = c["img"].shape
    H = cfg.dec_num_heads
    hd = C // H
    # the projection is shared by image and camera rows: both contributions land in the same dW
    d_oi2 = _lin_bwd(sd, g, key + ".proj", d_ai, c["oi2"])
    d_oc2 = _lin_bwd(sd, g, key + ".proj", d_ac, c["oc2"])
    d_oi = d_oi2.reshape(B, T * N, H, hd).transpose(1, 2)
    d_oc = d_oc2.reshape(B, T, H, hd).transpose(1, 2)
    dqi, dkf_i, dvf_i = sdpa_bwd(d_oi, c["qi"], c["kf"], c["vf"], c["Pi"])
    dqc, dkf_c, dvf_c = sdpa_bwd(d_oc, c["qc"], c["kf"], c["vf"], c["Pc"])
    dkf = (dkf_i + dkf_c).reshape(B, H, T, N + 1, hd)       # keys per frame: [camera | image tokens]
    dvf = (dvf_i + dvf_c).reshape(B, H, T, N + 1, hd)
    dkc, dki = dkf[:, :, :, 0], dkf[:, :, :, 1:].reshape(B, H, T * N, hd)
    dvc, dvi = dvf[:, :, :, 0], dvf[:, :, :, 1:].reshape(B, H, T * N, hd)
    dqi0, dki0 = rope2d_bwd(dqi, c["p"], cfg.rope_base), rope2d_bwd(dki, c["p"], cfg.rope_base)
    dqc0 = rope1d_bwd(dqc, c["fr"], cfg.temporal_rope_theta)
    dkc0 = rope1d_bwd(dkc, c["fr"], cfg.temporal_rope_theta)
    unheads = lambda q, k, v, L: torch.stack([q, k, v], 0).permute(1, 3, 0, 2, 4).reshape(B, L, 3 * C)
    d_img = _lin_bwd(sd, g, key + ".qkv", unheads(dqi0, dki0, dvi, T * N), c["img"].reshape(B, T * N, C))
    d_cam = _lin_bwd(sd, g, key + ".qkv", unheads(dqc0, dkc0, dvc, T), c["cam"])
    return d_img.reshape(B, T, N, C), d_cam


# ------------------------------------------------------------------ CrossNeighborAttention (backbone_vica.py:129-191)
def _neighbours(t, T):
    return [1 - t] if T == 2 else [t - 1 if t > 0 else 1, t + 1 if t < T - 1 else T - 2]


def neighbour_attn_fwd(sd, key, img, pos, cfg):
    B, T, N, C = img.shape
    H = cfg.dec_num_heads
    hd = C // H
    proj = lambda name: _lin_fwd(sd, f"{key}.{name}", img).reshape(B, T, N, H, hd).permute(0, 1, 3, 2, 4)
    q0, k0, v = proj("projq"), proj("projk"), proj("projv")
    p = pos.reshape(B * T, N, 2)
    rope = lambda t: er.rope2d(t.reshape(B * T, H, N, hd), p, cfg.rope_base).reshape(B, T, H, N, hd)
    q, k = rope(q0), rope(k0)
    outs, Ps = [], []
    for t in range(T):
        nb = _neighbours(t, T)
        o, P = sdpa_fwd(q[:, t], torch.cat([k[:, j] for j in nb], 2), torch.cat([v[:, j] for j in nb], 2))
        outs.append(o)
        Ps.append(P)
    o = torch.stack(outs, 1).permute(0, 1, 3, 2, 4).reshape(B, T, N, C)
    return _lin_fwd(sd, key + ".proj", o), dict(img=img, p=p, q=q, k=k, v=v, Ps=Ps, o=o)


def neighbour_attn_bwd(sd, g, key, dy, c, cfg):
    B, T, N, C = c["img"].shape
    H = cfg.dec_num_heads
    hd = C // H
    d_o = _lin_bwd(sd, g, key + ".proj", dy, c["o"]).reshape(B, T, N, H, hd).permute(0, 1, 3, 2, 4)
    q, k, v = c["q"], c["k"], c["v"]
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    for t in range(T):
        nb = _neighbours(t, T)
        kk, vv = torch.cat([k[:, j] for j in nb], 2), torch.cat([v[:, j] for j in nb], 2)
        dq_t, dkk, dvv = sdpa_bwd(d_o[:, t], q[:, t], kk, vv, c["Ps"][t])
        dq[:, t] = dq_t
        for i, j in enumerate(nb):      # a key frame is read by up to two query frames: accumulate
            dk[:, j] += dkk[:, :, i * N:(i + 1) * N]
            dv[:, j] += dvv[:, :, i * N:(i + 1) * N]
    unrope = lambda t: rope2d_bwd(t.reshape(B * T, H, N, hd), c["p"], cfg.rope_base).reshape(B, T, H, N, hd)
    flat = lambda t: t.permute(0, 1, 3, 2, 4).reshape(B, T, N, C)
    d_img = _lin_bwd(sd, g, key + ".projq", flat(unrope(dq)), c["img"])
    d_img = d_img + _lin_bwd(sd, g, key + ".projk", flat(unrope(dk)), c["img"])
    return d_img + _lin_bwd(sd, g, key + ".projv", flat(dv), c["img"])


# ------------------------------------------------------------------ the block
def dec_block_fwd(sd, key, img, cam, pos, cfg) -> Tuple[Tensor, Tensor, dict]:
    eps, c = cfg.ln_eps, {}
    ln = lambda name, x: ln_fwd(x, sd[f"{key}.{name}.weight"], sd[f"{key}.{name}.bias"], eps)
    cn, c["cam_norm1"] = ln("cam_norm1", cam)
    (sc, sh, gt), c["mod1"] = modulation_fwd(sd, key + ".modulation1", cn[:, :, None], 3)
    n1, c["norm1"] = ln("norm1", img)
    ai, ac, c["attn"] = video_attn_fwd(sd, key + ".attn", n1 * (1 + sc) + sh, cn, pos, cfg)
    c.update(n1=n1, sc=sc, gt=gt, ai=ai)
    img = img + (1 + gt) * ai
    cam = cam + ac
    cn2, c["cam_norm2"] = ln("cam_norm2", cam)
    (sc1, sh1, gt1, sc2, sh2, gt2), c["mod2"] = modulation_fwd(sd, key + ".modulation2", cn2[:, :, None], 6)
    n2, c["norm2"] = ln("norm2", img)
    xa, c["cross"] = neighbour_attn_fwd(sd, key + ".cross_attn", n2 * (1 + sc1) + sh1, pos, cfg)
    c.update(n2=n2, sc1=sc1, gt1=gt1, xa=xa)
    img = img + (1 + gt1) * xa
    n3, c["norm3"] = ln("norm3", img)
    m, c["mlp"] = mlp_fwd(sd, key + ".mlp", n3 * (1 + sc2) + sh2)
    c.update(n3=n3, sc2=sc2, gt2=gt2, m=m)
    img = img + (1 + gt2) * m
    mc, c["mlp_cam"] = mlp_fwd(sd, key + ".mlp_cam", cn2)
    return img, cam + mc, c


def dec_block_bwd(sd, key, d_img, d_cam, c, cfg) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """Gradients w.r.t. the block inputs (img, cam) and every parameter of the block (dict keyed by
    the reference's state_dict names)."""
    g: Dict[str, Tensor] = {}
    # cam = cam1 + mlp_cam(cn2)
    d_cn2 = mlp_bwd(sd, g, key + ".mlp_cam", d_cam, c["mlp_cam"])
    # img = img2 + (1 + gt2) * mlp(n3 (1 + sc2) + sh2)
    d_m, d_gt2 = gate_bwd(d_img, c["m"], c["gt2"])
    d_h3 = mlp_bwd(sd, g, key + ".mlp", d_m, c["mlp"])
    d_n3, d_sc2, d_sh2 = modulate_bwd(d_h3, c["n3"], c["sc2"])
    d_img = d_img + _ln_bwd(sd, g, key + ".norm3", d_n3, c["norm3"])
    # img2 = img1 + (1 + gt1) * cross(n2 (1 + sc1) + sh1)
    d_xa, d_gt1 = gate_bwd(d_img, c["xa"], c["gt1"])
    d_h2 = neighbour_attn_bwd(sd, g, key + ".cross_attn", d_xa, c["cross"], cfg)
    d_n2, d_sc1, d_sh1 = modulate_bwd(d_h2, c["n2"], c["sc1"])
    d_img = d_img + _ln_bwd(sd, g, key + ".norm2", d_n2, c["norm2"])
    # the six per-frame vectors come from modulation2(cn2)
    d_cn2 = d_cn2 + modulation_bwd(sd, g, key + ".modulation2",
                                   [d_sc1, d_sh1, d_gt1, d_sc2, d_sh2, d_gt2], c["mod2"])[:, :, 0]
    d_cam = d_cam + _ln_bwd(sd, g, key + ".cam_norm2", d_cn2, c["cam_norm2"])
    # img1 = img0 + (1 + gt) * ai ; cam1 = cam0 + ac ; (ai, ac) = attn(n1 (1 + sc) + sh, cn)
    d_ai, d_gt = gate_bwd(d_img, c["ai"], c["gt"])
    d_h, d_cn = video_attn_bwd(sd, g, key + ".attn", d_ai, d_cam, c["attn"], cfg)
    d_n1, d_sc, d_sh = modulate_bwd(d_h, c["n1"], c["sc"])
    d_img = d_img + _ln_bwd(sd, g, key + ".norm1", d_n1, c["norm1"])
    d_cn = d_cn + modulation_bwd(sd, g, key + ".modulation1", [d_sc, d_sh, d_gt], c["mod1"])[:, :, 0]
    d_cam = d_cam + _ln_bwd(sd, g, key + ".cam_norm1", d_cn, c["cam_norm1"])
    return d_img, d_cam, g
