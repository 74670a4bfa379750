"""CPU oracle for the VicaSplat encoder path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vicasplat_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs do, and only as the checker / baseline.

This is a *functional* fp32 (dtype-generic) torch restatement of ``VicaSplat.forward``: it takes a
plain ``state_dict`` with the reference's key names and evaluates the
network with ``torch.nn.functional`` calls.  Every function cites the reference lines it restates
(paths relative to /root/reference/src/model/encoder/).

PARITY PINNED: ``oracle/make_encoder_golden.py`` imports the real reference modules (in the build
container, where /root/reference exists), loads the same deterministic weights and commits
sub-sampled outputs under ``tests/golden/``; ``tests/test_oracle_encoder_cpu.py`` checks this
restatement against those vectors.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


@dataclass(frozen=True)
class EncoderConfig:
    """backbone/vica.yaml + experiment/re10k_8view.yaml values that shape the computation."""
    img_size: int = 256
    patch_size: int = 16
    enc_embed_dim: int = 1024
    enc_depth: int = 24
    enc_num_heads: int = 16
    dec_embed_dim: int = 768
    dec_depth: int = 12
    dec_num_heads: int = 12
    mlp_ratio: float = 4.0
    temporal_rope_theta: float = 30.0
    rope_base: float = 100.0
    sh_degree: int = 4
    ln_eps: float = 1e-6
    layer_dims: tuple = (96, 192, 384, 768)
    feature_dim: int = 256

    @property
    def d_sh(self) -> int:
        return (self.sh_degree + 1) ** 2

    @property
    def raw_gs_dim(self) -> int:     # vicasplat.py:73 (1 opacity + 7 + 3*d_sh)
        return 1 + 7 + 3 * self.d_sh

    @property
    def hooks(self) -> List[int]:    # heads/dpt_head.py:111
        l2 = self.dec_depth
        return [0, l2 * 2 // 4, l2 * 3 // 4, l2]


# --------------------------------------------------------------------------------- small pieces
def linear(sd: SD, key: str, x: Tensor) -> Tensor:
    return F.linear(x, sd[key + ".weight"], sd.get(key + ".bias"))


def layer_norm(sd: SD, key: str, x: Tensor, eps: float) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".weight"], sd[key + ".bias"], eps)


def mlp(sd: SD, key: str, x: Tensor) -> Tensor:
    """croco/blocks.py:58-79 (exact-erf GELU, dropout p=0)."""
    return linear(sd, key + ".fc2", F.gelu(linear(sd, key + ".fc1", x)))


def rope2d(t: Tensor, pos: Tensor, base: float) -> Tensor:
    """croco/pos_embed.py:112-159 == curope/kernels.cu:18-82.  t (B,H,N,D), pos (B,N,2) int (y,x).
    First half of D rotates with y, second half with x; inside a half, pairs are (d, d+D/4)."""
    D = t.shape[-1]
    Q = D // 4
    inv = base ** (-torch.arange(Q, dtype=torch.float32, device=t.device) / Q)   # (Q,)
    out = []
    for half, axis in ((t[..., : D // 2], 0), (t[..., D // 2:], 1)):
        ang = pos[..., axis].to(torch.float32)[:, None, :, None] * inv          # (B,1,N,Q)
        c, s = ang.cos().to(t.dtype), ang.sin().to(t.dtype)
        u, v = half[..., :Q], half[..., Q:]
        out += [u * c - v * s, v * c + u * s]
    return torch.cat(out, dim=-1)


def rope1d_interleaved(t: Tensor, frame: Tensor, theta: float) -> Tensor:
    """Temporal RoPE of the camera tokens: misc/rope_utils.py:133-137,297-305 (interleaved pairs
    (2i, 2i+1), frequency theta^(-2i/D), position = frame index).  t (B,H,T,D), frame (T,)."""
    D = t.shape[-1]
    inv = theta ** (-torch.arange(0, D, 2, dtype=torch.float32, device=t.device) / D)  # (D/2,)
    ang = frame.to(torch.float32)[:, None] * inv                                         # (T,D/2)
    c, s = ang.cos().to(t.dtype), ang.sin().to(t.dtype)
    a, b = t[..., 0::2], t[..., 1::2]
    return torch.stack([a * c - b * s, b * c + a * s], dim=-1).flatten(-2)


def sdpa(q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """softmax(q k^T / sqrt(d)) v with an optional boolean keep-mask (explicit, fp32-safe)."""
    s = (q @ k.transpose(-1, -2)) * (q.shape[-1] ** -0.5)
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    return s.softmax(dim=-1) @ v


def positions(frames: int, h: int, w: int, with_intrinsic: bool, device=None) -> Tensor:
    """croco/blocks.py:184-193 + backbone_vica.py:455-459: (y,x) per patch, row-major; the
    intrinsic token is appended at (h, 0)."""
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device),
                            indexing="ij")
    pos = torch.stack([ys, xs], dim=-1).reshape(1, h * w, 2)
    if with_intrinsic:
        pos = torch.cat([pos, torch.tensor([[[h, 0]]], device=device)], dim=1)
    return pos.expand(frames, -1, -1).contiguous()


# --------------------------------------------------------------------------------- ViT encoder
def enc_block(sd: SD, key: str, x: Tensor, pos: Tensor, cfg: EncoderConfig) -> Tensor:
    """croco/blocks.py:81-130."""
    Bf, N, C = x.shape
    H = cfg.enc_num_heads
    h = layer_norm(sd, key + ".norm1", x, cfg.ln_eps)
    qkv = linear(sd, key + ".attn.qkv", h).reshape(Bf, N, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = rope2d(qkv[0], pos, cfg.rope_base), rope2d(qkv[1], pos, cfg.rope_base), qkv[2]
    a = sdpa(q, k, v).transpose(1, 2).reshape(Bf, N, C)
    x = x + linear(sd, key + ".attn.proj", a)
    return x + mlp(sd, key + ".mlp", layer_norm(sd, key + ".norm2", x, cfg.ln_eps))


def encode_image(sd: SD, img: Tensor, K: Optional[Tensor], cfg: EncoderConfig):
    """backbone_vica.py:450-480,535-541.  img (F,3,H,W); K (F,3,3) or None -> (x (F,N,C), pos)."""
    P = cfg.patch_size
    x = F.conv2d(img, sd["backbone.patch_embed.proj.weight"], sd["backbone.patch_embed.proj.bias"],
                 stride=P)
    Fr, C, gh, gw = x.shape
    x = x.flatten(2).transpose(1, 2)
    if K is not None:
        tok = linear(sd, "backbone.intrinsic_encoder", K.flatten(1))[:, None]
        x = torch.cat([x, tok], dim=1)
    pos = positions(Fr, gh, gw, K is not None, img.device)
    for i in range(cfg.enc_depth):
        x = enc_block(sd, f"backbone.enc_blocks.{i}", x, pos, cfg)
    return layer_norm(sd, "backbone.enc_norm", x, cfg.ln_eps), pos


# --------------------------------------------------------------------------------- MixDecoder
def camera_mask(T: int, n_per_frame: int, device=None) -> Tensor:
    """backbone_vica.py:585-593 with first_token_full_attn=False: camera query t keeps every key
    (camera + image tokens) of frames <= t.  -> (T, T*(1+n_per_frame)) bool."""
    keep = torch.ones(T, T, dtype=torch.bool, device=device).tril()
    return keep[:, :, None].expand(T, T, 1 + n_per_frame).reshape(T, -1)


def modulation(sd: SD, key: str, emb: Tensor, n: int):
    """AdaLNModulation, backbone_vica.py:194-212."""
    return linear(sd, key + ".proj", F.silu(emb)).chunk(n, dim=-1)


def video_camera_attention(sd: SD, key: str, img: Tensor, cam: Tensor, pos: Tensor,
                           cfg: EncoderConfig) -> tuple:
    """backbone_vica.py:57-126.  img (B,T,N,C) modulated+normed, cam (B,T,C) normed."""
    B, T, N, C = img.shape
    H = cfg.dec_num_heads
    hd = C // H

    def heads(t, L):
        return t.reshape(B, L, 3, H, hd).permute(2, 0, 3, 1, 4)      # (3,B,H,L,hd)

    qi, ki, vi = heads(linear(sd, key + ".qkv", img), T * N)
    p = pos.reshape(B, T * N, 2)
    qi, ki = rope2d(qi, p, cfg.rope_base), rope2d(ki, p, cfg.rope_base)
    qc, kc, vc = heads(linear(sd, key + ".qkv", cam), T)
    fr = torch.arange(T, device=img.device)
    qc = rope1d_interleaved(qc, fr, cfg.temporal_rope_theta)
    kc = rope1d_interleaved(kc, fr, cfg.temporal_rope_theta)
    # keys / values per frame: [camera_t | image tokens of t]
    kf = torch.cat([kc[:, :, :, None], ki.reshape(B, H, T, N, hd)], dim=3).reshape(B, H, -1, hd)
    vf = torch.cat([vc[:, :, :, None], vi.reshape(B, H, T, N, hd)], dim=3).reshape(B, H, -1, hd)
    oi = sdpa(qi, kf, vf).transpose(1, 2).reshape(B, T, N, C)
    oc = sdpa(qc, kf, vf, camera_mask(T, N, img.device)).transpose(1, 2).reshape(B, T, C)
    return linear(sd, key + ".proj", oi), linear(sd, key + ".proj", oc)


def cross_neighbor_attention(sd: SD, key: str, img: Tensor, pos: Tensor, cfg: EncoderConfig):
    """backbone_vica.py:129-191: frame t attends to the image tokens of frames t-1 and t+1 (the
    end frames see their single neighbour twice, which softmax renders identical to once)."""
    B, T, N, C = img.shape
    H = cfg.dec_num_heads
    hd = C // H

    def proj(name):
        return linear(sd, f"{key}.{name}", img).reshape(B, T, N, H, hd).permute(0, 1, 3, 2, 4)

    q, k, v = proj("projq"), proj("projk"), proj("projv")            # (B,T,H,N,hd)
    p = pos.reshape(B * T, N, 2)
    q = rope2d(q.reshape(B * T, H, N, hd), p, cfg.rope_base).reshape(B, T, H, N, hd)
    k = rope2d(k.reshape(B * T, H, N, hd), p, cfg.rope_base).reshape(B, T, H, N, hd)
    outs = []
    for t in range(T):
        if T == 2:
            nb = [1 - t]
        else:
            nb = [t - 1 if t > 0 else 1, t + 1 if t < T - 1 else T - 2]
        kk = torch.cat([k[:, j] for j in nb], dim=2)
        vv = torch.cat([v[:, j] for j in nb], dim=2)
        outs.append(sdpa(q[:, t], kk, vv))                           # (B,H,N,hd)
    o = torch.stack(outs, dim=1).permute(0, 1, 3, 2, 4).reshape(B, T, N, C)
    return linear(sd, key + ".proj", o)


def dec_block(sd: SD, key: str, img: Tensor, cam: Tensor, pos: Tensor, cfg: EncoderConfig):
    """MixDecoderBlock.forward, backbone_vica.py:280-335."""
    eps = cfg.ln_eps
    cn = layer_norm(sd, key + ".cam_norm1", cam, eps)
    sc, sh, gt = modulation(sd, key + ".modulation1", cn[:, :, None], 3)
    h = layer_norm(sd, key + ".norm1", img, eps) * (1 + sc) + sh
    ai, ac = video_camera_attention(sd, key + ".attn", h, cn, pos, cfg)
    img = img + (1 + gt) * ai
    cam = cam + ac
    cn = layer_norm(sd, key + ".cam_norm2", cam, eps)
    sc1, sh1, gt1, sc2, sh2, gt2 = modulation(sd, key + ".modulation2", cn[:, :, None], 6)
    h = layer_norm(sd, key + ".norm2", img, eps) * (1 + sc1) + sh1
    img = img + (1 + gt1) * cross_neighbor_attention(sd, key + ".cross_attn", h, pos, cfg)
    h = layer_norm(sd, key + ".norm3", img, eps) * (1 + sc2) + sh2
    img = img + (1 + gt2) * mlp(sd, key + ".mlp", h)
    cam = cam + mlp(sd, key + ".mlp_cam", cn)
    return img, cam


def backbone(sd: SD, image: Tensor, intrinsics: Optional[Tensor], cfg: EncoderConfig):
    """VicaNet.forward, backbone_vica.py:526-583.  image (B,T,3,H,W) in [-1,1].
    Returns (intermediates: list of (B,T,N_patches,C), camera (B,T,C) after camera_dec_norm)."""
    B, T = image.shape[:2]
    K = intrinsics.reshape(B * T, 3, 3) if intrinsics is not None else None
    x, pos = encode_image(sd, image.flatten(0, 1), K, cfg)
    N = x.shape[1]
    x = x.reshape(B, T, N, -1)
    pos = pos.reshape(B, T, N, 2)
    inter = [x]
    img = linear(sd, "backbone.decoder_embed", x)
    it, et = sd["backbone.camera_intrinsic_token"], sd["backbone.camera_extrinsic_token"]
    cam = torch.stack([it] + [it + et] * (T - 1), dim=0)[None].expand(B, T, -1)
    for i in range(cfg.dec_depth):
        img, cam = dec_block(sd, f"backbone.dec_blocks.{i}", img, cam, pos, cfg)
        inter.append(img)
    inter[-1] = layer_norm(sd, "backbone.dec_norm", inter[-1], cfg.ln_eps)
    cam = layer_norm(sd, "backbone.camera_dec_norm", cam, cfg.ln_eps)
    if intrinsics is not None:
        inter = [t[:, :, :-1] for t in inter]
    return inter, cam


# --------------------------------------------------------------------------------- DPT heads
# Every ReLU of the plugin goes through relu(x, tag) (tag = the state_dict key of the layer it feeds or
# follows).  Tests may install RELU_HOOK(x, tag) -> Tensor, e.g. a ReLU with a PRESCRIBED mask: the
# gradient of a ReLU network is discontinuous in its activations, so comparing the gradients of two
# forward passes that differ by rounding is only sharp when both use the same masks.
RELU_HOOK = None


def relu(x: Tensor, tag: str) -> Tensor:
    return F.relu(x) if RELU_HOOK is None else RELU_HOOK(x, tag)


def conv(sd: SD, key: str, x: Tensor, stride=1, padding=0) -> Tensor:
    return F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias"), stride=stride, padding=padding)


def up2(x: Tensor) -> Tensor:
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


def rcu(sd: SD, key: str, x: Tensor) -> Tensor:
    """ResidualConvUnit_custom (pre-activation), heads/dpt_block.py:79-137."""
    y = conv(sd, key + ".conv1", relu(x, key + ".conv1"), padding=1)
    y = conv(sd, key + ".conv2", relu(y, key + ".conv2"), padding=1)
    return y + x


def fusion(sd: SD, key: str, x: Tensor, skip: Optional[Tensor]) -> Tensor:
    """FeatureFusionBlock_custom, heads/dpt_block.py:139-229 (width_ratio 1)."""
    if skip is not None:
        x = x + rcu(sd, key + ".resConfUnit1", skip)
    x = up2(rcu(sd, key + ".resConfUnit2", x))
    return conv(sd, key + ".out_conv", x)


def dpt_trunk(sd: SD, key: str, inter: List[Tensor], gh: int, gw: int, cfg: EncoderConfig):
    """heads/dpt_head.py:35-68 up to path_1.  inter: 13 tensors (F,N,C)."""
    layers = []
    for idx, hook in enumerate(cfg.hooks):
        t = inter[hook]
        t = t.transpose(1, 2).reshape(t.shape[0], t.shape[2], gh, gw)
        a = f"{key}.act_postprocess.{idx}"
        t = conv(sd, a + ".0", t)
        if idx == 0:
            t = F.conv_transpose2d(t, sd[a + ".1.weight"], sd[a + ".1.bias"], stride=4)
        elif idx == 1:
            t = F.conv_transpose2d(t, sd[a + ".1.weight"], sd[a + ".1.bias"], stride=2)
        elif idx == 3:
            t = conv(sd, a + ".1", t, stride=2, padding=1)
        layers.append(conv(sd, f"{key}.scratch.layer_rn.{idx}", t, padding=1))
    p4 = fusion(sd, key + ".scratch.refinenet4", layers[3], None)
    p4 = p4[:, :, : layers[2].shape[2], : layers[2].shape[3]]
    p3 = fusion(sd, key + ".scratch.refinenet3", p4, layers[2])
    p2 = fusion(sd, key + ".scratch.refinenet2", p3, layers[1])
    return fusion(sd, key + ".scratch.refinenet1", p2, layers[0])


def pts_head(sd: SD, inter, gh, gw, cfg: EncoderConfig) -> Tensor:
    """'regression' head + 'exp' postprocess: dpt_block.py:316-324, postprocess.py:10-61.
    -> (F,H,W,3)."""
    k = "downstream_head1.dpt"
    x = dpt_trunk(sd, k, inter, gh, gw, cfg)
    x = conv(sd, k + ".head.0", x, padding=1)
    x = relu(conv(sd, k + ".head.2", up2(x), padding=1), k + ".head.2")
    xyz = conv(sd, k + ".head.4", x).permute(0, 2, 3, 1)[..., :3]
    d = xyz.norm(dim=-1, keepdim=True)
    return xyz / d.clip(min=1e-8) * torch.expm1(d)


def gs_head(sd: SD, inter, imgs: Tensor, gh, gw, cfg: EncoderConfig) -> Tensor:
    """'gs_params' head: dpt_gs_head.py:113-157, dpt_block.py:335-343 (eval: dropout off).
    imgs (F,3,H,W) normalised.  -> (F,H,W,83)."""
    k = "gaussian_param_head.dpt"
    x = up2(dpt_trunk(sd, k, inter, gh, gw, cfg))
    x = x + relu(conv(sd, k + ".input_merger.0", imgs, padding=3), k + ".input_merger.0")
    x = relu(conv(sd, k + ".head.0", x, padding=1), k + ".head.0")
    return conv(sd, k + ".head.4", x).permute(0, 2, 3, 1)


# --------------------------------------------------------------------------------- pose + adapter
def quat_mul_xyzw(a: Tensor, b: Tensor) -> Tensor:
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw,
                        aw * bw - ax * bx - ay * by - az * bz], dim=-1)


def quat_to_matrix_xyzw(q: Tensor) -> Tensor:
    x, y, z, w = q.unbind(-1)
    return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
                       dim=-1).reshape(q.shape[:-1] + (3, 3))


def camera_head(sd: SD, cam: Tensor):
    """vicasplat.py:179-199 + misc/cam_utils.py:203-207 + misc/dq.py:224-262.
    cam (B,T,C) -> pred_extrins (B,T-1,8) unit dual quaternion, c2w (B,T,4,4), frame 0 identity."""
    B, T, _ = cam.shape
    dq = linear(sd, "camera_extrinsic_head.1", relu(cam[:, 1:], "camera_extrinsic_head.1")).clone()
    dq[..., 3] = dq[..., 3] + 1.0
    dq = dq / dq[..., :4].norm(dim=-1, keepdim=True)
    qr, qd = dq[..., :4], dq[..., 4:]
    conj = qr * torch.tensor([-1.0, -1.0, -1.0, 1.0], dtype=qr.dtype, device=qr.device)
    t = quat_mul_xyzw(2.0 * qd, conj)[..., :3]
    M = torch.zeros(B, T, 4, 4, dtype=cam.dtype, device=cam.device)
    M[:, :, 3, 3] = 1.0
    M[:, 0, :3, :3] = torch.eye(3, dtype=cam.dtype, device=cam.device)
    M[:, 1:, :3, :3] = quat_to_matrix_xyzw(qr)
    M[:, 1:, :3, 3] = t
    return dq, M


def sh_mask(cfg: EncoderConfig, device=None) -> Tensor:
    """common/gaussian_adapter.py:44-50."""
    m = torch.ones(cfg.d_sh, dtype=torch.float32, device=device)
    for deg in range(1, cfg.sh_degree + 1):
        m[deg * deg:(deg + 1) ** 2] = 0.1 * 0.25 ** deg
    return m


def gaussian_adapter(raw: Tensor, cfg: EncoderConfig) -> dict:
    """MyGaussianAdapter.forward (softplus scale act, identity opacity mapping):
    common/gaussian_adapter.py:167-212, common/gaussians.py:8-44.  raw (...,86)."""
    xyz, op, sc, rot = raw[..., :11].split((3, 1, 3, 4), dim=-1)
    sh = raw[..., 11:].reshape(raw.shape[:-1] + (3, cfg.d_sh)) * sh_mask(cfg, raw.device).to(raw.dtype)
    op = torch.sigmoid(op)
    sc = (0.001 * F.softplus(sc)).clamp_max(0.3)
    rot = F.normalize(rot, dim=-1)
    i, j, k, r = rot.unbind(-1)
    two_s = 2 / ((rot * rot).sum(-1) + 1e-8)
    R = torch.stack([1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)],
                    dim=-1).reshape(rot.shape[:-1] + (3, 3))
    S = torch.diag_embed(sc)
    cov = R @ S @ S.transpose(-1, -2) @ R.transpose(-1, -2)
    return dict(means=xyz, covariances=cov, harmonics=sh, opacities=op, scales=sc, rotations=rot)


# --------------------------------------------------------------------------------- whole forward
def forward(sd: SD, image: Tensor, intrinsics: Optional[Tensor], cfg: EncoderConfig,
            stages: bool = False) -> dict:
    """VicaSplat.forward (vicasplat.py:158-278) without the GT-extrinsics depth branch."""
    B, T, _, H, W = image.shape
    gh, gw = H // cfg.patch_size, W // cfg.patch_size
    inter, cam = backbone(sd, image, intrinsics, cfg)
    pred, c2w = camera_head(sd, cam)
    flat = [t.flatten(0, 1) for t in inter]
    centers = pts_head(sd, flat, gh, gw, cfg)
    params = gs_head(sd, flat, image.flatten(0, 1), gh, gw, cfg)
    raw = torch.cat([centers, params], dim=-1).reshape(B, T, H, W, -1)
    out = dict(pred_extrins=pred, gaussian_camera_extrins=c2w, raw_gaussians=raw,
               gaussian_centers=raw[..., :3], gaussians=gaussian_adapter(raw, cfg))
    if stages:
        out["intermediates"] = inter
        out["camera_tokens"] = cam
    return out


# --------------------------------------------------------------------------------- weights
def param_shapes(cfg: EncoderConfig, use_intrinsic_embedding: bool = True) -> Dict[str, tuple]:
    """The reference's state_dict contract, key -> shape, in a fixed order.
    The aliased ``scratch.layerK_rn`` names are listed after their ``layer_rn.K-1`` twins."""
    E, D, P = cfg.enc_embed_dim, cfg.dec_embed_dim, cfg.patch_size
    hid_e, hid_d = int(E * cfg.mlp_ratio), int(D * cfg.mlp_ratio)
    s: Dict[str, tuple] = {}

    def lin(k, o, i, bias=True):
        s[k + ".weight"] = (o, i)
        if bias:
            s[k + ".bias"] = (o,)

    def ln(k, c):
        s[k + ".weight"] = (c,)
        s[k + ".bias"] = (c,)

    def cv(k, o, i, kh, bias=True):
        s[k + ".weight"] = (o, i, kh, kh)
        if bias:
            s[k + ".bias"] = (o,)

    s["backbone.camera_extrinsic_token"] = (D,)
    s["backbone.camera_intrinsic_token"] = (D,)
    cv("backbone.patch_embed.proj", E, 3, P)
    for i in range(cfg.enc_depth):
        k = f"backbone.enc_blocks.{i}"
        ln(k + ".norm1", E); lin(k + ".attn.qkv", 3 * E, E); lin(k + ".attn.proj", E, E)
        ln(k + ".norm2", E); lin(k + ".mlp.fc1", hid_e, E); lin(k + ".mlp.fc2", E, hid_e)
    ln("backbone.enc_norm", E)
    lin("backbone.decoder_embed", D, E)
    for i in range(cfg.dec_depth):
        k = f"backbone.dec_blocks.{i}"
        ln(k + ".cam_norm1", D); lin(k + ".modulation1.proj", 3 * D, D); ln(k + ".norm1", D)
        lin(k + ".attn.qkv", 3 * D, D); lin(k + ".attn.proj", D, D)
        ln(k + ".cam_norm2", D); lin(k + ".modulation2.proj", 6 * D, D); ln(k + ".norm2", D)
        for n in ("projq", "projk", "projv", "proj"):
            lin(f"{k}.cross_attn.{n}", D, D)
        ln(k + ".norm3", D)
        lin(k + ".mlp.fc1", hid_d, D); lin(k + ".mlp.fc2", D, hid_d)
        lin(k + ".mlp_cam.fc1", hid_d, D); lin(k + ".mlp_cam.fc2", D, hid_d)
    ln("backbone.dec_norm", D)
    ln("backbone.camera_dec_norm", D)
    if use_intrinsic_embedding:
        lin("backbone.intrinsic_encoder", E, 9)
    Fd = cfg.feature_dim
    for head in ("downstream_head1", "gaussian_param_head"):
        k = head + ".dpt"
        for idx in range(4):
            s[f"{k}.scratch.layer{idx + 1}_rn.weight"] = (Fd, cfg.layer_dims[idx], 3, 3)
        for idx in range(4):
            s[f"{k}.scratch.layer_rn.{idx}.weight"] = (Fd, cfg.layer_dims[idx], 3, 3)
        for r in (1, 2, 3, 4):
            rk = f"{k}.scratch.refinenet{r}"
            cv(rk + ".out_conv", Fd, Fd, 1)
            for u in ("resConfUnit1", "resConfUnit2"):
                cv(f"{rk}.{u}.conv1", Fd, Fd, 3)
                cv(f"{rk}.{u}.conv2", Fd, Fd, 3)
        if head == "downstream_head1":
            cv(k + ".head.0", Fd // 2, Fd, 3); cv(k + ".head.2", Fd // 2, Fd // 2, 3)
            cv(k + ".head.4", 3, Fd // 2, 1)
        else:
            cv(k + ".head.0", Fd, Fd, 3, bias=False); cv(k + ".head.4", cfg.raw_gs_dim, Fd, 1)
        dims = [E, D, D, D]
        for idx in range(4):
            cv(f"{k}.act_postprocess.{idx}.0", cfg.layer_dims[idx], dims[idx], 1)
        c0, c1, c3 = cfg.layer_dims[0], cfg.layer_dims[1], cfg.layer_dims[3]
        s[f"{k}.act_postprocess.0.1.weight"] = (c0, c0, 4, 4); s[f"{k}.act_postprocess.0.1.bias"] = (c0,)
        s[f"{k}.act_postprocess.1.1.weight"] = (c1, c1, 2, 2); s[f"{k}.act_postprocess.1.1.bias"] = (c1,)
        cv(f"{k}.act_postprocess.3.1", c3, c3, 3)
        if head == "gaussian_param_head":
            cv(k + ".input_merger.0", Fd, 3, 7)
    lin("camera_extrinsic_head.1", 8, D)
    return s


def synth_state_dict(cfg: EncoderConfig, seed: int = 0, dtype=torch.float32) -> SD:
    """Deterministic random weights for tests / bench (there is no checkpoint to download).
    Matrix-like tensors ~ N(0, 1/fan_in) scaled so activations stay O(1); LayerNorm weights ~ 1,
    biases small; the reference's zero-initialised layers (AdaLN projections, camera head) get
    N(0, 0.02) so modulation, gating and the pose path are exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for k, shp in param_shapes(cfg).items():
        if ".scratch.layer" in k and "_rn.weight" in k and ".layer_rn." not in k:
            continue  # alias, filled below
        if len(shp) == 1:
            if k.endswith("weight"):                       # LayerNorm gain
                t = 1.0 + 0.1 * torch.randn(shp, generator=g)
            elif "token" in k:
                t = 0.02 * torch.randn(shp, generator=g)
            else:
                t = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = math.prod(shp[1:])
            if "modulation" in k or k.startswith("camera_extrinsic_head"):
                std = 0.02
            elif "act_postprocess.0.1" in k or "act_postprocess.1.1" in k:
                std = 1.0 / math.sqrt(shp[0])              # ConvTranspose: one tap per output
            else:
                std = 1.0 / math.sqrt(fan_in)
            t = std * torch.randn(shp, generator=g)
        sd[k] = t.to(dtype)
    for head in ("downstream_head1", "gaussian_param_head"):
        for idx in range(4):
            sd[f"{head}.dpt.scratch.layer{idx + 1}_rn.weight"] = \
                sd[f"{head}.dpt.scratch.layer_rn.{idx}.weight"]
    return sd
