"""CPU oracle for the Gaussian-splatting rasterizer.

TEST INFRASTRUCTURE ONLY.  Nothing under ``vicasplat_b200/`` may import this
file; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and only as the checker /
baseline.

PARITY UNPINNED.  The reference calls the third-party CUDA extension
``diff_gaussian_rasterization`` (pip ``git+https://github.com/rmurai0610/
diff-gaussian-rasterization-w-pose.git``, no commit pinned,
/root/reference/requirements.txt:17).  Its source is not under
/root/reference and the reference ships no golden vectors for it, so this file
restates the *published* 3D-Gaussian-splatting tile rasterizer (EWA projection,
16x16 tile binning, depth sort, front-to-back alpha compositing) and anchors
every calling convention on the reference's own call site:

* argument layout, transposed matrices, cov triu order, SH layout, degree rule:
  /root/reference/src/model/decoder/cuda_splatting.py:148-239
* projection matrix: cuda_splatting.py:18-45;  fov: src/geometry/projection.py:247-261
* flatten / per-view repeat: src/model/decoder/decoder_splatting_cuda.py:38-101
* pose deltas (theta, rho) as a left twist on W2C: src/misc/cam_utils.py:108-142

Every behavioural constant that could not be confirmed against the real
package is a named field of :class:`RasterConstants`.

The implementation is vectorised torch, dtype-generic (fp64 in the tests) and
differentiable end to end, so ``torch.autograd`` supplies the oracle gradients
for means / cov6 / SH / opacity / theta / rho.
"""
from __future__ import annotations

from dataclasses import dataclass
from math import isqrt
from typing import Optional

import torch

TILE = 16

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
         -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658,
         0.3731763325901154, -0.4570457994644658, 1.445305721320277,
         -0.5900435899266435)


@dataclass(frozen=True)
class RasterConstants:
    """Named switches for upstream behaviour that is recalled, not verified."""
    near_cull_z: float = 0.2          # view-space z <= this is culled, regardless of `near`
    lowpass: float = 0.3              # px^2 added to the 2-D covariance diagonal
    fov_clamp: float = 1.3            # tx/tz clamped to +-1.3*tanfov before the Jacobian
    radius_sigmas: float = 3.0        # radius = ceil(3*sqrt(lambda_max))
    lambda_floor: float = 0.1         # max(0.1, mid^2-det) inside the eigenvalue sqrt
    alpha_max: float = 0.99
    alpha_min: float = 1.0 / 255.0
    t_stop: float = 1e-4
    max_sh_band: int = 3              # bands above 3 are read past but not evaluated (H1)
    w_eps: float = 1e-7               # p_hom.w + eps before the perspective divide
    n_touched_t: float = 0.5          # a pixel counts toward n_touched while T > this


DEFAULT = RasterConstants()


# --------------------------------------------------------------------------
# camera helpers (restating cuda_splatting.py:18-45 and projection.py:247-261)
# --------------------------------------------------------------------------
def get_fov(intrinsics: torch.Tensor) -> torch.Tensor:
    inv = torch.linalg.inv(intrinsics)

    def ray(v):
        v = torch.tensor(v, dtype=intrinsics.dtype, device=intrinsics.device)
        d = torch.einsum("bij,j->bi", inv, v)
        return d / d.norm(dim=-1, keepdim=True)

    fov_x = (ray([0, 0.5, 1]) * ray([1, 0.5, 1])).sum(-1).acos()
    fov_y = (ray([0.5, 0, 1]) * ray([0.5, 1, 1])).sum(-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near, far, fov_x, fov_y):
    tx = (0.5 * fov_x).tan()
    ty = (0.5 * fov_y).tan()
    b = near.shape[0]
    P = torch.zeros((b, 4, 4), dtype=near.dtype, device=near.device)
    P[:, 0, 0] = 1.0 / tx          # 2n / (r-l) with r = -l = tx*n
    P[:, 1, 1] = 1.0 / ty
    P[:, 3, 2] = 1.0
    P[:, 2, 2] = far / (far - near)
    P[:, 2, 3] = -(far * near) / (far - near)
    return P


def se3_exp(rho: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """exp of the twist (rho, theta) as a 4x4, differentiable at 0 (cam_utils.py:108-121)."""
    z = torch.zeros((), dtype=rho.dtype, device=rho.device)
    A = torch.stack([
        torch.stack([z, -theta[2], theta[1], rho[0]]),
        torch.stack([theta[2], z, -theta[0], rho[1]]),
        torch.stack([-theta[1], theta[0], z, rho[2]]),
        torch.stack([z, z, z, z]),
    ])
    return torch.linalg.matrix_exp(A)


# --------------------------------------------------------------------------
# per-Gaussian preprocess
# --------------------------------------------------------------------------
def eval_sh(shs: torch.Tensor, dirs: torch.Tensor, degree: int, k: RasterConstants):
    """shs (G, M, 3), dirs (G,3) unit.  Returns un-clamped rgb+0.5 (G,3)."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    deg = min(degree, k.max_sh_band)
    res = SH_C0 * shs[:, 0]
    if deg > 0:
        res = res - SH_C1 * y * shs[:, 1] + SH_C1 * z * shs[:, 2] - SH_C1 * x * shs[:, 3]
    if deg > 1:
        xx, yy, zz = x * x, y * y, z * z
        xy, yz, xz = x * y, y * z, x * z
        res = (res + SH_C2[0] * xy * shs[:, 4] + SH_C2[1] * yz * shs[:, 5]
               + SH_C2[2] * (2.0 * zz - xx - yy) * shs[:, 6]
               + SH_C2[3] * xz * shs[:, 7] + SH_C2[4] * (xx - yy) * shs[:, 8])
        if deg > 2:
            res = (res + SH_C3[0] * y * (3.0 * xx - yy) * shs[:, 9]
                   + SH_C3[1] * xy * z * shs[:, 10]
                   + SH_C3[2] * y * (4.0 * zz - xx - yy) * shs[:, 11]
                   + SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * shs[:, 12]
                   + SH_C3[4] * x * (4.0 * zz - xx - yy) * shs[:, 13]
                   + SH_C3[5] * z * (xx - yy) * shs[:, 14]
                   + SH_C3[6] * x * (xx - 3.0 * yy) * shs[:, 15])
    return res + 0.5


def preprocess(means3D, cov6, shs, colors_precomp, opacities, W2C, P, campos,
               tanfovx, tanfovy, H, W, sh_degree, k: RasterConstants = DEFAULT):
    """Everything the per-Gaussian stage produces.  W2C, P are plain (not transposed) 4x4."""
    G = means3D.shape[0]
    dt = means3D.dtype
    ones = torch.ones((G, 1), dtype=dt, device=means3D.device)
    ph = torch.cat([means3D, ones], dim=1)
    t = (ph @ W2C.T)[:, :3]                          # view-space point
    full = P @ W2C
    hom = ph @ full.T
    pw = 1.0 / (hom[:, 3] + k.w_eps)
    ndc_x, ndc_y = hom[:, 0] * pw, hom[:, 1] * pw
    px = ((ndc_x + 1.0) * W - 1.0) * 0.5
    py = ((ndc_y + 1.0) * H - 1.0) * 0.5
    depth = t[:, 2]
    valid = depth > k.near_cull_z

    # EWA 2-D covariance
    fx = W / (2.0 * tanfovx)
    fy = H / (2.0 * tanfovy)
    tz = torch.where(valid, t[:, 2], torch.ones_like(t[:, 2]))
    limx, limy = k.fov_clamp * tanfovx, k.fov_clamp * tanfovy
    txc = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    tyc = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    zero = torch.zeros_like(tz)
    J = torch.stack([
        torch.stack([fx / tz, zero, -fx * txc / (tz * tz)], dim=-1),
        torch.stack([zero, fy / tz, -fy * tyc / (tz * tz)], dim=-1),
    ], dim=1)                                         # (G,2,3)
    S = torch.stack([
        torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2]], dim=-1),
        torch.stack([cov6[:, 1], cov6[:, 3], cov6[:, 4]], dim=-1),
        torch.stack([cov6[:, 2], cov6[:, 4], cov6[:, 5]], dim=-1),
    ], dim=1)                                         # (G,3,3)
    M = J @ W2C[:3, :3]
    cov2 = M @ S @ M.transpose(1, 2)
    a = cov2[:, 0, 0] + k.lowpass
    b = cov2[:, 0, 1]
    c = cov2[:, 1, 1] + k.lowpass
    det = a * c - b * b
    valid = valid & (det != 0)
    det_s = torch.where(det != 0, det, torch.ones_like(det))
    conic = torch.stack([c / det_s, -b / det_s, a / det_s], dim=-1)
    mid = 0.5 * (a + c)
    lam = mid + torch.sqrt(torch.clamp(mid * mid - det, min=k.lambda_floor))
    radius = torch.ceil(k.radius_sigmas * torch.sqrt(lam.detach()))
    radius = torch.where(valid, radius, torch.zeros_like(radius))

    gx = (W + TILE - 1) // TILE
    gy = (H + TILE - 1) // TILE
    pxd, pyd = px.detach(), py.detach()

    def lo(p, g):
        return torch.clamp(torch.trunc((p - radius) / TILE), 0, g)

    def hi(p, g):
        return torch.clamp(torch.trunc((p + radius + TILE - 1) / TILE), 0, g)

    rminx, rmaxx = lo(pxd, gx), hi(pxd, gx)
    rminy, rmaxy = lo(pyd, gy), hi(pyd, gy)
    tiles = (rmaxx - rminx) * (rmaxy - rminy)
    valid = valid & (tiles > 0)
    radius = torch.where(valid, radius, torch.zeros_like(radius))

    if colors_precomp is None:
        d = means3D - campos[None]
        d = d / d.norm(dim=-1, keepdim=True)
        rgb = torch.clamp(eval_sh(shs, d, sh_degree, k), min=0.0)
    else:
        rgb = colors_precomp

    return dict(valid=valid, px=px, py=py, depth=depth, conic=conic, rgb=rgb,
                opacity=opacities.reshape(G), radius=radius.to(torch.int32),
                rect=(rminx.long(), rmaxx.long(), rminy.long(), rmaxy.long()),
                tiles=torch.where(valid, tiles, torch.zeros_like(tiles)).long(),
                grid=(gx, gy))


# --------------------------------------------------------------------------
# tile compositing
# --------------------------------------------------------------------------
def composite(pre, H, W, bg, k: RasterConstants = DEFAULT, n_gauss: Optional[int] = None):
    dt = pre["px"].dtype
    dev = pre["px"].device
    gx, gy = pre["grid"]
    rminx, rmaxx, rminy, rmaxy = pre["rect"]
    valid = pre["valid"]
    G = valid.shape[0]
    color = torch.zeros((3, H, W), dtype=dt, device=dev)
    depth = torch.zeros((1, H, W), dtype=dt, device=dev)
    opac = torch.zeros((1, H, W), dtype=dt, device=dev)
    n_touched = torch.zeros((G,), dtype=torch.int32, device=dev)
    color = color + bg.reshape(3, 1, 1).to(dt)      # empty tiles show the background
    out_c, out_d, out_o = [], [], []
    depth_key = pre["depth"].detach()
    for ty in range(gy):
        for tx in range(gx):
            m = valid & (rminx <= tx) & (tx < rmaxx) & (rminy <= ty) & (ty < rmaxy)
            idx = torch.nonzero(m, as_tuple=False).flatten()
            y0, x0 = ty * TILE, tx * TILE
            y1, x1 = min(y0 + TILE, H), min(x0 + TILE, W)
            if idx.numel() == 0:
                continue
            # stable sort by depth keeps Gaussian-index order for ties
            order = torch.sort(depth_key[idx].float(), stable=True).indices
            idx = idx[order]
            ys = torch.arange(y0, y1, dtype=dt, device=dev)
            xs = torch.arange(x0, x1, dtype=dt, device=dev)
            pyy, pxx = torch.meshgrid(ys, xs, indexing="ij")
            pxx, pyy = pxx.reshape(-1, 1), pyy.reshape(-1, 1)         # (P,1)
            dx = pre["px"][idx][None] - pxx                            # (P,n)
            dy = pre["py"][idx][None] - pyy
            con = pre["conic"][idx]
            power = (-0.5 * (con[:, 0][None] * dx * dx + con[:, 2][None] * dy * dy)
                     - con[:, 1][None] * dx * dy)
            alpha = torch.clamp(pre["opacity"][idx][None] * torch.exp(torch.clamp(power, max=0.0)),
                                max=k.alpha_max)
            live = (power <= 0) & (alpha >= k.alpha_min)
            alpha = torch.where(live, alpha, torch.zeros_like(alpha))
            t_after = torch.cumprod(1.0 - alpha, dim=1)
            stop = (t_after.detach() < k.t_stop)
            keep = torch.cumsum(stop.to(torch.int32), dim=1) == 0
            alpha = torch.where(keep, alpha, torch.zeros_like(alpha))
            t_after = torch.cumprod(1.0 - alpha, dim=1)
            t_before = torch.cat([torch.ones_like(t_after[:, :1]), t_after[:, :-1]], dim=1)
            wgt = alpha * t_before
            t_fin = t_after[:, -1]
            c = wgt @ pre["rgb"][idx] + t_fin[:, None] * bg[None].to(dt)   # (P,3)
            d = wgt @ pre["depth"][idx]
            o = wgt.sum(dim=1)
            hh, ww = y1 - y0, x1 - x0
            out_c.append((y0, y1, x0, x1, c.T.reshape(3, hh, ww)))
            out_d.append(d.reshape(1, hh, ww))
            out_o.append(o.reshape(1, hh, ww))
            touched = ((wgt.detach() > 0) & (t_before.detach() > k.n_touched_t)).sum(dim=0)
            n_touched.index_add_(0, idx, touched.to(torch.int32))
    # assemble without in-place writes on a graph leaf
    if out_c:
        color_parts = color.clone()
        depth_parts = depth.clone()
        opac_parts = opac.clone()
        for (y0, y1, x0, x1, c), d, o in zip(out_c, out_d, out_o):
            color_parts[:, y0:y1, x0:x1] = c
            depth_parts[:, y0:y1, x0:x1] = d
            opac_parts[:, y0:y1, x0:x1] = o
        color, depth, opac = color_parts, depth_parts, opac_parts
    return color, depth, opac, n_touched


def rasterize_view(means3D, cov6, shs, colors_precomp, opacities, c2w, K_norm_fov,
                   near, far, H, W, bg, sh_degree, theta=None, rho=None,
                   k: RasterConstants = DEFAULT):
    """One view.  ``K_norm_fov`` = (tanfovx, tanfovy) python floats.

    Returns (image (3,H,W), radii (G,), depth (1,H,W), opacity (1,H,W), n_touched (G,)),
    the 5-tuple the reference unpacks at cuda_splatting.py:226.
    """
    tanfovx, tanfovy = K_norm_fov
    dt = means3D.dtype
    W2C = torch.linalg.inv(c2w.to(dt))
    if theta is not None or rho is not None:
        z3 = torch.zeros(3, dtype=dt, device=means3D.device)
        W2C = se3_exp(rho if rho is not None else z3, theta if theta is not None else z3) @ W2C
    campos = -(W2C[:3, :3].T @ W2C[:3, 3])
    nf = torch.tensor([1.0], dtype=dt)
    P = torch.zeros((4, 4), dtype=dt, device=means3D.device)
    P[0, 0] = 1.0 / tanfovx
    P[1, 1] = 1.0 / tanfovy
    P[3, 2] = 1.0
    P[2, 2] = far / (far - near)
    P[2, 3] = -(far * near) / (far - near)
    del nf
    pre = preprocess(means3D, cov6, shs, colors_precomp, opacities, W2C, P, campos,
                     tanfovx, tanfovy, H, W, sh_degree, k)
    color, depth, opac, n_touched = composite(pre, H, W, bg.to(dt), k)
    return color, pre["radius"], depth, opac, n_touched


def render_cuda_ref(extrinsics, intrinsics, near, far, image_shape, background_color,
                    gaussian_means, gaussian_covariances, gaussian_sh_coefficients,
                    gaussian_opacities, cam_rot_delta=None, cam_trans_delta=None,
                    use_sh=True, sh_degree=None, k: RasterConstants = DEFAULT):
    """Restates ``render_cuda`` (cuda_splatting.py:148-239), per-view loop included.

    extrinsics (B,4,4) c2w; intrinsics (B,3,3) normalised; Gaussians (B,G,...) or shared (G,...).
    Returns (color (B,3,H,W), depth (B,H,W)).
    """
    shared = gaussian_means.ndim == 2
    n = gaussian_sh_coefficients.shape[-1]
    degree = sh_degree or isqrt(n) - 1
    shs = gaussian_sh_coefficients.transpose(-1, -2)          # (..., n, xyz)
    b = extrinsics.shape[0]
    h, w = image_shape
    fov = get_fov(intrinsics)
    tfx, tfy = (0.5 * fov[:, 0]).tan(), (0.5 * fov[:, 1]).tan()
    iu = torch.triu_indices(3, 3)
    imgs, deps = [], []
    for i in range(b):
        g = (lambda t: t) if shared else (lambda t: t[i])
        cov = g(gaussian_covariances)[:, iu[0], iu[1]]
        img, _, dep, _, _ = rasterize_view(
            g(gaussian_means), cov, g(shs) if use_sh else None,
            None if use_sh else g(shs)[:, 0, :], g(gaussian_opacities)[..., None],
            extrinsics[i], (float(tfx[i]), float(tfy[i])), float(near[i]), float(far[i]),
            h, w, background_color[i], degree,
            theta=cam_rot_delta[i] if cam_rot_delta is not None else None,
            rho=cam_trans_delta[i] if cam_trans_delta is not None else None, k=k)
        imgs.append(img)
        deps.append(dep[0])
    return torch.stack(imgs), torch.stack(deps)


# --------------------------------------------------------------------------
# synthetic scene generator: pixel-aligned Gaussians on a line of cameras
# --------------------------------------------------------------------------
def synthetic_scene(n_ctx, h, w, n_tgt, seed=250307, d_sh=25, dtype=torch.float32,
                    depth_range=(1.0, 20.0), focal=0.86, n_gauss: Optional[int] = None):
    g = torch.Generator().manual_seed(seed)
    K = torch.tensor([[focal, 0, 0.5], [0, focal, 0.5], [0, 0, 1]], dtype=torch.float64)

    def cam(x):
        c = torch.eye(4, dtype=torch.float64)
        c[0, 3] = x
        return c

    ctx = torch.stack([cam(i / max(n_ctx - 1, 1)) for i in range(n_ctx)])
    tgt = torch.stack([cam((i + 0.5) / n_tgt) for i in range(n_tgt)])
    ys, xs = torch.meshgrid((torch.arange(h, dtype=torch.float64) + 0.5) / h,
                            (torch.arange(w, dtype=torch.float64) + 0.5) / w, indexing="ij")
    pix = torch.stack([xs, ys, torch.ones_like(xs)], dim=-1).reshape(-1, 3)      # (hw,3)
    rays = pix @ torch.linalg.inv(K).T
    G_per = h * w
    lo, hi = depth_range
    u = torch.rand((n_ctx, G_per), generator=g, dtype=torch.float64)
    z = lo * (hi / lo) ** u
    pts = rays[None] * z[..., None]
    pts = pts + ctx[:, None, :3, 3]
    px_world = z / (focal * w)                                    # one pixel at that depth
    sig = px_world * (0.5 + torch.rand((n_ctx, G_per), generator=g, dtype=torch.float64))
    aniso = 1.0 + 2.0 * torch.rand((n_ctx, G_per, 3), generator=g, dtype=torch.float64)
    scales = sig[..., None] * aniso / aniso.mean(-1, keepdim=True)
    q = torch.randn((n_ctx, G_per, 4), generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    i, j, kk, r = q.unbind(-1)
    R = torch.stack([1 - 2 * (j * j + kk * kk), 2 * (i * j - kk * r), 2 * (i * kk + j * r),
                     2 * (i * j + kk * r), 1 - 2 * (i * i + kk * kk), 2 * (j * kk - i * r),
                     2 * (i * kk - j * r), 2 * (j * kk + i * r), 1 - 2 * (i * i + j * j)],
                    dim=-1).reshape(n_ctx, G_per, 3, 3)
    cov = R @ torch.diag_embed(scales ** 2) @ R.transpose(-1, -2)
    opac = 0.05 + 0.9 * torch.rand((n_ctx, G_per), generator=g, dtype=torch.float64)
    sh = torch.randn((n_ctx, G_per, 3, d_sh), generator=g, dtype=torch.float64) * 0.5
    mask = torch.ones(d_sh, dtype=torch.float64)
    for deg in range(1, isqrt(d_sh)):
        mask[deg * deg:(deg + 1) ** 2] = 0.1 * 0.25 ** deg
    sh = sh * mask
    out = dict(means=pts.reshape(-1, 3), covariances=cov.reshape(-1, 3, 3),
               harmonics=sh.reshape(-1, 3, d_sh), opacities=opac.reshape(-1))
    if n_gauss is not None:
        sel = torch.randperm(out["means"].shape[0], generator=g)[:n_gauss]
        out = {k_: v[sel] for k_, v in out.items()}
    out = {k_: v.to(dtype).contiguous() for k_, v in out.items()}
    out.update(extrinsics=tgt.to(dtype), intrinsics=K.to(dtype)[None].repeat(n_tgt, 1, 1),
               near=torch.full((n_tgt,), 0.01, dtype=dtype), far=torch.full((n_tgt,), 100.0, dtype=dtype))
    return out
