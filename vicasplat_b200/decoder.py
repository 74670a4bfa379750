"""Render path of the reference behind its own plugin surface:

* ``render_cuda``            <-> src/model/decoder/cuda_splatting.py:148-239
* ``get_projection_matrix``  <-> cuda_splatting.py:18-45
* ``get_fov``                <-> src/geometry/projection.py:247-261
* ``DecoderSplattingCUDA``   <-> src/model/decoder/decoder_splatting_cuda.py:23-101
* ``DecoderOutput``          <-> src/model/decoder/decoder.py:18-21

Same names, argument meaning and return shapes; what changes is underneath: the V views of a
scene are rendered by ONE launch chain from ONE copy of the Gaussians (no per-view ``repeat``,
no Python loop, no ``.item()`` syncs, SH consumed in the encoder's (3, d_sh) layout).
"""
from __future__ import annotations

from dataclasses import dataclass
from math import isqrt
from typing import Literal, Optional

import torch
from torch import Tensor, nn

from .rasterizer import SceneStreams, rasterize_views

DepthRenderingMode = Literal["depth", "log", "disparity", "relative_disparity"]


@dataclass
class DecoderOutput:
    color: Optional[Tensor] = None   # (batch, view, 3, H, W)
    depth: Optional[Tensor] = None   # (batch, view, H, W)


@dataclass
class DecoderSplattingCUDACfg:
    name: Literal["splatting_cuda"]
    background_color: list
    make_scale_invariant: bool
    use_gsplat: bool = False


def _inv(m: Tensor) -> Tensor:
    """Batched inverse without the host synchronisation of torch.linalg.inv's error check."""
    return torch.linalg.inv_ex(m, check_errors=False)[0]


def get_fov(intrinsics: Tensor) -> Tensor:
    """(B,3,3) normalised intrinsics -> (B,2) full field of view (x, y) in radians: the angle
    between the rays through the mid-points of opposite image borders."""
    inv = _inv(intrinsics)
    # rays through (0, .5), (1, .5), (.5, 0), (.5, 1): columns of inv combined, no host tensors
    c0, c1, c2 = inv[..., 0], inv[..., 1], inv[..., 2]

    def unit(d):
        return d / d.norm(dim=-1, keepdim=True)

    cx = (unit(0.5 * c1 + c2) * unit(c0 + 0.5 * c1 + c2)).sum(-1)
    cy = (unit(0.5 * c0 + c2) * unit(0.5 * c0 + c1 + c2)).sum(-1)
    return torch.stack((cx.acos(), cy.acos()), dim=-1)


def get_projection_matrix(near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor) -> Tensor:
    """Symmetric-frustum projection: x,y -> [-1,1], z -> [0,1], w = z_view (+z forward)."""
    P = torch.zeros((near.shape[0], 4, 4), dtype=torch.float32, device=near.device)
    P[:, 0, 0] = 1.0 / (0.5 * fov_x).tan()
    P[:, 1, 1] = 1.0 / (0.5 * fov_y).tan()
    P[:, 2, 2] = far / (far - near)
    P[:, 2, 3] = -(far * near) / (far - near)
    P[:, 3, 2] = 1.0
    return P


def _cameras(extrinsics, intrinsics, near, far):
    fov = get_fov(intrinsics)
    tanfov = (0.5 * fov).tan()
    proj_t = get_projection_matrix(near, far, fov[:, 0], fov[:, 1]).transpose(1, 2)
    view_t = _inv(extrinsics).transpose(1, 2)
    return tanfov, view_t, view_t @ proj_t, extrinsics[:, :3, 3]


def _cov6(cov: Tensor) -> Tensor:
    return torch.stack([cov[..., 0, 0], cov[..., 0, 1], cov[..., 0, 2], cov[..., 1, 1],
                        cov[..., 1, 2], cov[..., 2, 2]], dim=-1)


def render_cuda(extrinsics, intrinsics, near, far, image_shape, background_color, gaussian_means,
                gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities,
                scale_invariant: bool = True, cam_rot_delta=None, cam_trans_delta=None,
                use_sh: bool = True, sh_degree: Optional[int] = None):
    """``batch`` cameras; Gaussians either per camera (batch, G, ...) or shared (G, ...).
    ``scale_invariant`` is accepted and has no effect, as in the reference (:170-178)."""
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    n = gaussian_sh_coefficients.shape[-1]
    degree = sh_degree or isqrt(n) - 1
    h, w = image_shape
    tanfov, view_t, full_t, campos = _cameras(extrinsics, intrinsics, near, far)
    if use_sh:
        shs, cols = gaussian_sh_coefficients, None
    else:
        shs, cols = None, gaussian_sh_coefficients[..., 0]
    color, _radii, depth, _alpha, _nt = rasterize_views(
        gaussian_means, _cov6(gaussian_covariances), gaussian_opacities, shs=shs,
        colors_precomp=cols, sh_degree=degree, sh_layout="chan_major", viewmatrix=view_t,
        projmatrix=full_t, campos=campos, tanfov=tanfov, bg=background_color, H=h, W=w,
        theta=cam_rot_delta, rho=cam_trans_delta, want_n_touched=False)
    return color, depth[:, 0]


class DecoderSplattingCUDA(nn.Module):
    """Same constructor / forward contract as the reference decoder plugin."""

    def __init__(self, cfg: DecoderSplattingCUDACfg) -> None:
        super().__init__()
        self.cfg = cfg
        self.make_scale_invariant = cfg.make_scale_invariant
        self.register_buffer("background_color",
                             torch.tensor(cfg.background_color, dtype=torch.float32),
                             persistent=False)

    def forward(self, gaussians, extrinsics, intrinsics, near, far, image_shape,
                depth_mode: DepthRenderingMode | None = None, cam_rot_delta=None,
                cam_trans_delta=None, use_sh: bool = True, active_sh_degree: Optional[int] = None,
                return_dict: bool = True, check_overflow=True):
        """``check_overflow``: see ``rasterize_views`` ("deferred" keeps this call free of host
        synchronisation; ScenePipeline verifies the queued counters with the batch's results)."""
        if self.cfg.use_gsplat:
            raise NotImplementedError("use_gsplat=True: gsplat is not part of this hot path "
                                      "(every shipped experiment sets use_gsplat: false)")
        b, v = extrinsics.shape[:2]
        means, cov, sh, opac = (gaussians.means, gaussians.covariances, gaussians.harmonics,
                                gaussians.opacities)
        if means.ndim > 3:  # (b, t, h, w, ...) -> (b, G, ...)
            means, cov, sh = means.flatten(1, 3), cov.flatten(1, 3), sh.flatten(1, 3)
            opac = opac.flatten(1)
        # camera set-up once for all b*v cameras (a handful of tiny kernels, no host sync), then one
        # launch chain per scene
        h, w = image_shape
        tanfov, view_t, full_t, campos = _cameras(extrinsics.flatten(0, 1), intrinsics.flatten(0, 1),
                                                  near.flatten(), far.flatten())
        n = sh.shape[-1]
        degree = active_sh_degree or isqrt(n) - 1
        assert use_sh or n == 1
        # the packed upper triangle the adapter already wrote (encoder.Gaussians.cov6), if it is there
        cov6 = getattr(gaussians, "cov6", None)
        cov6 = _cov6(cov) if cov6 is None else (cov6.flatten(1, 3) if cov6.ndim > 3 else cov6)
        bg = self.background_color[None].expand(v, 3)
        colors, depths = [], []
        # (one stream when autograd records the renders: its backward nodes would otherwise run on the side streams)
        needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in
                                                     (means, cov, sh, opac, cov6, cam_rot_delta, cam_trans_delta))
        with SceneStreams(means.device, n=1 if needs_grad else None) as ss:
            for i in range(b):
                sl = slice(i * v, (i + 1) * v)
                with ss.scene(i):
                    c, _r, d, _a, _n = rasterize_views(
                        means[i], cov6[i], opac[i], shs=sh[i] if use_sh else None,
                        colors_precomp=None if use_sh else sh[i][..., 0], sh_degree=degree,
                        sh_layout="chan_major", viewmatrix=view_t[sl], projmatrix=full_t[sl], campos=campos[sl],
                        tanfov=tanfov[sl], bg=bg, H=h, W=w,
                        theta=None if cam_rot_delta is None else cam_rot_delta[i],
                        rho=None if cam_trans_delta is None else cam_trans_delta[i], want_n_touched=False,
                        check_overflow=check_overflow)
                ss.keep(c, d)
                colors.append(c)
                depths.append(d[:, 0])
        color, depth = torch.stack(colors), torch.stack(depths)
        if not return_dict:
            return color, depth
        return DecoderOutput(color, depth)
