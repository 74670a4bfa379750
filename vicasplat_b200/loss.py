"""Loss consumers of the render behind the reference's ``Loss`` interface.

* ``LossMse`` <-> src/loss/loss_mse.py:12-31: ``weight * ((prediction.color - target) ** 2).mean()``.
  Forward value and dL/dcolor come from ONE pass over the render (``vs_mse_loss``); autograd then
  hands that gradient straight to the rasterizer's backward.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from . import ops


@dataclass
class LossMseCfg:
    weight: float


@dataclass
class LossMseCfgWrapper:
    mse: LossMseCfg


class _Mse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, weight):
        loss, grad = ops.mse_loss(pred.detach(), target.detach(), weight, want_grad=pred.requires_grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None, None
        return ctx.grad * g, None, None


def mse(pred: torch.Tensor, target: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    if not pred.is_cuda:
        raise RuntimeError("vicasplat_b200.loss.mse needs CUDA tensors (there is no CPU fallback)")
    return _Mse.apply(pred, target.to(pred.dtype), float(weight))


class LossMse(nn.Module):
    """Same constructor / forward contract as the reference loss (``Loss[LossMseCfg, ...]``)."""

    def __init__(self, cfg: LossMseCfgWrapper) -> None:
        super().__init__()
        self.cfg = cfg.mse if hasattr(cfg, "mse") else cfg
        self.name = "mse"

    def forward(self, prediction, batch, gaussians=None, global_step: int = 0) -> torch.Tensor:
        return mse(prediction.color, batch["target"]["image"], self.cfg.weight)


# ------------------------------------------------------------------ camera loss (dual quaternions)
# LossCamera <-> src/loss/loss_camera.py:30-80 with the algebra of src/misc/dq.py:38-41,100-131 (xyzw
# quaternions).  The operands are (B, T-1, 8) tensors -- a few hundred numbers -- so this is plain torch
# arithmetic on the device (autograd supplies d loss / d pred, which TrainStep hands to the encoder's
# backward pass); there is nothing for a kernel to do.
def _qmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def _qconj(a: torch.Tensor) -> torch.Tensor:
    return torch.cat([-a[..., :3], a[..., 3:]], -1)


def _dq_mul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """(r1 + eps d1)(r2 + eps d2) = r1 r2 + eps (r1 d2 + d1 r2)   (dq.py:38-41)"""
    return torch.cat([_qmul(a[..., :4], b[..., :4]),
                      _qmul(a[..., :4], b[..., 4:]) + _qmul(a[..., 4:], b[..., :4])], -1)


def _dq_conj(a: torch.Tensor) -> torch.Tensor:
    return torch.cat([_qconj(a[..., :4]), _qconj(a[..., 4:])], -1)


def camera_dq_loss(prediction: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """L1 distance of pred * conj(target) and of target * conj(pred) from the identity (loss_camera.py:30-44)."""
    ident = torch.zeros(8, dtype=prediction.dtype, device=prediction.device)
    ident[3] = 1.0
    a = _dq_mul(prediction, _dq_conj(target))
    b = _dq_mul(target, _dq_conj(prediction))
    return (a - ident).abs().mean() + (b - ident).abs().mean()


def quaternion_from_matrix(R: torch.Tensor) -> torch.Tensor:
    """Unit quaternion xyzw (w >= 0) of rotation matrices (..., 3, 3): the branch with the largest
    denominator of the four standard candidates (what pytorch3d.transforms.matrix_to_quaternion computes,
    cam_utils.py:200-201, re-ordered to xyzw as camera_dq_array_from_Rt does)."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = R.flatten(-2).unbind(-1)
    q_abs = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22,
                         1.0 - m00 - m11 + m22], -1).clamp_min(0).sqrt()              # 2|w|, 2|x|, 2|y|, 2|z|
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1)], -2)  # rows: (w, x, y, z) candidates
    cand = cand / (2.0 * q_abs[..., None].clamp_min(0.1))
    best = q_abs.argmax(-1)
    wxyz = torch.gather(cand, -2, best[..., None, None].expand(*best.shape, 1, 4))[..., 0, :]
    wxyz = torch.nn.functional.normalize(wxyz, dim=-1)
    wxyz = torch.where(wxyz[..., :1] < 0, -wxyz, wxyz)
    return wxyz[..., [1, 2, 3, 0]]


def dq_from_Rt(R: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    """camera_dq_array_from_Rt (cam_utils.py:209-215): sigma = r + eps/2 * t * r, xyzw, (..., 8)."""
    q = quaternion_from_matrix(R)
    tq = torch.cat([t, torch.zeros_like(t[..., :1])], -1)
    return torch.cat([q, 0.5 * _qmul(tq, q)], -1)


def camera_loss(pred_dq: torch.Tensor, context_extrinsics: torch.Tensor, weight: float = 1.0,
                use_dq_loss: bool = True) -> torch.Tensor:
    """weight * (dq loss + L1) between the predicted dual quaternions (B, T-1, 8) and the context cameras
    1.. relative to camera 0 (loss_camera.py:53-76; the data shim has already made camera 0 the identity)."""
    ext = context_extrinsics[:, 1:]
    tgt = dq_from_Rt(ext[..., :3, :3], ext[..., :3, 3]).to(pred_dq.dtype)
    loss = (pred_dq - tgt).abs().mean()
    if use_dq_loss:
        loss = loss + camera_dq_loss(pred_dq, tgt)
    return weight * loss


@dataclass
class LossCameraCfg:
    weight: float
    use_dq_loss: bool = True
    camera_type: str = "dq"


@dataclass
class LossCameraCfgWrapper:
    camera: LossCameraCfg


class LossCamera(nn.Module):
    """Same constructor / forward contract as the reference loss (``Loss[LossCameraCfg, ...]``): reads
    ``prediction.extrinsics`` (the encoder's pred_extrins) and ``batch['context']['extrinsics']``."""

    def __init__(self, cfg: LossCameraCfgWrapper) -> None:
        super().__init__()
        self.cfg = cfg.camera if hasattr(cfg, "camera") else cfg
        self.name = "camera"
        if self.cfg.camera_type != "dq":
            raise NotImplementedError("LossCamera: only camera_type='dq' (every shipped experiment)")

    def forward(self, prediction, batch, gaussians=None, global_step: int = 0) -> torch.Tensor:
        if getattr(prediction, "intrinsics", None) is not None:
            raise NotImplementedError("LossCamera: predicted intrinsics (use_intrinsic_embedding=False) are not on "
                                      "the shipped 8-view path")
        return camera_loss(prediction.extrinsics, batch["context"]["extrinsics"], self.cfg.weight,
                           self.cfg.use_dq_loss)
