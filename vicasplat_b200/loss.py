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
