"""LPIPS (VGG16) consumer of the render: ``LossLpips`` <-> src/loss/loss_lpips.py:27-54, i.e. the ``lpips``
package's ``LPIPS(net="vgg")`` called with ``normalize=True`` -- scaling layer, VGG16 features at relu1_2 /
relu2_2 / relu3_3 / relu4_3 / relu5_3, channel-unit-normalised squared difference, learned 1x1 ``lin``
weights, spatial mean, sum over the five layers; the loss is ``weight * mean over the (b v) images``.

Everything dense runs on the kernels of this library: the thirteen 3x3 convolutions (+ bias + ReLU epilogue)
on the implicit-GEMM conv of vs_gemm, their input gradients on the same kernel with flipped taps and the
ReLU-mask epilogue (the network is frozen: no weight gradients), 2x2 max pooling and the per-layer distance
(value + gradient in one pass) in csrc/train_ops.cu.  The loss VALUE and dL/d prediction come out of one
forward + backward sweep without autograd, like vs_mse_loss.

WEIGHTS.  The ``lpips`` package (VGG16 ImageNet weights + the learned ``lin`` layers) is not in this image and
there is no network.  ``LpipsVgg(weights=...)`` takes them as a dict of tensors (``vgg.{i}.weight / .bias`` in
torchvision's conv order, ``lin.{k}.weight``); ``LpipsVgg.stand_in(seed)`` builds a seeded RANDOM network of
the same architecture, clearly labelled (``.is_stand_in``): it exercises and times the real data path but is
not a perceptual metric."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import VS_ACT_RELU
from .encoder import _pack_conv
from .train import _flipped

CHANNELS = (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512)
TAPS = (1, 3, 6, 9, 12)                  # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 (conv indices)
POOL_AFTER = (1, 3, 6, 9)
SHIFT = (-0.030, -0.088, -0.188)         # lpips.ScalingLayer
SCALE = (0.458, 0.448, 0.450)
BF16, F32 = torch.bfloat16, torch.float32


class LpipsVgg:
    def __init__(self, weights: Dict[str, torch.Tensor], device, is_stand_in: bool = False):
        self.dev, self.is_stand_in = device, is_stand_in
        self.w, self.wf, self.b, self.lin = [], [], [], []
        cin = 3
        for i, cout in enumerate(CHANNELS):
            W = weights[f"vgg.{i}.weight"].to(device=device, dtype=F32)
            assert W.shape == (cout, cin, 3, 3), (i, W.shape)
            if i == 0:                    # 3 -> 8 zero-padded input channels, run as im2col + GEMM (K = 72)
                W8 = torch.zeros((cout, 8, 3, 3), dtype=F32, device=device)
                W8[:, :3] = W
                wp = W8.permute(0, 2, 3, 1).reshape(cout, 72)
                self.w.append(wp.to(BF16).contiguous())
                self.wf.append(wp.t().contiguous().to(BF16))          # (72, 64): d cols = dY W
            else:
                self.w.append(_pack_conv(W))
                self.wf.append(_flipped(W))
            self.b.append(weights[f"vgg.{i}.bias"].to(device=device, dtype=F32).contiguous())
            cin = cout
        for k, t in enumerate(TAPS):
            lw = weights[f"lin.{k}.weight"].to(device=device, dtype=F32).reshape(-1).contiguous()
            assert lw.numel() == CHANNELS[t]
            self.lin.append(lw)
        self.shift = torch.tensor(SHIFT, dtype=F32, device=device).view(1, 3, 1, 1)
        self.scale = torch.tensor(SCALE, dtype=F32, device=device).view(1, 3, 1, 1)

    @staticmethod
    def stand_in_weights(seed: int = 0) -> Dict[str, torch.Tensor]:
        """Seeded random VGG16-shaped weights (He-normal convs, small positive ``lin`` weights)."""
        g = torch.Generator().manual_seed(seed)
        w, cin = {}, 3
        for i, cout in enumerate(CHANNELS):
            w[f"vgg.{i}.weight"] = torch.randn((cout, cin, 3, 3), generator=g) * (2.0 / (9 * cin)) ** 0.5
            w[f"vgg.{i}.bias"] = torch.randn((cout,), generator=g) * 0.05
            cin = cout
        for k, t in enumerate(TAPS):
            w[f"lin.{k}.weight"] = torch.rand((CHANNELS[t],), generator=g) / CHANNELS[t]
        return w

    @classmethod
    def stand_in(cls, device, seed: int = 0) -> "LpipsVgg":
        return cls(cls.stand_in_weights(seed), device, is_stand_in=True)

    # ---- forward of the feature stack
    def _input(self, img: torch.Tensor) -> torch.Tensor:
        """(N,3,H,W) in [0,1] -> scaled NHWC8 bf16 (normalize=True: 2x - 1, then the scaling layer)."""
        x = ((2.0 * img.to(F32) - 1.0) - self.shift) / self.scale
        n, _, h, w = x.shape
        out = torch.zeros((n, h, w, 8), dtype=BF16, device=self.dev)
        out[..., :3] = x.permute(0, 2, 3, 1)
        return out

    def _features(self, img: torch.Tensor, keep: bool):
        x8 = self._input(img)
        n, h, w, _ = x8.shape
        cols = ops.im2col(x8, nchw_f32=False, n=n, h=h, w=w, c=8, k=3, stride=1, pad=1, kpad=72)
        a = ops.gemm(cols, self.w[0], bias=self.b[0], act=VS_ACT_RELU).view(n, h, w, CHANNELS[0])
        acts: List[torch.Tensor] = [a]
        pooled: Dict[int, torch.Tensor] = {}
        feats = []
        for i in range(1, len(CHANNELS)):
            src = acts[-1]
            if (i - 1) in POOL_AFTER:
                src = ops.maxpool2(src)
                pooled[i - 1] = src
            a = ops.conv_gemm(src, self.w[i], kh=3, kw=3, pad=1, N=CHANNELS[i], bias=self.b[i], act=VS_ACT_RELU)
            if keep:
                acts.append(a)
            else:
                if (i - 1) in TAPS:
                    feats.append(acts[-1])
                acts = [a]
        if keep:
            return acts, pooled
        feats.append(acts[-1])
        return feats, None

    # ---- value + gradient
    @torch.no_grad()
    def loss_and_grad(self, pred: torch.Tensor, target: torch.Tensor, weight: float = 1.0, want_grad: bool = True):
        """pred, target (N,3,H,W) in [0,1] -> (weight * mean_N lpips (device scalar), dL/dpred (N,3,H,W) fp32)."""
        n, _, h, w = pred.shape
        f1, _ = self._features(target, keep=False)
        acts, pooled = self._features(pred, keep=True)
        per_image = torch.zeros((n,), dtype=F32, device=self.dev)
        tap_grad: Dict[int, Optional[torch.Tensor]] = {}
        for k, t in enumerate(TAPS):
            f0 = acts[t]
            hw = f0.shape[1] * f0.shape[2]
            tap_grad[t] = ops.lpips_layer(f0, f1[k], self.lin[k], per_image, weight / (n * hw), want_grad)
        loss = per_image.mean() * weight
        if not want_grad:
            return loss, None
        G = tap_grad[12]                                       # gradient of conv 12's pre-activation
        for i in range(12, 0, -1):
            prev = i - 1
            if prev in POOL_AFTER:                             # conv i read pool(a_prev); a_prev is a tap
                d_pool = ops.conv_gemm(G, self.wf[i], kh=3, kw=3, pad=1, N=CHANNELS[prev])
                G = ops.maxpool2_backward(acts[prev], pooled[prev], d_pool, add=tap_grad[prev], relu_mask=True)
            else:
                G = ops.conv_gemm_masked(G, self.wf[i], kh=3, kw=3, pad=1, N=CHANNELS[prev], mask=acts[prev],
                                         res1=tap_grad.get(prev))
        d_cols = ops.gemm(G.view(-1, CHANNELS[0]), self.wf[0])          # (N h w, 72)
        d_x8 = ops.col2im(d_cols, n=n, h=h, w=w, c=8, k=3, stride=1, pad=1)
        d_pred = d_x8[..., :3].permute(0, 3, 1, 2).to(F32) * (2.0 / self.scale)
        return loss, d_pred.contiguous()


class LossLpips(torch.nn.Module):
    """Same forward contract as the reference loss (``Loss[LossLpipsCfg, ...]``); differentiable through a
    custom autograd node that returns the gradient computed alongside the value."""

    def __init__(self, cfg, net: Optional[LpipsVgg] = None) -> None:
        super().__init__()
        self.cfg = cfg.lpips if hasattr(cfg, "lpips") else cfg
        self.name = "lpips"
        self.net = net

    def forward(self, prediction, batch, gaussians=None, global_step: int = 0) -> torch.Tensor:
        image = batch["target"]["image"]
        if global_step < self.cfg.apply_after_step:
            return torch.tensor(0, dtype=torch.float32, device=image.device)
        if self.net is None:
            raise RuntimeError("LossLpips needs an LpipsVgg (the lpips package's weights are not in this image: "
                               "pass LpipsVgg(weights) or, for timing only, LpipsVgg.stand_in(device))")
        return _LpipsFn.apply(prediction.color.flatten(0, 1), image.flatten(0, 1), self.net, float(self.cfg.weight))


class _LpipsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, net, weight):
        loss, grad = net.loss_and_grad(pred.detach(), target.detach(), weight, want_grad=pred.requires_grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, g):
        return (None if ctx.grad is None else ctx.grad * g), None, None, None
