"""TEST INFRASTRUCTURE (oracle): plain torch restatement of ``lpips.LPIPS(net="vgg")`` called as the reference
does (src/loss/loss_lpips.py:49-54: ``normalize=True``, ``.mean()`` over the images), driven by an explicit
weight dict (``vgg.{i}.weight / .bias`` in torchvision's conv order, ``lin.{k}.weight``).

PARITY UNPINNED: the ``lpips`` package is a pip dependency of the reference (requirements.txt) that is absent
from /root/reference and from this image, and so are its weights; the algorithm is restated from the
package's published source (scaling layer constants, VGG16 slices at relu1_2 .. relu5_3, normalize_tensor
with eps 1e-10, squared difference, 1x1 ``lin`` layer, spatial average, sum over layers).  Nothing here is
imported by the product."""
from __future__ import annotations

import torch
import torch.nn.functional as F

CHANNELS = (64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512)
TAPS = (1, 3, 6, 9, 12)
POOL_AFTER = (1, 3, 6, 9)


def features(w, x):
    feats = []
    for i in range(len(CHANNELS)):
        x = F.relu(F.conv2d(x, w[f"vgg.{i}.weight"], w[f"vgg.{i}.bias"], padding=1))
        if i in TAPS:
            feats.append(x)
        if i in POOL_AFTER:
            x = F.max_pool2d(x, 2, 2)
    return feats


def lpips(w, pred, target):
    """(N,3,H,W) in [0,1] -> (N,) distances."""
    shift = torch.tensor([-0.030, -0.088, -0.188], dtype=pred.dtype, device=pred.device).view(1, 3, 1, 1)
    scale = torch.tensor([0.458, 0.448, 0.450], dtype=pred.dtype, device=pred.device).view(1, 3, 1, 1)
    prep = lambda t: ((2 * t - 1) - shift) / scale
    f0, f1 = features(w, prep(pred)), features(w, prep(target))
    total = 0
    for k in range(len(TAPS)):
        n0 = f0[k] / (f0[k].pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
        n1 = f1[k] / (f1[k].pow(2).sum(1, keepdim=True).sqrt() + 1e-10)
        d = ((n0 - n1) ** 2 * w[f"lin.{k}.weight"].view(1, -1, 1, 1)).sum(1)
        total = total + d.mean((1, 2))
    return total
