"""LossCamera (vicasplat_b200.loss: dual-quaternion L1, src/loss/loss_camera.py:30-80) against values and
gradients of the UNMODIFIED reference functions (tests/golden/loss_camera.npz, oracle/make_loss_golden.py).
Host-side torch arithmetic on a few hundred numbers: runs on the CPU."""
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

GOLD = Path(__file__).parent / "golden" / "loss_camera.npz"


def test_camera_loss_value_and_gradient_match_the_reference():
    from vicasplat_b200 import loss as L
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k])
    # ground-truth conversion: rotation matrix + translation -> dual quaternion
    dq = L.dq_from_Rt(t("R"), t("t"))
    assert torch.allclose(dq, t("target_dq"), atol=2e-6)
    pred = t("pred").clone().requires_grad_(True)
    loss = L.camera_dq_loss(pred, t("target_dq")) + (pred - t("target_dq")).abs().mean()
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    loss.backward()
    assert torch.allclose(pred.grad, t("grad"), atol=1e-7)
    # the plugin class, on 4x4 context extrinsics (camera 0 = identity)
    B, V = t("R").shape[:2]
    ext = torch.eye(4).repeat(B, V + 1, 1, 1)
    ext[:, 1:, :3, :3] = t("R")
    ext[:, 1:, :3, 3] = t("t")
    mod = L.LossCamera(L.LossCameraCfgWrapper(L.LossCameraCfg(weight=0.1)))
    got = mod(SimpleNamespace(extrinsics=t("pred"), intrinsics=None), {"context": {"extrinsics": ext}})
    assert abs(got.item() - 0.1 * float(g["loss"])) < 1e-6


def test_quaternion_from_matrix_all_branches():
    from vicasplat_b200 import loss as L
    gen = torch.Generator().manual_seed(0)
    q = torch.randn((2000, 4), generator=gen, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    q[:4] = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0], [0, 0, 0, 1.0]], dtype=torch.float64)  # 180 degree turns
    x, y, z, w = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    got = L.quaternion_from_matrix(R)
    same = (got - q).abs().max(-1).values
    flip = (got + q).abs().max(-1).values
    assert torch.minimum(same, flip).max() < 1e-9
