"""Golden values of the reference's camera loss (src/loss/loss_camera.py:30-80) from the UNMODIFIED reference
functions.  Run in the build container only (needs /root/reference):

    python oracle/make_loss_golden.py        # writes tests/golden/loss_camera.npz

``camera_dq_loss`` and ``DualQuaternion.from_quat_pose_array`` run as they are (pypose stubbed by the xyzw
Hamilton algebra of oracle/make_encoder_golden.py, made transparent to torch functions);
``matrix_to_quaternion`` (pytorch3d, absent) is only needed to turn the ground-truth rotations into
quaternions: the goldens therefore start from seeded unit quaternions and ALSO store the rotation matrices
built from them, so the test exercises vicasplat_b200.loss.dq_from_Rt on the matrices."""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def main():
    os.chdir("/tmp")
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT))
    from oracle import make_encoder_golden as mg
    mg.install_stubs()
    import pypose

    def tf(cls, func, types, args=(), kwargs=None):
        un = lambda a: a.t if isinstance(a, pypose.SO3) else a
        return func(*[un(a) for a in args], **{k: un(v) for k, v in (kwargs or {}).items()})
    pypose.SO3.__torch_function__ = classmethod(tf)
    from src.loss.loss import l1_loss
    from src.loss.loss_camera import camera_dq_loss
    from src.misc.dq import DualQuaternion
    g = torch.Generator().manual_seed(31)
    B, V = 3, 7
    pred = torch.randn((B, V, 8), generator=g) * 0.3
    pred[..., 3] += 1
    pred = pred / pred[..., :4].norm(dim=-1, keepdim=True)                 # what the pose head emits
    q = torch.randn((B, V, 4), generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    q = q * torch.sign(q[..., 3:])                                         # w >= 0, as matrix_to_quaternion returns
    t = torch.randn((B, V, 3), generator=g)
    target = DualQuaternion.from_quat_pose_array(torch.cat([q, t], -1))
    R = target.homogeneous_matrix[..., :3, :3]
    tgt = target.dq_array
    pred_r = pred.clone().requires_grad_(True)
    loss = camera_dq_loss(pred_r, tgt) + l1_loss(pred_r, tgt)              # loss_camera.py:69-70
    loss.backward()
    out = ROOT / "tests" / "golden" / "loss_camera.npz"
    np.savez_compressed(out, pred=pred.numpy(), q=q.numpy(), t=t.numpy(), R=R.numpy(), target_dq=tgt.numpy(),
                        loss=np.float64(loss.item()), grad=pred_r.grad.numpy())
    print("wrote", out, "loss", loss.item())


if __name__ == "__main__":
    main()
