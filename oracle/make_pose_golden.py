"""Golden vectors for the test-time pose update from the UNMODIFIED reference functions
(src/misc/cam_utils.py:127-148 ``update_pose``).  Build container only (needs /root/reference):

    python oracle/make_pose_golden.py          # writes tests/golden/pose_update.npz
"""
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def inputs(seed=7, n=24):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn((n, 4), generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    w, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(n, 3, 3)
    E = torch.eye(4).repeat(n, 1, 1)
    E[:, :3, :3] = R
    E[:, :3, 3] = torch.randn((n, 3), generator=g) * 2
    rho = torch.randn((n, 3), generator=g) * 0.05
    theta = torch.randn((n, 3), generator=g) * 0.05
    theta[:4] *= 1e-6            # small-angle branch (|theta| < 1e-5)
    theta[4] = 0.0
    rho[5] = 0.0
    theta[6:8] *= 20.0           # large rotations
    return rho, theta, E


def main():
    os.chdir("/tmp")
    sys.path.insert(0, str(ROOT))
    from oracle.make_encoder_golden import install_stubs
    install_stubs()
    if "cv2" not in sys.modules:
        try:
            import cv2  # noqa: F401
        except ImportError:
            sys.modules["cv2"] = types.ModuleType("cv2")
    sys.path.insert(0, str(REF))
    from src.misc.cam_utils import update_pose      # the reference's own code
    rho, theta, E = inputs()
    out = update_pose(cam_trans_delta=rho, cam_rot_delta=theta, extrinsics=E)
    path = ROOT / "tests" / "golden" / "pose_update.npz"
    np.savez_compressed(path, rho=rho.numpy(), theta=theta.numpy(), extrinsics=E.numpy(), out=out.numpy())
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
