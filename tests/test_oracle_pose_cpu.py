"""The pose-update oracle (oracle/pose_ref.py) against golden vectors produced by the reference's own
update_pose (src/misc/cam_utils.py:127-148; oracle/make_pose_golden.py)."""
from pathlib import Path

import numpy as np

from oracle import pose_ref

GOLD = Path(__file__).parent / "golden" / "pose_update.npz"


def test_update_pose_matches_reference_golden():
    g = np.load(GOLD)
    out = pose_ref.update_pose(g["rho"], g["theta"], g["extrinsics"], dtype=np.float64)
    assert np.abs(out - g["out"]).max() < 2e-5          # the golden is fp32 (two fp32 inverses)
    out32 = pose_ref.update_pose(g["rho"], g["theta"], g["extrinsics"], dtype=np.float32)
    assert np.abs(out32 - g["out"]).max() < 2e-5


def test_zero_delta_is_identity_and_small_angle_branch_is_continuous():
    g = np.load(GOLD)
    E = g["extrinsics"].astype(np.float64)
    z = np.zeros((E.shape[0], 3))
    assert np.abs(pose_ref.update_pose(z, z, E) - E).max() < 1e-12
    th = np.array([[0.99e-5, 0, 0]]); th2 = np.array([[1.01e-5, 0, 0]])
    a = pose_ref.update_pose(z[:1] + 0.1, th, E[:1]); b = pose_ref.update_pose(z[:1] + 0.1, th2, E[:1])
    assert np.abs(a - b).max() < 1e-6
