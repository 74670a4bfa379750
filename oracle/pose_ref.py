"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's pose update used by test-time
pose alignment -- ``update_pose`` / ``SE3_exp`` / ``SO3_exp`` / ``V`` of src/misc/cam_utils.py:52-142.
Pinned against the reference's own functions by oracle/make_pose_golden.py ->
tests/golden/pose_update.npz (tests/test_oracle_pose_cpu.py).  Only tests may import this."""
from __future__ import annotations

import numpy as np


def skew(x):                                    # cam_utils.py:60-71
    return np.array([[0.0, -x[2], x[1]], [x[2], 0.0, -x[0]], [-x[1], x[0], 0.0]], dtype=x.dtype)


def so3_exp(theta):                             # cam_utils.py:74-90
    W = skew(theta)
    W2 = W @ W
    angle = np.linalg.norm(theta)
    I = np.eye(3, dtype=theta.dtype)
    if angle < 1e-5:
        return I + W + 0.5 * W2
    return I + (np.sin(angle) / angle) * W + ((1 - np.cos(angle)) / angle ** 2) * W2


def v_mat(theta):                               # cam_utils.py:93-108
    W = skew(theta)
    W2 = W @ W
    angle = np.linalg.norm(theta)
    I = np.eye(3, dtype=theta.dtype)
    if angle < 1e-5:
        return I + 0.5 * W + (1.0 / 6.0) * W2
    return I + W * ((1.0 - np.cos(angle)) / angle ** 2) + W2 * ((angle - np.sin(angle)) / angle ** 3)


def se3_exp(tau):                               # cam_utils.py:111-124: tau = (rho, theta)
    T = np.eye(4, dtype=tau.dtype)
    T[:3, :3] = so3_exp(tau[3:])
    T[:3, 3] = v_mat(tau[3:]) @ tau[:3]
    return T


def update_pose(cam_trans_delta, cam_rot_delta, extrinsics, dtype=np.float64):
    """cam_utils.py:127-148: c2w' = inv( exp([rho, theta]) @ inv(c2w) ), per camera."""
    rho = np.asarray(cam_trans_delta, dtype=dtype)
    th = np.asarray(cam_rot_delta, dtype=dtype)
    E = np.asarray(extrinsics, dtype=dtype)
    out = np.empty_like(E)
    for i in range(E.shape[0]):
        w2c = np.linalg.inv(E[i])
        out[i] = np.linalg.inv(se3_exp(np.concatenate([rho[i], th[i]])) @ w2c)
    return out
