"""Test-time pose alignment: the consumer of the rasterizer's camera-twist
gradients.

* ``update_pose``      <-> src/misc/cam_utils.py:127-148 (host loop of SE3_exp + two batched inverses
                           in the reference; one kernel launch here, ``vs_update_pose``)
* ``test_step_align``  <-> ModelWrapper.test_step_align, src/model/model_wrapper.py:442-513: Adam on
                           per-view (rot, trans) deltas that are re-zeroed after every step while the
                           extrinsics absorb them; render -> losses -> backward each iteration.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import torch
from torch import Tensor, nn

from . import _lib
from ._lib import check, ptr, stream_ptr


def update_pose(cam_trans_delta: Tensor, cam_rot_delta: Tensor, extrinsics: Tensor) -> Tensor:
    """(n,3), (n,3), (n,4,4) c2w -> (n,4,4) c2w' = inv(SE3_exp([trans, rot]) @ inv(c2w))."""
    if not extrinsics.is_cuda:
        raise RuntimeError("vicasplat_b200.pose_align.update_pose needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    rho = cam_trans_delta.detach().to(torch.float32).contiguous()
    th = cam_rot_delta.detach().to(torch.float32).contiguous()
    E = extrinsics.detach().to(torch.float32).contiguous()
    n = E.shape[0]
    assert rho.shape == (n, 3) and th.shape == (n, 3) and E.shape == (n, 4, 4)
    out = torch.empty_like(E)
    check(lib.vs_update_pose(C.c_void_p(ptr(rho)), C.c_void_p(ptr(th)), C.c_void_p(ptr(E)),
                             C.c_void_p(ptr(out)), n, C.c_void_p(stream_ptr())), "vs_update_pose")
    return out


def test_step_align(decoder, gaussians, target: dict, losses: Sequence[Callable], *, steps: int,
                    rot_opt_lr: float, trans_opt_lr: float, global_step: int = 0):
    """target: {"image" (b,v,3,h,w), "extrinsics" (b,v,4,4), "intrinsics", "near", "far"}.
    losses: objects with ``.name`` and ``.forward(prediction, batch, gaussians, global_step)``; the
    one named "camera" is skipped, as in the reference.  Returns (final DecoderOutput, extrinsics)."""
    b, v, _, h, w = target["image"].shape
    dev = target["image"].device
    batch = {"target": target}
    with torch.enable_grad():
        cam_rot_delta = nn.Parameter(torch.zeros((b, v, 3), device=dev))
        cam_trans_delta = nn.Parameter(torch.zeros((b, v, 3), device=dev))
        opt = torch.optim.Adam([{"params": [cam_rot_delta], "lr": rot_opt_lr},
                                {"params": [cam_trans_delta], "lr": trans_opt_lr}])
        extrinsics = target["extrinsics"].clone()
        for _ in range(steps):
            opt.zero_grad()
            output = decoder.forward(gaussians, extrinsics, target["intrinsics"], target["near"],
                                     target["far"], (h, w), cam_rot_delta=cam_rot_delta,
                                     cam_trans_delta=cam_trans_delta)
            total = 0
            for loss_fn in losses:
                if getattr(loss_fn, "name", "") != "camera":
                    total = total + loss_fn.forward(output, batch, gaussians, global_step)
            total.backward()
            with torch.no_grad():
                opt.step()
                extrinsics = update_pose(cam_trans_delta.reshape(b * v, 3), cam_rot_delta.reshape(b * v, 3),
                                         extrinsics.reshape(b * v, 4, 4)).reshape(b, v, 4, 4)
                cam_rot_delta.data.fill_(0)
                cam_trans_delta.data.fill_(0)
    with torch.no_grad():
        output = decoder.forward(gaussians, extrinsics, target["intrinsics"], target["near"],
                                 target["far"], (h, w))
    return output, extrinsics


test_step_align.__test__ = False   # not a pytest test despite the reference's name
