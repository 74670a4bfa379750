"""vs_update_pose against the oracle / the reference golden, and the alignment loop end to end:
perturbed target cameras are pulled back towards the poses the target images were rendered from."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import pose_ref

GOLD = Path(__file__).parent / "golden" / "pose_update.npz"


def test_update_pose_kernel_matches_reference_golden_and_oracle(cuda, lib):
    from vicasplat_b200.pose_align import update_pose
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    out = update_pose(t("rho"), t("theta"), t("extrinsics")).cpu().numpy()
    assert np.abs(out - g["out"]).max() < 2e-5
    ref = pose_ref.update_pose(g["rho"], g["theta"], g["extrinsics"], dtype=np.float64)
    assert np.abs(out - ref).max() < 2e-5
    z = torch.zeros((g["rho"].shape[0], 3), device=cuda)
    assert (update_pose(z, z, t("extrinsics")).cpu() - torch.from_numpy(g["extrinsics"])).abs().max() < 2e-6
    assert update_pose(z[:0], z[:0], t("extrinsics")[:0]).shape == (0, 4, 4)


def test_alignment_loop_reduces_loss_and_pose_error(cuda, lib):
    from vicasplat_b200 import decoder as dec, synthetic
    from vicasplat_b200.encoder import Gaussians
    from vicasplat_b200.loss import LossMse, LossMseCfg, LossMseCfgWrapper
    from vicasplat_b200.pose_align import test_step_align, update_pose
    S, V = 64, 3
    sc = {k: v.to(cuda) for k, v in synthetic.gaussian_scene(2, S, S, V, seed=11).items()}
    decoder = dec.DecoderSplattingCUDA(dec.DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], False)).to(cuda)
    gs = Gaussians(means=sc["means"][None], covariances=sc["covariances"][None],
                   harmonics=sc["harmonics"][None], opacities=sc["opacities"][None])
    cams = dict(intrinsics=sc["intrinsics"][None], near=sc["near"][None], far=sc["far"][None])
    with torch.no_grad():
        gt = decoder.forward(gs, sc["extrinsics"][None], cams["intrinsics"], cams["near"], cams["far"], (S, S))
    gen = torch.Generator().manual_seed(0)
    rho = (torch.randn((V, 3), generator=gen) * 0.01).to(cuda)
    th = (torch.randn((V, 3), generator=gen) * 0.004).to(cuda)
    start = update_pose(rho, th, sc["extrinsics"])[None]
    target = dict(image=gt.color, extrinsics=start, **cams)
    loss = LossMse(LossMseCfgWrapper(LossMseCfg(weight=1.0)))

    def err(E):
        return (E[0, :, :3, 3] - sc["extrinsics"][:, :3, 3]).norm(dim=-1).mean().item()

    with torch.no_grad():
        l0 = loss.forward(decoder.forward(gs, start, cams["intrinsics"], cams["near"], cams["far"], (S, S)),
                          {"target": target}).item()
    out, E = test_step_align(decoder, gs, target, [loss], steps=40, rot_opt_lr=5e-4, trans_opt_lr=1e-3)
    l1 = loss.forward(out, {"target": target}).item()
    assert l1 < 0.6 * l0, (l0, l1)
    assert err(E) < err(start), (err(start), err(E))
