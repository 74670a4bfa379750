"""GPU parity of the rasterizer BACKWARD pass (vs_raster_backward through the autograd surface)
against torch.autograd on the oracle: gradients w.r.t. means3D, cov6, SH, opacity and the camera
twist (theta, rho).  As for the forward pass the fp32 oracle is the stand-in for the (fp32)
upstream extension: tolerance 3e-3 relative L2 per tensor against it, 2e-2 against the fp64 oracle
(whose forward image already differs by threshold flips on ~1 % of the pixels)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import raster_ref as rr


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _setup(hw, n_ctx, n_tgt, seed, n_gauss=None):
    sc = rr.synthetic_scene(n_ctx, hw, hw, n_tgt, seed=seed, n_gauss=n_gauss)
    g = torch.Generator().manual_seed(seed + 100)
    wc = torch.randn((n_tgt, 3, hw, hw), generator=g)
    wd = 0.1 * torch.randn((n_tgt, hw, hw), generator=g)
    return sc, wc, wd


def _oracle_grads(sc, wc, wd, hw, with_pose, dtype=torch.float64):
    f = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sc.items()}
    means = f["means"].clone().requires_grad_(True)
    cov = f["covariances"].clone().requires_grad_(True)
    sh = f["harmonics"].clone().requires_grad_(True)
    op = f["opacities"].clone().requires_grad_(True)
    V = f["extrinsics"].shape[0]
    theta = torch.zeros((V, 3), dtype=dtype, requires_grad=True)
    rho = torch.zeros((V, 3), dtype=dtype, requires_grad=True)
    c, d = rr.render_cuda_ref(f["extrinsics"], f["intrinsics"], f["near"], f["far"], (hw, hw),
                              torch.zeros((V, 3), dtype=dtype), means, cov, sh, op,
                              cam_rot_delta=theta if with_pose else None,
                              cam_trans_delta=rho if with_pose else None)
    loss = (c * wc.to(dtype)).sum() + (d * wd.to(dtype)).sum()
    loss.backward()
    gc = cov.grad
    # gradient w.r.t. the 6 packed entries: off-diagonals collect both symmetric positions
    g6 = torch.stack([gc[:, 0, 0], gc[:, 0, 1] + gc[:, 1, 0], gc[:, 0, 2] + gc[:, 2, 0], gc[:, 1, 1],
                      gc[:, 1, 2] + gc[:, 2, 1], gc[:, 2, 2]], dim=-1)
    return dict(means=means.grad, cov6=g6, sh=sh.grad, op=op.grad,
                theta=theta.grad if with_pose else None, rho=rho.grad if with_pose else None,
                loss=loss.detach())


def _our_grads(sc, wc, wd, hw, dev, with_pose):
    from vicasplat_b200 import decoder as dec
    from vicasplat_b200.rasterizer import rasterize_views
    d = {k: v.to(dev) for k, v in sc.items()}
    V = d["extrinsics"].shape[0]
    means = d["means"].clone().requires_grad_(True)
    cov6 = dec._cov6(d["covariances"]).clone().requires_grad_(True)
    sh = d["harmonics"].clone().requires_grad_(True)
    op = d["opacities"].clone().requires_grad_(True)
    theta = torch.zeros((V, 3), device=dev, requires_grad=True) if with_pose else None
    rho = torch.zeros((V, 3), device=dev, requires_grad=True) if with_pose else None
    tanfov, view_t, full_t, campos = dec._cameras(d["extrinsics"], d["intrinsics"], d["near"], d["far"])
    color, radii, depth, alpha, nt = rasterize_views(
        means, cov6, op, shs=sh, sh_degree=4, sh_layout="chan_major", viewmatrix=view_t,
        projmatrix=full_t, campos=campos, tanfov=tanfov, bg=torch.zeros(3, device=dev), H=hw, W=hw,
        theta=theta, rho=rho)
    loss = (color * wc.to(dev)).sum() + (depth[:, 0] * wd.to(dev)).sum()
    loss.backward()
    return dict(means=means.grad, cov6=cov6.grad, sh=sh.grad, op=op.grad,
                theta=None if theta is None else theta.grad, rho=None if rho is None else rho.grad,
                loss=loss.detach())


@pytest.mark.parametrize("hw,n_ctx,n_tgt,seed", [(32, 2, 2, 3), (48, 1, 3, 5), (48, 1, 1, 5)])
def test_gradients_match_oracle_autograd(cuda, lib, hw, n_ctx, n_tgt, seed):
    sc, wc, wd = _setup(hw, n_ctx, n_tgt, seed)
    got = _our_grads(sc, wc, wd, hw, cuda, with_pose=True)
    for dtype, tol in ((torch.float32, 3e-3), (torch.float64, 2e-2)):
        ref = _oracle_grads(sc, wc, wd, hw, with_pose=True, dtype=dtype)
        assert abs(got["loss"].item() - ref["loss"].item()) < 1e-2 * max(1.0, abs(ref["loss"].item()))
        for k in ("means", "cov6", "sh", "op", "theta", "rho"):
            assert got[k].shape == ref[k].shape, k
            assert _rel(got[k], ref[k]) < tol, (dtype, k, _rel(got[k], ref[k]))
    # bands above 3 receive no gradient (they are not evaluated)
    assert (got["sh"][..., 16:] == 0).all()


def test_backward_without_pose_and_with_precomputed_colors(cuda, lib):
    from vicasplat_b200 import decoder as dec
    from vicasplat_b200.rasterizer import rasterize_views
    hw = 32
    sc, wc, wd = _setup(hw, 1, 2, 9)
    d = {k: v.to(cuda) for k, v in sc.items()}
    cols = torch.rand((d["means"].shape[0], 3), device=cuda, requires_grad=True)
    means = d["means"].clone().requires_grad_(True)
    tanfov, view_t, full_t, campos = dec._cameras(d["extrinsics"], d["intrinsics"], d["near"], d["far"])
    color, *_ = rasterize_views(means, dec._cov6(d["covariances"]), d["opacities"], colors_precomp=cols,
                                viewmatrix=view_t, projmatrix=full_t, campos=campos, tanfov=tanfov,
                                bg=torch.zeros(3, device=cuda), H=hw, W=hw)
    (color * wc.to(cuda)).sum().backward()
    # oracle with the same precomputed colours
    f = {k: (v.double() if v.is_floating_point() else v) for k, v in sc.items()}
    rc = cols.detach().double().cpu().requires_grad_(True)
    rm = f["means"].clone().requires_grad_(True)
    iu = torch.triu_indices(3, 3)
    imgs = []
    for i in range(2):
        fov = rr.get_fov(f["intrinsics"][i:i + 1])
        img, *_ = rr.rasterize_view(rm, f["covariances"][:, iu[0], iu[1]], None, rc, f["opacities"][:, None],
                                    f["extrinsics"][i], (float((0.5 * fov[0, 0]).tan()), float((0.5 * fov[0, 1]).tan())),
                                    0.01, 100.0, hw, hw, torch.zeros(3, dtype=torch.float64), 0)
        imgs.append(img)
    (torch.stack(imgs) * wc.double()).sum().backward()
    assert _rel(cols.grad, rc.grad) < 2e-3
    assert _rel(means.grad, rm.grad) < 2e-3


def test_decoder_plugin_forward_backward_with_pose_deltas(cuda, lib):
    """DecoderSplattingCUDA.forward as ModelWrapper.test_step_align drives it
    (model_wrapper.py:473-482): batch of scenes, per-view pose deltas, gradients to the deltas."""
    from vicasplat_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg, DecoderOutput
    from vicasplat_b200.encoder import Gaussians
    hw, V = 32, 3
    sc, wc, wd = _setup(hw, 1, V, 13)
    d = {k: v.to(cuda) for k, v in sc.items()}
    B = 2
    rep = lambda t: t[None].expand(B, *t.shape).contiguous()
    g = Gaussians(means=rep(d["means"]).view(B, 1, hw, hw, 3), covariances=rep(d["covariances"]).view(B, 1, hw, hw, 3, 3),
                  harmonics=rep(d["harmonics"]).view(B, 1, hw, hw, 3, 25), opacities=rep(d["opacities"]).view(B, 1, hw, hw))
    dec_ = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], False)).to(cuda)
    rot = torch.zeros((B, V, 3), device=cuda, requires_grad=True)
    trans = torch.zeros((B, V, 3), device=cuda, requires_grad=True)
    out = dec_.forward(g, rep(d["extrinsics"]), rep(d["intrinsics"]), rep(d["near"]), rep(d["far"]), (hw, hw),
                       cam_rot_delta=rot, cam_trans_delta=trans)
    assert isinstance(out, DecoderOutput)
    assert out.color.shape == (B, V, 3, hw, hw) and out.depth.shape == (B, V, hw, hw)
    assert torch.equal(out.color[0], out.color[1])
    (out.color * wc.to(cuda)[None]).sum().backward()
    ref = _oracle_grads(sc, wc, torch.zeros_like(wd), hw, with_pose=True, dtype=torch.float32)
    assert _rel(rot.grad[0], ref["theta"]) < 5e-3 and _rel(trans.grad[1], ref["rho"]) < 5e-3
    c2, d2 = dec_.forward(g, rep(d["extrinsics"]), rep(d["intrinsics"]), rep(d["near"]), rep(d["far"]), (hw, hw),
                          return_dict=False)
    assert torch.equal(c2, out.color.detach())
