"""The training step end to end on the GPU (BASELINE configs[2]): the fast path (vicasplat_b200.train_step.
TrainStep: encoder forward, per-scene render + fused MSE + rasterizer backward, hand-written encoder
backward, fused AdamW) against the DROP-IN path a user of the reference takes -- VicaSplat.forward in training
mode (differentiable through torch.autograd), DecoderSplattingCUDA.forward, LossMse, loss.backward().
Both run the same kernels; they differ in plumbing only, so the parameter gradients agree to the
summation order of the atomics."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder_ref as er
from oracle.make_encoder_golden import CASES, synth_inputs


def _setup(dev):
    from vicasplat_b200 import decoder as dec, synthetic
    from vicasplat_b200.encoder import VicaSplat, VicaSplatCfg, default_backbone_cfg
    kw, _, T, _ = CASES["small"]
    cfg = er.EncoderConfig(**kw)
    bb = dict(default_backbone_cfg(), img_size=cfg.img_size, enc_depth=cfg.enc_depth, dec_depth=cfg.dec_depth)
    model = VicaSplat(VicaSplatCfg(backbone=bb)).to(dev)
    model.load_state_dict(er.synth_state_dict(cfg, seed=0), strict=True)
    model.gs_head_dropout = 0.0        # the goldens / the oracle are dropout-free (eval mode)
    B, V, S = 2, 2, cfg.img_size
    image, K = synth_inputs(B, T, S)
    scenes = []
    for b in range(B):
        sc = {k: v.to(dev) for k, v in synthetic.gaussian_scene(T, S, S, V, seed=10 + b).items()}
        sc["cov6"] = dec._cov6(sc["covariances"]).contiguous()
        scenes.append(sc)
    stk = lambda k: torch.stack([s[k] for s in scenes])
    g = torch.Generator().manual_seed(5)
    target = dict(extrinsics=stk("extrinsics"), intrinsics=stk("intrinsics"), near=stk("near"), far=stk("far"),
                  image=torch.rand((B, V, 3, S, S), generator=g).to(dev))
    return model, dict(image=image.to(dev), intrinsics=K.to(dev)), target, scenes


def test_fast_path_equals_the_autograd_drop_in_path(cuda, lib):
    from vicasplat_b200 import decoder as dec
    from vicasplat_b200.encoder import Gaussians
    from vicasplat_b200.loss import mse
    from vicasplat_b200.train_step import TrainStep
    model, context, target, scenes = _setup(cuda)
    B, V = target["image"].shape[:2]
    S = target["image"].shape[-1]
    stk = lambda k: torch.stack([s[k] for s in scenes])
    # ---- drop-in path: the calls the reference's training_step makes (model_wrapper.py:207-246)
    model.train()
    out = model(context, compute_viewspace_depth=False)
    assert out["raw_gaussians"].requires_grad and out["pred_extrins"].requires_grad
    g = out["gaussians"]
    fl = lambda t: t.flatten(1, 3)
    used = Gaussians(means=stk("means") + fl(g.means), covariances=fl(g.covariances),
                     harmonics=stk("harmonics") + fl(g.harmonics),
                     opacities=stk("opacities") + (fl(g.opacities)[..., 0] - 0.5),
                     cov6=stk("cov6") + fl(g.cov6))
    decoder = dec.DecoderSplattingCUDA(dec.DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], False)).to(cuda)
    ren = decoder.forward(used, target["extrinsics"], target["intrinsics"], target["near"], target["far"], (S, S))
    loss_a = mse(ren.color, target["image"], 1.0)
    loss_a.backward()
    grads_a = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    for p in model.parameters():
        p.grad = None
    # ---- fast path
    ts = TrainStep(model, micro_batch=1)          # two micro-batches: exercises the accumulation too

    def override(b, gz):
        s = scenes[b]
        return dict(means=s["means"] + gz["means"], cov6=s["cov6"] + gz["cov6"], sh=s["harmonics"] + gz["sh"],
                    opac=s["opacities"] + (gz["opac"] - 0.5))
    from vicasplat_b200.rasterizer import RasterOverflow
    try:
        loss_b = ts.accumulate(context, target, override_gaussians=override)
    except RasterOverflow:            # a busier scene than the capacity hint had seen: the hint was raised
        loss_b = ts.accumulate(context, target, override_gaussians=override)
    assert abs(loss_a.item() - loss_b.item()) <= 1e-5 * abs(loss_a.item()) + 1e-7
    assert loss_a.item() > 1e-3                   # the synthetic splats are visible: a real render
    worst = 0.0
    for n, p in model.named_parameters():
        if n in grads_a:
            err = ((p.grad - grads_a[n]).norm() / grads_a[n].norm().clamp_min(1e-20)).item()
            worst = max(worst, err)
            assert err < 2e-2, (n, err)
        else:
            assert p.grad is None, n
    assert len(grads_a) == 499
    print(f"[train step] fast path vs autograd path: worst rel-L2 of a parameter gradient {worst:.2e}")
    # ---- and two optimisation steps
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    l1 = ts.step(context, target, override_gaussians=override)
    l2 = ts.step(context, target, override_gaussians=override)
    assert torch.isfinite(l1) and torch.isfinite(l2) and ts.opt.step_count == 2
    moved = [n for n, p in model.named_parameters() if not torch.equal(p, before[n])]
    assert len(moved) == 499, len(moved)
    assert all(torch.isfinite(p).all() for p in model.parameters())
    # the inference engine of the same module sees the updated weights
    model.eval()
    model.invalidate()
    with torch.no_grad():
        o2 = model(context, compute_viewspace_depth=False)
    assert torch.isfinite(o2["raw_gaussians"]).all()


def test_lpips_camera_losses_and_micro_batching_agree(cuda, lib):
    """All three losses of the 8-view experiment (MSE + LPIPS + dual-quaternion camera loss): the gradients of
    a 2-scene batch taken as ONE micro-batch (LPIPS as one sweep over both scenes' images) equal those of two
    micro-batches of one scene (accumulated), to the order of the fp32 atomics / bf16 LPIPS maps."""
    from vicasplat_b200.lpips import LpipsVgg
    from vicasplat_b200.rasterizer import RasterOverflow
    from vicasplat_b200.train_step import TrainStep
    model, context, target, scenes = _setup(cuda)
    model.train()
    B, T = context["image"].shape[:2]
    ext = torch.eye(4, device=cuda).repeat(B, T, 1, 1)        # camera 0 = identity (the data shim's convention)
    ext[:, 1:, :3, 3] = 0.1 * torch.randn((B, T - 1, 3), generator=torch.Generator().manual_seed(2)).to(cuda)
    context = dict(context, extrinsics=ext)
    net = LpipsVgg.stand_in(cuda, seed=3)

    def override(b, gz):
        s = scenes[b]
        return dict(means=s["means"] + gz["means"], cov6=s["cov6"] + gz["cov6"], sh=s["harmonics"] + gz["sh"],
                    opac=s["opacities"] + (gz["opac"] - 0.5))

    def grads(mb):
        ts = TrainStep(model, micro_batch=mb, lpips=net, lpips_weight=0.5, camera_weight=0.1)
        for _ in range(3):
            try:
                loss = ts.accumulate(context, target, override_gaussians=override)
                break
            except RasterOverflow:
                continue
        return loss.item(), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    l2, g2 = grads(2)
    l1, g1 = grads(1)
    assert abs(l1 - l2) <= 2e-3 * abs(l2), (l1, l2)
    assert len(g1) == len(g2) == 499
    errs = sorted(((g1[n] - g2[n]).norm() / g2[n].norm().clamp_min(1e-20)).item() for n in g2)
    worst, median = errs[-1], errs[len(errs) // 2]
    print(f"[train step] MSE + LPIPS + camera loss: one micro-batch vs two, loss {l2:.5f} / {l1:.5f}, gradient rel-L2 "
          f"worst {worst:.2e} median {median:.2e}")
    # Two outcomes, run to run (also between two IDENTICAL calls, scripts/diag_micro_batching.py): ~1e-6 everywhere, or
    # ~1e-3 median / ~1e-2 worst.  The per-frame AdaLN sums are fp32 atomics; their order moves a value by ~1e-7, which
    # now and then flips the bf16 rounding of ONE element of a (frames x 6 D) modulation gradient (2e-5 of that
    # tensor); the decoder's backward amplifies it block by block, most on tensors whose true gradient is ~0 (the key
    # bias of a softmax: cross_attn.projk.bias).  Both are far below the bf16 error of the gradients themselves
    # (1.4e-2 median against the fp32 oracle, tests/test_gpu_model_grad.py).
    assert worst < 5e-2 and median < 5e-3
