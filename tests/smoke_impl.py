"""One small invocation of the hot path on cuda:0, checked against the oracle."""
import torch


def run_smoke():
    from oracle import raster_ref as rr
    from vicasplat_b200.decoder import render_cuda
    dev = torch.device("cuda:0")
    hw = 48
    sc = rr.synthetic_scene(1, hw, hw, 2, seed=4)
    rc, rd = rr.render_cuda_ref(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"],
                                (hw, hw), torch.zeros((2, 3)), sc["means"],
                                sc["covariances"], sc["harmonics"], sc["opacities"])
    d = {k: v.to(dev) for k, v in sc.items()}
    c, dep = render_cuda(d["extrinsics"], d["intrinsics"], d["near"], d["far"], (hw, hw),
                         torch.zeros((2, 3), device=dev), d["means"], d["covariances"],
                         d["harmonics"], d["opacities"])
    torch.cuda.synchronize()
    ec = (c.cpu().double() - rc).abs()
    ok = (ec <= 1e-4).double().mean().item()
    assert ok > 0.999, f"raster smoke mismatch: {ok}"
    print(f"smoke: raster ok ({ok:.5f} of pixels within 1e-4, max {ec.max():.2e})")


    # encoder: a shallow-depth VicaSplat on a 3-frame 64x64 clip against the fp32 oracle restatement
    from oracle import encoder_ref as er
    from vicasplat_b200.encoder import VicaSplat, VicaSplatCfg, default_backbone_cfg
    cfg = er.EncoderConfig(img_size=64, enc_depth=2, dec_depth=10)
    bb = dict(default_backbone_cfg(), img_size=64, enc_depth=2, dec_depth=10)
    model = VicaSplat(VicaSplatCfg(backbone=bb)).to(dev).eval()
    sd = er.synth_state_dict(cfg, seed=0)
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(7)
    image = torch.rand((1, 3, 3, 64, 64), generator=g) * 2 - 1
    K = torch.tensor([[0.86, 0, 0.5], [0, 0.86, 0.5], [0, 0, 1.0]]).expand(1, 3, 3, 3).clone()
    out = model({"image": image.to(dev), "intrinsics": K.to(dev)}, compute_viewspace_depth=False)
    with torch.no_grad():
        ref = er.forward(sd, image, K, cfg)
    torch.cuda.synchronize()
    e_pose = (out["pred_extrins"].cpu() - ref["pred_extrins"]).abs().max().item()
    a, b = out["raw_gaussians"][..., 3:].cpu(), ref["raw_gaussians"][..., 3:]
    e_raw = ((a - b).norm() / b.norm()).item()
    assert e_pose < 2e-2 and e_raw < 3e-2, (e_pose, e_raw)
    print(f"smoke: encoder ok (pose abs err {e_pose:.2e}, Gaussian-parameter rel-L2 {e_raw:.2e})")

    # encoder training path: one ViT block forward + hand-written backward against torch.autograd
    # over the oracle block (fp32, CPU)
    from vicasplat_b200 import encoder_grad as eg
    key = "backbone.enc_blocks.0"
    bsd = {k: v.clone() for k, v in sd.items() if k.startswith(key + ".")}
    w = eg.pack_block(bsd, key, dev)
    gacc = eg.zero_grads(w)
    lay = eg.FrameLayout.make(2, 4, 4, cfg.enc_num_heads, dev)
    x = torch.randn((2 * lay.n, cfg.enc_embed_dim), generator=g)
    dy = torch.randn((2 * lay.n, cfg.enc_embed_dim), generator=g)
    saved = eg.Saved()
    eg.block_forward(x.to(dev), w, lay, saved)
    dx = eg.block_backward(dy.to(dev), w, gacc, lay, saved)
    for v in bsd.values():
        v.requires_grad_(True)
    xr = x.view(2, lay.n, -1).clone().requires_grad_(True)
    er.enc_block(bsd, key, xr, er.positions(2, 4, 4, True), cfg).backward(dy.view(2, lay.n, -1))
    torch.cuda.synchronize()
    rel = lambda a, b: ((a.cpu().float() - b).norm() / b.norm()).item()
    e_dx = rel(dx, xr.grad.reshape(dx.shape))
    e_dw = rel(gacc["attn.qkv.weight"], bsd[key + ".attn.qkv.weight"].grad)
    assert e_dx < 2e-2 and e_dw < 3e-2, (e_dx, e_dw)
    print(f"smoke: encoder block backward ok (dx rel-L2 {e_dx:.2e}, d qkv.weight rel-L2 {e_dw:.2e})")

    # the whole training path: one TrainStep (encoder fwd, render, MSE, raster bwd, hand-written encoder bwd,
    # fused AdamW) on a 2-scene toy batch (its gradients are held to the oracle by tests/test_gpu_model_grad.py)
    from vicasplat_b200 import decoder as dec, synthetic
    from vicasplat_b200.rasterizer import RasterOverflow
    from vicasplat_b200.train_step import TrainStep
    model.train()
    model.gs_head_dropout = 0.0
    ts = TrainStep(model, micro_batch=1)
    B, V, S, T = 2, 2, 64, 3
    image2 = torch.rand((B, T, 3, S, S), generator=g) * 2 - 1
    K2 = K.expand(B, T, 3, 3).contiguous()
    scenes = []
    for b_ in range(B):
        sc_ = {k: v.to(dev) for k, v in synthetic.gaussian_scene(T, S, S, V, seed=20 + b_).items()}
        sc_["cov6"] = dec._cov6(sc_["covariances"]).contiguous()
        scenes.append(sc_)
    stk = lambda k: torch.stack([s_[k] for s_ in scenes])
    target = dict(extrinsics=stk("extrinsics"), intrinsics=stk("intrinsics"), near=stk("near"), far=stk("far"),
                  image=torch.rand((B, V, 3, S, S), generator=g).to(dev))
    ctx = dict(image=image2.to(dev), intrinsics=K2.to(dev))

    def override(b_, gz):
        s_ = scenes[b_]
        return dict(means=s_["means"] + gz["means"], cov6=s_["cov6"] + gz["cov6"], sh=s_["harmonics"] + gz["sh"],
                    opac=s_["opacities"] + (gz["opac"] - 0.5))
    before = [p.detach().clone() for p in model.parameters()]
    for attempt in range(3):
        try:
            loss = ts.step(ctx, target, override_gaussians=override)
            break
        except RasterOverflow:
            continue
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and loss.item() > 0
    moved = sum(int(not torch.equal(p, q)) for p, q in zip(model.parameters(), before))
    assert moved >= 499, moved
    assert all(torch.isfinite(p).all() for p in model.parameters())
    print(f"smoke: training step ok (loss {loss.item():.4f}, {ts.opt.step_count} optimizer step, "
          f"{len(ts.eng.buckets)} gradient buckets)")
