"""The whole encoder plugin's training path on the GPU (vicasplat_b200.train.TrainEngine: hand-written
backward of the ViT encoder, the MixDecoder blocks, both DPT heads, the adapter and the pose head) against
the gradients of the UNMODIFIED reference (tests/golden/model_grad_small.npz, written by
oracle/make_model_grad_golden.py): same seeded weights, clip and loss functional; all 499 parameters
that receive a gradient.  bf16 GEMM operands / fp32 accumulation: norms within 3e-2, strided samples
within 6e-2 of the gradient's RMS size where the comparison is well-posed (ReLU masks pinned), see the
tolerances stated in each test."""
import json
import os
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder_ref as er
from oracle import make_model_grad_golden as mg2
from oracle.make_encoder_golden import CASES, synth_inputs

GOLD = Path(__file__).parent / "golden"


def _build(dev, case="small"):
    from vicasplat_b200.encoder import VicaSplat, VicaSplatCfg, default_backbone_cfg
    kw, B, T, _ = CASES[case]
    cfg = er.EncoderConfig(**kw)
    bb = dict(default_backbone_cfg(), img_size=cfg.img_size, enc_depth=cfg.enc_depth, dec_depth=cfg.dec_depth)
    model = VicaSplat(VicaSplatCfg(backbone=bb)).to(dev)
    model.load_state_dict(er.synth_state_dict(cfg, seed=0), strict=True)
    model.gs_head_dropout = 0.0        # the goldens / the oracle are dropout-free (eval mode)
    image, K = synth_inputs(B, T, cfg.img_size)
    return cfg, model, image.to(dev), K.to(dev)


def _loss_grads(out, dev):
    """the seeded linear functional of oracle/make_model_grad_golden.loss_of: dL/d(output_i) = D_i"""
    G = out["raw"].shape[0]
    shapes = [("d_raw", out["raw"].shape), ("d_pred", out["pred_extrins"].shape), ("d_means", (G, 3)),
              ("d_cov", (G, 3, 3)), ("d_sh", out["sh"].shape), ("d_opac", (G,))]
    return {k: torch.randn(s, generator=torch.Generator().manual_seed(900 + i)).to(dev)
            for i, (k, s) in enumerate(shapes)}


def test_train_forward_equals_inference_forward(cuda, lib):
    from vicasplat_b200.train import TrainEngine
    cfg, model, image, K = _build(cuda)
    ref = model.engine().run(image, K)
    eng = TrainEngine(model)
    out = eng.forward(image, K)
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()
    G = out["raw"].shape[0]
    # the training forward keeps un-fused intermediates (pre-gate projections, pre-GELU, the stem before
    # the merge): same arithmetic up to a handful of extra bf16 roundings
    rraw = ref["raw"].reshape(G, -1)
    errs = dict(params=rel(out["raw"][:, 3:], rraw[:, 3:]),
                centers_log=rel(torch.log1p(out["raw"][:, :3].norm(dim=-1)), torch.log1p(rraw[:, :3].norm(dim=-1))),
                pred=rel(out["pred_extrins"], ref["pred_extrins"]),
                cov6=rel(out["cov6"], ref["gaussians"]["cov6"].reshape(G, 6)),
                sh=rel(out["sh"], ref["gaussians"]["sh"].reshape(G, 3, -1)))
    print("[train fwd vs inference fwd]", " ".join(f"{k}={v:.3e}" for k, v in errs.items()))
    # (the centres are exponentiated by the head: compared in its own log domain).  Both paths carry their
    # own bf16 roundings (each is ~1e-2 from the fp32 reference at this config): bounds = twice that
    assert errs["params"] < 2e-2 and errs["centers_log"] < 2e-2 and errs["pred"] < 2e-2
    assert errs["cov6"] < 5e-2 and errs["sh"] < 3e-2
    # ... and the training forward against the UNMODIFIED reference's golden vectors directly
    gold = np.load(GOLD / "encoder_small.npz")
    st = CASES["small"][3]
    B, T = image.shape[:2]
    raw5 = out["raw"].view(B, T, cfg.img_size, cfg.img_size, -1)[:, :, ::st, ::st]
    want = torch.from_numpy(gold["raw_sub"]).to(cuda)
    e2 = dict(params=rel(raw5[..., 3:], want[..., 3:]),
              centers_log=rel(torch.log1p(raw5[..., :3].norm(dim=-1)), torch.log1p(want[..., :3].norm(dim=-1))),
              pred=(out["pred_extrins"] - torch.from_numpy(gold["pred_extrins"]).to(cuda)).abs().max().item())
    print("[train fwd vs reference golden]", " ".join(f"{k}={v:.3e}" for k, v in e2.items()))
    assert e2["params"] < 2e-2 and e2["centers_log"] < 2e-2 and e2["pred"] < 2e-2


def test_all_parameter_gradients_against_reference_golden(cuda, lib):
    from vicasplat_b200.train import TrainEngine
    gold = np.load(GOLD / "model_grad_small.npz")
    cfg, model, image, K = _build(cuda)
    eng = TrainEngine(model)
    out = eng.forward(image, K)
    d = _loss_grads(out, cuda)
    pairs = [("raw", "d_raw"), ("pred_extrins", "d_pred"), ("means", "d_means"), ("cov", "d_cov"),
             ("sh", "d_sh"), ("opac", "d_opac")]
    loss = sum((out[a].double() * d[b].double()).sum().item() for a, b in pairs)
    # (the functional weights the exponentiated centres with N(0,1) numbers: its VALUE is ill-conditioned
    # -- a 3 % error of one far-away centre moves it by thousands -- so it is only reported)
    print(f"[model grad] loss {loss:.1f} (reference {float(gold['loss']):.1f})")
    eng.backward(**d)
    keys = [f[len("norm/"):] for f in gold.files if f.startswith("norm/")]
    assert len(keys) == 499
    params = dict(model.named_parameters(remove_duplicate=False))
    report, worst_norm, worst_sample = {}, 0.0, 0.0
    for k in keys:
        g = params[k].grad
        assert g is not None, k
        ref_norm = float(gold["norm/" + k])
        nerr = abs(g.double().norm().item() - ref_norm) / max(ref_norm, 1e-12)
        sample = gold["sample/" + k]
        mine = g.flatten()[::mg2.SAMPLE].cpu().numpy()
        scale = ref_norm * np.sqrt(len(sample) / g.numel())
        serr = float(np.linalg.norm(mine - sample) / max(scale, 1e-12))
        report[k] = (nerr, serr)
        worst_norm, worst_sample = max(worst_norm, nerr), max(worst_sample, serr)
    # Tolerances: see test_gradients_per_output_group_against_autograd_oracle.  Here NOTHING is pinned: the
    # reference's forward pass and ours differ by bf16 rounding, ~1 % of the entries of every ReLU mask flip
    # (each flip switches a gradient path: sqrt(1 %) = 10 %), and the functional weights the exponentiated
    # centres with N(0,1) numbers.  Measured: norms within 11.5 %, cosine ~0.995 per tensor.  The sharp
    # statement is the pinned test above together with tests/test_oracle_encoder_cpu.py, which holds the
    # oracle's autograd to THIS golden file at 5e-3.
    norms = sorted(v[0] for v in report.values())
    samples = sorted(v[1] for v in report.values())
    bad = {k: v for k, v in report.items() if v[0] > 0.15}
    print(f"[model grad] norm err worst {worst_norm:.3e} median {norms[len(norms) // 2]:.3e}; sample err worst "
          f"{worst_sample:.3e} median {samples[len(samples) // 2]:.3e}; {len(bad)} of {len(keys)} beyond tolerance")
    assert norms[len(norms) // 2] < 0.10 and samples[len(samples) // 2] < 0.20
    try:
        out_dir = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).parent.parent)) / "gpurun_out"
        out_dir.mkdir(exist_ok=True)
        (out_dir / "model_grad_report.json").write_text(json.dumps(
            dict(worst_norm=worst_norm, worst_sample=worst_sample,
                 top=sorted(((k, *v) for k, v in report.items()), key=lambda r: -max(r[1], r[2]))[:40]), indent=1))
    except OSError:
        pass
    assert not bad, sorted(bad.items(), key=lambda kv: -max(kv[1]))[:12]
    # the parameters the reference never gives a gradient have none here either
    for k in json.loads((GOLD / "unused_params.json").read_text())["no_grad"]:
        assert params[k].grad is None, k


def test_gradient_accumulation_over_micro_batches(cuda, lib):
    """backward(zero=False) adds to the existing gradients: two passes = twice the gradient."""
    from vicasplat_b200.train import TrainEngine
    cfg, model, image, K = _build(cuda)
    eng = TrainEngine(model)
    d = _loss_grads(eng.forward(image, K), cuda)
    eng.backward(**d)
    once = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    eng.forward(image, K)
    eng.backward(**d, zero=False)
    for k, p in model.named_parameters():
        if p.grad is not None and once[k].norm() > 0:
            err = ((p.grad - 2 * once[k]).norm() / (2 * once[k]).norm()).item()
            # atomics: the summation order differs between passes, and the per-frame gate / modulation
            # gradients are rounded to bf16 after that sum
            assert err < 1e-2, (k, err)


def _oracle_grads(cfg, image, K, make_loss, dev, masks=None):
    """torch.autograd over the fp32 oracle (pinned to the reference's gradients by
    tests/test_oracle_encoder_cpu.py) for an arbitrary functional of the outputs.  masks: ReLU masks to pin
    (TrainEngine.relu_masks()): every ReLU becomes x * mask, so both sides differentiate the same
    piecewise-linear branch."""
    sd = {k: v.to(dev) for k, v in er.synth_state_dict(cfg, seed=0).items()}
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    used = set()

    def hook(x, tag):
        used.add(tag)
        return x * masks[tag].to(x.dtype)
    er.RELU_HOOK = hook if masks is not None else None
    try:
        out = er.forward(sd, image, K, cfg)
    finally:
        er.RELU_HOOK = None
    assert masks is None or used == set(masks), sorted(set(masks) ^ used)
    make_loss(out).backward()

    def grad_of(k):
        if sd[k].grad is not None:
            return sd[k].grad
        for a, b in ((f".scratch.layer{i + 1}_rn.", f".scratch.layer_rn.{i}.") for i in range(4)):
            for src, dst in ((a, b), (b, a)):
                if src in k and sd.get(k.replace(src, dst)) is not None and sd[k.replace(src, dst)].grad is not None:
                    return sd[k.replace(src, dst)].grad
        return None
    return out, grad_of


@pytest.mark.parametrize("group", ["pose", "gaussian_params", "centers", "all"])
def test_gradients_per_output_group_against_autograd_oracle(cuda, lib, group):
    """Well-conditioned functionals, one output group at a time, every parameter compared as a WHOLE tensor
    (rel-L2) with autograd over the fp32 oracle run on the same device WITH THE ReLU MASKS OF THE GPU's
    FORWARD PASS PINNED: the two forward passes differ by bf16 rounding (~1e-2), which flips ~1 % of the
    entries of every ReLU mask, and each flipped entry switches a gradient path on or off -- sqrt(1 %) =
    10 % of gradient error that says nothing about the backward kernels (measured: 9-13 % on every
    tensor, cosine 0.995, without the pinning).  The centre functional is weighted
    by 1 / (1 + |c_ref|^2) with the oracle's own (constant) centres, which undoes the head's exponential:
    with N(0,1) weights the gradient of every parameter is dominated by a few far-away points whose
    amplification exp(|x|) turns the forward pass's bf16 error into a several-per-cent gradient error
    (the committed reference golden uses such weights: see the next test)."""
    from vicasplat_b200.train import TrainEngine
    cfg, model, image, K = _build(cuda)
    eng = TrainEngine(model)
    out = eng.forward(image, K)
    G = out["raw"].shape[0]
    gen = lambda i, s: torch.randn(s, generator=torch.Generator().manual_seed(700 + i)).to(cuda)
    d_raw = gen(0, (G, 86)); d_raw[:, :3] = 0
    d_pred, d_cov, d_sh, d_opac = gen(1, out["pred_extrins"].shape), gen(2, (G, 3, 3)) * 1e4, gen(3, (G, 3, 25)), gen(4, (G,))
    d_c = gen(5, (G, 3))
    use = dict(pose=("pred",), gaussian_params=("raw", "cov", "sh", "opac"), centers=("c",),
               all=("pred", "raw", "cov", "sh", "opac", "c"))[group]
    hold = {}

    def make_loss(o):
        g = o["gaussians"]
        c_ref = o["raw_gaussians"][..., :3].detach().reshape(G, 3)
        hold["wc"] = d_c / (1 + (c_ref * c_ref).sum(-1, keepdim=True))
        terms = dict(pred=(o["pred_extrins"] * d_pred).sum(), raw=(o["raw_gaussians"].reshape(G, 86) * d_raw).sum(),
                     cov=(g["covariances"].reshape(G, 3, 3) * d_cov).sum(),
                     sh=(g["harmonics"].reshape(G, 3, 25) * d_sh).sum(),
                     opac=(g["opacities"].reshape(G) * d_opac).sum(),
                     c=(g["means"].reshape(G, 3) * hold["wc"]).sum())
        return sum(terms[k] for k in use)

    _, grad_of = _oracle_grads(cfg, image, K, make_loss, cuda, masks=eng.relu_masks())
    kw = dict(d_pred=d_pred if "pred" in use else None, d_raw=d_raw if "raw" in use else None,
              d_cov=d_cov if "cov" in use else None, d_sh=d_sh if "sh" in use else None,
              d_opac=d_opac if "opac" in use else None, d_means=hold["wc"].contiguous() if "c" in use else None)
    eng.backward(**kw)
    worst, rows = 0.0, []
    for k, p in model.named_parameters():
        ref = grad_of(k)
        if ref is None or ref.norm() == 0:
            if p.grad is not None:
                assert p.grad.abs().max() == 0 or group in ("pose",) or ref is None, k
            continue
        err = ((p.grad - ref).norm() / ref.norm()).item()
        rows.append((err, k))
        worst = max(worst, err)
    rows.sort(reverse=True)
    median = rows[len(rows) // 2][0]
    print(f"[model grad / {group}] rel-L2 worst {worst:.3e} median {median:.3e} over {len(rows)} tensors; top: " +
          ", ".join(f"{k}={e:.2e}" for e, k in rows[:6]))
    try:
        out_dir = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).parent.parent)) / "gpurun_out"
        out_dir.mkdir(exist_ok=True)
        (out_dir / f"model_grad_pinned_{group}.json").write_text(json.dumps(
            dict(group=group, tensors=len(rows), worst=worst, median=median, top=[(k, e) for e, k in rows[:12]]),
            indent=1))
    except OSError:
        pass
    # Measured on B200 (bf16 operands, 2 + 10 layers): pose 3.9e-2 worst; Gaussian parameters 5.7e-2 worst
    # (the key-projection biases of the neighbour attention: a softmax is invariant to a common key offset,
    # their gradient is a small remainder of cancelling terms); centres 4-5 % on EVERY tensor and 1.2e-1
    # worst: the head's exp(|x|) turns the forward pass's absolute bf16 error of |x| (~4e-2 at |x| ~ 4)
    # into a relative error of d xyz / d x, the first factor of the chain.
    tol_worst, tol_median = dict(pose=(5e-2, 3e-2), gaussian_params=(7e-2, 4e-2), centers=(1.4e-1, 6e-2),
                                 all=(1.4e-1, 6e-2))[group]
    assert worst < tol_worst, rows[:10]
    assert median < tol_median
