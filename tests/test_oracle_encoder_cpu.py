"""Pins oracle/encoder_ref.py (the functional restatement) to golden vectors produced by the
UNMODIFIED reference modules (oracle/make_encoder_golden.py, run in the build container)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import encoder_ref as er
from oracle.make_encoder_golden import CASES, synth_inputs

GOLD = Path(__file__).parent / "golden"


def _run(name):
    kw, B, T, stride = CASES[name]
    cfg = er.EncoderConfig(**kw)
    sd = er.synth_state_dict(cfg, seed=0)
    image, K = synth_inputs(B, T, cfg.img_size)
    with torch.no_grad():
        out = er.forward(sd, image, K, cfg, stages=True)
    return cfg, sd, out, np.load(GOLD / f"encoder_{name}.npz"), stride


def _close(a, b, rtol, atol, what):
    a = a.numpy() if isinstance(a, torch.Tensor) else a
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    assert (err <= tol).all(), f"{what}: max err {err.max():.3e} (max |ref| {np.abs(b).max():.3e})"


@pytest.mark.parametrize("name", ["small", "full2v"])
def test_restatement_matches_reference_golden(name):
    cfg, sd, out, g, s = _run(name)
    assert int(g["n_keys"]) == len(sd)                       # the state_dict key contract
    if name == "full2v":
        assert len(sd) == 847
    _close(out["pred_extrins"], g["pred_extrins"], 1e-4, 1e-5, "pred_extrins")
    _close(out["gaussian_camera_extrins"], g["gaussian_camera_extrins"], 1e-4, 1e-5, "c2w")
    _close(out["camera_tokens"][:, 1:], g["camera_tokens"], 2e-4, 2e-4, "camera tokens")
    raw = out["raw_gaussians"]
    _close(raw[:, :, ::s, ::s], g["raw_sub"], 2e-3, 2e-4, "raw_gaussians")
    _close(raw.mean(dim=(0, 1, 2, 3)), g["raw_mean"], 1e-3, 1e-4, "raw mean")
    _close(raw.std(dim=(0, 1, 2, 3)), g["raw_std"], 1e-3, 1e-4, "raw std")
    gs = out["gaussians"]
    _close(gs["covariances"][:, :, ::s, ::s], g["cov_sub"], 2e-3, 1e-9, "covariances")
    _close(gs["harmonics"][:, :, ::s, ::s], g["sh_sub"], 2e-3, 2e-5, "harmonics")
    _close(gs["opacities"][:, :, ::s, ::s], g["opac_sub"], 1e-3, 1e-5, "opacities")
    inter = out["intermediates"]
    assert len(inter) == cfg.dec_depth + 1
    _close(np.array([t.mean().item() for t in inter]), g["inter_mean"], 1e-3, 1e-4, "inter mean")
    _close(np.array([t.std().item() for t in inter]), g["inter_std"], 1e-3, 1e-4, "inter std")
    _close(torch.stack([t[0, -1, -1, :64] for t in inter]), g["inter_last_row"], 2e-3, 5e-4,
           "intermediate rows")


def test_golden_is_not_degenerate():
    g = np.load(GOLD / "encoder_small.npz")
    # modulation / pose paths are exercised: poses differ from identity, features have spread
    assert np.abs(g["pred_extrins"] - np.array([0, 0, 0, 1, 0, 0, 0, 0])).max() > 1e-3
    assert (g["raw_std"] > 1e-3).all()
    assert (g["inter_std"] > 0.1).all()


def test_oracle_autograd_matches_reference_gradients():
    """The oracle of the BACKWARD path: torch.autograd over oracle/encoder_ref.encode_image against
    the gradients of the unmodified reference modules (oracle/make_encoder_grad_golden.py) for all 30
    parameters of the image encoder in the small case.  The GPU tests of the hand-written backward
    pass (tests/test_gpu_encoder_grad.py) compare against this autograd-over-the-oracle."""
    from oracle import make_encoder_grad_golden as gg
    gold = np.load(GOLD / "encoder_grad_small.npz")
    cfg = er.EncoderConfig(**gg.CASE)
    sd = er.synth_state_dict(cfg, seed=0)
    keys = gg.path_keys(sd)
    assert len(keys) == 30 and all("norm/" + k in gold.files for k in keys)
    for k in keys:
        sd[k].requires_grad_(True)
    image, K = synth_inputs(1, gg.FRAMES, cfg.img_size)
    x, _ = er.encode_image(sd, image[0], K[0], cfg)
    _close(x.detach()[:, ::4, ::16], gold["out_sub"], 1e-4, 1e-4, "encoder output")
    (x * gg.output_grad(x.shape)).sum().backward()
    for k in keys:
        g = sd[k].grad
        assert g is not None, k
        ref_norm = float(gold["norm/" + k])
        assert abs(g.double().norm().item() - ref_norm) <= 1e-4 * ref_norm + 1e-7, k
        sample = gold["sample/" + k]
        _close(g.flatten()[::gg.SAMPLE], sample, 1e-3, 1e-4 * max(np.abs(sample).max(), 1e-6), k)


def test_oracle_autograd_matches_reference_gradients_whole_model():
    """Oracle of the decoder / DPT-head / adapter backward (the next rows to be built): autograd over
    oracle/encoder_ref.forward against the gradients of the unmodified reference's VicaSplat.forward
    for all 499 parameters that receive one (oracle/make_model_grad_golden.py)."""
    from oracle import make_model_grad_golden as mg2
    gold = np.load(GOLD / "model_grad_small.npz")
    kw, B, T, _ = CASES["small"]
    cfg = er.EncoderConfig(**kw)
    sd = er.synth_state_dict(cfg, seed=0)
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    image, K = synth_inputs(B, T, cfg.img_size)
    loss = mg2.loss_of(er.forward(sd, image, K, cfg))
    assert abs(loss.item() - float(gold["loss"])) <= 1e-4 * abs(float(gold["loss"]))
    loss.backward()
    # the reference shares scratch.layerK_rn with scratch.layer_rn.(K-1): the oracle reads one of the
    # two equal state_dict entries, so the gradient of a shared parameter may sit on its twin
    def grad_of(k):
        if sd[k].grad is not None:
            return sd[k].grad
        for a, b in ((f".scratch.layer{i + 1}_rn.", f".scratch.layer_rn.{i}.") for i in range(4)):
            for src, dst in ((a, b), (b, a)):
                if src in k and sd.get(k.replace(src, dst)) is not None and sd[k.replace(src, dst)].grad is not None:
                    return sd[k.replace(src, dst)].grad
        return None
    keys = [f[len("norm/"):] for f in gold.files if f.startswith("norm/")]
    assert len(keys) == 499
    worst = 0.0
    for k in keys:
        g = grad_of(k)
        assert g is not None, k
        ref_norm = float(gold["norm/" + k])
        assert abs(g.double().norm().item() - ref_norm) <= 2e-3 * ref_norm + 1e-6, (k, g.norm().item(), ref_norm)
        sample = gold["sample/" + k]
        mine = g.flatten()[::mg2.SAMPLE].numpy()
        # error of the sample relative to the RMS size of this gradient (a sample of a small tensor
        # is one or two elements, which may happen to be near zero)
        scale = ref_norm * np.sqrt(len(sample) / g.numel())
        err = np.linalg.norm(mine - sample) / max(scale, 1e-12)
        worst = max(worst, err)
        assert err <= 5e-3, (k, err)
