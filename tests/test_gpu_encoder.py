"""GPU parity of the encoder path (VicaSplat.forward through the C-ABI kernels) against the fp32
oracle restatement (oracle/encoder_ref.py, itself pinned to the reference's golden vectors) and
against those golden vectors directly.

Numerics: bf16 GEMM operands, fp32 accumulation and fp32 residual streams; the reference runs
TF32 (backbone_vica.py:9).  Tolerances (relative L2 unless noted): per-stage residual streams
<= 1.5e-2, raw Gaussian parameters <= 3e-2, poses <= 2e-2 absolute, centres <= 2e-2 in the head's
pre-exp domain (see _centers_close)."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import encoder_ref as er
from oracle.make_encoder_golden import CASES, synth_inputs

GOLD = Path(__file__).parent / "golden"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def _centers_close(got, want, tol=2e-2):
    """The centre head ends in xyz = dir * expm1(|x|) (postprocess.py:52-61): compare in the head's
    own (pre-exp) domain -- log1p(|c|) and the direction -- because exp turns an absolute error of
    the bf16 pipeline into a relative one of the same size."""
    dg, dw = torch.log1p(got.norm(dim=-1)), torch.log1p(want.norm(dim=-1))
    assert _rel(dg, dw) < tol, _rel(dg, dw)
    cos = torch.nn.functional.cosine_similarity(got, want, dim=-1)
    assert cos.mean() > 0.999 and (cos > 0.98).float().mean() > 0.999, cos.mean()


def _build(name, dev):
    from vicasplat_b200.encoder import VicaSplat, VicaSplatCfg, default_backbone_cfg
    kw, B, T, stride = CASES[name]
    cfg = er.EncoderConfig(**kw)
    bb = dict(default_backbone_cfg(), img_size=cfg.img_size, enc_depth=cfg.enc_depth,
              dec_depth=cfg.dec_depth)
    model = VicaSplat(VicaSplatCfg(backbone=bb)).to(dev).eval()
    sd = er.synth_state_dict(cfg, seed=0)
    assert set(model.state_dict().keys()) == set(sd.keys())
    model.load_state_dict(sd, strict=True)
    image, K = synth_inputs(B, T, cfg.img_size)
    return cfg, model, {k: v.to(dev) for k, v in sd.items()}, image.to(dev), K.to(dev), stride


def test_small_config_stage_by_stage(cuda, lib):
    cfg, model, sd, image, K, _ = _build("small", cuda)
    B, T = image.shape[:2]
    taps = {}
    eng = model.engine()
    out = eng.run(image, K, taps=taps)
    with torch.no_grad():
        ref = er.forward(sd, image, K, cfg, stages=True)
        # oracle encoder stream, block by block
        x, pos = None, None
        Kf = K.reshape(B * T, 3, 3)
        xr, pos = er.encode_image(sd, image.flatten(0, 1), Kf, cfg)
    N = xr.shape[1]
    assert _rel(taps["inter0"].view(B * T, N, -1), xr) < 1e-2
    rpf = N + 1
    for layer in range(1, cfg.dec_depth + 1):
        got = taps[f"dec{layer}"].view(B, T, rpf, -1)[:, :, 1:-1]       # image patches only
        want = ref["intermediates"][layer]
        if layer == cfg.dec_depth:                                       # oracle applied dec_norm
            continue
        assert _rel(got, want) < 1.5e-2, layer
    assert _rel(taps["cam_out"].view(B, T, -1), ref["camera_tokens"]) < 1.5e-2
    assert (out["pred_extrins"] - ref["pred_extrins"]).abs().max() < 2e-2
    assert (out["c2w"] - ref["gaussian_camera_extrins"]).abs().max() < 3e-2
    _centers_close(out["raw"][..., :3], ref["raw_gaussians"][..., :3])
    assert _rel(out["raw"][..., 3:], ref["raw_gaussians"][..., 3:]) < 3e-2
    g, rg = out["gaussians"], ref["gaussians"]
    assert _rel(g["sh"], rg["harmonics"]) < 3e-2
    assert (g["opac"] - rg["opacities"][..., 0]).abs().max() < 6e-2
    assert _rel(g["cov"], rg["covariances"]) < 5e-2


def _record(name, errs):
    """achieved errors -> stdout (pytest -s / -rP) and gpurun_out/encoder_parity_<case>.json"""
    import json
    import os
    print(f"[encoder parity {name}] " + " ".join(f"{k}={v:.3e}" for k, v in errs.items()))
    out = Path(os.environ.get("GRAFT_REPO_ROOT", Path(__file__).parent.parent)) / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        (out / f"encoder_parity_{name}.json").write_text(json.dumps(errs, indent=1))
    except OSError:
        pass


@pytest.mark.parametrize("name", ["small", "full2v", "full8v", "full4v_b2"])
def test_forward_matches_reference_golden(cuda, lib, name):
    """the public plugin call, graph replay path, against vectors from the UNMODIFIED reference:
    toy depth, 2 views, the HEADLINE shape (8 views, 24 + 12 layers, 256 x 256: two-segment neighbour
    attention, 2 064-key video attention, 8 camera rows) and a batch of two different clips."""
    cfg, model, sd, image, K, s = _build(name, cuda)
    g = np.load(GOLD / f"encoder_{name}.npz")
    for rep in range(2):                                                  # 2nd call = graph replay
        out = model({"image": image, "intrinsics": K}, compute_viewspace_depth=False)
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    assert set(out.keys()) >= {"gaussians", "pred_extrins", "pred_intrins", "raw_gaussians",
                               "gaussian_camera_extrins", "gaussian_camera_intrins",
                               "gaussian_centers", "confidence", "context_view_depths"}
    raw = out["raw_gaussians"]
    gs = out["gaussians"]
    errs = dict(
        pred_extrins_abs=(out["pred_extrins"] - t("pred_extrins")).abs().max().item(),
        c2w_abs=(out["gaussian_camera_extrins"] - t("gaussian_camera_extrins")).abs().max().item(),
        raw_params_rel=_rel(raw[:, :, ::s, ::s, 3:], t("raw_sub")[..., 3:]),
        centers_log_rel=_rel(torch.log1p(raw[:, :, ::s, ::s, :3].norm(dim=-1)),
                             torch.log1p(t("raw_sub")[..., :3].norm(dim=-1))),
        sh_rel=_rel(gs.harmonics[:, :, ::s, ::s], t("sh_sub")),
        cov_rel=_rel(gs.covariances[:, :, ::s, ::s], t("cov_sub")),
        opac_abs_max=(gs.opacities[:, :, ::s, ::s] - t("opac_sub")).abs().max().item(),
        opac_abs_mean=(gs.opacities[:, :, ::s, ::s] - t("opac_sub")).abs().mean().item())
    _record(name, errs)
    assert errs["pred_extrins_abs"] < 2e-2
    assert errs["c2w_abs"] < 3e-2
    assert raw.shape == (image.shape[0], image.shape[1], cfg.img_size, cfg.img_size, 86)
    _centers_close(raw[:, :, ::s, ::s, :3], t("raw_sub")[..., :3])
    # SURVEY §7's bf16 budget: rel-L2 <= 2e-2 on the raw Gaussian parameters
    assert errs["raw_params_rel"] < 2e-2
    assert gs.means.shape[-1] == 3 and gs.covariances.shape[-2:] == (3, 3)
    assert gs.harmonics.shape[-2:] == (3, 25) and gs.opacities.shape[-1] == 1
    assert errs["sh_rel"] < 2e-2
    # covariances are QUADRATIC in the scales (R diag(s^2) R^T): twice the relative error of the raw
    # parameters they come from (measured on B200: raw 1.1-1.4e-2 -> cov 3.1-3.9e-2 at full depth)
    assert errs["cov_rel"] < 4e-2
    assert errs["opac_abs_max"] < 6e-2
    assert errs["opac_abs_mean"] < 5e-3


def test_viewspace_depth_and_distill_subset(cuda, lib):
    cfg, model, sd, image, K, _ = _build("small", cuda)
    B, T = image.shape[:2]
    ext = torch.eye(4, device=cuda).expand(B, T, 4, 4).clone()
    ext[:, :, 0, 3] = torch.arange(T, device=cuda) * 0.1
    out = model({"image": image, "intrinsics": K, "extrinsics": ext})
    assert torch.allclose(out["context_view_depths"], out["gaussian_centers"][..., 2], atol=1e-6)
    d = model({"image": image, "intrinsics": K, "extrinsics": ext}, distill=True)
    assert "gaussians" not in d and d["gaussian_centers"].shape == out["gaussian_centers"].shape
    assert _rel(d["gaussian_centers"], out["gaussian_centers"]) < 1e-6


@pytest.mark.parametrize("name", ["small", "full2v", "full8v", "full4v_b2"])
def test_parity_mode_fp16_matches_reference_golden(cuda, lib, name):
    """precision="fp16": the engine's parity mode -- fp16 GEMM / attention operands carry the 10-bit mantissa
    that the reference's TF32 matmuls round to (backbone_vica.py:9), everything else as in the speed mode
    (fp32 accumulation, fp32 residual streams).  Against the UNMODIFIED reference's fp32 golden vectors at
    the full 24 + 12-layer depth; SURVEY.md §7's TF32 budget is rel-L2 <= 2e-3 on the raw Gaussians."""
    cfg, model, sd, image, K, s = _build(name, cuda)
    model.set_precision("fp16")
    g = np.load(GOLD / f"encoder_{name}.npz")
    for rep in range(2):
        out = model({"image": image, "intrinsics": K}, compute_viewspace_depth=False)
    t = lambda k: torch.from_numpy(g[k]).to(cuda)
    raw, gs = out["raw_gaussians"], out["gaussians"]
    assert torch.isfinite(raw).all()
    errs = dict(
        pred_extrins_abs=(out["pred_extrins"] - t("pred_extrins")).abs().max().item(),
        c2w_abs=(out["gaussian_camera_extrins"] - t("gaussian_camera_extrins")).abs().max().item(),
        raw_params_rel=_rel(raw[:, :, ::s, ::s, 3:], t("raw_sub")[..., 3:]),
        centers_log_rel=_rel(torch.log1p(raw[:, :, ::s, ::s, :3].norm(dim=-1)),
                             torch.log1p(t("raw_sub")[..., :3].norm(dim=-1))),
        centers_rel=_rel(raw[:, :, ::s, ::s, :3], t("raw_sub")[..., :3]),
        sh_rel=_rel(gs.harmonics[:, :, ::s, ::s], t("sh_sub")),
        cov_rel=_rel(gs.covariances[:, :, ::s, ::s], t("cov_sub")),
        opac_abs_max=(gs.opacities[:, :, ::s, ::s] - t("opac_sub")).abs().max().item())
    _record(name + "_fp16", errs)
    # measured on B200 at full depth (8 views): raw 1.1e-3, SH 9.4e-4, poses 1.4e-3 abs, covariances 3.5e-3
    # (quadratic in the scales) -- ten times closer than the bf16 speed mode; the toy config is noisier
    assert errs["raw_params_rel"] < 2e-3 and errs["centers_log_rel"] < 2e-3
    assert errs["pred_extrins_abs"] < 2e-3 and errs["c2w_abs"] < 5e-3
    assert errs["sh_rel"] < (2e-3 if name != "small" else 3e-3)
    assert errs["cov_rel"] < 5e-3 and errs["opac_abs_max"] < 6e-3
