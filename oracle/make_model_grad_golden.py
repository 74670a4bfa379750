"""Golden GRADIENTS of the whole encoder plugin from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_model_grad_golden.py         # writes tests/golden/model_grad_small.npz

``VicaSplat.forward`` of the reference (small case of make_encoder_golden.py, seeded weights and clip)
runs under torch.autograd on CPU; the loss is a seeded linear functional of everything the training
step consumes (raw Gaussian parameters, adapter outputs, predicted dual quaternions), see ``loss_of``.
For every parameter that receives a gradient: L2 norm + a strided sample.  This pins the oracle of
the decoder / DPT-head / adapter backward (``tests/test_oracle_encoder_cpu.py``), the next rows to be
built on the GPU; the image-encoder part is also covered by make_encoder_grad_golden.py.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
SAMPLE = 4099                                   # prime stride of the stored gradient samples


def loss_of(out) -> torch.Tensor:
    """sum_i <output_i, D_i> with seeded D_i ~ N(0, 1); outputs by name so that the reference's dict
    and the oracle's dict give the same functional."""
    g = out["gaussians"]
    get = (lambda k: getattr(g, k)) if not isinstance(g, dict) else (lambda k: g[k])
    terms = [("raw_gaussians", out["raw_gaussians"]), ("pred_extrins", out["pred_extrins"]),
             ("means", get("means")), ("covariances", get("covariances")),
             ("harmonics", get("harmonics")), ("opacities", get("opacities"))]
    total = 0.0
    for i, (_, t) in enumerate(terms):
        d = torch.randn(t.shape, generator=torch.Generator().manual_seed(900 + i), dtype=t.dtype)
        total = total + (t * d).sum()
    return total


def main():
    os.chdir("/tmp")
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT))
    from oracle import make_encoder_golden as mg
    mg.install_stubs()
    from oracle import encoder_ref as er
    kw, B, T, _ = mg.CASES["small"]
    cfg = er.EncoderConfig(**kw)
    model = mg.build_reference(cfg)
    model.load_state_dict(er.synth_state_dict(cfg, seed=0), strict=True)
    image, K = mg.synth_inputs(B, T, cfg.img_size)
    out = model({"image": image, "intrinsics": K}, compute_viewspace_depth=False)
    loss = loss_of(out)
    loss.backward()
    data = {"loss": np.float64(loss.item())}
    n = 0
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        n += 1
        data["norm/" + k] = np.float64(p.grad.double().norm().item())
        data["sample/" + k] = p.grad.flatten()[::SAMPLE].numpy().copy()
    path = ROOT / "tests" / "golden" / "model_grad_small.npz"
    np.savez_compressed(path, **data)
    print(n, "parameter gradients,", path.stat().st_size, "bytes, loss", loss.item())


if __name__ == "__main__":
    main()
