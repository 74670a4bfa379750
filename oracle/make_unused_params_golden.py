"""Which parameters of the reference model never receive a gradient?  (the reference
trains under DDP with find_unused_parameters=True, src/main.py:111.)

Run in the build container only (needs /root/reference):

    python oracle/make_unused_params_golden.py      # writes tests/golden/unused_params.json

The UNMODIFIED reference VicaSplat (small case of make_encoder_golden.py, seeded weights) runs forward
under torch.autograd on CPU; the loss touches every output the training step consumes (Gaussian
parameters, predicted poses).  Parameters whose ``.grad`` stays None are recorded: a static-bucket
gradient all-reduce has to leave them out (or feed zeros) while keeping them in ``state_dict``.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path


ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


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
    g = out["gaussians"]
    loss = (out["raw_gaussians"].sum() + out["pred_extrins"].sum() + g.means.sum() + g.covariances.sum()
            + g.harmonics.sum() + g.opacities.sum())
    loss.backward()
    unused = sorted(k for k, p in model.named_parameters() if p.grad is None)
    zero = sorted(k for k, p in model.named_parameters() if p.grad is not None and not p.grad.any())
    n = sum(1 for _ in model.named_parameters())
    (ROOT / "tests" / "golden" / "unused_params.json").write_text(
        json.dumps({"n_parameters": n, "no_grad": unused, "zero_grad": zero}, indent=1) + "\n")
    print(n, "parameters;", len(unused), "without a gradient;", len(zero), "with an all-zero gradient")
    for k in unused:
        print("  ", k)


if __name__ == "__main__":
    main()
