"""Golden GRADIENTS of the image encoder from the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python oracle/make_encoder_grad_golden.py       # writes tests/golden/encoder_grad_small.npz

The reference's ``VicaNet.intrinsic_encoder`` + ``_encode_image`` (backbone_vica.py:450-480,535-541)
run under ``torch.autograd`` on CPU with the seeded weights / inputs of ``make_encoder_golden.py``; the
loss is ``sum(output * D)`` with a seeded D, so ``D`` is the output gradient the tests feed to the
hand-written backward pass.  Stored: the output (sub-sampled), and for EVERY parameter of the path its
gradient's L2 norm plus a strided sample -- ``tests/test_oracle_encoder_cpu.py`` checks autograd over
the oracle restatement against them, and the GPU tests check the CUDA path against that oracle.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

CASE = dict(img_size=64, enc_depth=2, dec_depth=10)     # the "small" case of make_encoder_golden.py
FRAMES, SEED_D, SAMPLE = 3, 1234, 257                     # gradient sample stride (prime)


def path_keys(sd):
    return [k for k in sd if k.startswith(("backbone.enc_blocks.", "backbone.enc_norm.",
                                           "backbone.patch_embed.", "backbone.intrinsic_encoder."))]


def output_grad(shape):
    return torch.randn(shape, generator=torch.Generator().manual_seed(SEED_D))


def main():
    os.chdir("/tmp")
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT))
    from oracle import make_encoder_golden as mg
    mg.install_stubs()
    from oracle import encoder_ref as er
    cfg = er.EncoderConfig(**CASE)
    sd = er.synth_state_dict(cfg, seed=0)
    model = mg.build_reference(cfg)
    model.load_state_dict(sd, strict=True)
    image, K = mg.synth_inputs(1, FRAMES, cfg.img_size)
    bb = model.backbone.train(False)
    for p in bb.parameters():
        p.requires_grad_(True)
    emb = bb.intrinsic_encoder(K.flatten(2)).reshape(FRAMES, 1, -1)
    x, _ = bb._encode_image(image[0], emb)
    D = output_grad(x.shape)
    (x * D).sum().backward()
    named = dict(model.named_parameters())
    data = dict(out_sub=x.detach()[:, ::4, ::16].numpy())
    keys = path_keys(sd)
    for k in keys:
        g = named[k].grad
        assert g is not None, k
        data["norm/" + k] = np.float64(g.double().norm().item())
        data["sample/" + k] = g.flatten()[::SAMPLE].numpy().copy()
    out = ROOT / "tests" / "golden" / "encoder_grad_small.npz"
    np.savez_compressed(out, **data)
    print(len(keys), "parameters", out.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
