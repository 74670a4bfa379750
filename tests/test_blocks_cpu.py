"""Host-side contract of the drop-in trainable Block (no GPU): constructor surface, state_dict names
of croco/blocks.py:115-130 (as pinned by the oracle's 847-key table), loud failure without CUDA."""
from functools import partial

import pytest
import torch
from torch import nn

from oracle import encoder_ref as er


class _Rope:
    base = 100.0


def _block():
    from vicasplat_b200.blocks import Block
    return Block(1024, 16, 4.0, qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), rope=_Rope())


def test_state_dict_names_and_shapes_match_the_reference_block():
    blk = _block()
    cfg = er.EncoderConfig(enc_depth=1, dec_depth=4)
    want = {k[len("backbone.enc_blocks.0."):]: v for k, v in er.param_shapes(cfg).items()
            if k.startswith("backbone.enc_blocks.0.")}
    got = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    assert got == {k: tuple(v) for k, v in want.items()}
    sd = {k[len("backbone.enc_blocks.0."):]: v for k, v in er.synth_state_dict(cfg, seed=0).items()
          if k.startswith("backbone.enc_blocks.0.")}
    blk.load_state_dict(sd, strict=True)


def test_unsupported_configurations_raise():
    from vicasplat_b200.blocks import Block
    with pytest.raises(NotImplementedError):
        Block(1024, 16, drop_path=0.1, qkv_bias=True, rope=_Rope())
    with pytest.raises(NotImplementedError):
        Block(768, 16, qkv_bias=True, rope=_Rope())              # head_dim 48
    with pytest.raises(NotImplementedError):
        Block(1024, 16, qkv_bias=True, rope=None)


def test_cpu_tensors_fail_loudly():
    blk = _block()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blk(torch.zeros(1, 5, 1024), torch.zeros(1, 5, 2, dtype=torch.long))


def test_synthetic_vit_weights_follow_the_reference_names_and_shapes():
    """scripts/bench_*train*.py draw their weights from vicasplat_b200.synthetic (not from oracle/):
    names and shapes are the reference's for the image encoder."""
    from vicasplat_b200 import synthetic
    from vicasplat_b200.encoder_train import ViTEncoderConfig
    cfg = er.EncoderConfig(enc_depth=2, dec_depth=4)
    want = {k: tuple(v) for k, v in er.param_shapes(cfg).items()
            if k.startswith(("backbone.enc_blocks.", "backbone.enc_norm.", "backbone.patch_embed.",
                             "backbone.intrinsic_encoder."))}
    sd = synthetic.vit_encoder_state_dict(depth=2, seed=1)
    assert {k: tuple(v.shape) for k, v in sd.items()} == want
    again = synthetic.vit_encoder_state_dict(depth=2, seed=1)
    assert all(torch.equal(sd[k], again[k]) for k in sd)               # seeded
    v = ViTEncoderConfig(enc_depth=2)
    assert (v.enc_embed_dim, v.enc_num_heads, v.patch_size, v.ln_eps) == (cfg.enc_embed_dim, cfg.enc_num_heads,
                                                                         cfg.patch_size, cfg.ln_eps)
