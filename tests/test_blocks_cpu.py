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
