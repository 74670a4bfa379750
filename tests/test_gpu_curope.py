"""GPU parity of the curope drop-in (vicasplat_b200.curope) against the reference's own CPU
implementation (golden vectors from oracle/_ref) -- fp32 tolerance 1e-5 as SURVEY.md §7 asks."""
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = np.load(Path(__file__).parent / "golden" / "rope_2d.npz")


@pytest.mark.parametrize("name", ["enc", "dec", "d32"])
def test_rope_2d_matches_reference_cpu(cuda, lib, name):
    from vicasplat_b200 import curope
    tok = torch.from_numpy(GOLD[f"{name}_tok"]).to(cuda)
    pos = torch.from_numpy(GOLD[f"{name}_pos"]).to(cuda)
    t = tok.clone()
    assert curope.rope_2d(t, pos, 100.0, 1.0) is None            # in place, returns nothing
    assert (t.cpu() - torch.from_numpy(GOLD[f"{name}_fwd"])).abs().max() < 1e-5
    curope.rope_2d(t, pos, 100.0, -1.0)                            # backward = inverse rotation
    assert (t - tok).abs().max() < 1e-5
    for dt, tol in ((torch.float16, 4e-3), (torch.bfloat16, 3e-2)):
        th = tok.to(dt)
        curope.rope_2d(th, pos, 100.0, 1.0)
        assert (th.float().cpu() - torch.from_numpy(GOLD[f"{name}_fwd"])).abs().max() < tol


def test_module_on_strided_qkv_view_with_autograd(cuda, lib):
    """used exactly like croco/blocks.py:95-103: q = qkv[:, :, 0] of a fused (B,N,3,H,D) buffer."""
    from vicasplat_b200.curope import cuRoPE2D
    from oracle import encoder_ref as er
    g = torch.Generator().manual_seed(1)
    B, N, H, D = 2, 17, 16, 64
    x = torch.randn((B, N, 3 * H * D), generator=g).to(cuda).requires_grad_(True)
    pos = er.positions(B, 4, 4, True).to(cuda)
    rope = cuRoPE2D(freq=100.0)
    qkv = (x * 1.0).reshape(B, N, 3, H, D).transpose(1, 3)           # (B,H,3,N,D)
    q = qkv[:, :, 0]                                                   # (B,H,N,D) strided view
    out = rope(q, pos)
    ref = er.rope2d((x.detach().reshape(B, N, 3, H, D).transpose(1, 3))[:, :, 0], pos, 100.0)
    assert (out - ref).abs().max() < 1e-5
    w = torch.randn(out.shape, generator=g).to(cuda)
    (out * w).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    (er.rope2d(xr.reshape(B, N, 3, H, D).transpose(1, 3)[:, :, 0], pos, 100.0) * w).sum().backward()
    assert (x.grad - xr.grad).abs().max() < 1e-5


def test_argument_checks_mirror_torch_check(cuda, lib):
    from vicasplat_b200 import curope
    tok = torch.zeros((1, 4, 2, 64), device=cuda)
    pos = torch.zeros((1, 4, 2), dtype=torch.int64, device=cuda)
    with pytest.raises(RuntimeError, match="4 dimensions"):
        curope.rope_2d(tok[0], pos, 100.0, 1.0)
    with pytest.raises(RuntimeError, match="seq_length differs"):
        curope.rope_2d(tok, pos[:, :3], 100.0, 1.0)
    with pytest.raises(RuntimeError, match="must be equal to 2"):
        curope.rope_2d(tok, torch.zeros((1, 4, 3), dtype=torch.int64, device=cuda), 100.0, 1.0)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        curope.rope_2d(torch.zeros((1, 4, 2, 6), device=cuda), pos, 100.0, 1.0)
    with pytest.raises(RuntimeError, match="same device"):
        curope.rope_2d(tok, pos.cpu(), 100.0, 1.0)
    curope.rope_2d(tok[:, :0], pos[:, :0], 100.0, 1.0)                 # empty: no-op
