"""Pins the rope oracles (plain-C oracle/rope_ref.c and the torch restatement in
oracle/encoder_ref.py) to golden vectors from the reference's own curope.cpp."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import encoder_ref as er

ROOT = Path(__file__).resolve().parent.parent
GOLD = np.load(ROOT / "tests" / "golden" / "rope_2d.npz")
CASES = ["enc", "dec", "d32"]


@pytest.fixture(scope="module")
def cref():
    so = ROOT / "oracle" / "_build" / "librope_ref.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "_build/librope_ref.so"], check=True)
    lib = C.CDLL(str(so))
    lib.rope_2d_ref.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
    return lib


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag,fwd", [("fwd", 1.0), ("bwd", -1.0)])
def test_c_oracle_matches_reference(cref, name, tag, fwd):
    tok = np.ascontiguousarray(GOLD[f"{name}_tok"]).copy()
    pos = np.ascontiguousarray(GOLD[f"{name}_pos"])
    B, N, H, D = tok.shape
    cref.rope_2d_ref(tok.ctypes.data, pos.ctypes.data, B, N, H, D, 100.0, fwd)
    assert np.array_equal(tok, GOLD[f"{name}_{tag}"])          # same libm calls: bit-exact


@pytest.mark.parametrize("name", CASES)
def test_torch_restatement_matches_reference(name):
    tok, pos = torch.from_numpy(GOLD[f"{name}_tok"]), torch.from_numpy(GOLD[f"{name}_pos"])
    out = er.rope2d(tok.permute(0, 2, 1, 3), pos, 100.0).permute(0, 2, 1, 3)
    assert np.abs(out.numpy() - GOLD[f"{name}_fwd"]).max() < 5e-6


def test_forward_then_backward_is_identity():
    for name in CASES:
        tok = torch.from_numpy(GOLD[f"{name}_fwd"]).clone().numpy()
        # the 'bwd' vectors rotate the ORIGINAL tokens by -angle; rotating fwd output back recovers them
        t = torch.from_numpy(tok)
        pos = torch.from_numpy(GOLD[f"{name}_pos"])
        D = t.shape[-1]
        back = er.rope2d(t.permute(0, 2, 1, 3), -pos, 100.0).permute(0, 2, 1, 3)
        assert np.abs(back.numpy() - GOLD[f"{name}_tok"]).max() < 1e-5
