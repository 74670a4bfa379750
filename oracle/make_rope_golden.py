"""Golden vectors for rope_2d from the REFERENCE's own curope.cpp (CPU branch), compiled by
oracle/Makefile into oracle/_ref.  Run in the build container: python oracle/make_rope_golden.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "oracle" / "_ref"))
import curope_ref  # noqa: E402  (the reference extension, CPU build)


def cases():
    g = torch.Generator().manual_seed(20251017)
    out = {}
    for name, (B, N, H, D, hw) in {"enc": (2, 17, 16, 64, 4), "dec": (1, 37, 12, 64, 6),
                                   "d32": (3, 5, 2, 32, 9)}.items():
        tok = torch.randn((B, N, H, D), generator=g)
        pos = torch.randint(0, hw + 1, (B, N, 2), generator=g, dtype=torch.int64)
        out[name] = (tok, pos)
    return out


def main():
    data = {}
    for name, (tok, pos) in cases().items():
        for fwd, tag in ((1.0, "fwd"), (-1.0, "bwd")):
            t = tok.clone()
            curope_ref.rope_2d(t, pos, 100.0, fwd)
            data[f"{name}_{tag}"] = t.numpy()
        data[f"{name}_tok"], data[f"{name}_pos"] = tok.numpy(), pos.numpy()
    path = ROOT / "tests" / "golden" / "rope_2d.npz"
    np.savez_compressed(path, **data)
    print("wrote", path, path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
