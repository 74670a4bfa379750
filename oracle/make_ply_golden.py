"""Golden records of the reference's ``export_ply`` (src/model/ply_export.py:31-90) run UNMODIFIED on seeded
Gaussians.  ``plyfile`` is absent from the image, so it is replaced by a recorder that keeps the structured
array the reference hands to ``PlyElement.describe`` (field names + values = everything that reaches the
file).  Run in the build container only:

    python oracle/make_ply_golden.py          # writes tests/golden/ply_export.npz
"""
from __future__ import annotations

import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")


def seeded_gaussians(n=500, d_sh=25, seed=3):
    g = torch.Generator().manual_seed(seed)
    means = torch.randn((n, 3), generator=g)
    scales = torch.rand((n, 3), generator=g) * 0.05 + 1e-3
    rot = torch.randn((n, 4), generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)
    sh = torch.randn((n, 3, d_sh), generator=g) * 0.3
    opac = torch.rand((n,), generator=g) * 0.9
    opac[::7] = 0.001                                  # pruned
    return means, scales, rot, sh, opac


def main():
    os.chdir("/tmp")
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT))
    captured = {}
    pf = types.ModuleType("plyfile")

    class PlyElement:
        @staticmethod
        def describe(arr, name):
            captured["elements"], captured["name"] = arr.copy(), name
            return arr

    class PlyData:
        def __init__(self, elements):
            pass

        def write(self, path):
            captured["path"] = str(path)
    pf.PlyElement, pf.PlyData = PlyElement, PlyData
    sys.modules["plyfile"] = pf
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_ply_export", REF / "src" / "model" / "ply_export.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    means, scales, rot, sh, opac = seeded_gaussians()
    out = {}
    for dc_only in (False, True):
        mod.export_ply(torch.eye(4), means, scales, rot, sh, opac, Path("/tmp/_ref_ply/x.ply"), save_sh_dc_only=dc_only)
        el = captured["elements"]
        tag = "dc" if dc_only else "full"
        out[f"{tag}/names"] = np.array(el.dtype.names)
        out[f"{tag}/records"] = np.stack([el[n] for n in el.dtype.names], axis=1)
    path = ROOT / "tests" / "golden" / "ply_export.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
