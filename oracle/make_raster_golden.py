"""Golden vectors for the rasterizer from the REAL upstream extension -- for the day it is reachable.

The reference renders through ``diff_gaussian_rasterization`` (pip git dependency, no commit pinned:
/root/reference/requirements.txt:17; call site src/model/decoder/cuda_splatting.py:207-235).  Its source is
not under /root/reference and the package is not in this image, so ``oracle/raster_ref.py`` is a restatement
whose eight behavioural constants are UNPINNED.  On any machine where

    python -c "import diff_gaussian_rasterization"

works (a CUDA box with the package installed), run

    python oracle/make_raster_golden.py        # writes tests/golden/raster_upstream.npz

and commit the file: tests/test_gpu_raster.py::test_upstream_golden then holds BOTH the CUDA kernels and the
CPU oracle to it (it is skipped while the file is absent), and each probe below pins one of the open
constants:

    probe            pins
    sh_degree4       H1: are SH bands > 3 evaluated when sh_degree = 4 is passed with 25 coefficients
    near_cull        the hard-coded view-space cull (z <= 0.2) vs the `near` plane
    depth            depth = sum depth * alpha * T, un-normalised
    thresholds       alpha < 1/255 skipped, alpha clamp 0.99, stop at T < 1e-4
    lowpass          +0.3 px^2 on the 2-D covariance diagonal
    radius           ceil(3 sigma) radius / tile rectangle
    tie_order        blend order of exactly equal depths (Gaussian index)
    n_touched        definition of the per-Gaussian touch counter

Every probe is a small seeded scene rendered through the reference's own ``render_cuda`` argument
conventions (restated in oracle.raster_ref.render_cuda_ref); inputs are regenerated from the seeds by the
test, only the upstream outputs are stored.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def probes():
    """name -> dict(scene kwargs for oracle.raster_ref.synthetic_scene, tweak(scene) -> scene)."""
    def ident(sc):
        return sc

    def big(sc):
        sc["covariances"] = sc["covariances"] * 400.0
        return sc

    def near(sc):      # half of the Gaussians between the near plane (0.01) and the suspected 0.2 cull
        sc["means"][::2, 2] = 0.05 + 0.1 * torch.rand(sc["means"][::2].shape[0], generator=torch.Generator().manual_seed(3))
        return sc

    def ties(sc):      # pairs of Gaussians at exactly the same depth
        sc["means"][1::2] = sc["means"][0::2]
        return sc

    def opaque(sc):    # saturating opacities: alpha clamp and early stop
        sc["opacities"] = torch.full_like(sc["opacities"], 0.999)
        return sc

    def band4(sc):     # only band-4 coefficients non-zero: any colour other than 0.5 means they are used
        sc["harmonics"][..., :16] = 0
        sc["harmonics"][..., 16:] = 1.0
        return sc

    base = dict(n_ctx=1, h=48, w=48, n_tgt=2)
    return {
        "plain": (dict(base, seed=21), ident),
        "sh_degree4": (dict(base, seed=22), band4),
        "near_cull": (dict(base, seed=23), near),
        "depth": (dict(base, seed=24), ident),
        "thresholds": (dict(base, seed=25), opaque),
        "lowpass": (dict(base, seed=26, depth_range=(15.0, 20.0)), ident),       # sub-pixel splats
        "radius": (dict(base, seed=27, depth_range=(0.5, 3.0)), big),
        "tie_order": (dict(base, seed=28), ties),
        "n_touched": (dict(base, seed=29), ident),
    }


def build_scene(name):
    from oracle import raster_ref as rr
    kw, tweak = probes()[name]
    kw = dict(kw)
    h = kw.pop("h"); w = kw.pop("w")
    return tweak(rr.synthetic_scene(kw.pop("n_ctx"), h, w, kw.pop("n_tgt"), **kw)), h, w


def render_upstream(sc, h, w, dev):
    """one view at a time through the upstream package, with the reference's conventions
    (cuda_splatting.py:180-235)."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    from oracle import raster_ref as rr
    d = {k: v.to(dev) for k, v in sc.items()}
    fov = rr.get_fov(d["intrinsics"])
    proj = rr.get_projection_matrix(d["near"], d["far"], fov[:, 0], fov[:, 1]).transpose(1, 2)
    view = torch.linalg.inv(d["extrinsics"]).transpose(1, 2)
    full = view @ proj
    row, col = torch.triu_indices(3, 3)
    outs = []
    for i in range(view.shape[0]):
        st = GaussianRasterizationSettings(
            image_height=h, image_width=w, tanfovx=(0.5 * fov[i, 0]).tan().item(),
            tanfovy=(0.5 * fov[i, 1]).tan().item(), bg=torch.zeros(3, device=dev), scale_modifier=1.0,
            viewmatrix=view[i], projmatrix=full[i], projmatrix_raw=proj[i], sh_degree=4,
            campos=d["extrinsics"][i, :3, 3], prefiltered=False, debug=False)
        image, radii, depth, opacity, n_touched = GaussianRasterizer(st)(
            means3D=d["means"], means2D=torch.zeros_like(d["means"]), shs=d["harmonics"].transpose(-1, -2).contiguous(),
            colors_precomp=None, opacities=d["opacities"][..., None], cov3D_precomp=d["covariances"][:, row, col],
            theta=torch.zeros(3, device=dev), rho=torch.zeros(3, device=dev))
        outs.append(dict(image=image, radii=radii, depth=depth, opacity=opacity, n_touched=n_touched))
    return {k: torch.stack([o[k] for o in outs]).cpu().numpy() for k in outs[0]}


def main():
    try:
        import diff_gaussian_rasterization  # noqa: F401
    except ImportError:
        print("diff_gaussian_rasterization is not importable here: nothing written (parity stays UNPINNED)")
        return 1
    dev = torch.device("cuda:0")
    data = {}
    for name in probes():
        sc, h, w = build_scene(name)
        for k, v in render_upstream(sc, h, w, dev).items():
            data[f"{name}/{k}"] = v
    path = ROOT / "tests" / "golden" / "raster_upstream.npz"
    np.savez_compressed(path, **data)
    print("wrote", path, path.stat().st_size, "bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
