"""GPU parity of the splatting rasterizer (through the C-ABI and the reference-shaped Python
surface) against the fp64 CPU oracle (oracle/raster_ref.py).  Tolerance: 1e-4 absolute on RGB and
depth (BASELINE.json north_star), on the same fp32 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import raster_ref as rr


def _scene(n_ctx, hw, n_tgt, seed, n_gauss=None, depth_range=(1.0, 20.0)):
    return rr.synthetic_scene(n_ctx, hw, hw, n_tgt, seed=seed, n_gauss=n_gauss,
                              depth_range=depth_range)


def _oracle(sc, hw, dtype=torch.float32, bg=None):
    """The oracle in fp32 is the stand-in for the (fp32) upstream extension; in fp64 it is the
    mathematical truth.  On these scenes fp32-oracle vs fp64-oracle already differ by up to ~1e-3
    on ~1% of the pixels (conic = inverse of a nearly singular 2x2 in fp32, then 1/255 threshold
    flips), so the 1e-4 bar is applied against the fp32 oracle and a looser one against fp64."""
    f = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sc.items()}
    V = f["extrinsics"].shape[0]
    bg = torch.zeros((V, 3), dtype=dtype) if bg is None else bg.to(dtype)
    c, d = rr.render_cuda_ref(f["extrinsics"], f["intrinsics"], f["near"], f["far"], (hw, hw), bg,
                              f["means"], f["covariances"], f["harmonics"], f["opacities"])
    return c.double(), d.double()


def _ours(sc, hw, dev, bg=None):
    from vicasplat_b200.decoder import render_cuda
    d = {k: v.to(dev) for k, v in sc.items()}
    V = d["extrinsics"].shape[0]
    bg = torch.zeros((V, 3), device=dev) if bg is None else bg.to(dev)
    return render_cuda(d["extrinsics"], d["intrinsics"], d["near"], d["far"], (hw, hw), bg,
                       d["means"], d["covariances"], d["harmonics"], d["opacities"])


def _check(c, d, rc, rd, atol=1e-4, frac=0.999, worst=5e-3):
    ec = (c.cpu().double() - rc).abs()
    ed = (d.cpu().double() - rd).abs() / rd.abs().clamp_min(1)
    # threshold flips (alpha < 1/255, T < 1e-4) may move isolated pixels by up to ~1/255
    assert (ec <= atol).double().mean() >= frac, f"color: max {ec.max():.2e}, ok {(ec <= atol).double().mean():.5f}"
    assert (ed <= atol).double().mean() >= frac, f"depth: max {ed.max():.2e}, ok {(ed <= atol).double().mean():.5f}"
    assert ec.max() < worst and ed.max() < worst, (ec.max(), ed.max())


def _check64(c, d, sc, hw, bg=None):
    rc, rd = _oracle(sc, hw, torch.float64, bg)
    _check(c, d, rc, rd, atol=1e-4, frac=0.95, worst=5e-2)


@pytest.mark.parametrize("hw,n_ctx,n_tgt,seed", [(64, 2, 3, 1), (48, 1, 2, 2), (80, 2, 1, 3)])
def test_render_matches_oracle(cuda, lib, hw, n_ctx, n_tgt, seed):
    sc = _scene(n_ctx, hw, n_tgt, seed)
    rc, rd = _oracle(sc, hw)
    c, d = _ours(sc, hw, cuda)
    assert c.shape == (n_tgt, 3, hw, hw) and d.shape == (n_tgt, hw, hw)
    _check(c, d, rc, rd)
    _check64(c, d, sc, hw)
    assert rc.abs().max() > 0.2 and rd.max() > 1.0          # the scene is not empty


def test_render_big_splats_and_background(cuda, lib):
    """large Gaussians (many tiles each), close to the camera, non-zero background."""
    sc = _scene(1, 32, 2, 5, depth_range=(0.5, 3.0))
    sc["covariances"] = sc["covariances"] * 400.0
    bg = torch.tensor([[0.2, 0.5, 0.9], [1.0, 0.0, 0.3]])
    rc, rd = _oracle(sc, 32, bg=bg)
    c, d = _ours(sc, 32, cuda, bg=bg)
    _check(c, d, rc, rd)
    _check64(c, d, sc, 32, bg)


def test_per_view_sets_equal_shared(cuda, lib):
    """the reference repeats the Gaussians per view (decoder_splatting_cuda.py:79-95); both modes
    must give bit-identical images."""
    from vicasplat_b200.decoder import render_cuda
    sc = _scene(2, 32, 3, 7)
    d = {k: v.to(cuda) for k, v in sc.items()}
    bg = torch.zeros((3, 3), device=cuda)
    args = (d["extrinsics"], d["intrinsics"], d["near"], d["far"], (32, 32), bg)
    c0, d0 = render_cuda(*args, d["means"], d["covariances"], d["harmonics"], d["opacities"])
    rep = lambda t: t[None].expand(3, *t.shape).contiguous()
    c1, d1 = render_cuda(*args, rep(d["means"]), rep(d["covariances"]), rep(d["harmonics"]),
                         rep(d["opacities"]))
    assert torch.equal(c0, c1) and torch.equal(d0, d1)


def test_reference_shaped_rasterizer_call(cuda, lib):
    """GaussianRasterizationSettings / GaussianRasterizer used exactly as cuda_splatting.py:207-235."""
    from vicasplat_b200.rasterizer import GaussianRasterizationSettings, GaussianRasterizer
    from vicasplat_b200.decoder import get_fov, get_projection_matrix
    hw = 48
    sc = _scene(1, hw, 1, 9)
    rc, rd = _oracle(sc, hw)
    d = {k: v.to(cuda) for k, v in sc.items()}
    fov = get_fov(d["intrinsics"])
    proj = get_projection_matrix(d["near"], d["far"], fov[:, 0], fov[:, 1]).transpose(1, 2)
    view = torch.linalg.inv(d["extrinsics"]).transpose(1, 2)
    settings = GaussianRasterizationSettings(
        image_height=hw, image_width=hw, tanfovx=(0.5 * fov[0, 0]).tan().item(),
        tanfovy=(0.5 * fov[0, 1]).tan().item(), bg=torch.zeros(3, device=cuda), scale_modifier=1.0,
        viewmatrix=view[0], projmatrix=(view @ proj)[0], projmatrix_raw=proj[0], sh_degree=4,
        campos=d["extrinsics"][0, :3, 3], prefiltered=False, debug=False)
    row, col = torch.triu_indices(3, 3)
    shs = d["harmonics"].transpose(-1, -2).contiguous()                    # (G, 25, 3)
    image, radii, depth, opacity, n_touched = GaussianRasterizer(settings)(
        means3D=d["means"], means2D=torch.zeros_like(d["means"]), shs=shs, colors_precomp=None,
        opacities=d["opacities"][..., None], cov3D_precomp=d["covariances"][:, row, col])
    assert image.shape == (3, hw, hw) and depth.shape == (1, hw, hw) and radii.shape == (d["means"].shape[0],)
    _check(image[None], depth, rc, rd)
    # radii / n_touched agree with the oracle's per-Gaussian outputs
    iu = torch.triu_indices(3, 3)
    _, r_radii, _, r_op, r_nt = rr.rasterize_view(
        sc["means"], sc["covariances"][:, iu[0], iu[1]], sc["harmonics"].transpose(-1, -2), None,
        sc["opacities"][:, None], sc["extrinsics"][0], (settings.tanfovx, settings.tanfovy), 0.01,
        100.0, hw, hw, torch.zeros(3), 4)
    assert (radii.cpu() != r_radii).double().mean() < 1e-3
    assert (opacity.cpu() - r_op).abs().max() < 5e-3
    assert (n_touched.cpu() != r_nt).double().mean() < 2e-2


def test_empty_and_culled(cuda, lib):
    from vicasplat_b200.decoder import render_cuda
    sc = _scene(1, 32, 2, 11)
    d = {k: v.to(cuda) for k, v in sc.items()}
    bg = torch.tensor([[0.1, 0.2, 0.3]], device=cuda).expand(2, 3)
    behind = d["means"].clone()
    behind[:, 2] = -behind[:, 2]                                           # all behind the cameras
    c, dep = render_cuda(d["extrinsics"], d["intrinsics"], d["near"], d["far"], (32, 32), bg,
                         behind, d["covariances"], d["harmonics"], d["opacities"])
    assert torch.allclose(c, bg[:, :, None, None].expand_as(c)) and (dep == 0).all()
    c, dep = render_cuda(d["extrinsics"], d["intrinsics"], d["near"], d["far"], (32, 32), bg,
                         d["means"][:0], d["covariances"][:0], d["harmonics"][:0], d["opacities"][:0])
    assert torch.allclose(c, bg[:, :, None, None].expand_as(c)) and (dep == 0).all()


def test_capacity_overflow_is_detected_and_recovered(cuda, lib):
    """a workspace sized too small is reported through num_pairs and the call re-runs larger."""
    from vicasplat_b200.rasterizer import rasterize_views
    from vicasplat_b200 import decoder as dec
    sc = _scene(1, 32, 1, 5, depth_range=(0.5, 3.0))
    sc["covariances"] = sc["covariances"] * 400.0
    d = {k: v.to(cuda) for k, v in sc.items()}
    tanfov, view_t, full_t, campos = dec._cameras(d["extrinsics"], d["intrinsics"], d["near"], d["far"])
    kw = dict(shs=d["harmonics"], sh_degree=4, sh_layout="chan_major", viewmatrix=view_t,
              projmatrix=full_t, campos=campos, tanfov=tanfov, bg=torch.zeros(3, device=cuda),
              H=32, W=32)
    big = rasterize_views(d["means"], dec._cov6(d["covariances"]), d["opacities"], **kw)
    small = rasterize_views(d["means"], dec._cov6(d["covariances"]), d["opacities"], max_pairs=64, **kw)
    assert torch.equal(big[0], small[0]) and torch.equal(big[2], small[2])


def test_stress_size_properties(cuda, lib):
    """BASELINE configs[4] at full size (16 views of 512x512, G = 2^21): the oracle cannot run this,
    so the render is checked through size-independent properties -- finite output, alpha in [0,1],
    invariance under a permutation of the Gaussians (the depth sort decides the order, not the
    input order), and bit-identity of the per-tile shared-memory sort path with the global 64-bit
    radix sort path."""
    from vicasplat_b200 import decoder as dec, synthetic
    from vicasplat_b200.rasterizer import rasterize_views
    V, S, G = 16, 512, 1 << 21
    sc = {k: v.to(cuda) for k, v in synthetic.gaussian_scene(16, S, S, V, seed=4, n_gauss=G).items()}
    tanfov, view_t, full_t, campos = dec._cameras(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
    cov6 = dec._cov6(sc["covariances"]).contiguous()
    kw = dict(sh_degree=4, sh_layout="chan_major", viewmatrix=view_t, projmatrix=full_t, campos=campos,
              tanfov=tanfov, bg=torch.full((V, 3), 0.25, device=cuda), H=S, W=S)

    def render(idx=None, **extra):
        pick = (lambda t: t) if idx is None else (lambda t: t[idx].contiguous())
        with torch.no_grad():
            return rasterize_views(pick(sc["means"]), pick(cov6), pick(sc["opacities"]),
                                   shs=pick(sc["harmonics"]), **kw, **extra)

    color, radii, depth, alpha, touched = render()
    assert color.shape == (V, 3, S, S) and depth.shape == (V, 1, S, S)
    assert torch.isfinite(color).all() and torch.isfinite(depth).all()
    assert alpha.min() >= 0 and alpha.max() <= 1.0 + 1e-6
    assert (radii > 0).float().mean() > 0.2 and int(touched.sum()) > G      # the scene is really drawn
    # tile-sort path (hinted by the first, checked call) == global radix-sort path
    c_glob = render(max_tile_pairs=0)[0]
    assert torch.equal(color, c_glob)
    # permutation invariance: only exact depth ties (broken by Gaussian id, like the stable upstream
    # sort) can reorder the blend, so a few pixels may move by a rounding-order amount
    perm = torch.randperm(G, device=cuda, generator=torch.Generator(device=cuda).manual_seed(0))
    c_perm, _, d_perm, _, _ = render(perm)
    dc = (c_perm - color).abs()
    assert dc.max() < 2e-3 and dc.mean() < 1e-6 and (dc > 1e-5).float().mean() < 1e-3
    assert (d_perm - depth).abs().max() < 2e-2 and (d_perm - depth).abs().mean() < 1e-5


def test_headline_scene_against_oracle(cuda, lib):
    """The scene of bench.py / BASELINE configs[1] at FULL size -- 524 288 pixel-aligned Gaussians, 256 x 256
    target views -- against the fp32 oracle: two of the twelve views (~1 s of CPU each; first and a middle
    camera).  Measured on B200: 99.71 % of the pixels within 1e-4 absolute (colour), mean error 2.5e-6, worst
    pixel 9.7e-3.  The small scenes above reach >= 99.9 %; here ~100 splats pile up on every pixel, so many
    more alpha evaluations sit within an ulp of the hard thresholds (alpha < 1/255 skipped, stop at
    T < 1e-4), and one flipped decision moves a pixel by up to ~1/255 of a colour.  Asserted: >= 99.5 %
    within 1e-4, mean <= 1e-5, worst <= 2e-2."""
    from vicasplat_b200 import synthetic
    from vicasplat_b200.decoder import render_cuda
    T, V, S = 8, 12, 256
    sc = synthetic.gaussian_scene(T, S, S, V, seed=1)
    pick = [0, 7]
    sub = {k: (v[pick] if k in ("extrinsics", "intrinsics", "near", "far") else v) for k, v in sc.items()}
    rc, rd = rr.render_cuda_ref(sub["extrinsics"], sub["intrinsics"], sub["near"], sub["far"], (S, S),
                                torch.zeros((len(pick), 3)), sub["means"], sub["covariances"], sub["harmonics"],
                                sub["opacities"])
    d = {k: v.to(cuda) for k, v in sc.items()}
    c, dep = render_cuda(d["extrinsics"], d["intrinsics"], d["near"], d["far"], (S, S), torch.zeros((V, 3), device=cuda),
                         d["means"], d["covariances"], d["harmonics"], d["opacities"])
    c, dep = c[pick].cpu().double(), dep[pick].cpu().double()
    ec = (c - rc.double()).abs()
    ed = (dep - rd.double()).abs() / rd.double().abs().clamp_min(1)
    ok_c, ok_d = (ec <= 1e-4).double().mean().item(), (ed <= 1e-4).double().mean().item()
    print(f"[raster headline scene] colour: {ok_c * 100:.4f} % of pixels within 1e-4, max {ec.max():.3e}, mean "
          f"{ec.mean():.3e}; depth (relative): {ok_d * 100:.4f} % within 1e-4, max {ed.max():.3e}")
    assert rc.abs().max() > 0.2 and rd.max() > 1.0
    assert ok_c >= 0.995 and ok_d >= 0.995
    assert ec.mean() <= 1e-5 and ed.mean() <= 1e-5
    assert ec.max() < 2e-2 and ed.max() < 2e-2


@pytest.mark.parametrize("probe", ["plain", "sh_degree4", "near_cull", "depth", "thresholds", "lowpass", "radius",
                                   "tie_order", "n_touched"])
def test_upstream_golden(cuda, lib, probe):
    """Pins the rasterizer (kernels AND oracle) to the real diff_gaussian_rasterization the day its goldens
    exist: oracle/make_raster_golden.py writes tests/golden/raster_upstream.npz on a box with the package.
    Until then the parity of rows R1-R3 is UNPINNED and this test is skipped."""
    from pathlib import Path
    import numpy as np
    from oracle import make_raster_golden as mg
    path = Path(__file__).parent / "golden" / "raster_upstream.npz"
    if not path.exists():
        pytest.skip("no upstream goldens (diff_gaussian_rasterization is not installable here): parity UNPINNED")
    gold = np.load(path)
    sc, h, w = mg.build_scene(probe)
    want_c = torch.from_numpy(gold[f"{probe}/image"]).double()
    want_d = torch.from_numpy(gold[f"{probe}/depth"]).double()[:, 0]
    rc, rd = _oracle(sc, h)
    _check(rc, rd, want_c, want_d)                      # the oracle restates upstream
    c, d = _ours(sc, h, cuda)
    _check(c, d, want_c, want_d)                        # and so do the kernels


def test_stress_size_view_against_oracle(cuda, lib):
    """BASELINE configs[4] at full size -- 16 views of 512 x 512, G = 2^21 Gaussians -- with ONE view held to
    the fp32 oracle (the CPU needs ~10 s per view at this size): same bar as the headline scene."""
    from vicasplat_b200 import synthetic
    from vicasplat_b200.decoder import render_cuda
    V, S, G = 16, 512, 1 << 21
    sc = synthetic.gaussian_scene(16, S, S, V, seed=4, n_gauss=G)
    pick = [5]
    rc, rd = rr.render_cuda_ref(sc["extrinsics"][pick], sc["intrinsics"][pick], sc["near"][pick], sc["far"][pick],
                                (S, S), torch.zeros((1, 3)), sc["means"], sc["covariances"], sc["harmonics"],
                                sc["opacities"])
    d = {k: v.to(cuda) for k, v in sc.items()}
    c, dep = render_cuda(d["extrinsics"], d["intrinsics"], d["near"], d["far"], (S, S), torch.zeros((V, 3), device=cuda),
                         d["means"], d["covariances"], d["harmonics"], d["opacities"])
    ec = (c[pick].cpu().double() - rc.double()).abs()
    ed = (dep[pick].cpu().double() - rd.double()).abs() / rd.double().abs().clamp_min(1)
    ok_c, ok_d = (ec <= 1e-4).double().mean().item(), (ed <= 1e-4).double().mean().item()
    print(f"[raster stress scene] colour: {ok_c * 100:.4f} % of pixels within 1e-4, max {ec.max():.3e}, mean "
          f"{ec.mean():.3e}; depth (relative): {ok_d * 100:.4f} % within 1e-4, max {ed.max():.3e}")
    assert ok_c >= 0.995 and ok_d >= 0.995 and ec.mean() <= 1e-5 and ec.max() < 2e-2
