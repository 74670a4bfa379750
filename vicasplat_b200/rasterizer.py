"""Drop-in for the ``diff_gaussian_rasterization`` extension the reference imports at
src/model/decoder/cuda_splatting.py:5-8 and calls at :207-235 / :304-330, plus the batched
multi-view entry (`rasterize_views`) that `DecoderSplattingCUDA` uses to render all V views of a
scene from ONE copy of the Gaussians.

Everything runs in the hand-written sm_100a kernels behind ``vs_raster_forward`` /
``vs_raster_backward``; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional

import torch
import torch.nn as nn

from . import _lib
from ._lib import RasterBwdParams, RasterParams, check, ptr, stream_ptr


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False
    debug: bool = False


class RasterOverflow(RuntimeError):
    pass


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


# (V, G, H, W) -> (pair capacity, per-tile bound) that was enough last time (+25 %): later calls of
# the same shape size their binning workspace from it instead of the worst-case default, and sort
# per tile in shared memory (bound <= 16384) instead of one global 64-bit radix sort (bound 0)
_capacity_hint: dict = {}
MAX_TILE_SORT = 16384
SMALL_TILE_SORT = 4096     # tiles up to here are sorted by tile_sort_small_kernel alone


# check_overflow="deferred": the render runs unchecked with the hinted capacities and its device-side
# counters (pairs, largest tile) are queued here; the caller copies them to the host with the rest
# of its results and calls `verify_deferred` on them -- no host synchronisation in the render call
_deferred: list = []


def take_deferred() -> list:
    """[(counters int64[2] device tensor, max_pairs, max_tile, shape key)] since the last call."""
    out = list(_deferred)
    _deferred.clear()
    return out


def verify_deferred(counts, records) -> None:
    """counts: host int64 (n, 2) copies of the queued counters.  Raises RasterOverflow (after
    raising the per-shape capacity hint, so that a re-submission succeeds) if any render of the
    batch had more (tile, splat) pairs than its workspace held."""
    bad = None
    for (n, tmax), (_t, max_pairs, max_tile, key) in zip(counts.tolist(), records):
        if n > max_pairs or (max_tile > 0 and tmax > max_tile):
            next_tile = 0 if tmax > MAX_TILE_SORT else min(MAX_TILE_SORT, int(tmax * 1.25) + 64)
            _raise_hint(key, max(int(n * 1.25) + 4096, max_pairs), next_tile)
            bad = (n, tmax, max_pairs, max_tile)
    if bad is not None:
        raise RasterOverflow("render exceeded its binning capacity (pairs=%d, largest tile=%d, capacity=%d/%d); "
                             "the capacity hint has been raised: re-submit the batch" % bad)


def _raise_hint(key, pairs: int, tile: int) -> None:
    """The per-shape hint only ever GROWS (scenes of one shape differ: sizing for the last one alone makes
    the next, busier one overflow): pairs = the most any scene needed; tile bound = the largest seen, or
    0 (global sort) as soon as one scene needed it."""
    old = _capacity_hint.get(key)
    if old is not None:
        pairs = max(pairs, old[0])
        tile = 0 if (tile == 0 or old[1] == 0) else max(tile, old[1])
    _capacity_hint[key] = (pairs, tile)


class _Ctx:
    """Forward state kept for the backward pass (workspace holds the sorted splat lists)."""
    __slots__ = ("params", "keep", "num_pairs", "max_pairs", "max_tile")


def _run_forward(V, G, H, W, shared, means, cov6, opac, shs, sh_M, sh_degree, sh_strides, colors,
                 viewm, projm, campos, tanfov, bg, max_pairs, max_tile_pairs=0, want_aux=True):
    lib = _lib.load()
    dev = means.device
    f32, i32 = torch.float32, torch.int32
    color = torch.empty((V, 3, H, W), dtype=f32, device=dev)
    depth = torch.empty((V, 1, H, W), dtype=f32, device=dev)
    alpha = torch.empty((V, 1, H, W), dtype=f32, device=dev)
    radii = torch.empty((V, max(G, 1)), dtype=i32, device=dev)
    n_touched = torch.empty((V, max(G, 1)), dtype=i32, device=dev) if want_aux else None
    final_T = torch.empty((V, H, W), dtype=f32, device=dev)
    n_contrib = torch.empty((V, H, W), dtype=i32, device=dev)
    num_pairs = torch.zeros((2,), dtype=torch.int64, device=dev)
    ws_bytes = lib.vs_raster_workspace_bytes(V, max(G, 1), H, W, max_pairs)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    p = RasterParams()
    p.V, p.G, p.H, p.W, p.gaussians_shared = V, G, H, W, int(shared)
    p.means3D, p.cov3D, p.opacities = ptr(means), ptr(cov6), ptr(opac)
    p.shs, p.sh_M, p.sh_degree = ptr(shs), sh_M, sh_degree
    p.sh_stride_coef, p.sh_stride_chan = sh_strides
    p.colors_precomp = ptr(colors)
    p.viewmatrix, p.projmatrix, p.campos = ptr(viewm), ptr(projm), ptr(campos)
    p.tanfov, p.bg, p.scale_modifier = ptr(tanfov), ptr(bg), 1.0
    p.out_color, p.out_depth, p.out_alpha = ptr(color), ptr(depth), ptr(alpha)
    p.radii, p.n_touched, p.final_T, p.n_contrib = ptr(radii), ptr(n_touched), ptr(final_T), ptr(n_contrib)
    p.workspace, p.workspace_bytes, p.max_pairs = ptr(ws), ws_bytes, max_pairs
    p.num_pairs_out = ptr(num_pairs)
    p.max_tile_pairs = int(max_tile_pairs)
    check(lib.vs_raster_forward(C.byref(p), C.c_void_p(stream_ptr())), "vs_raster_forward")
    ctx = _Ctx()
    ctx.params = p
    ctx.keep = (means, cov6, opac, shs, colors, viewm, projm, campos, tanfov, bg, color, depth,
                alpha, radii, n_touched, final_T, n_contrib, ws, num_pairs)
    ctx.num_pairs, ctx.max_pairs = num_pairs, max_pairs
    return color, depth, alpha, radii, n_touched, ctx


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, cov6, opac, shs, colors, theta, rho, cfg):
        (V, G, H, W, shared, sh_M, sh_degree, sh_strides, viewm, projm, campos, tanfov, bg,
         max_pairs, max_tile, check_overflow, want_aux) = cfg
        m, c6, o = _f32c(means), _f32c(cov6), _f32c(opac).reshape(-1)
        s = _f32c(shs) if shs is not None else None
        cp = _f32c(colors) if colors is not None else None
        while True:
            color, depth, alpha, radii, n_touched, st = _run_forward(
                V, G, H, W, shared, m, c6, o, s, sh_M, sh_degree, sh_strides, cp, viewm, projm,
                campos, tanfov, bg, max_pairs, max_tile, want_aux)
            if check_overflow == "deferred":
                _deferred.append((st.num_pairs, max_pairs, max_tile, (V, G, H, W)))
                break
            if not check_overflow:
                break
            n, tmax = st.num_pairs.tolist()   # the upstream extension syncs here too (num_rendered)
            ok = n <= max_pairs and (max_tile == 0 or tmax <= max_tile)
            if max_tile == 0:
                next_tile = 0 if _capacity_hint.get((V, G, H, W), (0, 1))[1] == 0 else MAX_TILE_SORT
            else:
                next_tile = min(MAX_TILE_SORT, int(tmax * 1.25) + 64) if tmax <= MAX_TILE_SORT else 0
                if tmax <= SMALL_TILE_SORT < next_tile:
                    next_tile = SMALL_TILE_SORT   # stay within the one-kernel size classes
            if (V, G, H, W) in _capacity_hint and max_tile > 0:
                _raise_hint((V, G, H, W), int(n * 1.25) + 4096, next_tile)
            else:   # first sight of this shape (or leaving the global-sort default): take what this scene needs
                _capacity_hint[(V, G, H, W)] = (int(n * 1.25) + 4096, next_tile)
            if ok:
                break
            # capacity or per-tile bound was too small: re-run with what this scene needs
            max_pairs = max(max_pairs, int(n * 1.05) + 1024)
            max_tile = 0 if tmax > MAX_TILE_SORT else min(MAX_TILE_SORT, int(tmax * 1.05) + 64)
        ctx.state = st
        ctx.shapes = (means.shape, cov6.shape, opac.shape, None if shs is None else shs.shape,
                      None if colors is None else colors.shape)
        ctx.has_pose = (theta is not None, rho is not None)
        ctx.pose_shapes = (None if theta is None else theta.shape, None if rho is None else rho.shape)
        if n_touched is None:
            n_touched = radii.new_zeros(())
        ctx.mark_non_differentiable(radii, n_touched)
        return color, radii, depth, alpha, n_touched

    @staticmethod
    def backward(ctx, g_color, g_radii, g_depth, g_alpha, g_touched):
        st = ctx.state
        g = render_backward(st, g_color, g_depth, g_alpha, want_tau=any(ctx.has_pose))
        sm, sc, so, ss, scol = ctx.shapes
        has_sh = st.params.shs is not None
        g_shs = g["d_sh"].reshape(ss) if has_sh else None
        g_cols = g["d_colors"].reshape(scol) if not has_sh else None
        d_tau = g.get("d_tau")
        g_theta = d_tau[:, 3:].reshape(ctx.pose_shapes[0]) if ctx.has_pose[0] else None
        g_rho = d_tau[:, :3].reshape(ctx.pose_shapes[1]) if ctx.has_pose[1] else None
        return (g["d_means"].reshape(sm), g["d_cov6"].reshape(sc), g["d_opac"].reshape(so), g_shs, g_cols,
                g_theta, g_rho, None)


def render_backward(state: _Ctx, g_color, g_depth=None, g_alpha=None, out: Optional[dict] = None,
                    want_tau: bool = True) -> dict:
    """vs_raster_backward for a kept forward state (``_run_forward``): gradients w.r.t. means (G,3) /
    (V,G,3), cov6, opacity, SH (in the layout the forward call was given) or colours, and the camera
    twist (V,6: rho, theta).  ``out`` may hold preallocated, ZEROED fp32 buffers under the same keys
    (d_means, d_cov6, d_opac, d_sh) -- e.g. one scene's slice of a batch-wide gradient tensor, so the
    training step needs no per-scene copies; the kernels accumulate into them."""
    lib = _lib.load()
    fp = state.params
    V, G = fp.V, fp.G
    f32 = torch.float32
    dev = (g_color if g_color is not None else g_depth).device
    shared = bool(fp.gaussians_shared)
    lead = (G,) if shared else (V, G)
    out = dict(out or {})
    has_sh = fp.shs is not None

    def buf(key, shape):
        t = out.get(key)
        if t is None:
            t = out[key] = torch.zeros(shape, dtype=f32, device=dev)
        assert t.dtype == f32 and t.is_contiguous() and t.numel() == int(torch.Size(shape).numel()), key
        return t

    d_means, d_cov, d_opac = buf("d_means", lead + (3,)), buf("d_cov6", lead + (6,)), buf("d_opac", lead)
    d_shs = buf("d_sh", lead + (3 * fp.sh_M,)) if has_sh else None
    d_col = buf("d_colors", lead + (3,)) if not has_sh else None
    # the camera-twist gradient costs a block reduction per view and component: only when asked for
    d_tau = buf("d_tau", (V, 6)) if want_tau else None
    gc = _f32c(g_color) if g_color is not None else torch.zeros((V, 3, fp.H, fp.W), dtype=f32, device=dev)
    gd = _f32c(g_depth) if g_depth is not None else None
    ga = _f32c(g_alpha) if g_alpha is not None else None
    bws = torch.empty((V * G * 10,), dtype=f32, device=dev)
    bp = RasterBwdParams()
    bp.fwd = fp
    bp.bwd_workspace, bp.bwd_workspace_bytes = ptr(bws), bws.numel() * 4
    bp.dL_dcolor, bp.dL_ddepth, bp.dL_dalpha = ptr(gc), ptr(gd), ptr(ga)
    bp.dL_dmeans3D, bp.dL_dcov3D, bp.dL_dopacity = ptr(d_means), ptr(d_cov), ptr(d_opac)
    bp.dL_dshs, bp.dL_dcolors, bp.dL_dtau = ptr(d_shs), ptr(d_col), ptr(d_tau)
    check(lib.vs_raster_backward(C.byref(bp), C.c_void_p(stream_ptr())), "vs_raster_backward")
    return out


def render_forward(means, cov6, opac, shs, *, sh_degree, sh_layout, viewmatrix, projmatrix, campos, tanfov, bg,
                   H, W, max_pairs=None, max_tile_pairs=None):
    """Plain (non-autograd) forward of V views of ONE shared Gaussian set with the per-shape capacity
    hint: -> (color (V,3,H,W), depth (V,1,H,W), alpha, state).  ``state.num_pairs`` (device int64[2]:
    pairs, largest tile) lets the caller verify the capacity when it next synchronises
    (``verify_deferred``); ``render_backward(state, ...)`` differentiates it."""
    V, G = viewmatrix.shape[0], means.shape[0]
    sh_M, strides = (shs.shape[-2], (3, 1)) if sh_layout == "coef_major" else (shs.shape[-1], (1, shs.shape[-1]))
    hint = _capacity_hint.get((V, G, H, W), (max(4 * V * G, 1 << 16), MAX_TILE_SORT))
    max_pairs = hint[0] if max_pairs is None else max_pairs
    max_tile = hint[1] if max_tile_pairs is None else max_tile_pairs
    color, depth, alpha, _radii, _nt, st = _run_forward(
        V, G, H, W, True, means, cov6, opac.reshape(-1), shs, sh_M, int(sh_degree), strides, None,
        _f32c(viewmatrix).reshape(V, 16), _f32c(projmatrix).reshape(V, 16), _f32c(campos).reshape(V, 3),
        _f32c(tanfov).reshape(V, 2), _f32c(bg).reshape(-1, 3).expand(V, 3).contiguous(), int(max_pairs),
        int(max_tile), want_aux=False)
    st.max_tile = int(max_tile)
    return color, depth, alpha, st


class SceneStreams:
    """The scenes of a batch are independent launch chains (preprocess -> bin -> sort -> blend): alternating
    them between the current stream and side streams lets the block scheduler co-run one scene's issue-bound
    blend with the next scene's memory- / latency-bound binning and sort kernels.

        with SceneStreams(dev) as ss:
            for i, scene in enumerate(batch):
                with ss.scene(i):
                    out.append(rasterize_views(...))

    On exit the current stream waits for the side streams; tensors allocated inside are marked as used by the
    current stream (``keep``).  VS_RASTER_STREAMS=1 turns the interleaving off (A/B measurements)."""
    _pool: dict = {}

    def __init__(self, dev, n: Optional[int] = None):
        import os
        self.n = n if n is not None else int(os.environ.get("VS_RASTER_STREAMS", "4"))
        self.dev = dev
        self.cur = torch.cuda.current_stream(dev)
        key = (dev.index if dev.index is not None else torch.cuda.current_device(), self.n)
        if key not in SceneStreams._pool:
            SceneStreams._pool[key] = [torch.cuda.Stream(dev) for _ in range(max(self.n - 1, 0))]
        self.side = SceneStreams._pool[key]
        self.used = set()

    def __enter__(self):
        return self

    def scene(self, i: int):
        k = i % max(self.n, 1)
        if k == 0:
            return torch.cuda.stream(self.cur)
        st = self.side[k - 1]
        if k not in self.used:
            st.wait_stream(self.cur)
            self.used.add(k)
        return torch.cuda.stream(st)

    def keep(self, *tensors):
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(self.cur)

    def __exit__(self, *exc):
        for k in self.used:
            self.cur.wait_stream(self.side[k - 1])
        return False


def rasterize_views(means3D, cov6, opacities, *, shs=None, colors_precomp=None, sh_degree=0,
                    sh_layout="coef_major", viewmatrix, projmatrix, campos, tanfov, bg, H, W,
                    theta=None, rho=None, max_pairs: Optional[int] = None,
                    max_tile_pairs: Optional[int] = None, check_overflow=True,
                    want_n_touched: bool = True):
    """Render V views.  Gaussians are shared by all views when means3D is (G,3), per-view when
    (V,G,3).  viewmatrix/projmatrix (V,4,4) are the *transposed* matrices the reference passes
    (cuda_splatting.py:192-194); tanfov (V,2); bg (V,3) or (3,).

    sh_layout: "coef_major" = (G, M, 3), the layout the reference hands the extension
    (cuda_splatting.py:182); "chan_major" = (G, 3, M), the encoder's own layout
    (gaussian_adapter.py:180), consumed without the transpose copy.

    check_overflow: True = read the pair counters back and re-run larger if the workspace was too
    small (one host sync, like the upstream extension's num_rendered); False = trust the
    per-shape hint; "deferred" = trust it now, queue the counters for `verify_deferred`.

    Returns color (V,3,H,W), radii (V,G), depth (V,1,H,W), alpha (V,1,H,W), n_touched (V,G)
    (a 0-d placeholder when want_n_touched=False: the reference's render_cuda discards it,
    cuda_splatting.py:226-239, and counting costs one atomic per (warp, splat) hit).
    """
    if not means3D.is_cuda:
        raise RuntimeError("rasterize_views needs CUDA tensors (there is no CPU fallback)")
    V = viewmatrix.shape[0]
    shared = means3D.dim() == 2
    G = means3D.shape[-2]
    dev = means3D.device
    vm = _f32c(viewmatrix).reshape(V, 16)
    pm = _f32c(projmatrix).reshape(V, 16)
    cp = _f32c(campos).reshape(V, 3)
    tf = _f32c(tanfov.to(dev) if isinstance(tanfov, torch.Tensor) else
               torch.tensor(tanfov, dtype=torch.float32, device=dev)).reshape(V, 2)
    bgc = _f32c(bg.to(dev)).reshape(-1, 3).expand(V, 3).contiguous()
    sh_M = 0
    strides = (0, 0)
    if shs is not None:
        if sh_layout == "coef_major":
            sh_M, strides = shs.shape[-2], (3, 1)
        elif sh_layout == "chan_major":
            sh_M, strides = shs.shape[-1], (1, shs.shape[-1])
        else:
            raise ValueError(f"bad sh_layout {sh_layout!r}")
    hint = _capacity_hint.get((V, G, H, W), (max(4 * V * G, 1 << 16), MAX_TILE_SORT))
    if max_pairs is None:
        max_pairs = hint[0]
    if max_tile_pairs is None:
        max_tile_pairs = hint[1]
    cfg = (V, G, H, W, shared, sh_M, int(sh_degree), strides, vm, pm, cp, tf, bgc, int(max_pairs),
           int(max_tile_pairs), check_overflow, bool(want_n_touched))
    return _Rasterize.apply(means3D, cov6, opacities, shs, colors_precomp, theta, rho, cfg)


class GaussianRasterizer(nn.Module):
    """Same call signature and 5-tuple return as the reference's extension
    (cuda_splatting.py:226-235, :323-330)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D=None, shs=None, colors_precomp=None, opacities=None,
                scales=None, rotations=None, cov3D_precomp=None, theta=None, rho=None):
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if cov3D_precomp is None:
            raise Exception("vicasplat_b200 rasterizer: only cov3D_precomp is supported "
                            "(the reference never passes scales/rotations, cuda_splatting.py:231-232)")
        tanfov = torch.tensor([[float(rs.tanfovx), float(rs.tanfovy)]], dtype=torch.float32)
        color, radii, depth, alpha, n_touched = rasterize_views(
            means3D, cov3D_precomp, opacities, shs=shs, colors_precomp=colors_precomp,
            sh_degree=rs.sh_degree, viewmatrix=rs.viewmatrix[None], projmatrix=rs.projmatrix[None],
            campos=rs.campos[None], tanfov=tanfov, bg=rs.bg, H=int(rs.image_height),
            W=int(rs.image_width), theta=theta, rho=rho)
        return color[0], radii[0], depth[0], alpha[0], n_touched[0]
