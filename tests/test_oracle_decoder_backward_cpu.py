"""The hand-derived backward pass of MixDecoderBlock (oracle/decoder_backward_ref.py: the formulas the
decoder's CUDA training kernels will implement) against torch.autograd over encoder_ref.dec_block,
which tests/golden/model_grad_small.npz pins to the unmodified reference's own gradients."""
import pytest
import torch

from oracle import decoder_backward_ref as db
from oracle import encoder_ref as er


def _setup(T, dtype):
    cfg = er.EncoderConfig(img_size=64, enc_depth=1, dec_depth=2)
    key = "backbone.dec_blocks.1"
    sd = {k: v.to(dtype) for k, v in er.synth_state_dict(cfg, seed=3).items() if k.startswith(key + ".")}
    g = torch.Generator().manual_seed(T)
    for k in sd:                                   # non-trivial biases / norm weights
        if k.endswith(".bias"):
            sd[k] = (0.1 * torch.randn(sd[k].shape, generator=g)).to(dtype)
        elif "norm" in k:
            sd[k] = (1 + 0.1 * torch.randn(sd[k].shape, generator=g)).to(dtype)
    B, N, C = 2, 17, cfg.dec_embed_dim
    img = torch.randn((B, T, N, C), generator=g).to(dtype)
    cam = torch.randn((B, T, C), generator=g).to(dtype)
    pos = er.positions(B * T, 4, 4, True).reshape(B, T, N, 2)
    d_img = torch.randn((B, T, N, C), generator=g).to(dtype)
    d_cam = torch.randn((B, T, C), generator=g).to(dtype)
    return cfg, key, sd, img, cam, pos, d_img, d_cam


@pytest.mark.parametrize("T", [2, 3, 4])            # T = 2: single neighbour; T >= 3: end frames see one frame twice
def test_manual_decoder_block_backward_matches_autograd(T):
    cfg, key, sd, img, cam, pos, d_img, d_cam = _setup(T, torch.float64)
    out_img, out_cam, cache = db.dec_block_fwd(sd, key, img, cam, pos, cfg)
    gi, gc, grads = db.dec_block_bwd(sd, key, d_img, d_cam, cache, cfg)
    for v in sd.values():
        v.requires_grad_(True)
    ri, rc = img.clone().requires_grad_(True), cam.clone().requires_grad_(True)
    ref_img, ref_cam = er.dec_block(sd, key, ri, rc, pos, cfg)
    assert torch.allclose(out_img, ref_img.detach(), rtol=1e-9, atol=1e-9)
    assert torch.allclose(out_cam, ref_cam.detach(), rtol=1e-9, atol=1e-9)
    ((ref_img * d_img).sum() + (ref_cam * d_cam).sum()).backward()
    rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert rel(gi, ri.grad) < 1e-9 and rel(gc, rc.grad) < 1e-9
    assert set(grads) == set(sd), sorted(set(sd) ^ set(grads))
    for k in sd:
        assert sd[k].grad is not None, k
        assert rel(grads[k], sd[k].grad) < 1e-8, (k, rel(grads[k], sd[k].grad))


def test_stage_formulas():
    g = torch.Generator().manual_seed(0)
    x = torch.randn((5, 7, 32), generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn((32,), generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn((32,), generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn((5, 7, 32), generator=g, dtype=torch.float64)
    y, cache = db.ln_fwd(x.detach(), w.detach(), b.detach(), 1e-6)
    torch.nn.functional.layer_norm(x, (32,), w, b, 1e-6).backward(dy)
    dx, dw, dbias = db.ln_bwd(dy, cache, w.detach())
    assert torch.allclose(dx, x.grad) and torch.allclose(dw, w.grad) and torch.allclose(dbias, b.grad)
    z = torch.randn((64,), generator=g, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.gelu(z).sum().backward()
    assert torch.allclose(db.gelu_grad(z.detach()), z.grad)
    z.grad = None
    torch.nn.functional.silu(z).sum().backward()
    assert torch.allclose(db.silu_grad(z.detach()), z.grad)
    # rope transposes: <R a, b> == <a, R^T b>
    a = torch.randn((2, 3, 9, 64), generator=g, dtype=torch.float64)
    bb = torch.randn((2, 3, 9, 64), generator=g, dtype=torch.float64)
    pos = torch.randint(0, 5, (2, 9, 2), generator=g)
    assert torch.allclose((er.rope2d(a, pos, 100.0) * bb).sum(), (a * db.rope2d_bwd(bb, pos, 100.0)).sum())
    fr = torch.arange(9)
    assert torch.allclose((er.rope1d_interleaved(a, fr, 30.0) * bb).sum(),
                          (a * db.rope1d_bwd(bb, fr, 30.0)).sum())


def test_adapter_postprocess_and_pose_tail_backward():
    from oracle import adapter_backward_ref as ab
    cfg = er.EncoderConfig()
    g = torch.Generator().manual_seed(9)
    raw = torch.randn((2, 3, 5, 86), generator=g, dtype=torch.float64)
    raw[..., 4:7] += torch.tensor([0.0, 6.5, -3.0], dtype=torch.float64)   # one scale channel beyond the 0.3 clamp
    raw[0, 0, 0, 4:7] = 400.0                                               # clamped: zero gradient
    rr = raw.clone().requires_grad_(True)
    out = er.gaussian_adapter(rr, cfg)
    D = {k: torch.randn(out[k].shape, generator=g, dtype=torch.float64)
         for k in ("means", "covariances", "harmonics", "opacities", "scales", "rotations")}
    sum((out[k] * D[k]).sum() for k in D).backward()
    got = ab.adapter_backward(raw, cfg, D["means"], D["covariances"], D["harmonics"], D["opacities"],
                              d_scales=D["scales"], d_rot=D["rotations"])
    assert torch.allclose(got, rr.grad, rtol=1e-9, atol=1e-12)
    assert (got[0, 0, 0, 4:7] == 0).all()
    # raster-only case: no direct gradient on scales / rotations
    rr.grad = None
    out = er.gaussian_adapter(rr, cfg)
    sum((out[k] * D[k]).sum() for k in ("means", "covariances", "harmonics", "opacities")).backward()
    got = ab.adapter_backward(raw, cfg, D["means"], D["covariances"], D["harmonics"], D["opacities"])
    assert torch.allclose(got, rr.grad, rtol=1e-9, atol=1e-12)

    x = torch.randn((4, 6, 3), generator=g, dtype=torch.float64, requires_grad=True)
    dxyz = torch.randn((4, 6, 3), generator=g, dtype=torch.float64)
    d = x.norm(dim=-1, keepdim=True)
    (x / d.clip(min=1e-8) * torch.expm1(d) * dxyz).sum().backward()          # encoder_ref.pts_head's tail
    assert torch.allclose(ab.exp_postprocess_backward(x.detach(), dxyz), x.grad, rtol=1e-10, atol=1e-12)

    v = torch.randn((2, 7, 8), generator=g, dtype=torch.float64)
    v[..., 3] += 1.0
    vr = v.clone().requires_grad_(True)
    dp = torch.randn((2, 7, 8), generator=g, dtype=torch.float64)
    ((vr / vr[..., :4].norm(dim=-1, keepdim=True)) * dp).sum().backward()    # encoder_ref.camera_head's tail
    assert torch.allclose(ab.dq_normalise_backward(v, dp), vr.grad, rtol=1e-10, atol=1e-12)


def test_dpt_operator_backward_formulas():
    import torch.nn.functional as F
    from oracle import dpt_backward_ref as dp
    g = torch.Generator().manual_seed(4)
    rnd = lambda *s: torch.randn(s, generator=g, dtype=torch.float64)
    # stride-1 convolutions: 3x3 pad 1, 1x1 pad 0, 7x7 pad 3
    for k, pad in ((3, 1), (1, 0), (7, 3)):
        x, W, b = rnd(2, 5, 9, 8).requires_grad_(True), rnd(6, 5, k, k).requires_grad_(True), rnd(6).requires_grad_(True)
        dy = rnd(2, 6, 9, 8)
        F.conv2d(x, W, b, padding=pad).backward(dy)
        assert torch.allclose(dp.conv_dgrad_s1(dy, W.detach(), pad), x.grad, atol=1e-11)
        assert torch.allclose(dp.conv_wgrad(dy, x.detach(), k, pad), W.grad, atol=1e-11)
        assert torch.allclose(dy.sum((0, 2, 3)), b.grad, atol=1e-11)
    # stride-2 3x3 pad 1 (act_postprocess.3.1), even and odd input sizes
    for hw in ((8, 8), (7, 9)):
        x, W = rnd(2, 4, *hw).requires_grad_(True), rnd(3, 4, 3, 3).requires_grad_(True)
        y = F.conv2d(x, W, stride=2, padding=1)
        dy = rnd(*y.shape)
        y.backward(dy)
        assert torch.allclose(dp.conv_dgrad_strided(dy, W.detach(), 1, 2, hw), x.grad, atol=1e-11)
        assert torch.allclose(dp.conv_wgrad(dy, x.detach(), 3, 1, stride=2), W.grad, atol=1e-11)
    # ConvTranspose2d kernel = stride
    for s in (2, 4):
        x, Wt, b = rnd(2, 5, 3, 4).requires_grad_(True), rnd(5, 6, s, s).requires_grad_(True), rnd(6).requires_grad_(True)
        y = F.conv_transpose2d(x, Wt, b, stride=s)
        dy = rnd(*y.shape)
        y.backward(dy)
        dx, dW, db_ = dp.conv_transpose_ks_backward(dy, x.detach(), Wt.detach())
        assert torch.allclose(dx, x.grad, atol=1e-11) and torch.allclose(dW, Wt.grad, atol=1e-11)
        assert torch.allclose(db_, b.grad, atol=1e-11)
    # bilinear x2, align_corners=True
    x = rnd(2, 3, 5, 4).requires_grad_(True)
    y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    assert torch.allclose(dp.up2_forward(x.detach()), y.detach(), atol=1e-12)
    dy = rnd(*y.shape)
    y.backward(dy)
    assert torch.allclose(dp.up2_backward(dy), x.grad, atol=1e-12)
    # residual conv unit (pre-activation), against the oracle's rcu
    sd = {"u.conv1.weight": rnd(4, 4, 3, 3), "u.conv1.bias": rnd(4), "u.conv2.weight": rnd(4, 4, 3, 3),
          "u.conv2.bias": rnd(4)}
    for v in sd.values():
        v.requires_grad_(True)
    x = rnd(2, 4, 6, 5).requires_grad_(True)
    dy = rnd(2, 4, 6, 5)
    er.rcu(sd, "u", x).backward(dy)
    y1 = er.conv(sd, "u.conv1", F.relu(x), padding=1).detach()
    dx, (dW1, db1), (dW2, db2) = dp.rcu_backward(dy, x.detach(), sd["u.conv1.weight"].detach(),
                                                 sd["u.conv2.weight"].detach(), y1)
    assert torch.allclose(dx, x.grad, atol=1e-11)
    assert torch.allclose(dW1, sd["u.conv1.weight"].grad, atol=1e-11) and torch.allclose(db1, sd["u.conv1.bias"].grad, atol=1e-11)
    assert torch.allclose(dW2, sd["u.conv2.weight"].grad, atol=1e-11) and torch.allclose(db2, sd["u.conv2.bias"].grad, atol=1e-11)


def test_shared_host_device_tail_math_against_the_fp64_oracle():
    """vicasplat_b200/csrc/tail_math.h (the per-element chain rules the backward kernels of the
    encoder tails run; host build oracle/_build/libtail_math_host.so) against
    oracle/adapter_backward_ref.py."""
    import ctypes as C
    import subprocess
    from pathlib import Path
    from oracle import adapter_backward_ref as ab
    root = Path(__file__).resolve().parent.parent
    so = root / "oracle" / "_build" / "libtail_math_host.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(root / "oracle"), "_build/libtail_math_host.so"], check=True)
    lib = C.CDLL(str(so))
    fp = lambda t: C.c_void_p(t.data_ptr())
    cfg = er.EncoderConfig()
    g = torch.Generator().manual_seed(21)
    n = 4096
    raw = torch.randn((n, 86), generator=g)
    raw[:, 4:7] += torch.tensor([0.0, 6.5, -3.0])
    raw[:7, 4:7] = 400.0                                         # clamped scales
    d_means, d_opac = torch.randn((n, 3), generator=g), torch.randn((n, 1), generator=g)
    d_cov6 = torch.randn((n, 6), generator=g)
    iu = torch.triu_indices(3, 3)
    d_cov = torch.zeros((n, 3, 3), dtype=torch.float64)
    d_cov[:, iu[0], iu[1]] = d_cov6.double()                     # gradient of the packed upper triangle
    want = ab.adapter_backward(raw.double(), cfg, d_means.double(), d_cov,
                               torch.zeros((n, 3, 25), dtype=torch.float64), d_opac.double())[:, :11]
    got = torch.empty((n, 11))
    lib.tm_adapter_backward(fp(raw), C.c_longlong(86), fp(d_means), fp(d_cov6), fp(d_opac), fp(got),
                            C.c_longlong(n))
    err = (got.double() - want).norm(dim=0) / want.norm(dim=0).clamp_min(1e-30)
    assert (err < 2e-5).all(), err
    assert (got[:7, 4:7] == 0).all()

    x, gx = torch.randn((n, 3), generator=g), torch.randn((n, 3), generator=g)
    x[:5] *= 1e-3                                                # small norms: f(d) -> 1, f'(d) -> 1/2
    got = torch.empty((n, 3))
    lib.tm_exp_postprocess_backward(fp(x), fp(gx), fp(got), C.c_longlong(n))
    want = ab.exp_postprocess_backward(x.double(), gx.double())
    assert ((got.double() - want).norm() / want.norm()).item() < 1e-5
    assert torch.allclose(got[:5].double(), want[:5], rtol=2e-3, atol=1e-6)   # cancellation in f32 at tiny |x|

    v, dp = torch.randn((n, 8), generator=g), torch.randn((n, 8), generator=g)
    v[:, 3] += 1.0
    got = torch.empty((n, 8))
    lib.tm_dq_normalise_backward(fp(v), fp(dp), fp(got), C.c_longlong(n))
    want = ab.dq_normalise_backward(v.double(), dp.double())
    assert ((got.double() - want).norm() / want.norm()).item() < 1e-5


@pytest.mark.parametrize("T", [2, 3, 5])
def test_neighbour_backward_tables_write_every_key_row_once(T):
    """Key-frame-centric item tables of the neighbour attention's dK / dV pass
    (vicasplat_b200.encoder_grad.neighbour_backward_tables): emulating the pass item by item -- key frame
    j with its (<= 2) reading query frames, probabilities from each query row's OWN softmax -- reproduces
    the hand-derived oracle's dK / dV, which accumulates over the reference's duplicated key lists."""
    from vicasplat_b200.encoder_grad import neighbour_backward_tables
    cfg = er.EncoderConfig(img_size=64, enc_depth=1, dec_depth=2)
    B, N, H, hd = 2, 9, cfg.dec_num_heads, 64
    rpf = N + 1
    g = torch.Generator().manual_seed(40 + T)
    rnd = lambda: torch.randn((B * T * rpf, H, hd), generator=g, dtype=torch.float64)
    q, k, v, do = rnd(), rnd(), rnd(), rnd()
    img = lambda t: t.view(B, T, rpf, H, hd)[:, :, 1:].permute(0, 1, 3, 2, 4)      # (B,T,H,N,hd) image rows
    # the oracle's way: per query frame over its (possibly duplicated) neighbour list, scatter-add
    want_dk, want_dv = torch.zeros_like(img(k)), torch.zeros_like(img(v))
    P_rows = {}
    for t in range(T):
        nb = db._neighbours(t, T)
        kk = torch.cat([img(k)[:, j] for j in nb], 2)
        vv = torch.cat([img(v)[:, j] for j in nb], 2)
        o, P = db.sdpa_fwd(img(q)[:, t], kk, vv)
        _, dkk, dvv = db.sdpa_bwd(img(do)[:, t], img(q)[:, t], kk, vv, P)
        for i, j in enumerate(nb):
            want_dk[:, j] += dkk[:, :, i * N:(i + 1) * N]
            want_dv[:, j] += dvv[:, :, i * N:(i + 1) * N]
        # statistics the kernels keep per query row: lse over the DEDUPLICATED key set, delta = rowsum(dO O)
        uniq = sorted(set(nb))
        s = img(q)[:, t] @ torch.cat([img(k)[:, j] for j in uniq], 2).transpose(-1, -2) * hd ** -0.5
        P_rows[t] = (torch.logsumexp(s, -1), (img(do)[:, t] * o).sum(-1))
    tab = neighbour_backward_tables(B, T, rpf, N)
    got_dk, got_dv = torch.zeros_like(q), torch.zeros_like(v)
    written = torch.zeros(B * T * rpf, dtype=torch.int32)
    flat = lambda t: t.permute(1, 0, 2)                                              # (H, rows, hd)
    for item in range(B * T):
        ks = slice(int(tab["kv_start"][item]), int(tab["kv_start"][item]) + int(tab["kv_len"][item]))
        kj, vj = flat(k[ks]), flat(v[ks])
        dk_j, dv_j = torch.zeros_like(kj), torch.zeros_like(vj)
        for s0, ln in ((tab["q_start0"][item], tab["q_len0"][item]), (tab["q_start1"][item], tab["q_len1"][item])):
            if int(ln) == 0:
                continue
            qs = slice(int(s0), int(s0) + int(ln))
            frame = int(s0) // rpf
            b, t = divmod(frame, T)
            lse, delta = P_rows[t][0][b], P_rows[t][1][b]                            # (H, N)
            qq, dd = flat(q[qs]), flat(do[qs])
            P = torch.exp(qq @ kj.transpose(-1, -2) * hd ** -0.5 - lse[..., None])
            dS = P * (dd @ vj.transpose(-1, -2) - delta[..., None]) * hd ** -0.5
            dk_j += dS.transpose(-1, -2) @ qq
            dv_j += P.transpose(-1, -2) @ dd
        got_dk[ks] = dk_j.permute(1, 0, 2)
        got_dv[ks] = dv_j.permute(1, 0, 2)
        written[ks] += 1
    assert (written.view(B * T, rpf)[:, 1:] == 1).all() and (written.view(B * T, rpf)[:, 0] == 0).all()
    assert torch.allclose(img(got_dk), want_dk, rtol=1e-9, atol=1e-10)
    assert torch.allclose(img(got_dv), want_dv, rtol=1e-9, atol=1e-10)
