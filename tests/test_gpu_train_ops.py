"""GPU parity of the kernels of the decoder / DPT-head / tail training path (csrc/train_ops.cu, the conv
wgrad / split-K / ReLU-mask modes of vs_gemm, the key-centric dK/dV pass of vs_attention_backward) against
torch.autograd in fp32 -- which is how the reference obtains these gradients -- and against the
hand-derived oracles (oracle/decoder_backward_ref.py, adapter_backward_ref.py).  bf16 operands / fp32
accumulation; tolerances stated per test."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import adapter_backward_ref as ab
from oracle import decoder_backward_ref as db
from oracle import encoder_ref as er


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def _bf(t):
    return t.to(torch.bfloat16)


def _gen(seed):
    return torch.Generator().manual_seed(seed)


# ------------------------------------------------------------------------------------ vs_gemm: wgrad forms
@pytest.mark.parametrize("M,N,K,split", [(4096, 256, 320, 0), (1000, 96, 1024, 0), (16448, 1024, 1024, 0),
                                          (130, 64, 72, 1), (5, 768, 2304, 0)])
def test_gemm_tn_accumulate_split_k(cuda, lib, M, N, K, split):
    from vicasplat_b200 import ops
    g = _gen(M + N)
    dy = _bf(torch.randn((M, N), generator=g)).to(cuda)
    x = _bf(torch.randn((M, K), generator=g)).to(cuda)
    dW = torch.full((N, K), 0.25, device=cuda)
    ops.gemm_tn_acc(dy, x, dW, split_k=split)
    ref = dy.float().t() @ x.float() + 0.25
    assert _rel(dW, ref) < 2e-5          # fp32 accumulation, order differs between the K splits


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(2, 16, 16, 64, 128, 3), (3, 8, 8, 96, 256, 3), (1, 64, 64, 256, 256, 3),
                                               (3, 2, 2, 768, 256, 3), (2, 32, 32, 128, 128, 1), (5, 4, 4, 384, 256, 3)])
def test_conv_wgrad_matches_autograd(cuda, lib, n, h, w, cin, cout, k):
    from vicasplat_b200 import ops
    g = _gen(n * h + cin)
    x = _bf(torch.randn((n, h, w, cin), generator=g)).to(cuda)
    dy = _bf(torch.randn((n, h, w, cout), generator=g)).to(cuda)
    cin_pad = (cin + 63) // 64 * 64
    dW = torch.zeros((cout, k * k * cin_pad), device=cuda)
    ops.conv_wgrad(dy, x, dW, kh=k, kw=k, pad=k // 2)
    wt = torch.zeros((cout, cin, k, k), device=cuda, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wt, padding=k // 2).backward(dy.float().permute(0, 3, 1, 2))
    got = dW.view(cout, k, k, cin_pad)[..., :cin].permute(0, 3, 1, 2)
    assert _rel(got, wt.grad) < 2e-5
    assert torch.count_nonzero(dW.view(cout, k * k, cin_pad)[..., cin:]) == 0


def test_conv_wgrad_of_the_7x7_stem_view(cuda, lib):
    """the image stem (dpt_gs_head.py:113-118) runs as kh = 7, kw = 1 over windows of 8 px x 8 channels of a
    zero-bordered NHWC8 image: its weight gradient through the same overlapping view."""
    from vicasplat_b200 import ops
    g = _gen(7)
    n, H, W = 2, 32, 32
    img = (torch.rand((n, 3, H, W), generator=g) * 2 - 1).to(cuda)
    dy = _bf(torch.randn((n, H, W, 256), generator=g)).to(cuda)
    img8 = ops.image_nhwc8(img, pad=3)
    view = (n, H, W, 64, H + 6, 8, (W + 8) * 8, (H + 6) * (W + 8) * 8)
    dW = torch.zeros((256, 7 * 64), device=cuda)
    ops.conv_wgrad(dy, img8, dW, kh=7, kw=1, pad=0, view=view)
    wt = torch.zeros((256, 3, 7, 7), device=cuda, requires_grad=True)
    F.conv2d(_bf(img).float(), wt, padding=3).backward(dy.float().permute(0, 3, 1, 2))
    got = dW.view(256, 7, 8, 8)[:, :, :7, :3].permute(0, 3, 1, 2)
    assert _rel(got, wt.grad) < 2e-5


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 256, 256), (1, 64, 64, 256, 128), (3, 8, 8, 96, 256)])
def test_conv_dgrad_with_relu_mask_and_skip(cuda, lib, n, h, w, cin, cout):
    """dX of y = conv3x3(relu(x)): the flipped-tap conv of dY with the mask (x > 0) and the skip added in
    the epilogue (ResidualConvUnit backward, dpt_block.py:129-137)."""
    from vicasplat_b200 import ops
    from vicasplat_b200.train import _flipped
    g = _gen(cin + h)
    xp = torch.randn((n, h, w, cin), generator=g).to(cuda)
    wt = (torch.randn((cout, cin, 3, 3), generator=g) * 0.05).to(cuda)
    dy = _bf(torch.randn((n, h, w, cout), generator=g)).to(cuda)
    skip = _bf(torch.randn((n, h, w, cin), generator=g)).to(cuda)
    x_relu = _bf(F.relu(xp))
    xr = xp.clone().requires_grad_(True)
    F.conv2d(F.relu(xr).permute(0, 3, 1, 2), _bf(wt).float(), padding=1).backward(dy.float().permute(0, 3, 1, 2))
    wf = _flipped(wt)
    got = ops.conv_gemm_masked(dy, wf, kh=3, kw=3, pad=1, N=cin, mask=x_relu, res1=skip)
    # the mask comes from the bf16 copy: entries that round to zero are masked out
    ref = xr.grad * (x_relu.float() > 0) + skip.float()
    assert _rel(got, ref) < 6e-3
    got2 = ops.conv_gemm_masked(dy, wf, kh=3, kw=3, pad=1, N=cin, mask=x_relu)
    assert _rel(got2, xr.grad * (x_relu.float() > 0)) < 6e-3
    plain = ops.conv_gemm(dy, wf, kh=3, kw=3, pad=1, N=cin)
    xl = xp.clone().requires_grad_(True)
    F.conv2d(xl.permute(0, 3, 1, 2), _bf(wt).float(), padding=1).backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(plain, xl.grad) < 6e-3


def test_gemm_masked_rows_mode(cuda, lib):
    from vicasplat_b200 import ops
    g = _gen(3)
    M, N, K = 1000, 96, 256
    dy = _bf(torch.randn((M, N), generator=g)).to(cuda)
    w_t = _bf(torch.randn((K, N), generator=g) * 0.1).to(cuda)
    y = _bf(F.relu(torch.randn((M, K), generator=g))).to(cuda)
    got = ops.gemm_masked(dy, w_t, mask=y, out_scale=2.0)
    ref = (dy.float() @ w_t.float().t()) * 2.0 * (y.float() > 0)
    assert _rel(got, ref) < 6e-3


# ------------------------------------------------------------------------------------ decoder block kernels
@pytest.mark.parametrize("frames,rpf,C,skip", [(6, 18, 768, True), (3, 258, 768, True), (4, 257, 1024, False)])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_layernorm_mod_backward(cuda, lib, frames, rpf, C, skip, dt):
    from vicasplat_b200 import ops
    g = _gen(frames * rpf)
    rows = frames * rpf
    x = (torch.randn((rows, C), generator=g) * 2 + 0.3).to(cuda)
    gamma = (1 + 0.1 * torch.randn((C,), generator=g)).to(cuda)
    beta = (0.1 * torch.randn((C,), generator=g)).to(cuda)
    mod = (0.2 * torch.randn((frames, 2 * C), generator=g)).to(cuda)
    dh = torch.randn((rows, C), generator=g).to(dt).to(cuda)
    dres = torch.randn((rows, C), generator=g).to(cuda)
    # reference through autograd
    xr, gr, br, mr = (t.clone().requires_grad_(True) for t in (x, gamma, beta, mod))
    ln = F.layer_norm(xr, (C,), gr, br, 1e-6).view(frames, rpf, C)
    h = ln * (1 + mr[:, None, :C]) + mr[:, None, C:]
    w = torch.ones((frames, rpf, 1), device=cuda)
    if skip:
        w[:, 0] = 0
    (h * w * dh.float().view(frames, rpf, C)).sum().backward()
    fa, fb = torch.zeros((frames, C), device=cuda), torch.zeros((frames, C), device=cuda)
    dx = ops.layernorm_mod_backward(x, dh, gamma, frame_a=fa, frame_b=fb, frames=frames, rows_per_frame=rpf,
                                    skip_first=skip, scale=mod[:, :C], dres=dres)
    assert _rel(dx, xr.grad + dres) < 2e-5
    dmod = torch.zeros_like(mod)
    dgam, dbet = torch.full((C,), 0.5, device=cuda), torch.full((C,), -0.5, device=cuda)
    ops.adaln_reduce(fa, fb, gamma, beta, scale=mod[:, :C], dscale=dmod[:, :C], dshift=dmod[:, C:], dgamma=dgam,
                     dbeta=dbet)
    assert _rel(dmod, mr.grad) < 2e-5
    assert _rel(dgam - 0.5, gr.grad) < 2e-5 and _rel(dbet + 0.5, br.grad) < 2e-5
    # in place (dx aliases dres), no modulation
    fa.zero_(); fb.zero_()
    buf = dres.clone()
    ops.layernorm_mod_backward(x, dh, gamma, frame_a=fa, frame_b=fb, frames=frames, rows_per_frame=rpf,
                               skip_first=skip, dres=buf, dx=buf)
    x2 = x.clone().requires_grad_(True)
    (F.layer_norm(x2, (C,), gamma, beta, 1e-6).view(frames, rpf, C) * w * dh.float().view(frames, rpf, C)).sum().backward()
    assert _rel(buf, x2.grad + dres) < 2e-5


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_gate_residual_and_backward(cuda, lib, mode):
    from vicasplat_b200 import ops
    g = _gen(mode)
    frames, rpf, C = 5, 18, 768
    rows = frames * rpf
    x = torch.randn((rows, C), generator=g).to(cuda)
    br = _bf(torch.randn((rows, C), generator=g)).to(cuda)
    gate = (0.3 * torch.randn((frames, 3 * C), generator=g)).to(cuda)[:, C:2 * C]
    dout = torch.randn((rows, C), generator=g).to(cuda)
    brr, gr = br.float().requires_grad_(True), gate.clone().requires_grad_(True)
    fac = (1 + gr)[:, None, :].expand(frames, rpf, C)
    if mode == 1:
        fac = torch.cat([torch.ones((frames, 1, C), device=cuda), fac[:, 1:]], 1)
    elif mode == 2:
        fac = torch.cat([torch.zeros((frames, 1, C), device=cuda), fac[:, 1:]], 1)
    ref = x + (fac * brr.view(frames, rpf, C)).reshape(rows, C)
    ref.backward(dout)
    out = ops.gate_residual(x, br, gate=gate, rows_per_frame=rpf, first_row_mode=mode, out=torch.empty_like(x))
    assert torch.allclose(out, ref.detach(), rtol=1e-6, atol=1e-6)
    inplace = ops.gate_residual(x.clone(), br, gate=gate, rows_per_frame=rpf, first_row_mode=mode)
    assert torch.equal(inplace, out)
    dgate = torch.zeros((frames, C), device=cuda)
    cs = torch.zeros((C,), device=cuda)
    dbr = ops.gate_backward(dout, frames=frames, rows_per_frame=rpf, branch=br, gate=gate, dgate=dgate, colsum=cs,
                            first_row_mode=mode)
    assert _rel(dbr, brr.grad) < 4e-3                 # one bf16 rounding
    assert _rel(dgate, gr.grad) < 2e-5
    assert _rel(cs, brr.grad.sum(0)) < 2e-4


def test_silu_backward(cuda, lib):
    from vicasplat_b200 import ops
    g = _gen(0)
    x = torch.randn((7, 768), generator=g).to(cuda)
    dy = torch.randn((7, 768), generator=g).to(cuda)
    acc = torch.ones((7, 768), device=cuda)
    ops.silu_backward(x, dy, acc, accumulate=True)
    assert torch.allclose(acc, 1 + dy * db.silu_grad(x), rtol=1e-5, atol=1e-6)


def test_neighbour_attention_backward_key_centric(cuda, lib):
    """CrossNeighborAttention backward (backbone_vica.py:152-191): frames share key frames, so the dK/dV
    pass runs over key-centric items with two query segments.  Against autograd over fp32 attention."""
    from vicasplat_b200 import encoder_grad as eg, ops
    g = _gen(11)
    B, T, N, H, rpf = 2, 4, 33, 3, 35
    C = H * 64
    rows = B * T * rpf
    qkv = _bf(torch.randn((rows, 3 * C), generator=g)).to(cuda)
    do = _bf(torch.randn((rows, C), generator=g)).to(cuda)
    i32 = dict(dtype=torch.int32, device=cuda)
    t_idx = torch.arange(T)
    prev = torch.where(t_idx > 0, t_idx - 1, t_idx + 1)
    nxt = torch.where(t_idx < T - 1, t_idx + 1, t_idx - 1)
    base = torch.arange(B)[:, None] * T
    nb_q = ((base + t_idx[None]) * rpf + 1).reshape(-1).to(**i32)
    nb_k0 = ((base + prev[None]) * rpf + 1).reshape(-1).to(**i32)
    nb_k1 = ((base + nxt[None]) * rpf + 1).reshape(-1).to(**i32)
    nb_len = torch.full((B * T,), N, **i32)
    nb_len1 = ((prev != nxt)[None].expand(B, T).reshape(-1) * N).to(**i32)
    o = torch.zeros((rows, C), dtype=torch.bfloat16, device=cuda)
    lse = torch.zeros((rows, H), device=cuda)
    kw = dict(heads=H, q_start=nb_q, q_len=nb_len, kv_start0=nb_k0, kv_len0=nb_len, kv_start1=nb_k1,
              kv_len1=nb_len1, max_q_len=N, max_kv_len=2 * N, scale=0.125)
    ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, lse=lse, **kw)
    tb = {k: v.to(cuda) for k, v in eg.neighbour_backward_tables(B, T, rpf, N, 1).items()}
    tb["max_kv_len"] = N
    dqkv = torch.zeros_like(qkv)
    with pytest.raises(ValueError):
        ops.attention_backward(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, do, lse, dqkv[:, :C],
                               dqkv[:, C:2 * C], dqkv[:, 2 * C:], **kw)
    ops.attention_backward(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], o, do, lse, dqkv[:, :C], dqkv[:, C:2 * C],
                           dqkv[:, 2 * C:], dkv_tables=tb, **kw)
    # reference
    qr = qkv.float().clone().requires_grad_(True)
    v5 = qr.view(B, T, rpf, 3, H, 64)[:, :, 1:1 + N]
    outs = []
    for t in range(T):
        nb = sorted({int(prev[t]), int(nxt[t])})
        q = v5[:, t, :, 0].transpose(1, 2)
        k = torch.cat([v5[:, j, :, 1] for j in nb], 1).transpose(1, 2)
        v = torch.cat([v5[:, j, :, 2] for j in nb], 1).transpose(1, 2)
        outs.append(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C))
    ref_o = torch.stack(outs, 1)
    (ref_o * do.float().view(B, T, rpf, C)[:, :, 1:1 + N]).sum().backward()
    assert _rel(o.view(B, T, rpf, C)[:, :, 1:1 + N], ref_o) < 1e-2
    assert _rel(dqkv, qr.grad) < 2e-2
    assert torch.count_nonzero(dqkv.view(B * T, rpf, -1)[:, 0]) == 0     # camera rows are not items


# ------------------------------------------------------------------------------------ DPT operators
@pytest.mark.parametrize("n,h,w,c", [(2, 8, 8, 256), (1, 2, 2, 64), (3, 16, 12, 128), (1, 128, 128, 8)])
def test_upsample2x_backward_is_the_transpose(cuda, lib, n, h, w, c):
    from vicasplat_b200 import ops
    g = _gen(h * w)
    dy = _bf(torch.randn((n, 2 * h, 2 * w, c), generator=g)).to(cuda)
    got = ops.upsample2x_backward(dy)
    x = torch.zeros((n, c, h, w), device=cuda, requires_grad=True)
    F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True).backward(dy.float().permute(0, 3, 1, 2))
    assert _rel(got, x.grad.permute(0, 2, 3, 1)) < 4e-3
    # <up(a), b> == <a, up^T(b)> with the kernel's own forward (same interpolation matrix)
    a = _bf(torch.randn((n, h, w, c), generator=g)).to(cuda)
    lhs = (ops.upsample2x(a).float() * dy.float()).sum()
    rhs = (a.float() * got.float()).sum()
    assert abs(lhs - rhs) <= 2e-2 * max(abs(lhs).item(), 1.0) + 1e-2 * (a.numel() ** 0.5)


def test_upsample2x_add(cuda, lib):
    from vicasplat_b200 import ops
    g = _gen(1)
    a = _bf(torch.randn((2, 8, 8, 64), generator=g)).to(cuda)
    e = _bf(torch.randn((2, 16, 16, 64), generator=g)).to(cuda)
    assert _rel(ops.upsample2x(a, add=e), ops.upsample2x(a).float() + e.float()) < 4e-3


def test_pixel_unshuffle_and_col2im(cuda, lib):
    from vicasplat_b200 import ops
    g = _gen(2)
    n, h, w, c, k = 3, 4, 4, 96, 4
    rows = _bf(torch.randn((n * h * w, k * k * c), generator=g)).to(cuda)
    assert torch.equal(ops.pixel_unshuffle(ops.pixel_shuffle(rows, n, h, w, c, k), k), rows)
    # col2im = transpose of im2col (3x3, stride 2, pad 1)
    n, h, w, c = 2, 16, 16, 64
    x = _bf(torch.randn((n, h, w, c), generator=g)).to(cuda)
    cols = ops.im2col(x, nchw_f32=False, n=n, h=h, w=w, c=c, k=3, stride=2, pad=1, kpad=9 * c)
    dcols = _bf(torch.randn(cols.shape, generator=g)).to(cuda)
    got = ops.col2im(dcols, n=n, h=h, w=w, c=c, k=3, stride=2, pad=1)
    xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    unf = F.unfold(xr, 3, padding=1, stride=2)                    # (n, c*9, L), channel-major
    d = dcols.float().view(n, -1, 9, c).permute(0, 3, 2, 1).reshape(n, c * 9, -1)
    (unf * d).sum().backward()
    assert _rel(got, xr.grad.permute(0, 2, 3, 1)) < 4e-3


def test_relu_backward_and_colsum(cuda, lib):
    from vicasplat_b200 import ops
    g = _gen(4)
    for rows, C in ((1000, 256), (77, 96), (4096, 128)):
        dy = _bf(torch.randn((rows, C), generator=g)).to(cuda)
        y = _bf(F.relu(torch.randn((rows, C), generator=g))).to(cuda)
        cs = torch.zeros((C,), device=cuda)
        dx = ops.relu_backward(dy, y, colsum=cs)
        ref = dy.float() * (y.float() > 0)
        assert torch.equal(dx.float(), ref)
        assert torch.allclose(cs, ref.sum(0), rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------------------------ tails
def test_pts_tail_backward(cuda, lib):
    from vicasplat_b200 import ops
    g = _gen(5)
    px, Cf = 1000, 128
    feat = _bf(F.relu(torch.randn((px, Cf), generator=g))).to(cuda)
    w = (torch.randn((3, Cf), generator=g) * 0.05).to(cuda)
    b = (torch.randn((3,), generator=g) * 0.1).to(cuda)
    dgs = torch.zeros((px, 96), device=cuda)
    dgs[:, 84:87] = torch.randn((px, 3), generator=g).to(cuda)
    fr, wr, br = feat.float().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    a = fr @ wr.t() + br
    d = a.norm(dim=-1, keepdim=True)
    (a / d.clamp_min(1e-8) * torch.expm1(d) * dgs[:, 84:87]).sum().backward()
    dw, dbias = torch.zeros((3, Cf), device=cuda), torch.zeros((3,), device=cuda)
    dfeat = ops.pts_tail_backward(feat, Cf, w, b, dgs[:, 84:], dw, dbias)
    assert _rel(dfeat, fr.grad * (feat.float() > 0)) < 4e-3
    assert _rel(dw, wr.grad) < 1e-4 and _rel(dbias, br.grad) < 1e-4


def test_gaussian_adapter_backward_against_oracle(cuda, lib):
    from vicasplat_b200 import ops
    cfg = er.EncoderConfig()
    g = _gen(6)
    G = 3000
    gsp = torch.zeros((G, 96), device=cuda)
    gsp[:, :83] = torch.randn((G, 83), generator=g).to(cuda)
    gsp[:, 84:87] = torch.randn((G, 3), generator=g).to(cuda)
    mask = er.sh_mask(cfg, cuda).float()
    d = {k: torch.randn(s, generator=g).to(cuda) for k, s in
         dict(raw=(G, 86), means=(G, 3), cov=(G, 3, 3), cov6=(G, 6), sh=(G, 3, 25), opac=(G,)).items()}
    raw = torch.cat([gsp[:, 84:87], gsp[:, :83]], -1).double()
    full = d["cov"].double().clone()
    iu = torch.triu_indices(3, 3)
    c6 = torch.zeros((G, 3, 3), dtype=torch.float64, device=cuda)
    c6[:, iu[0], iu[1]] = d["cov6"].double()
    ref = ab.adapter_backward(raw, cfg, d["means"].double(), full + c6, d["sh"].double(),
                              d["opac"].double()[:, None]) + d["raw"].double()
    out = torch.zeros((G, 96), device=cuda)
    ops.gaussian_adapter_backward(gsp, 25, mask, out, center_col=84, param_col=0, d_raw=d["raw"], d_means=d["means"],
                                  d_cov=d["cov"], d_cov6=d["cov6"], d_shs=d["sh"], d_opac=d["opac"])
    assert _rel(out[:, 84:87], ref[:, :3]) < 1e-5
    assert _rel(out[:, :83], ref[:, 3:]) < 2e-4
    assert torch.count_nonzero(out[:, 83]) == 0 and torch.count_nonzero(out[:, 87:]) == 0


def test_camera_head_backward(cuda, lib):
    from vicasplat_b200 import ops
    g = _gen(8)
    B, T, C = 2, 4, 768
    feat = torch.randn((B * T, C), generator=g).to(cuda)
    w = (torch.randn((8, C), generator=g) * 0.02).to(cuda)
    b = (torch.randn((8,), generator=g) * 0.1).to(cuda)
    dp = torch.randn((B, T - 1, 8), generator=g).to(cuda)
    fr, wr, br = feat.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    v = F.relu(fr).view(B, T, C)[:, 1:] @ wr.t() + br
    v = v + torch.tensor([0, 0, 0, 1.0, 0, 0, 0, 0], device=cuda)
    pred = v / v[..., :4].norm(dim=-1, keepdim=True)
    (pred * dp).sum().backward()
    got_pred, _ = ops.camera_head(feat, C, w, b, B, T, C)
    assert torch.allclose(got_pred, pred.detach(), rtol=1e-4, atol=1e-5)
    dw, dbias = torch.zeros((8, C), device=cuda), torch.zeros((8,), device=cuda)
    dfeat = ops.camera_head_backward(feat, w, b, B, T, C, dp, dw, dbias)
    assert _rel(dfeat, fr.grad) < 1e-4 and _rel(dw, wr.grad) < 1e-4 and _rel(dbias, br.grad) < 1e-4


def test_dropout_in_place(cuda, lib):
    """nn.Dropout(0.1) of the gs head (dpt_block.py:341): kept fraction, scale, determinism per seed."""
    from vicasplat_b200 import ops
    x = torch.ones((1 << 20,), dtype=torch.bfloat16, device=cuda)
    y = ops.dropout_(x.clone(), 0.1, 7)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.9) < 2e-3
    assert torch.allclose(y[y != 0].float(), torch.tensor(1 / 0.9), rtol=4e-3)
    assert torch.equal(y, ops.dropout_(x.clone(), 0.1, 7)) and not torch.equal(y, ops.dropout_(x.clone(), 0.1, 8))
    assert torch.equal(ops.dropout_(x.clone(), 0.0, 1), x)
    assert abs((y.float().mean().item()) - 1.0) < 3e-3          # unbiased
