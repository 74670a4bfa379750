"""GPU parity of the encoder's training (backward) path -- the ViT block of SURVEY.md §8 E2 -- against
torch.autograd over the fp32 oracle pieces (oracle/encoder_ref.py), which is how the reference itself
obtains these gradients.  bf16 operands / fp32 accumulation: tolerances are stated per test."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import encoder_ref as er


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


def _bf(t):
    return t.to(torch.bfloat16)


# ------------------------------------------------------------------------------------ grad_prep
@pytest.mark.parametrize("rows,cols", [(64, 64), (300, 132), (257, 1024), (2056, 3072), (1, 4), (70, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_grad_prep_copy_transpose_colsum(cuda, lib, rows, cols, dtype):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(rows * 7 + cols)
    src = torch.randn((rows, cols), generator=g).to(dtype).to(cuda)
    colsum = torch.full((cols,), 0.5, device=cuda)          # accumulated on top of what is there
    copy, tr = ops.grad_prep(src, colsum=colsum)
    ref = src.to(torch.bfloat16)
    assert torch.equal(copy, ref)                             # bit-exact: one bf16 rounding
    assert tr.shape == (cols, rows) and tr.stride(0) % 8 == 0
    assert torch.equal(tr, ref.t())
    base = tr.as_strided((cols, tr.stride(0)), (tr.stride(0), 1))
    assert torch.count_nonzero(base[:, rows:]) == 0           # pad columns are zeroed
    want = src.float().sum(0) + 0.5
    assert torch.allclose(colsum, want, rtol=1e-4, atol=1e-3 * math.sqrt(rows))
    only_t = ops.grad_prep(src, want_copy=False)[1]
    assert torch.equal(only_t, ref.t())


def test_grad_prep_strided_source_and_gelu_factor(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(5)
    big = torch.randn((200, 256), generator=g).to(cuda)
    src = big[:, 64:192]                                      # row stride 256, 128 columns
    z = _bf(torch.randn((200, 128), generator=g) * 2).to(cuda)
    colsum = torch.zeros((128,), device=cuda)
    copy, tr = ops.grad_prep(src, z=z, colsum=colsum)
    zf = z.float().requires_grad_(True)
    F.gelu(zf).backward(src)                                  # d/dz of gelu(z) . src
    ref = zf.grad
    assert _rel(copy, ref) < 4e-3                             # one bf16 rounding
    assert torch.equal(tr, copy.t())
    assert torch.allclose(colsum, ref.sum(0), rtol=1e-3, atol=1e-3)


def test_gelu_bf16(cuda, lib):
    from vicasplat_b200 import ops
    z = _bf(torch.linspace(-8, 8, 4096 * 8).reshape(8, 4096)).to(cuda)
    a = ops.gelu_bf16(z)
    ref = F.gelu(z.float())
    assert (a.float() - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item()
    assert _rel(a, ref) < 3e-3


# ------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,C", [(257, 1024), (2056, 768), (9, 128), (5000, 1024)])
@pytest.mark.parametrize("dy_dtype", [torch.float32, torch.bfloat16])
def test_layernorm_backward(cuda, lib, rows, C, dy_dtype):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn((rows, C), generator=g) * 1.5 + 0.3).to(cuda)
    gamma = (1 + 0.2 * torch.randn((C,), generator=g)).to(cuda)
    beta = (0.1 * torch.randn((C,), generator=g)).to(cuda)
    dy = torch.randn((rows, C), generator=g).to(dy_dtype).to(cuda)
    dres = torch.randn((rows, C), generator=g).to(cuda)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    F.layer_norm(xr, (C,), gr, br, 1e-6).backward(dy.float())
    dgamma, dbeta = torch.zeros_like(gamma), torch.zeros_like(beta)
    dx = ops.layernorm_backward(x, dy, gamma, dres=dres, dgamma=dgamma, dbeta=dbeta, eps=1e-6)
    assert _rel(dx - dres, xr.grad) < 2e-5
    assert _rel(dgamma, gr.grad) < 2e-5
    assert _rel(dbeta, br.grad) < 2e-5
    # in place on the residual-stream gradient, without parameter gradients
    buf = dres.clone()
    ops.layernorm_backward(x, dy, gamma, dres=buf, dx=buf, eps=1e-6)
    assert torch.allclose(buf, dx, rtol=1e-6, atol=1e-6)


# ------------------------------------------------------------------------------------ rope backward
def test_rope_rows_backward_is_the_transpose(cuda, lib):
    from vicasplat_b200 import ops
    from vicasplat_b200.encoder_grad import FrameLayout
    H, E = 16, 1024
    lay = FrameLayout.make(2, 16, 16, H, cuda)
    rows = lay.pos.shape[0]
    g = torch.Generator().manual_seed(3)
    a = _bf(torch.randn((rows, 3 * E), generator=g)).to(cuda)
    b = _bf(torch.randn((rows, 3 * E), generator=g)).to(cuda)
    Ra = ops.rope_rows(a.clone(), lay.pos, heads=H, q_col=0, k_col=E)
    Rtb = ops.rope_rows_backward(b.clone(), lay.pos, heads=H, q_col=0, k_col=E)
    # <R a, b> == <a, R^T b> on the rotated columns; the v columns are untouched by both
    lhs = (Ra.float() * b.float()).sum().item()
    rhs = (a.float() * Rtb.float()).sum().item()
    assert abs(lhs - rhs) <= 2e-3 * (a.float().norm() * b.float().norm()).item()
    assert torch.equal(Rtb[:, 2 * E:], b[:, 2 * E:])
    back = ops.rope_rows_backward(Ra.clone(), lay.pos, heads=H, q_col=0, k_col=E)
    assert _rel(back, a) < 8e-3                               # two bf16 roundings
    # against the oracle's rope under autograd
    q = a[:, :E].float().reshape(2, lay.n, H, 64).permute(0, 2, 1, 3).clone().requires_grad_(True)
    pos = lay.pos.view(2, lay.n, 2).long()
    er.rope2d(q, pos, 100.0).backward(b[:, :E].float().reshape(2, lay.n, H, 64).permute(0, 2, 1, 3))
    ref = q.grad.permute(0, 2, 1, 3).reshape(rows, E)
    assert _rel(Rtb[:, :E], ref) < 6e-3


# ------------------------------------------------------------------------------------ wgrad-form GEMM
@pytest.mark.parametrize("K,rows,N", [(64, 64, 64), (300, 200, 132), (2056, 1024, 1024), (514, 3072, 1024),
                                      (2056, 1024, 4096), (1000, 4096, 1024), (257, 128, 83), (72, 520, 8)])
@pytest.mark.parametrize("block_n", [0, 64])
def test_gemm_tn_mn_major_operands(cuda, lib, K, rows, N, block_n):
    """a_mode 2: C = A^T W with A (K, rows), W (K, N) used as stored (MN-major shared-memory tiles)."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(K + rows + N)
    # row strides must be multiples of 8 elements (TMA): odd widths are column slices of padded storage
    A = _bf(torch.randn((K, (rows + 7) // 8 * 8), generator=g)).to(cuda)[:, :rows]
    W = _bf(torch.randn((K, (N + 7) // 8 * 8), generator=g) / math.sqrt(K)).to(cuda)[:, :N]
    ref = A.float().t() @ W.float()
    out = ops.gemm(A, W, tn=True, out_dtype=torch.float32, block_n=block_n)
    assert out.shape == (rows, N)
    assert _rel(out, ref) < 2e-5
    acc = torch.full((rows, N), 0.5, device=cuda)              # accumulate in place (the wgrad use)
    ops.gemm(A, W, tn=True, out=acc, res1=acc, block_n=block_n)
    assert _rel(acc - 0.5, ref) < 2e-5


def test_gemm_tn_strided_views(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(77)
    bigA = _bf(torch.randn((700, 3 * 256), generator=g)).to(cuda)
    bigW = _bf(torch.randn((700, 512), generator=g)).to(cuda)
    A, W = bigA[:, 256:512], bigW[:, 128:328]                   # column slices: row strides 768 / 512
    out = ops.gemm(A, W, tn=True, out_dtype=torch.float32)
    assert _rel(out, A.float().t() @ W.float()) < 2e-5


# ------------------------------------------------------------------------------------ linear layer
@pytest.mark.parametrize("M,K,N", [(2056, 1024, 3072), (514, 768, 768), (2056, 4096, 1024), (300, 128, 64)])
def test_linear_backward(cuda, lib, M, K, N):
    from vicasplat_b200 import encoder_grad as eg, ops
    g = torch.Generator().manual_seed(M + K + N)
    x = _bf(torch.randn((M, K), generator=g)).to(cuda)
    W = _bf(torch.randn((N, K), generator=g) / math.sqrt(K)).to(cuda)
    dy32 = torch.randn((M, N), generator=g).to(cuda)
    db = torch.zeros((N,), device=cuda)
    dW = torch.full((N, K), 0.25, device=cuda)               # accumulated in place
    dy, _ = ops.grad_prep(dy32, want_t=False, colsum=db)
    dx = eg.linear_backward(dy, x, W.t().contiguous(), dW)
    dyf = dy.float()                                          # the operands the GEMMs really see
    assert _rel(dx, dyf @ W.float()) < 2e-5
    assert _rel(dW - 0.25, dyf.t() @ x.float()) < 2e-5
    assert _rel(db, dy32.sum(0)) < 1e-4
    # and against autograd on unrounded gradients: one bf16 rounding of dy
    xr, Wr = x.float().requires_grad_(True), W.float().requires_grad_(True)
    F.linear(xr, Wr).backward(dy32)
    assert _rel(dx, xr.grad) < 5e-3 and _rel(dW - 0.25, Wr.grad) < 5e-3


# ------------------------------------------------------------------------------------ MLP half of the block
def _block_sd(cfg, seed, device):
    sd = {k: v.to(device) for k, v in er.synth_state_dict(cfg, seed=seed).items()
          if k.startswith("backbone.enc_blocks.0.")}
    g = torch.Generator().manual_seed(seed + 1)               # non-trivial LayerNorm parameters and biases
    for k in sd:
        if k.endswith("bias"):
            sd[k] = (0.1 * torch.randn(sd[k].shape, generator=g)).to(device)
        elif ".norm" in k:
            sd[k] = (1 + 0.1 * torch.randn(sd[k].shape, generator=g)).to(device)
    return sd


def _cfg():
    return er.EncoderConfig(enc_depth=1, dec_depth=4)   # only enc_blocks.0 is used


def _check_grads(g, sd, names, tol):
    for name in names:
        ref = sd["backbone.enc_blocks.0." + name].grad
        assert ref is not None, name
        assert _rel(g[name], ref) < tol, (name, _rel(g[name], ref))


def test_mlp_half_forward_backward(cuda, lib):
    from vicasplat_b200 import encoder_grad as eg
    cfg = _cfg()
    sd = _block_sd(cfg, 0, cuda)
    key = "backbone.enc_blocks.0"
    w = eg.pack_block(sd, key, cuda)
    g = eg.zero_grads(w)
    gen = torch.Generator().manual_seed(11)
    M, E = 2 * 257, cfg.enc_embed_dim
    x = torch.randn((M, E), generator=gen).to(cuda)
    dout = torch.randn((M, E), generator=gen).to(cuda)
    saved = eg.Saved()
    out = eg.mlp_half_forward(x, w, saved)
    dx = eg.mlp_half_backward(dout, w, g, saved)
    for v in sd.values():
        v.requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = xr + er.mlp(sd, key + ".mlp", er.layer_norm(sd, key + ".norm2", xr, cfg.ln_eps))
    ref.backward(dout)
    assert _rel(out, ref) < 5e-3
    assert _rel(dx, xr.grad) < 1e-2
    _check_grads(g, sd, ["mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias",
                         "norm2.weight", "norm2.bias"], 2e-2)


# ------------------------------------------------------------------------------------ attention backward
def _sdpa_grads(q, k, v, dO, mask=None):
    """fp32 autograd through the oracle's softmax attention on the bf16-rounded operands.
    q, k, v, dO: (..., rows, 64) float."""
    q, k, v = (t.clone().requires_grad_(True) for t in (q, k, v))
    o = er.sdpa(q, k, v, mask)
    o.backward(dO)
    s = (q.detach() @ k.detach().transpose(-1, -2)) * 0.125
    if mask is not None:
        s = s.masked_fill(~mask, float("-inf"))
    return o.detach(), q.grad, k.grad, v.grad, torch.logsumexp(s, -1)


def test_attention_backward_encoder_style(cuda, lib):
    """per-frame self-attention on a packed qkv buffer, 257 tokens (partial query and key tiles)."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(21)
    Fr, N, H = 3, 257, 16
    C = H * 64
    qkv = _bf(torch.randn((Fr * N, 3 * C), generator=g)).to(cuda)
    dO = _bf(torch.randn((Fr * N, C), generator=g)).to(cuda)
    O = torch.zeros((Fr * N, C), dtype=torch.bfloat16, device=cuda)
    lse = torch.zeros((Fr * N, H), device=cuda)
    st = torch.arange(Fr, dtype=torch.int32, device=cuda) * N
    ln = torch.full((Fr,), N, dtype=torch.int32, device=cuda)
    kw = dict(heads=H, q_start=st, q_len=ln, kv_start0=st, kv_len0=ln, max_q_len=N, max_kv_len=N, scale=0.125)
    ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], O, lse=lse, **kw)
    dqkv = torch.full_like(qkv, float("nan"))                 # every element must be written
    ops.attention_backward(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], O, dO, lse,
                           dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], **kw)
    t = qkv.float().view(Fr, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    o_ref, dq, dk, dv, lse_ref = _sdpa_grads(t[0], t[1], t[2], dO.float().view(Fr, N, H, 64).permute(0, 2, 1, 3))
    back = lambda x: x.permute(0, 2, 1, 3).reshape(Fr * N, C)
    assert _rel(O, back(o_ref)) < 1e-2
    assert (lse * math.log(2.0) - lse_ref.permute(0, 2, 1).reshape(Fr * N, H)).abs().max().item() < 2e-3
    assert torch.isfinite(dqkv.float()).all()
    assert _rel(dqkv[:, :C], back(dq)) < 2e-2
    assert _rel(dqkv[:, C:2 * C], back(dk)) < 2e-2
    assert _rel(dqkv[:, 2 * C:], back(dv)) < 2e-2
    tail = torch.arange(Fr, device=cuda) * N + N - 1          # the 257-th row of every frame
    assert _rel(dqkv[tail], torch.cat([back(dq), back(dk), back(dv)], 1)[tail]) < 2e-2


def test_attention_backward_video_with_camera_mask(cuda, lib):
    """one scene: T frames x (1 camera + N image) rows; camera rows see frames <= t only."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(22)
    T, N, H = 4, 65, 12
    rpf = N + 1
    rows = T * rpf
    C = H * 64
    qkv = _bf(torch.randn((rows, 3 * C), generator=g)).to(cuda)
    dO = _bf(torch.randn((rows, C), generator=g)).to(cuda)
    O = torch.zeros((rows, C), dtype=torch.bfloat16, device=cuda)
    lse = torch.zeros((rows, H), device=cuda)
    i32 = dict(dtype=torch.int32, device=cuda)
    kw = dict(heads=H, q_start=torch.zeros(1, **i32), q_len=torch.full((1,), rows, **i32),
              kv_start0=torch.zeros(1, **i32), kv_len0=torch.full((1,), rows, **i32),
              max_q_len=rows, max_kv_len=rows, causal_block=rpf, scale=0.125)
    ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], O, lse=lse, **kw)
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attention_backward(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], O, dO, lse,
                           dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], **kw)
    t = qkv.float().view(rows, 3, H, 64).permute(1, 2, 0, 3)                  # (3,H,rows,64)
    mask = torch.ones((rows, rows), dtype=torch.bool, device=cuda)
    for f in range(T):
        mask[f * rpf, (f + 1) * rpf:] = False
    o_ref, dq, dk, dv, _ = _sdpa_grads(t[0], t[1], t[2], dO.float().view(rows, H, 64).permute(1, 0, 2), mask)
    back = lambda x: x.permute(1, 0, 2).reshape(rows, C)
    assert _rel(O, back(o_ref)) < 1e-2
    assert _rel(dqkv[:, :C], back(dq)) < 2e-2
    assert _rel(dqkv[:, C:2 * C], back(dk)) < 2e-2
    assert _rel(dqkv[:, 2 * C:], back(dv)) < 2e-2
    cam = torch.arange(T, device=cuda) * rpf
    assert _rel(dqkv[cam, :C], back(dq)[cam]) < 2e-2


def test_attention_backward_two_segments(cuda, lib):
    """keys of an item = two disjoint row segments (the layout of the neighbour attention, with
    key rows private to the item as this version of the backward pass requires)."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(23)
    Fr, N, H = 3, 150, 4
    C = H * 64
    q = _bf(torch.randn((Fr * N, C), generator=g)).to(cuda)
    k = _bf(torch.randn((2 * Fr * N, C), generator=g)).to(cuda)
    v = _bf(torch.randn((2 * Fr * N, C), generator=g)).to(cuda)
    dO = _bf(torch.randn((Fr * N, C), generator=g)).to(cuda)
    O = torch.zeros((Fr * N, C), dtype=torch.bfloat16, device=cuda)
    lse = torch.zeros((Fr * N, H), device=cuda)
    fr = torch.arange(Fr)
    i32 = dict(dtype=torch.int32, device=cuda)
    len1 = torch.tensor([N, 0, 70])                            # full, absent and partial second segment
    kw = dict(heads=H, q_start=(fr * N).to(**i32), q_len=torch.full((Fr,), N, **i32),
              kv_start0=(fr * N).to(**i32), kv_len0=torch.full((Fr,), N, **i32),
              kv_start1=((Fr + fr) * N).to(**i32), kv_len1=len1.to(**i32),
              max_q_len=N, max_kv_len=2 * N, scale=0.125)
    ops.attention(q, k, v, O, lse=lse, **kw)
    dq = torch.full_like(q, float("nan"))
    dk, dv = torch.zeros_like(k), torch.zeros_like(v)
    ops.attention_backward(q, k, v, O, dO, lse, dq, dk, dv, **kw)
    for i in range(Fr):
        rows0 = slice(i * N, (i + 1) * N)
        rows1 = slice((Fr + i) * N, (Fr + i) * N + int(len1[i]))
        heads = lambda x: x.float().view(-1, H, 64).permute(1, 0, 2)
        kk = torch.cat([heads(k[rows0]), heads(k[rows1])], 1)
        vv = torch.cat([heads(v[rows0]), heads(v[rows1])], 1)
        o_ref, rq, rk, rv, _ = _sdpa_grads(heads(q[rows0]), kk, vv, heads(dO[rows0]))
        back = lambda x: x.permute(1, 0, 2).reshape(-1, C)
        assert _rel(O[rows0], back(o_ref)) < 1e-2, i
        assert _rel(dq[rows0], back(rq)) < 2e-2, i
        assert _rel(torch.cat([dk[rows0], dk[rows1]]), back(rk)) < 2e-2, i
        assert _rel(torch.cat([dv[rows0], dv[rows1]]), back(rv)) < 2e-2, i
    # rows of K / V that belong to no item are left alone
    unused = slice((Fr + 1) * N, (Fr + 2) * N)
    assert (dk[unused] == 0).all() and (dv[unused] == 0).all()


# ------------------------------------------------------------------------------------ the whole ViT block
def test_vit_block_forward_backward(cuda, lib):
    """croco/blocks.py:81-130 end to end: every parameter gradient and the input gradient of one
    encoder block against torch.autograd over the fp32 oracle block."""
    from vicasplat_b200 import encoder_grad as eg
    cfg = _cfg()
    sd = _block_sd(cfg, 1, cuda)
    key = "backbone.enc_blocks.0"
    w = eg.pack_block(sd, key, cuda)
    g = eg.zero_grads(w)
    Fr, gh = 3, 16
    lay = eg.FrameLayout.make(Fr, gh, gh, cfg.enc_num_heads, cuda)
    M, E = Fr * lay.n, cfg.enc_embed_dim
    gen = torch.Generator().manual_seed(12)
    x = torch.randn((M, E), generator=gen).to(cuda)
    dout = torch.randn((M, E), generator=gen).to(cuda)
    saved = eg.Saved()
    out = eg.block_forward(x, w, lay, saved)
    dx = eg.block_backward(dout, w, g, lay, saved)
    for v in sd.values():
        v.requires_grad_(True)
    xr = x.view(Fr, lay.n, E).clone().requires_grad_(True)
    pos = er.positions(Fr, gh, gh, True, device=cuda)
    ref = er.enc_block(sd, key, xr, pos, cfg)
    ref.backward(dout.view(Fr, lay.n, E))
    assert _rel(out, ref.reshape(M, E)) < 5e-3
    assert _rel(dx, xr.grad.reshape(M, E)) < 1.5e-2
    _check_grads(g, sd, ["mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias",
                         "norm2.weight", "norm2.bias", "attn.proj.weight", "attn.proj.bias",
                         "attn.qkv.weight", "attn.qkv.bias", "norm1.weight", "norm1.bias"], 3e-2)


# ------------------------------------------------------------------------------------ the ViT encoder (E1-E3)
@pytest.mark.parametrize("frames,hw,depth", [(2, 64, 2), (3, 256, 2)])
def test_vit_encoder_trainer_gradients(cuda, lib, frames, hw, depth):
    """patch embedding + intrinsic token + blocks + enc_norm (backbone_vica.py:450-480,535-541):
    every parameter gradient against torch.autograd over the oracle's encode_image."""
    from vicasplat_b200.encoder_train import VitEncoderTrainer
    cfg = er.EncoderConfig(enc_depth=depth, dec_depth=4)
    sd = {k: v.to(cuda) for k, v in er.synth_state_dict(cfg, seed=3).items()
          if k.startswith(("backbone.enc_", "backbone.patch_embed", "backbone.intrinsic_encoder"))}
    g = torch.Generator().manual_seed(4)
    for k in sd:                                               # non-trivial biases / LayerNorm parameters
        if k.endswith("bias"):
            sd[k] = (0.1 * torch.randn(sd[k].shape, generator=g)).to(cuda)
        elif "norm" in k:
            sd[k] = (1 + 0.1 * torch.randn(sd[k].shape, generator=g)).to(cuda)
    img = (torch.rand((frames, 3, hw, hw), generator=g) * 2 - 1).to(cuda)
    K = torch.tensor([[0.86, 0, 0.5], [0, 0.86, 0.5], [0, 0, 1.0]]).repeat(frames, 1, 1)
    K = (K + 0.05 * torch.randn(K.shape, generator=g)).to(cuda)
    tr = VitEncoderTrainer(sd, cfg, frames, (hw, hw), cuda)
    out = tr.forward(img, K)
    d_out = torch.randn(out.shape, generator=g).to(cuda)
    tr.backward(d_out)
    for v in sd.values():
        v.requires_grad_(True)
    ref, _ = er.encode_image(sd, img, K, cfg)
    ref.backward(d_out.view_as(ref))
    assert _rel(out, ref.reshape(out.shape)) < 1e-2
    worst = {}
    for name, prm in tr.params.items():
        worst[name] = _rel(prm.grad, sd[name].grad)
    bad = {k: v for k, v in worst.items() if not v < 4e-2}
    assert not bad, bad
    assert set(tr.params) == set(sd)                           # nothing of E1-E3 is left without a gradient


def test_trainer_step_with_fused_adamw(cuda, lib):
    """forward / backward / FusedAdamW / repack: the loss of a fixed regression target goes down and
    the re-packed bf16 operands follow the fp32 masters."""
    from vicasplat_b200.encoder_train import VitEncoderTrainer
    from vicasplat_b200.optim import FusedAdamW
    cfg = er.EncoderConfig(enc_depth=2, dec_depth=4)
    sd = {k: v.to(cuda) for k, v in er.synth_state_dict(cfg, seed=5).items()}
    g = torch.Generator().manual_seed(6)
    img = (torch.rand((2, 3, 64, 64), generator=g) * 2 - 1).to(cuda)
    K = torch.tensor([[0.86, 0, 0.5], [0, 0.86, 0.5], [0, 0, 1.0]]).repeat(2, 1, 1).to(cuda)
    tr = VitEncoderTrainer(sd, cfg, 2, (64, 64), cuda)
    opt = FusedAdamW(tr.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05, max_grad_norm=0.5)
    target = torch.randn((2 * tr.lay.n, cfg.enc_embed_dim), generator=g).to(cuda)
    losses = []
    for _ in range(6):
        out = tr.forward(img, K).float()
        losses.append(((out - target) ** 2).mean().item())
        tr.backward(2 * (out - target) / out.numel())
        opt.step()
        tr.repack()
    assert losses[-1] < losses[0]
    k = "backbone.enc_blocks.1.mlp.fc1.weight"
    assert torch.equal(tr.w[1]["mlp.fc1"], tr.params[k].detach().to(torch.bfloat16))
    assert torch.equal(tr.w[1]["mlp.fc1.t"], tr.params[k].detach().to(torch.bfloat16).t())
    assert not torch.equal(tr.params[k].detach(), sd[k])       # the masters moved


# ------------------------------------------------------------------------------------ drop-in nn.Module
def test_dropin_block_trains_through_torch_autograd(cuda, lib):
    """vicasplat_b200.blocks.Block in the place of croco/blocks.py:115-130: ordinary nn.Parameters,
    gradients through torch.autograd, parity with the oracle block; operand copies follow the
    parameters after an in-place update."""
    from functools import partial
    from torch import nn
    from vicasplat_b200.blocks import Block

    class Rope:
        base = 100.0

    cfg = _cfg()
    sd = _block_sd(cfg, 2, cuda)
    key = "backbone.enc_blocks.0"
    blk = Block(cfg.enc_embed_dim, cfg.enc_num_heads, cfg.mlp_ratio, qkv_bias=True,
                norm_layer=partial(nn.LayerNorm, eps=cfg.ln_eps), rope=Rope()).to(cuda)
    blk.load_state_dict({k[len(key) + 1:]: v for k, v in sd.items()}, strict=True)
    Fr, gh = 2, 16
    pos = er.positions(Fr, gh, gh, True, device=cuda)
    gen = torch.Generator().manual_seed(31)
    x = torch.randn((Fr, pos.shape[1], cfg.enc_embed_dim), generator=gen).to(cuda).requires_grad_(True)
    dout = torch.randn(x.shape, generator=gen).to(cuda)
    out = blk(x, pos)
    out.backward(dout)
    for v in sd.values():
        v.requires_grad_(True)
    xr = x.detach().clone().requires_grad_(True)
    ref = er.enc_block(sd, key, xr, pos, cfg)
    ref.backward(dout)
    assert _rel(out, ref) < 5e-3
    assert _rel(x.grad, xr.grad) < 1.5e-2
    for name, p in blk.named_parameters():
        assert _rel(p.grad, sd[f"{key}.{name}"].grad) < 3e-2, name
    # an optimizer step moves the parameters in place -> the bf16 operand copies are rebuilt
    before = blk._packed()["mlp.fc1"].clone()
    with torch.no_grad():
        blk.mlp.fc1.weight.add_(0.5)
    out2 = blk(x.detach(), pos)
    assert not torch.equal(blk._packed()["mlp.fc1"], before)
    assert not torch.allclose(out2, out.detach())


def test_vit_encoder_trainer_against_reference_gradient_golden(cuda, lib):
    """The CUDA backward pass against gradients of the UNMODIFIED reference modules
    (tests/golden/encoder_grad_small.npz, written by oracle/make_encoder_grad_golden.py): same seeded
    weights, clip, intrinsics and output gradient; every one of the 30 parameters."""
    from pathlib import Path
    import numpy as np
    from oracle import make_encoder_grad_golden as gg
    from oracle.make_encoder_golden import synth_inputs
    from vicasplat_b200.encoder_train import VitEncoderTrainer
    gold = np.load(Path(__file__).parent / "golden" / "encoder_grad_small.npz")
    cfg = er.EncoderConfig(**gg.CASE)
    sd = er.synth_state_dict(cfg, seed=0)
    image, K = synth_inputs(1, gg.FRAMES, cfg.img_size)
    tr = VitEncoderTrainer(sd, cfg, gg.FRAMES, (cfg.img_size, cfg.img_size), cuda)
    out = tr.forward(image[0].to(cuda), K[0].to(cuda))
    n = tr.lay.n
    got = out.float().view(gg.FRAMES, n, -1)[:, ::4, ::16].cpu().numpy()
    assert np.abs(got - gold["out_sub"]).max() <= 2e-2 * np.abs(gold["out_sub"]).max()
    tr.backward(gg.output_grad((gg.FRAMES, n, cfg.enc_embed_dim)).reshape(gg.FRAMES * n, -1).to(cuda))
    keys = gg.path_keys(sd)
    assert set(keys) == set(tr.params)
    for k in keys:
        g = tr.params[k].grad
        ref_norm = float(gold["norm/" + k])
        assert abs(g.double().norm().item() - ref_norm) <= 2e-2 * ref_norm, (k, g.norm().item(), ref_norm)
        sample = torch.from_numpy(gold["sample/" + k])
        mine = g.flatten()[::gg.SAMPLE].cpu()
        assert ((mine - sample).norm() / sample.norm().clamp_min(1e-12)).item() < 4e-2, k
