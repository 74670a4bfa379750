"""GPU parity of the individual sm_100a kernels (through the C-ABI) against fp32 torch / the oracle
pieces in oracle/encoder_ref.py.  bf16 operands, fp32 accumulation: tolerances are stated per test."""
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


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,K,N", [(2056, 1024, 3072), (257, 768, 2304), (100, 768, 768),
                                   (2056, 4096, 1024), (4112, 256, 83), (7, 768, 8),
                                   (1, 768, 4608)])
@pytest.mark.parametrize("block_n", [0, 64])
def test_gemm_rows(cuda, lib, M, K, N, block_n):
    from vicasplat_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + K + N)
    A = _bf(torch.randn((M, K), generator=g)).to(cuda)
    W = _bf(torch.randn((N, K), generator=g) / math.sqrt(K)).to(cuda)
    bias = torch.randn((N,), generator=g).to(cuda)
    out = ops.gemm(A, W, bias=bias, out_dtype=torch.float32, block_n=block_n)
    ref = A.float() @ W.float().T + bias
    assert out.shape == (M, N)
    assert _rel(out, ref) < 2e-5                      # fp32 accumulate of exact bf16 products
    out16 = ops.gemm(A, W, bias=bias, out_dtype=torch.bfloat16, block_n=block_n)
    assert _rel(out16, ref) < 4e-3                    # one bf16 rounding of the result


@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_gemm_tile_widths(cuda, lib, block_n):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(block_n)
    A = _bf(torch.randn((300, 320), generator=g)).to(cuda)
    W = _bf(torch.randn((512, 320), generator=g)).to(cuda)
    out = ops.gemm(A, W, out_dtype=torch.float32, block_n=block_n)
    assert _rel(out, A.float() @ W.float().T) < 2e-5


def test_gemm_epilogue_gelu_residual_gate_mapping(cuda, lib):
    from vicasplat_b200 import ops, _lib
    g = torch.Generator().manual_seed(7)
    frames, n_in, n_out, K, N = 3, 257, 258, 768, 768
    A = _bf(torch.randn((frames * n_in, K), generator=g)).to(cuda)
    W = _bf(torch.randn((N, K), generator=g) / math.sqrt(K)).to(cuda)
    bias = torch.randn((N,), generator=g).to(cuda)
    gate = (0.5 * torch.randn((frames, N), generator=g)).to(cuda)
    x = torch.randn((frames * n_out, N), generator=g).to(cuda)
    x0 = x.clone()
    # rows of frame f land at f*258 + 1 + r, gated by gate[f], added to the residual stream in place
    ops.gemm(A, W, bias=bias, act=_lib.VS_ACT_GELU, gate=gate, gate_rows=n_out, res1=x, out=x,
             out_gin=n_in, out_gout=n_out, out_off=1)
    ref = x0.clone().view(frames, n_out, N)
    y = F.gelu(A.float() @ W.float().T + bias).view(frames, n_in, N) * (1 + gate[:, None])
    ref[:, 1:] += y
    assert _rel(x, ref.view(-1, N)) < 2e-5
    assert torch.equal(x.view(frames, n_out, N)[:, 0], x0.view(frames, n_out, N)[:, 0])


def test_gemm_first_row_modes_and_relu_copy(cuda, lib):
    from vicasplat_b200 import ops, _lib
    g = torch.Generator().manual_seed(8)
    frames, rpf, K, N = 4, 130, 256, 192
    A = _bf(torch.randn((frames * rpf, K), generator=g)).to(cuda)
    W = _bf(torch.randn((N, K), generator=g) / math.sqrt(K)).to(cuda)
    gate = torch.randn((frames, N), generator=g).to(cuda)
    res = torch.randn((frames * rpf, N), generator=g).to(cuda)
    y = (A.float() @ W.float().T).view(frames, rpf, N)
    # mode 1: first row of each frame is not gated
    out = ops.gemm(A, W, gate=gate, gate_rows=rpf, first_row_mode=1, res1=res,
                   out_dtype=torch.float32)
    ref = y * (1 + gate[:, None])
    ref[:, 0] = y[:, 0]
    assert _rel(out, ref.view(-1, N) + res) < 2e-5
    # mode 2: first row of each frame is not written at all
    out2 = torch.full((frames * rpf, N), 123.0, device=cuda)
    relu_copy = torch.zeros((frames * rpf, N), dtype=torch.bfloat16, device=cuda)
    ops.gemm(A, W, gate=gate, gate_rows=rpf, first_row_mode=2, out=out2, out2=relu_copy)
    o = out2.view(frames, rpf, N)
    assert (o[:, 0] == 123.0).all()
    assert _rel(o[:, 1:], (y * (1 + gate[:, None]))[:, 1:]) < 2e-5
    assert _rel(relu_copy.view(frames, rpf, N)[:, 1:], F.relu(o[:, 1:])) < 4e-3


def _pack_conv_weight(w, cin_pad):
    n, cin, kh, kw = w.shape
    p = torch.zeros((n, kh * kw, cin_pad), dtype=torch.float32)
    p[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(n, kh * kw, cin)
    return p.reshape(n, -1)


@pytest.mark.parametrize("n,h,w,cin,cout,k", [(2, 16, 16, 64, 128, 3), (3, 8, 8, 96, 256, 3),
                                              (1, 64, 64, 256, 256, 3), (2, 32, 32, 256, 83, 1),
                                              (5, 4, 4, 192, 256, 3), (1, 128, 128, 128, 128, 3)])
def test_gemm_conv(cuda, lib, n, h, w, cin, cout, k):
    from vicasplat_b200 import ops, _lib
    g = torch.Generator().manual_seed(n * h + cin)
    x = _bf(torch.randn((n, cin, h, w), generator=g))
    wt = _bf(torch.randn((cout, cin, k, k), generator=g) / math.sqrt(cin * k * k))
    bias = torch.randn((cout,), generator=g)
    cin_pad = (cin + 63) // 64 * 64
    Wp = _bf(_pack_conv_weight(wt.float(), cin_pad)).to(cuda)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(cuda)
    res = _bf(torch.randn((n, h, w, cout), generator=g)).to(cuda)
    out = ops.conv_gemm(x_nhwc, Wp, kh=k, kw=k, pad=k // 2, N=cout, bias=bias.to(cuda),
                        act=_lib.VS_ACT_NONE, res1=res, out_dtype=torch.float32)
    ref = F.conv2d(x.float(), wt.float(), bias, padding=k // 2).permute(0, 2, 3, 1).to(cuda)
    assert _rel(out, ref + res.float()) < 2e-5


def test_conv7_stem_as_overlapping_view_with_upsampled_residual(cuda, lib):
    """dpt_gs_head.py:113-118,148-150: relu(conv7x7(img)) + bilinear_x2(path_1), no im2col buffer."""
    from vicasplat_b200 import ops, _lib
    g = torch.Generator().manual_seed(21)
    n, H, W, Cout = 3, 32, 48, 256
    img = torch.randn((n, 3, H, W), generator=g).to(cuda)
    w7 = (torch.randn((Cout, 3, 7, 7), generator=g) / 12.0).to(cuda)
    bias = torch.randn((Cout,), generator=g).to(cuda)
    p1 = _bf(torch.randn((n, H // 2, W // 2, Cout), generator=g)).to(cuda)
    wp = torch.zeros((Cout, 7, 8, 8), device=cuda)
    wp[:, :, :7, :3] = w7.permute(0, 2, 3, 1)
    img8 = ops.image_nhwc8(img, pad=3)
    assert img8.shape == (n, H + 6, W + 8, 8)
    assert (img8[:, :3] == 0).all() and (img8[:, :, :3] == 0).all() and (img8[..., 3:] == 0).all()
    out = ops.conv_gemm(img8, _bf(wp.reshape(Cout, -1)), kh=7, kw=1, pad=0, N=Cout, bias=bias,
                        act=_lib.VS_ACT_RELU, res1=p1, res_up2=True, out_dtype=torch.float32,
                        view=(n, H, W, 64, H + 6, 8, (W + 8) * 8, (H + 6) * (W + 8) * 8))
    conv = F.conv2d(_bf(img).float(), _bf(w7).float(), bias, padding=3).relu()
    up = F.interpolate(p1.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear",
                       align_corners=True)
    ref = (conv + _bf(up).float()).permute(0, 2, 3, 1)
    assert _rel(out, ref) < 1e-3
    assert (out - ref).abs().max() < 0.05     # bf16 rounding of the upsampled map may flip by 1 ulp


# ------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("C", [768, 1024, 256])
def test_layernorm_plain_and_modulated(cuda, lib, C):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(C)
    frames, rpf = 3, 258
    x = (torch.randn((frames * rpf, C), generator=g) * 3 + 0.5).to(cuda)
    w, b = torch.randn((C,), generator=g).to(cuda), torch.randn((C,), generator=g).to(cuda)
    w0, b0 = torch.randn((C,), generator=g).to(cuda), torch.randn((C,), generator=g).to(cuda)
    sc, sh = (0.3 * torch.randn((frames, C), generator=g)).to(cuda), torch.randn((frames, C), generator=g).to(cuda)
    y16, y32 = ops.layernorm(x, w, b, eps=1e-6, want_f32=True)
    ref = F.layer_norm(x, (C,), w, b, 1e-6)
    assert (y32 - ref).abs().max() < 2e-5 * ref.abs().max().clamp_min(1)
    assert _rel(y16, ref) < 4e-3
    # camera rows (first of each frame) use (w0, b0) and are not modulated
    _, y = ops.layernorm(x, w, b, eps=1e-6, w0=w0, b0=b0, scale=sc, shift=sh, rows_per_frame=rpf,
                         want_bf16=False, want_f32=True)
    xr = x.view(frames, rpf, C)
    ref = F.layer_norm(xr, (C,), w, b, 1e-6) * (1 + sc[:, None]) + sh[:, None]
    ref[:, 0] = F.layer_norm(xr[:, 0], (C,), w0, b0, 1e-6)
    assert (y.view(frames, rpf, C) - ref).abs().max() < 5e-5 * ref.abs().max()


# ------------------------------------------------------------------------------------ RoPE
def test_rope_rows_image_and_camera(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(3)
    T, N, H = 3, 17, 12
    rows = T * (N + 1)
    qkv = _bf(torch.randn((rows, 3 * H * 64), generator=g)).to(cuda)
    pos = er.positions(T, 4, 4, True)                                    # (T,17,2)
    pos_rows = torch.zeros((T, N + 1, 2), dtype=torch.int32)
    pos_rows[:, 1:] = pos.to(torch.int32)
    for t in range(T):
        pos_rows[t, 0, 0] = -1 - t                                        # camera token of frame t
    ref = qkv.float().view(T, N + 1, 3, H, 64).clone()
    for which in (0, 1):
        img = ref[:, 1:, which].permute(0, 2, 1, 3)                       # (T,H,N,64)
        ref[:, 1:, which] = er.rope2d(img, pos.to(cuda), 100.0).permute(0, 2, 1, 3)
        cam = ref[:, 0, which].permute(1, 0, 2)[None]                     # (1,H,T,64)
        ref[:, 0, which] = er.rope1d_interleaved(cam, torch.arange(T, device=cuda), 30.0)[0].permute(1, 0, 2)
    ops.rope_rows(qkv, pos_rows.view(rows, 2).to(cuda), heads=H, q_col=0, k_col=H * 64,
                  base=100.0, cam_theta=30.0)
    assert _rel(qkv, ref.view(rows, -1)) < 4e-3
    assert torch.equal(qkv.view(T, N + 1, 3, H, 64)[:, :, 2].float(), ref[:, :, 2])  # v untouched


def test_gemm_with_fused_rope_matches_oracle_rope(cuda, lib):
    """qkv projection with the rotary embedding applied in the epilogue (vs_gemm_params.rope_pos)
    against fp32 GEMM + the oracle's rope2d / interleaved camera rope (croco/blocks.py:101-103,
    rope_utils.py:297-305)."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(5)
    T, N, H, K = 3, 17, 12, 256
    rows = T * (N + 1)
    A = _bf(torch.randn((rows, K), generator=g)).to(cuda)
    W = _bf(torch.randn((3 * H * 64, K), generator=g) / math.sqrt(K)).to(cuda)
    bias = torch.randn((3 * H * 64,), generator=g).to(cuda)
    pos = er.positions(T, 4, 4, True)
    pos_rows = torch.zeros((T, N + 1, 2), dtype=torch.int32)
    pos_rows[:, 1:] = pos.to(torch.int32)
    for t in range(T):
        pos_rows[t, 0, 0] = -1 - t
    pos_rows = pos_rows.view(rows, 2).to(cuda).contiguous()
    ref = (A.float() @ W.float().t() + bias).view(T, N + 1, 3, H, 64).clone()
    for which in (0, 1):
        img = ref[:, 1:, which].permute(0, 2, 1, 3)
        ref[:, 1:, which] = er.rope2d(img, pos.to(cuda), 100.0).permute(0, 2, 1, 3)
        cam = ref[:, 0, which].permute(1, 0, 2)[None]
        ref[:, 0, which] = er.rope1d_interleaved(cam, torch.arange(T, device=cuda), 30.0)[0].permute(1, 0, 2)
    for dt, tol in ((torch.float32, 2e-5), (torch.bfloat16, 4e-3)):
        out = torch.empty((rows, 3 * H * 64), dtype=dt, device=cuda)
        ops.gemm(A, W, bias=bias, out=out, rope=(pos_rows, 0, H * 64, H, 100.0, 30.0))
        assert _rel(out, ref.view(rows, -1)) < tol, (dt, _rel(out, ref.view(rows, -1)))
    plain = ops.gemm(A, W, bias=bias, out_dtype=torch.float32)
    assert torch.equal(out.view(T, N + 1, 3, H, 64)[:, :, 2].float(),
                       plain.view(T, N + 1, 3, H, 64)[:, :, 2].to(torch.bfloat16).float())   # v untouched


# ------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, mask=None):
    return er.sdpa(q.float(), k.float(), v.float(), mask)


def test_attention_encoder_style(cuda, lib):
    """per-frame self-attention on a packed qkv buffer, 257 tokens (odd length -> tail tiles)."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(11)
    Fr, N, H = 3, 257, 16
    qkv = _bf(torch.randn((Fr * N, 3 * H * 64), generator=g)).to(cuda)
    O = torch.zeros((Fr * N, H * 64), dtype=torch.bfloat16, device=cuda)
    st = torch.arange(Fr, dtype=torch.int32, device=cuda) * N
    ln = torch.full((Fr,), N, dtype=torch.int32, device=cuda)
    ops.attention(qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:], O, heads=H,
                  q_start=st, q_len=ln, kv_start0=st, kv_len0=ln, max_q_len=N, scale=0.125)
    t = qkv.view(Fr, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = _attn_ref(t[0], t[1], t[2]).transpose(1, 2).reshape(Fr * N, H * 64)
    assert _rel(O, ref) < 1e-2
    # same call with the key bound given: the 257th row goes through the CUDA-core tail kernel
    O2 = torch.zeros_like(O)
    ops.attention(qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:], O2, heads=H,
                  q_start=st, q_len=ln, kv_start0=st, kv_len0=ln, max_q_len=N, max_kv_len=N,
                  scale=0.125)
    assert _rel(O2, ref) < 1e-2
    tail = torch.arange(Fr, device=cuda) * N + N - 1
    assert _rel(O2[tail], ref[tail]) < 1e-2


def test_attention_video_with_camera_mask(cuda, lib):
    """one scene: T frames x (1 camera + N image) rows; camera rows see frames <= t only."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(12)
    T, N, H = 4, 65, 12
    rpf = N + 1
    rows = T * rpf
    qkv = _bf(torch.randn((rows, 3 * H * 64), generator=g)).to(cuda)
    O = torch.zeros((rows, H * 64), dtype=torch.bfloat16, device=cuda)
    i32 = dict(dtype=torch.int32, device=cuda)
    ops.attention(qkv[:, :H * 64], qkv[:, H * 64:2 * H * 64], qkv[:, 2 * H * 64:], O, heads=H,
                  q_start=torch.zeros(1, **i32), q_len=torch.full((1,), rows, **i32),
                  kv_start0=torch.zeros(1, **i32), kv_len0=torch.full((1,), rows, **i32),
                  max_q_len=rows, causal_block=rpf, scale=0.125)
    t = qkv.view(rows, 3, H, 64).permute(1, 2, 0, 3)                      # (3,H,rows,64)
    mask = torch.ones((rows, rows), dtype=torch.bool, device=cuda)
    for f in range(T):
        mask[f * rpf, (f + 1) * rpf:] = False
    ref = _attn_ref(t[0], t[1], t[2], mask).transpose(0, 1).reshape(rows, H * 64)
    assert _rel(O, ref) < 1e-2
    cam_rows = torch.arange(T, device=cuda) * rpf
    assert _rel(O[cam_rows], ref[cam_rows]) < 1e-2


def test_attention_two_segments(cuda, lib):
    """neighbour cross-attention: queries of frame t, keys = rows of frame t-1 and t+1."""
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(13)
    T, N, H = 4, 257, 12
    rpf = N + 1
    rows = T * rpf
    C = H * 64
    q = _bf(torch.randn((rows, C), generator=g)).to(cuda)
    k = _bf(torch.randn((rows, C), generator=g)).to(cuda)
    v = _bf(torch.randn((rows, C), generator=g)).to(cuda)
    O = torch.zeros((rows, C), dtype=torch.bfloat16, device=cuda)
    fr = torch.arange(T)
    prev = torch.where(fr > 0, fr - 1, fr + 1)
    nxt = torch.where(fr < T - 1, fr + 1, fr - 1)
    two = prev != nxt
    i32 = dict(dtype=torch.int32, device=cuda)
    ops.attention(q, k, v, O, heads=H,
                  q_start=(fr * rpf + 1).to(**i32), q_len=torch.full((T,), N, **i32),
                  kv_start0=(prev * rpf + 1).to(**i32), kv_len0=torch.full((T,), N, **i32),
                  kv_start1=(nxt * rpf + 1).to(**i32), kv_len1=(two * N).to(**i32),
                  max_q_len=N, max_kv_len=2 * N, scale=0.125)
    qh = q.view(T, rpf, H, 64)[:, 1:].permute(0, 2, 1, 3)
    kh = k.view(T, rpf, H, 64)[:, 1:].permute(0, 2, 1, 3)
    vh = v.view(T, rpf, H, 64)[:, 1:].permute(0, 2, 1, 3)
    for t in range(T):
        nb = [prev[t].item(), nxt[t].item()]
        ref = _attn_ref(qh[t], torch.cat([kh[j] for j in nb], 1), torch.cat([vh[j] for j in nb], 1))
        got = O.view(T, rpf, H, 64)[t, 1:].permute(1, 0, 2)
        assert _rel(got, ref) < 1e-2, t
    assert (O.view(T, rpf, C)[:, 0] == 0).all()                           # camera rows untouched


# ------------------------------------------------------------------------------------ small ops
def test_patchify_matches_conv_weight_flattening(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(5)
    img = torch.randn((2, 3, 64, 48), generator=g).to(cuda)
    w = torch.randn((32, 3, 16, 16), generator=g).to(cuda)
    cols = ops.patchify(img, 16).float()
    ref = F.conv2d(_bf(img).float(), w, stride=16).flatten(2).transpose(1, 2).reshape(-1, 32)
    assert _rel(cols @ w.flatten(1).T, ref) < 1e-5


def test_im2col_strided_and_nchw(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = _bf(torch.randn((2, 8, 8, 64), generator=g)).to(cuda)              # NHWC
    w = torch.randn((16, 64, 3, 3), generator=g).to(cuda)
    cols = ops.im2col(x, nchw_f32=False, n=2, h=8, w=8, c=64, k=3, stride=2, pad=1, kpad=576)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=2, padding=1).permute(0, 2, 3, 1)
    got = cols.float() @ w.permute(0, 2, 3, 1).reshape(16, -1).T
    assert _rel(got, ref.reshape(-1, 16)) < 1e-5
    img = torch.randn((1, 3, 16, 16), generator=g).to(cuda)                # NCHW fp32, 7x7
    w7 = torch.randn((8, 3, 7, 7), generator=g).to(cuda)
    cols = ops.im2col(img, nchw_f32=True, n=1, h=16, w=16, c=3, k=7, stride=1, pad=3, kpad=192)
    wp = torch.zeros((8, 192), device=cuda)
    wp[:, :147] = w7.permute(0, 2, 3, 1).reshape(8, -1)
    ref = F.conv2d(_bf(img).float(), w7, padding=3).permute(0, 2, 3, 1).reshape(-1, 8)
    assert _rel(cols.float() @ wp.T, ref) < 1e-5


def test_upsample_and_pixel_shuffle(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = _bf(torch.randn((2, 5, 7, 16), generator=g)).to(cuda)
    up = ops.upsample2x(x)
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2, mode="bilinear",
                        align_corners=True).permute(0, 2, 3, 1)
    assert _rel(up, ref) < 4e-3
    # ConvTranspose2d(k == stride) == GEMM to (k*k*c) columns + pixel shuffle
    c, k = 16, 4
    wt = torch.randn((c, c, k, k), generator=g).to(cuda)                   # [in, out, kh, kw]
    cols = x.float().reshape(-1, c) @ wt.permute(0, 2, 3, 1).reshape(c, k * k * c)
    shuf = ops.pixel_shuffle(_bf(cols), 2, 5, 7, c, k)
    ref = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wt, stride=k).permute(0, 2, 3, 1)
    assert _rel(shuf, ref) < 6e-3


def test_tokens_silu_camera_head(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(10)
    B, T, C = 2, 4, 768
    feat = torch.randn((B * T, C), generator=g).to(cuda)
    sd = {"camera_extrinsic_head.1.weight": (0.05 * torch.randn((8, C), generator=g)).to(cuda),
          "camera_extrinsic_head.1.bias": (0.05 * torch.randn((8,), generator=g)).to(cuda)}
    pred, c2w = ops.camera_head(feat, C, sd["camera_extrinsic_head.1.weight"],
                                sd["camera_extrinsic_head.1.bias"], B, T, C)
    rp, rm = er.camera_head(sd, feat.view(B, T, C))
    assert (pred - rp).abs().max() < 1e-5
    assert (c2w - rm).abs().max() < 1e-5
    y = ops.silu_bf16(feat, B * T, C)
    assert _rel(y, F.silu(feat)) < 4e-3
    x = torch.zeros((B * T * 5, C), device=cuda)
    it, et = torch.randn((C,), generator=g).to(cuda), torch.randn((C,), generator=g).to(cuda)
    ops.camera_tokens(it, et, x, B * T, T, C, 5)
    xr = x.view(B, T, 5, C)
    assert torch.equal(xr[:, 0, 0], it.expand(B, C)) and torch.allclose(xr[:, 1:, 0], (it + et).expand(B, T - 1, C))
    assert (xr[:, :, 1:] == 0).all()
    K9 = torch.randn((B * T, 9), generator=g).to(cuda)
    w, b = torch.randn((1024, 9), generator=g).to(cuda), torch.randn((1024,), generator=g).to(cuda)
    xe = torch.zeros((B * T * 17, 1024), device=cuda)
    ops.intrinsic_token(K9, w, b, xe, B * T, 1024, 17, 16)
    assert (xe.view(B * T, 17, 1024)[:, 16] - (K9 @ w.T + b)).abs().max() < 1e-5


def test_pts_tail_and_gaussian_adapter(cuda, lib):
    from vicasplat_b200 import ops
    g = torch.Generator().manual_seed(14)
    cfg = er.EncoderConfig()
    px, Cf = 1000, 128
    feat = _bf(torch.randn((px, Cf), generator=g).relu()).to(cuda)
    w, b = (0.1 * torch.randn((3, Cf), generator=g)).to(cuda), torch.randn((3,), generator=g).to(cuda)
    raw = torch.randn((px, 86), generator=g).to(cuda)
    raw[:, 4:7] *= 8                                                       # exercise softplus/clamp range
    ops.pts_tail(feat, Cf, w, b, raw, px)
    xyz = feat.float() @ w.T + b
    d = xyz.norm(dim=-1, keepdim=True)
    assert (raw[:, :3] - xyz / d.clip(min=1e-8) * torch.expm1(d)).abs().max() < 1e-4 * torch.expm1(d).max()
    out = ops.gaussian_adapter(raw, cfg.d_sh, er.sh_mask(cfg, cuda))
    ref = er.gaussian_adapter(raw, cfg)
    assert torch.equal(out["means"], raw[:, :3])
    assert (out["opac"] - ref["opacities"][:, 0]).abs().max() < 1e-6
    assert (out["scales"] - ref["scales"]).abs().max() < 1e-7
    assert (out["rot"] - ref["rotations"]).abs().max() < 1e-6
    assert (out["sh"] - ref["harmonics"]).abs().max() < 1e-6
    assert (out["cov"] - ref["covariances"]).abs().max() < 1e-9 + 1e-5 * ref["covariances"].abs().max()
    iu = torch.triu_indices(3, 3)
    assert torch.equal(out["cov6"], out["cov"][:, iu[0], iu[1]])
    # padded head-output layout (parameters at column 0, xyz at 84) + reference-layout raw output
    gsp = torch.full((px, 96), float("nan"), device=cuda)
    gsp[:, :83], gsp[:, 84:87] = raw[:, 3:], raw[:, :3]
    raw_out = torch.zeros_like(raw)
    out2 = ops.gaussian_adapter(gsp, cfg.d_sh, er.sh_mask(cfg, cuda), center_col=84, param_col=0,
                                raw_out=raw_out)
    assert torch.equal(raw_out, raw)
    for k in out:
        assert torch.equal(out[k], out2[k]), k
