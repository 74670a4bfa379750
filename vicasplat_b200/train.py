"""Training step of the WHOLE encoder plugin: ``VicaSplat.forward`` kept
differentiable by a hand-written backward pass, so that the reference's training step
(src/model/model_wrapper.py:184-321: encoder -> rasterizer -> losses -> backward -> AdamW) runs on the
sm_100a kernels end to end.  The reference gets every gradient below from ``torch.autograd`` over
``backbone_vica.py:280-335`` (MixDecoderBlock), ``heads/dpt_block.py:79-229,264-459``,
``heads/dpt_gs_head.py:98-157``, ``common/gaussian_adapter.py:167-212`` and ``vicasplat.py:158-278``.

``TrainEngine`` owns
  * gradient BUCKETS (one flat fp32 buffer per bucket of ``encoder_train.plan_buckets``: heads, decoder
    blocks, encoder blocks, stems -- in the order their gradients become final); ``param.grad`` of the
    module's fp32 master parameters are views into them; the parameters the reference never gives a
    gradient (refinenet4.resConfUnit1) stay without one;
  * bf16 operand copies of every weight in the layouts of its forward, dgrad and wgrad GEMMs;
  * ``forward(image, intrinsics)`` (keeps activations) and ``backward(grads of the outputs)``, which
    hands each bucket to the ``GradReducer`` (asynchronous NCCL all-reduce) as soon as it is final.

Numerics follow the forward path: bf16 GEMM operands, fp32 accumulation, fp32 residual streams and
fp32 parameter gradients.  Per operator:
    linear   dgrad dX = dY W (vs_gemm), wgrad dW += dY^T X (a_mode 2, split-K, atomic), bias = column sums
    conv     dgrad = the same implicit-GEMM conv with flipped taps, ReLU mask (+ skip) in its epilogue;
             wgrad = a_mode 3 (pixels as K, taps as shifted TMA boxes of the input map)
    LN+AdaLN vs_layernorm_mod_backward + vs_adaln_reduce;  gate: vs_gate_backward
    attention vs_attention_backward (video: blocked-causal camera rows; neighbour: key-centric dK/dV)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import encoder_grad as eg, ops
from ._lib import VS_ACT_NONE, VS_ACT_RELU
from .encoder import FEAT, LAYER_DIMS, VicaSplat, _pack_conv
from .encoder_train import GradReducer, plan_buckets

BF16, F32 = torch.bfloat16, torch.float32


class Bucket:
    """One flat fp32 gradient buffer; 1-D tensors first.  ``views[name]`` are the gradients."""

    def __init__(self, name: str, shapes: Dict[str, torch.Size], device):
        self.name = name
        small = [n for n, s in shapes.items() if len(s) == 1]
        big = [n for n, s in shapes.items() if len(s) != 1]
        self.names = small + big
        off, offs = 0, {}
        for n in self.names:
            offs[n] = off
            off += (shapes[n].numel() + 3) // 4 * 4
        self.flat = torch.zeros((off,), dtype=F32, device=device)
        self.views = {n: self.flat[offs[n]: offs[n] + shapes[n].numel()].view(shapes[n]) for n in self.names}
        self.offs = offs


def _flipped(w: torch.Tensor) -> torch.Tensor:
    """dgrad weights of a stride-1 conv: [O, C, kh, kw] -> packed [C, kh*kw*O_pad], taps flipped."""
    return _pack_conv(w.detach().flip(2, 3).transpose(0, 1))


class TrainEngine:
    def __init__(self, model: VicaSplat, reducer: Optional[GradReducer] = None, attach_grads: bool = True):
        """attach_grads: make every ``param.grad`` a view of its bucket (the fast path: backward() writes
        the gradients in place and the reducer / FusedAdamW work on the buckets).  False: the buckets are
        private scratch and the caller (the autograd.Function of VicaSplat.forward) hands them on."""
        self.m = model
        self.bb = model._bb
        self.dev = next(model.parameters()).device
        assert self.dev.type == "cuda", "move the module to CUDA first (there is no CPU fallback)"
        self.reducer = reducer or GradReducer()
        self.p: Dict[str, torch.nn.Parameter] = dict(model.named_parameters())
        plan, self.unused = plan_buckets(list(self.p.keys()), self.bb["enc_depth"], self.bb["dec_depth"])
        self.buckets: Dict[str, Bucket] = {}
        self.g: Dict[str, torch.Tensor] = {}
        for bname, members in plan:
            b = Bucket(bname, {n: self.p[n].shape for n in members}, self.dev)
            self.buckets[bname] = b
            for n, v in b.views.items():
                if attach_grads:
                    self.p[n].grad = v
                self.g[n] = v
        self.bucket_order = [bname for bname, _ in plan]
        D = self.bb["dec_embed_dim"]
        # the three projections of CrossNeighborAttention run as ONE GEMM: their gradients are adjacent in
        # the bucket, so the (3D, D) / (3D) views below alias them
        for i in range(self.bb["dec_depth"]):
            k = f"backbone.dec_blocks.{i}.cross_attn."
            gw, gb = self.g[k + "projq.weight"], self.g[k + "projq.bias"]
            fw = torch.as_strided(gw, (3 * D, D), (D, 1))
            fb = torch.as_strided(gb, (3 * D,), (1,))
            assert fw[D].data_ptr() == self.g[k + "projk.weight"].data_ptr() and \
                fw[2 * D].data_ptr() == self.g[k + "projv.weight"].data_ptr() and \
                fb[D:].data_ptr() == self.g[k + "projk.bias"].data_ptr() and \
                fb[2 * D:].data_ptr() == self.g[k + "projv.bias"].data_ptr()
            self.g[k + "qkv.weight"], self.g[k + "qkv.bias"] = fw, fb
        # nn.Dropout(0.1) between the gs head's ReLU and its 1x1 projection is active in training mode
        # (dpt_block.py:341); `model.gs_head_dropout` overrides it (0 for the parity tests: the reference's
        # goldens are taken in eval mode)
        self.dropout_p = float(getattr(model, "gs_head_dropout", 0.1))
        self._drop_seed = 0x5eed
        self.w: Dict[str, torch.Tensor] = {}
        self.dwp: Dict[str, torch.Tensor] = {}      # packed-layout conv weight gradients (scratch)
        self._plans: Dict[tuple, dict] = {}
        self._saved: Optional[dict] = None
        self.repack()

    # ------------------------------------------------------------------ parameters
    def trainable_parameters(self) -> List[torch.nn.Parameter]:
        """Everything that receives a gradient (the reference's optimizer skips grad-less tensors)."""
        return [p for n, p in self.p.items() if n not in set(self.unused)]

    def zero_grad(self) -> None:
        for b in self.buckets.values():
            b.flat.zero_()

    @torch.no_grad()
    def repack(self) -> None:
        """bf16 operand copies of every weight: W (forward / wgrad layout) and W^T (dgrad) of the linear
        layers in one vs_grad_prep pass each; packed (forward) and flipped-packed (dgrad) conv weights."""
        p, w, bb = self.p, self.w, self.bb
        f = lambda k: p[k].detach()
        self.enc_w: List[dict] = []
        for i in range(bb["enc_depth"]):
            k, wi = f"backbone.enc_blocks.{i}.", {}
            for n in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                wi[n], wi[n + ".t"] = ops.grad_prep(f(k + n + ".weight"))
                wi[n + ".bias"] = f(k + n + ".bias")
            for n in ("norm1", "norm2"):
                wi[n + ".weight"], wi[n + ".bias"] = f(k + n + ".weight"), f(k + n + ".bias")
            self.enc_w.append(wi)
        w["patch"], _ = ops.grad_prep(f("backbone.patch_embed.proj.weight").flatten(1), want_t=False)
        w["decoder_embed"], w["decoder_embed.t"] = ops.grad_prep(f("backbone.decoder_embed.weight"))
        for i in range(bb["dec_depth"]):
            k = f"backbone.dec_blocks.{i}."
            for n in ("modulation1.proj", "modulation2.proj", "attn.qkv", "attn.proj", "cross_attn.proj",
                      "mlp.fc1", "mlp.fc2", "mlp_cam.fc1", "mlp_cam.fc2"):
                w[k + n], w[k + n + ".t"] = ops.grad_prep(f(k + n + ".weight"))
            cq = torch.cat([f(f"{k}cross_attn.{n}.weight") for n in ("projq", "projk", "projv")], 0)
            w[k + "cross_qkv"], w[k + "cross_qkv.t"] = ops.grad_prep(cq)
            w[k + "cross_qkv.bias"] = torch.cat([f(f"{k}cross_attn.{n}.bias") for n in ("projq", "projk", "projv")], 0)
        for head in ("downstream_head1", "gaussian_param_head"):
            k = head + ".dpt."
            for idx in range(4):
                w[f"{k}ap{idx}.0"], w[f"{k}ap{idx}.0.t"] = ops.grad_prep(f(f"{k}act_postprocess.{idx}.0.weight").flatten(1))
                wr = f(f"{k}scratch.layer_rn.{idx}.weight")
                w[f"{k}rn{idx}"], w[f"{k}rn{idx}.f"] = _pack_conv(wr), _flipped(wr)
            for idx, kk in ((0, 4), (1, 2)):
                wt = f(f"{k}act_postprocess.{idx}.1.weight")                       # [in, out, k, k]
                wp = wt.permute(2, 3, 1, 0).reshape(kk * kk * wt.shape[1], wt.shape[0]).contiguous()
                w[f"{k}ap{idx}.1"], w[f"{k}ap{idx}.1.t"] = ops.grad_prep(wp)
                w[f"{k}ap{idx}.1.bias"] = f(f"{k}act_postprocess.{idx}.1.bias").repeat(kk * kk).contiguous()
            w3 = f(f"{k}act_postprocess.3.1.weight")
            w[f"{k}ap3.1"], w[f"{k}ap3.1.t"] = ops.grad_prep(w3.permute(0, 2, 3, 1).reshape(w3.shape[0], -1).contiguous())
            for r in (1, 2, 3, 4):
                rk = f"{k}scratch.refinenet{r}."
                w[rk + "out"], w[rk + "out.t"] = ops.grad_prep(f(rk + "out_conv.weight").flatten(1))
                for u in ("resConfUnit1", "resConfUnit2"):
                    if r == 4 and u == "resConfUnit1":
                        continue
                    for c in ("conv1", "conv2"):
                        wc = f(f"{rk}{u}.{c}.weight")
                        w[f"{rk}{u}.{c}"], w[f"{rk}{u}.{c}.f"] = _pack_conv(wc), _flipped(wc)
            wh = f(k + "head.0.weight")
            w[k + "head.0"], w[k + "head.0.f"] = _pack_conv(wh), _flipped(wh)
            if head == "downstream_head1":
                wh2 = f(k + "head.2.weight")
                w[k + "head.2"], w[k + "head.2.f"] = _pack_conv(wh2), _flipped(wh2)
            else:
                w4 = torch.zeros((96, FEAT), dtype=F32, device=self.dev)       # 83 rows used: 16-byte rows of gsp
                w4[: self.m.raw_gs_dim] = f(k + "head.4.weight").flatten(1)
                w[k + "head.4"], w[k + "head.4.t"] = ops.grad_prep(w4)
                b4 = torch.zeros((96,), dtype=F32, device=self.dev)
                b4[: self.m.raw_gs_dim] = f(k + "head.4.bias")
                w[k + "head.4.bias"] = b4
                w7 = f(k + "input_merger.0.weight")                             # [256,3,7,7]
                pk = torch.zeros((w7.shape[0], 7, 8, 8), dtype=F32, device=self.dev)
                pk[:, :, :7, :3] = w7.permute(0, 2, 3, 1)
                w[k + "merger"] = pk.reshape(w7.shape[0], -1).to(BF16).contiguous()

    # ------------------------------------------------------------------ per-shape tables
    def _plan(self, B, T, H, W) -> dict:
        key = (B, T, H, W)
        if key in self._plans:
            return self._plans[key]
        bb, dev = self.bb, self.dev
        P = bb["patch_size"]
        gh, gw = H // P, W // P
        assert H % P == 0 and W % P == 0 and gh % 2 == 0 and gw % 2 == 0
        Fr, Np = B * T, gh * gw
        N, rpf = Np + 1, Np + 2
        i32 = dict(dtype=torch.int32, device=dev)
        pl = dict(B=B, T=T, H=H, W=W, gh=gh, gw=gw, Fr=Fr, Np=Np, N=N, rpf=rpf)
        pl["lay"] = eg.FrameLayout.make(Fr, gh, gw, bb["enc_num_heads"], dev)
        ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
        pos = torch.cat([torch.stack([ys, xs], -1).reshape(Np, 2), torch.tensor([[gh, 0]])], 0)
        pos_d = torch.zeros((B, T, rpf, 2), dtype=torch.int32)
        pos_d[:, :, 1:] = pos.to(torch.int32)
        pos_d[:, :, 0, 0] = -1 - torch.arange(T, dtype=torch.int32)[None]
        pl["pos_dec"] = pos_d.reshape(Fr * rpf, 2).to(dev).contiguous()
        sc = torch.arange(B, **i32)
        pl["vid_start"], pl["vid_len"] = sc * (T * rpf), torch.full((B,), T * rpf, **i32)
        t_idx = torch.arange(T)
        prev = torch.where(t_idx > 0, t_idx - 1, t_idx + 1) if T > 1 else t_idx
        nxt = torch.where(t_idx < T - 1, t_idx + 1, t_idx - 1) if T > 1 else t_idx
        base = (torch.arange(B)[:, None] * T)
        pl["nb_q"] = ((base + t_idx[None]) * rpf + 1).reshape(-1).to(**i32)
        pl["nb_k0"] = ((base + prev[None]) * rpf + 1).reshape(-1).to(**i32)
        pl["nb_k1"] = ((base + nxt[None]) * rpf + 1).reshape(-1).to(**i32)
        pl["nb_len"] = torch.full((Fr,), N, **i32)
        pl["nb_len1"] = ((prev != nxt)[None].expand(B, T).reshape(-1) * N).to(**i32)
        tb = eg.neighbour_backward_tables(B, T, rpf, N, first_img_row=1)
        pl["nb_dkv"] = {k: v.to(dev) for k, v in tb.items()}
        pl["nb_dkv"]["max_kv_len"] = N
        d = bb["dec_depth"]
        pl["hooks"] = [0, d * 2 // 4, d * 3 // 4, d]
        self._plans[key] = pl
        return pl

    # ------------------------------------------------------------------ forward (keeps activations)
    @torch.no_grad()
    def forward(self, image: torch.Tensor, intrinsics: torch.Tensor) -> dict:
        """image (B,T,3,H,W) fp32 in [-1,1], intrinsics (B,T,3,3) -> dict(raw (G,86), means, cov, cov6, sh,
        opac, scales, rot, pred_extrins (B,T-1,8), c2w (B,T,4,4)) with G = B*T*H*W rows; everything the
        backward pass needs is kept until ``backward``."""
        B, T, _, H, W = image.shape
        pl = self._plan(B, T, H, W)
        s: dict = dict(pl=pl)
        Fr = pl["Fr"]
        img = image.reshape(Fr, 3, H, W).to(F32).contiguous()
        K9 = intrinsics.reshape(Fr, 9).to(F32).contiguous()
        inter0 = self._encoder_fwd(pl, s, img, K9)
        hooks = self._decoder_fwd(pl, s, inter0)
        out = self._heads_fwd(pl, s, img, inter0, hooks)
        self._saved = s
        return out

    def _encoder_fwd(self, pl, s, img, K9):
        p, bb = self.p, self.bb
        E, lay = bb["enc_embed_dim"], pl["lay"]
        Fr, Np, N = pl["Fr"], pl["Np"], pl["N"]
        cols = ops.patchify(img, bb["patch_size"])
        x = torch.empty((Fr * N, E), dtype=F32, device=self.dev)
        ops.gemm(cols, self.w["patch"], bias=p["backbone.patch_embed.proj.bias"].detach(), out=x,
                 out_gin=Np, out_gout=N, out_off=0)
        ops.intrinsic_token(K9, p["backbone.intrinsic_encoder.weight"].detach(),
                            p["backbone.intrinsic_encoder.bias"].detach(), x, Fr, E, N, Np)
        blocks = []
        for i in range(bb["enc_depth"]):
            sv = eg.Saved()
            x = eg.block_forward(x, self.enc_w[i], lay, sv)
            blocks.append(sv)
        inter0, _ = ops.layernorm(x, p["backbone.enc_norm.weight"].detach(), p["backbone.enc_norm.bias"].detach())
        s.update(enc_cols=cols, K9=K9, enc_blocks=blocks, enc_x_last=x, inter0=inter0)
        return inter0

    def _decoder_fwd(self, pl, s, inter0):
        p, w, bb = self.p, self.w, self.bb
        D, Hh = bb["dec_embed_dim"], bb["dec_num_heads"]
        theta = float(bb["temporal_rope_theta"])
        Fr, T, N, rpf = pl["Fr"], pl["T"], pl["N"], pl["rpf"]
        f = lambda k: p[k].detach()
        M = Fr * rpf
        x = torch.empty((M, D), dtype=F32, device=self.dev)
        ops.gemm(inter0, w["decoder_embed"], bias=f("backbone.decoder_embed.bias"), out=x,
                 out_gin=N, out_gout=rpf, out_off=1)
        ops.camera_tokens(f("backbone.camera_intrinsic_token"), f("backbone.camera_extrinsic_token"), x, Fr, T, D, rpf)
        cam = lambda t: t.view(Fr, rpf * D)[:, :D]
        blocks, hooks = [], {}
        for i in range(bb["dec_depth"]):
            k = f"backbone.dec_blocks.{i}."
            c: dict = dict(x0=x)
            # --- video + camera self attention
            _, c["cn1"] = ops.layernorm(cam(x), f(k + "cam_norm1.weight"), f(k + "cam_norm1.bias"),
                                        want_bf16=False, want_f32=True)
            c["sil1"] = ops.silu_bf16(c["cn1"], Fr, D)
            m1 = c["mod1"] = ops.gemm(c["sil1"], w[k + "modulation1.proj"], bias=f(k + "modulation1.proj.bias"),
                                      out_dtype=F32)
            c["h1"], _ = ops.layernorm(x, f(k + "norm1.weight"), f(k + "norm1.bias"), w0=f(k + "cam_norm1.weight"),
                                       b0=f(k + "cam_norm1.bias"), scale=m1[:, :D], shift=m1[:, D:2 * D],
                                       rows_per_frame=rpf)
            qkv = c["qkv1"] = ops.gemm(c["h1"], w[k + "attn.qkv"], bias=f(k + "attn.qkv.bias"),
                                       rope=(pl["pos_dec"], 0, D, Hh, 100.0, theta))
            c["o1"] = torch.empty((M, D), dtype=BF16, device=self.dev)
            c["lse1"] = torch.empty((M, Hh), dtype=F32, device=self.dev)
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], c["o1"], heads=Hh, q_start=pl["vid_start"],
                          q_len=pl["vid_len"], kv_start0=pl["vid_start"], kv_len0=pl["vid_len"],
                          max_q_len=T * rpf, max_kv_len=T * rpf, causal_block=rpf, scale=0.125, lse=c["lse1"])
            c["br1"] = ops.gemm(c["o1"], w[k + "attn.proj"], bias=f(k + "attn.proj.bias"))
            x = c["x1"] = ops.gate_residual(x, c["br1"], gate=m1[:, 2 * D:], rows_per_frame=rpf, first_row_mode=1,
                                            out=torch.empty_like(x))
            # --- neighbour cross attention (image rows only)
            c["cn2b"], c["cn2"] = ops.layernorm(cam(x), f(k + "cam_norm2.weight"), f(k + "cam_norm2.bias"),
                                                want_bf16=True, want_f32=True)
            c["sil2"] = ops.silu_bf16(c["cn2"], Fr, D)
            m2 = c["mod2"] = ops.gemm(c["sil2"], w[k + "modulation2.proj"], bias=f(k + "modulation2.proj.bias"),
                                      out_dtype=F32)
            c["h2"], _ = ops.layernorm(x, f(k + "norm2.weight"), f(k + "norm2.bias"), scale=m2[:, :D],
                                       shift=m2[:, D:2 * D], rows_per_frame=rpf)
            qkv = c["qkv2"] = ops.gemm(c["h2"], w[k + "cross_qkv"], bias=w[k + "cross_qkv.bias"],
                                       rope=(pl["pos_dec"], 0, D, Hh, 100.0, theta))
            c["o2"] = torch.zeros((M, D), dtype=BF16, device=self.dev)     # camera rows are not attention items
            c["lse2"] = torch.zeros((M, Hh), dtype=F32, device=self.dev)
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], c["o2"], heads=Hh, q_start=pl["nb_q"],
                          q_len=pl["nb_len"], kv_start0=pl["nb_k0"], kv_len0=pl["nb_len"], kv_start1=pl["nb_k1"],
                          kv_len1=pl["nb_len1"], max_q_len=N, max_kv_len=2 * N, scale=0.125, lse=c["lse2"])
            c["br2"] = ops.gemm(c["o2"], w[k + "cross_attn.proj"], bias=f(k + "cross_attn.proj.bias"))
            x = c["x2"] = ops.gate_residual(x, c["br2"], gate=m2[:, 2 * D:3 * D], rows_per_frame=rpf,
                                            first_row_mode=2, out=torch.empty_like(x))
            # --- MLPs
            c["h3"], _ = ops.layernorm(x, f(k + "norm3.weight"), f(k + "norm3.bias"), scale=m2[:, 3 * D:4 * D],
                                       shift=m2[:, 4 * D:5 * D], rows_per_frame=rpf)
            c["z3"] = ops.gemm(c["h3"], w[k + "mlp.fc1"], bias=f(k + "mlp.fc1.bias"))
            c["a3"] = ops.gelu_bf16(c["z3"])
            c["br3"] = ops.gemm(c["a3"], w[k + "mlp.fc2"], bias=f(k + "mlp.fc2.bias"))
            x = ops.gate_residual(x, c["br3"], gate=m2[:, 5 * D:], rows_per_frame=rpf, first_row_mode=2,
                                  out=torch.empty_like(x))
            c["zc"] = ops.gemm(c["cn2b"], w[k + "mlp_cam.fc1"], bias=f(k + "mlp_cam.fc1.bias"))
            c["ac"] = ops.gelu_bf16(c["zc"])
            ops.gemm(c["ac"], w[k + "mlp_cam.fc2"], bias=f(k + "mlp_cam.fc2.bias"), res1=x, out=x,
                     out_gin=1, out_gout=rpf, out_off=0)
            blocks.append(c)
            layer = i + 1
            if layer in pl["hooks"][1:]:
                if layer == bb["dec_depth"]:
                    hooks[layer], _ = ops.layernorm(x, f("backbone.dec_norm.weight"), f("backbone.dec_norm.bias"))
                else:
                    hooks[layer], _ = ops.layernorm(x, normalize=False)
        _, cam_out = ops.layernorm(cam(x), f("backbone.camera_dec_norm.weight"), f("backbone.camera_dec_norm.bias"),
                                   want_bf16=False, want_f32=True)
        pred, c2w = ops.camera_head(cam_out, D, f("camera_extrinsic_head.1.weight"),
                                    f("camera_extrinsic_head.1.bias"), pl["B"], T, D)
        s.update(dec_blocks=blocks, dec_x_last=x, cam_out=cam_out, hooks=hooks, pred=pred, c2w=c2w)
        return hooks

    # ---- DPT pieces (NHWC bf16 maps)
    def _conv(self, x, key, *, k=3, N=FEAT, bias=None, act=VS_ACT_NONE, res1=None, res2=None, relu_copy=False):
        out2 = torch.empty(x.shape[:3] + (N,), dtype=BF16, device=x.device) if relu_copy else None
        out = ops.conv_gemm(x, self.w[key], kh=k, kw=k, pad=k // 2, N=N, bias=bias, act=act, res1=res1,
                            res2=res2, out2=out2)
        return (out, out2) if relu_copy else out

    def _trunk_fwd(self, pl, head, inter0, hooks, t: dict):
        p, w, bb = self.p, self.w, self.bb
        f = lambda k: p[k].detach()
        k = head + ".dpt."
        Fr, Np, N, rpf, gh, gw = pl["Fr"], pl["Np"], pl["N"], pl["rpf"], pl["gh"], pl["gw"]
        E, D = bb["enc_embed_dim"], bb["dec_embed_dim"]
        layers = []
        for idx, hook in enumerate(pl["hooks"]):
            if idx == 0:
                A, Cc, gs = inter0, E, N * E
            else:
                A, Cc, gs = hooks[hook][1:], D, rpf * D
            t0 = ops.gemm(A, w[f"{k}ap{idx}.0"], K=Cc, a_rows=Np, a_groups=Fr, a_row_stride=Cc,
                          a_group_stride=gs, bias=f(f"{k}act_postprocess.{idx}.0.bias"))
            c = LAYER_DIMS[idx]
            t[f"t0_{idx}"] = t0
            if idx in (0, 1):
                kk = 4 if idx == 0 else 2
                t1 = ops.gemm(t0, w[f"{k}ap{idx}.1"], bias=w[f"{k}ap{idx}.1.bias"])
                m = ops.pixel_shuffle(t1, Fr, gh, gw, c, kk)
            elif idx == 2:
                m = t0.view(Fr, gh, gw, c)
            else:
                cols = ops.im2col(t0, nchw_f32=False, n=Fr, h=gh, w=gw, c=c, k=3, stride=2, pad=1, kpad=9 * c)
                t["cols3"] = cols
                m = ops.gemm(cols, w[f"{k}ap3.1"], bias=f(f"{k}act_postprocess.3.1.bias")).view(Fr, gh // 2, gw // 2, c)
            t[f"m{idx}"] = m
            layers.append(self._conv(m, f"{k}rn{idx}", relu_copy=True))
        t["layers"] = layers
        sk = k + "scratch.refinenet"
        path = self._fusion_fwd(sk + "4.", layers[3][0], layers[3][1], t, 4)
        for r in (3, 2, 1):
            l, lr = layers[r - 1]
            rk = f"{sk}{r}.resConfUnit1."
            u1 = self._conv(lr, rk + "conv1", bias=f(rk + "conv1.bias"), act=VS_ACT_RELU)
            x, xr = self._conv(u1, rk + "conv2", bias=f(rk + "conv2.bias"), res1=l, res2=path, relu_copy=True)
            t[f"u1_{r}"], t[f"xr_{r}"] = u1, xr
            path = self._fusion_fwd(f"{sk}{r}.", x, xr, t, r)
        return path

    def _fusion_fwd(self, rk, x, xr, t, r):
        f = lambda k: self.p[k].detach()
        u = rk + "resConfUnit2."
        y1 = self._conv(xr, u + "conv1", bias=f(u + "conv1.bias"), act=VS_ACT_RELU)
        y = self._conv(y1, u + "conv2", bias=f(u + "conv2.bias"), res1=x)
        n, h, w_, c = y.shape
        oc = ops.gemm(y.view(-1, c), self.w[rk + "out"], bias=f(rk + "out_conv.bias"))
        t[f"y1_{r}"], t[f"y_{r}"], t[f"fx_{r}"] = y1, y, xr
        return ops.upsample2x(oc.view(n, h, w_, FEAT))

    def _heads_fwd(self, pl, s, img, inter0, hooks):
        p, w = self.p, self.w
        f = lambda k: p[k].detach()
        Fr, H, W = pl["Fr"], pl["H"], pl["W"]
        G = Fr * H * W
        gsp = torch.zeros((G, 96), dtype=F32, device=self.dev)
        # --- Gaussian centres
        k = "downstream_head1.dpt."
        tp: dict = {}
        p1 = self._trunk_fwd(pl, "downstream_head1", inter0, hooks, tp)
        y0 = self._conv(p1, k + "head.0", N=FEAT // 2, bias=f(k + "head.0.bias"))
        u = ops.upsample2x(y0)
        y2 = self._conv(u, k + "head.2", N=FEAT // 2, bias=f(k + "head.2.bias"), act=VS_ACT_RELU)
        w4 = f(k + "head.4.weight").flatten(1).contiguous()
        ops.pts_tail(y2, FEAT // 2, w4, f(k + "head.4.bias"), gsp[:, 84:], G)
        tp.update(p1=p1, u=u, y2=y2)
        # --- Gaussian parameters
        k = "gaussian_param_head.dpt."
        tg: dict = {}
        g1 = self._trunk_fwd(pl, "gaussian_param_head", inter0, hooks, tg)
        img8 = ops.image_nhwc8(img, pad=3)
        view = (Fr, H, W, 64, H + 6, 8, (W + 8) * 8, (H + 6) * (W + 8) * 8)
        c7 = ops.conv_gemm(img8, w[k + "merger"], kh=7, kw=1, pad=0, N=FEAT, bias=f(k + "input_merger.0.bias"),
                           act=VS_ACT_RELU, view=view)
        merged = ops.upsample2x(g1, add=c7)
        y = self._conv(merged, k + "head.0", act=VS_ACT_RELU)
        if self.dropout_p > 0:
            self._drop_seed += 1
            ops.dropout_(y, self.dropout_p, self._drop_seed)      # y now holds the kept, re-scaled activations
        ops.gemm(y.view(-1, FEAT), w[k + "head.4"], bias=w[k + "head.4.bias"], out=gsp, N=self.m.raw_gs_dim)
        tg.update(img8=img8, view=view, c7=c7, merged=merged, y=y)
        raw = torch.empty((G, 3 + self.m.raw_gs_dim), dtype=F32, device=self.dev)
        gq = ops.gaussian_adapter(gsp, self.m.d_sh, self.m.sh_mask.to(self.dev), center_col=84, param_col=0,
                                  raw_out=raw)
        s.update(tp=tp, tg=tg, gsp=gsp)
        return dict(raw=raw, pred_extrins=s["pred"], c2w=s["c2w"], **gq)

    def relu_masks(self) -> Dict[str, torch.Tensor]:
        """The ReLU masks of the kept forward pass, keyed by the reference's layer names (bool, NCHW for the
        DPT maps; (B, T-1, C) for the pose head).  A checker that pins these masks in its own forward pass
        compares gradients at the SAME linearisation point (tests/test_gpu_model_grad.py)."""
        s = self._saved
        assert s is not None, "relu_masks() needs a forward() first"
        pl = s["pl"]
        nchw = lambda t: (t > 0).permute(0, 3, 1, 2)
        out = {}
        for head, t in (("downstream_head1", s["tp"]), ("gaussian_param_head", s["tg"])):
            k = f"{head}.dpt.scratch.refinenet"
            for r in (1, 2, 3, 4):
                if r < 4:
                    out[f"{k}{r}.resConfUnit1.conv1"] = nchw(t["layers"][r - 1][1])
                    out[f"{k}{r}.resConfUnit1.conv2"] = nchw(t[f"u1_{r}"])
                out[f"{k}{r}.resConfUnit2.conv1"] = nchw(t[f"fx_{r}"])
                out[f"{k}{r}.resConfUnit2.conv2"] = nchw(t[f"y1_{r}"])
        out["downstream_head1.dpt.head.2"] = nchw(s["tp"]["y2"])
        out["gaussian_param_head.dpt.input_merger.0"] = nchw(s["tg"]["c7"])
        out["gaussian_param_head.dpt.head.0"] = nchw(s["tg"]["y"])
        out["camera_extrinsic_head.1"] = (s["cam_out"] > 0).view(pl["B"], pl["T"], -1)[:, 1:]
        return out

    # ------------------------------------------------------------------ backward
    @torch.no_grad()
    def backward(self, *, d_raw=None, d_means=None, d_cov=None, d_cov6=None, d_sh=None, d_opac=None,
                 d_pred=None, zero: bool = True) -> None:
        """Gradients of the loss w.r.t. forward()'s outputs (any may be None; shapes as returned, flattened
        over (B,T,H,W)) -> every ``param.grad``.  zero=False accumulates on top of the existing gradients
        (micro-batches).  Buckets go to the reducer in the order they become final."""
        s = self._saved
        assert s is not None, "backward() needs a forward() first"
        pl = s["pl"]
        if zero:
            self.zero_grad()
        for v in self.dwp.values():
            v.zero_()
        d_hooks = self._heads_bwd(pl, s, d_raw, d_means, d_cov, d_cov6, d_sh, d_opac)
        for b in ("gs_head", "pts_head"):
            self.reducer.bucket_ready(self.buckets[b])
        d_inter0 = self._decoder_bwd(pl, s, d_hooks, d_pred)
        self._encoder_bwd(pl, s, d_inter0)
        self.reducer.finish()
        self._saved = None

    # ---- helpers
    def _colsum(self, d2d, name):
        ops.grad_prep(d2d, want_copy=False, want_t=False, colsum=self.g[name])

    def _dwp(self, key, cout):
        if key not in self.dwp:
            self.dwp[key] = torch.zeros((cout, self.w[key].shape[1]), dtype=F32, device=self.dev)
        return self.dwp[key]

    def _conv_bwd(self, dy, x_in, key, *, k=3, bias=None, mask=None, res1=None, need_dx=True):
        """dy (n,h,w,Cout) bf16: gradient of a stride-1 conv's output; x_in: its input map.  Weight gradient
        into the packed scratch, bias gradient into its bucket view; returns the input gradient with the
        ReLU mask (`mask` = the kept post-ReLU input of the conv's producer) and the skip (`res1`) fused."""
        cout = dy.shape[-1]
        if bias is not None:
            self._colsum(dy.view(-1, cout), bias)
        ops.conv_wgrad(dy, x_in, self._dwp(key, cout), kh=k, kw=k, pad=k // 2)
        if not need_dx:
            return None
        cin = x_in.shape[-1]
        if mask is not None:
            return ops.conv_gemm_masked(dy, self.w[key + ".f"], kh=k, kw=k, pad=k // 2, N=cin, mask=mask, res1=res1)
        return ops.conv_gemm(dy, self.w[key + ".f"], kh=k, kw=k, pad=k // 2, N=cin, res1=res1)

    def _lin_bwd(self, dy, x, key, gname, *, bias=None, need_dx=True, dx_dtype=BF16, mask=None, **kw):
        """dy (M, N) bf16, x (M, K) bf16: wgrad into g[gname] (N, K), optional bias colsum, returns dy W."""
        if bias is not None:
            self._colsum(dy, bias)
        gw = self.g[gname] if isinstance(gname, str) else gname
        ops.gemm_tn_acc(dy, x, gw.view(gw.shape[0], -1))
        if not need_dx:
            return None
        if mask is not None:
            return ops.gemm_masked(dy, self.w[key + ".t"], mask=mask, out_dtype=dx_dtype, **kw)
        return ops.gemm(dy, self.w[key + ".t"], out_dtype=dx_dtype, **kw)

    def _fusion_bwd(self, rk, d_up, t, r):
        """backward of _fusion_fwd: d(path) at the doubled resolution -> d x (the fusion block's input)."""
        u = rk + "resConfUnit2."
        y1, y, xr = t[f"y1_{r}"], t[f"y_{r}"], t[f"fx_{r}"]
        d_oc = ops.upsample2x_backward(d_up)
        n, h, w_, c = d_oc.shape
        d_y = self._lin_bwd(d_oc.view(-1, c), y.view(-1, c), rk + "out", rk + "out_conv.weight",
                            bias=rk + "out_conv.bias").view(n, h, w_, c)
        d_y1 = self._conv_bwd(d_y, y1, u + "conv2", bias=u + "conv2.bias", mask=y1)
        return self._conv_bwd(d_y1, xr, u + "conv1", bias=u + "conv1.bias", mask=xr, res1=d_y)

    def _trunk_bwd(self, pl, head, d_path, t, d_hooks):
        w, bb = self.w, self.bb
        k = head + ".dpt."
        sk = k + "scratch.refinenet"
        Fr, Np, N, rpf, gh, gw = pl["Fr"], pl["Np"], pl["N"], pl["rpf"], pl["gh"], pl["gw"]
        layers = t["layers"]
        d_layers = [None] * 4
        for r in (1, 2, 3):
            d_x = self._fusion_bwd(f"{sk}{r}.", d_path, t, r)
            rk = f"{sk}{r}.resConfUnit1."
            l, lr = layers[r - 1]
            d_u1 = self._conv_bwd(d_x, t[f"u1_{r}"], rk + "conv2", bias=rk + "conv2.bias", mask=t[f"u1_{r}"])
            d_layers[r - 1] = self._conv_bwd(d_u1, lr, rk + "conv1", bias=rk + "conv1.bias", mask=lr, res1=d_x)
            d_path = d_x
        d_layers[3] = self._fusion_bwd(sk + "4.", d_path, t, 4)
        for idx, hook in enumerate(pl["hooks"]):
            c = LAYER_DIMS[idx]
            m = t[f"m{idx}"]
            rows_out = N if idx == 0 else rpf
            off = 0 if idx == 0 else 1
            d_t0x = torch.zeros((Fr * rows_out, c), dtype=BF16, device=self.dev)   # token-row layout of the hook
            omap = dict(out=d_t0x, out_gin=Np, out_gout=rows_out, out_off=off)
            if idx == 2:   # the 1x1 output IS the map: the layer_rn dgrad writes the token rows directly
                ops.conv_wgrad(d_layers[idx], m, self._dwp(f"{k}rn{idx}", FEAT), kh=3, kw=3, pad=1)
                ops.conv_gemm(d_layers[idx], w[f"{k}rn{idx}.f"], kh=3, kw=3, pad=1, N=c, **_conv_omap(omap))
            else:
                d_m = self._conv_bwd(d_layers[idx], m, f"{k}rn{idx}")
                if idx in (0, 1):
                    kk = 4 if idx == 0 else 2
                    d_t1 = ops.pixel_unshuffle(d_m, kk)
                    cs = torch.zeros((kk * kk * c,), dtype=F32, device=self.dev)
                    ops.grad_prep(d_t1, want_copy=False, want_t=False, colsum=cs)
                    self.g[f"{k}act_postprocess.{idx}.1.bias"].add_(cs.view(kk * kk, c).sum(0))
                    ops.gemm_tn_acc(d_t1, t[f"t0_{idx}"], self._dwp(f"{k}ap{idx}.1", kk * kk * c))
                    ops.gemm(d_t1, w[f"{k}ap{idx}.1.t"], **omap)
                else:
                    d_o = d_m.view(-1, c)
                    self._colsum(d_o, f"{k}act_postprocess.3.1.bias")
                    ops.gemm_tn_acc(d_o, t["cols3"], self._dwp(f"{k}ap3.1", c))
                    d_cols = ops.gemm(d_o, w[f"{k}ap3.1.t"])
                    d_map = ops.col2im(d_cols, n=Fr, h=gh, w=gw, c=c, k=3, stride=2, pad=1)
                    # re-map the dense (Fr*Np, c) rows to the hook's token-row layout
                    d_t0x.view(Fr, rows_out, c)[:, off:off + Np].copy_(d_map.view(Fr, Np, c))
            # 1x1 token projection: tokens (hook rows) -> c channels
            A = pl["_tok"][hook]
            self._colsum(d_t0x, f"{k}act_postprocess.{idx}.0.bias")
            ops.gemm_tn_acc(d_t0x, A, self.g[f"{k}act_postprocess.{idx}.0.weight"].view(c, -1))
            ops.gemm(d_t0x, w[f"{k}ap{idx}.0.t"], res1=d_hooks[hook], out=d_hooks[hook])

    def _heads_bwd(self, pl, s, d_raw, d_means, d_cov, d_cov6, d_sh, d_opac):
        w, p = self.w, self.p
        f = lambda k: p[k].detach()
        Fr, H, W, rpf, N = pl["Fr"], pl["H"], pl["W"], pl["rpf"], pl["N"]
        E, D = self.bb["enc_embed_dim"], self.bb["dec_embed_dim"]
        G = Fr * H * W
        hooks = pl["hooks"]
        pl["_tok"] = {0: s["inter0"], **s["hooks"]}
        d_hooks = {0: torch.zeros((Fr * N, E), dtype=F32, device=self.dev)}
        for h in hooks[1:]:
            d_hooks[h] = torch.zeros((Fr * rpf, D), dtype=F32, device=self.dev)
        if all(t is None for t in (d_raw, d_means, d_cov, d_cov6, d_sh, d_opac)):
            return d_hooks                       # only the pose output carries a gradient
        gsp = s["gsp"]
        d_gsp = torch.zeros((G, 96), dtype=F32, device=self.dev)
        ops.gaussian_adapter_backward(gsp, self.m.d_sh, self.m.sh_mask.to(self.dev), d_gsp, center_col=84,
                                      param_col=0, d_raw=d_raw, d_means=d_means, d_cov=d_cov, d_cov6=d_cov6,
                                      d_shs=d_sh, d_opac=d_opac)
        # ---- gs head: gsp[:, :83] = y W4^T + b4, y = relu(conv3x3(merged)), merged = relu(conv7(img)) + up2(g1)
        k = "gaussian_param_head.dpt."
        t = s["tg"]
        nraw = self.m.raw_gs_dim
        cs = torch.zeros((96,), dtype=F32, device=self.dev)
        dgs, _ = ops.grad_prep(d_gsp, want_t=False, colsum=cs)
        dgs[:, nraw:].zero_()                   # columns 84..86 hold the centre gradient of the other head
        self.g[k + "head.4.bias"].add_(cs[:nraw])
        dW4 = torch.zeros((96, FEAT), dtype=F32, device=self.dev)
        y2d = t["y"].view(-1, FEAT)
        ops.gemm_tn_acc(dgs, y2d, dW4)
        self.g[k + "head.4.weight"].view(nraw, FEAT).add_(dW4[:nraw])
        # through the 1x1, the dropout (kept elements carry 1 / (1 - p)) and the ReLU: one masked epilogue
        scale = 1.0 / (1.0 - self.dropout_p) if self.dropout_p > 0 else 0.0
        d_y = ops.gemm_masked(dgs, w[k + "head.4.t"], mask=y2d, out_scale=scale).view(Fr, H, W, FEAT)
        d_merged = self._conv_bwd(d_y, t["merged"], k + "head.0")
        del d_y
        d_g1 = ops.upsample2x_backward(d_merged)
        d_c7 = ops.relu_backward(d_merged.view(-1, FEAT), t["c7"].view(-1, FEAT), out=d_merged.view(-1, FEAT),
                                 colsum=self.g[k + "input_merger.0.bias"]).view(Fr, H, W, FEAT)
        ops.conv_wgrad(d_c7, t["img8"], self._dwp(k + "merger", FEAT), kh=7, kw=1, pad=0, view=t["view"])
        del d_c7, d_merged
        self._trunk_bwd(pl, "gaussian_param_head", d_g1, t, d_hooks)
        # ---- centre head
        k = "downstream_head1.dpt."
        t = s["tp"]
        Cf = FEAT // 2
        w4 = f(k + "head.4.weight").flatten(1).contiguous()
        dw4 = torch.zeros((3, Cf), dtype=F32, device=self.dev)
        d_y2 = ops.pts_tail_backward(t["y2"], Cf, w4, f(k + "head.4.bias"), d_gsp[:, 84:], dw4,
                                     self.g[k + "head.4.bias"]).view(Fr, H, W, Cf)
        self.g[k + "head.4.weight"].view(3, Cf).add_(dw4)
        d_u = self._conv_bwd(d_y2, t["u"], k + "head.2", bias=k + "head.2.bias")
        d_y0 = ops.upsample2x_backward(d_u)
        del d_u, d_y2
        d_p1 = self._conv_bwd(d_y0, t["p1"], k + "head.0", bias=k + "head.0.bias")
        self._trunk_bwd(pl, "downstream_head1", d_p1, t, d_hooks)
        self._unpack_conv_grads()
        return d_hooks

    def _unpack_conv_grads(self) -> None:
        """packed-layout conv / ConvTranspose weight gradients -> the parameters' own layouts."""
        g = self.g
        for key, dwp in self.dwp.items():
            head, rest = key.split(".dpt.")
            k = head + ".dpt."
            if rest == "merger":
                g[k + "input_merger.0.weight"].add_(dwp.view(-1, 7, 8, 8)[:, :, :7, :3].permute(0, 3, 1, 2))
            elif rest.startswith("ap") and rest.endswith(".1") and rest[2] in "01":
                idx = int(rest[2])
                kk = 4 if idx == 0 else 2
                gw = g[f"{k}act_postprocess.{idx}.1.weight"]                  # [in, out, k, k]
                gw.add_(dwp.view(kk, kk, gw.shape[1], gw.shape[0]).permute(3, 2, 0, 1))
            elif rest == "ap3.1":
                gw = g[f"{k}act_postprocess.3.1.weight"]                      # [out, in, 3, 3]
                gw.add_(dwp.view(gw.shape[0], 3, 3, gw.shape[1]).permute(0, 3, 1, 2))
            else:
                name = rest.replace("rn", "scratch.layer_rn.") if rest.startswith("rn") else rest
                gw = g[f"{k}{name}.weight"]                                   # [out, in, kh, kw]
                o, c, kh, kw = gw.shape
                gw.add_(dwp.view(o, kh, kw, -1)[..., :c].permute(0, 3, 1, 2))

    def _decoder_bwd(self, pl, s, d_hooks, d_pred):
        p, w, bb, g = self.p, self.w, self.bb, self.g
        f = lambda k: p[k].detach()
        D, Hh = bb["dec_embed_dim"], bb["dec_num_heads"]
        theta = float(bb["temporal_rope_theta"])
        Fr, B, T, N, rpf = pl["Fr"], pl["B"], pl["T"], pl["N"], pl["rpf"]
        M = Fr * rpf
        cam = lambda t: t.view(Fr, rpf * D)[:, :D]
        z = lambda *shape: torch.zeros(shape, dtype=F32, device=self.dev)
        fa, fb = z(Fr, D), z(Fr, D)
        depth = bb["dec_depth"]
        # ---- final norms: hook[depth] = dec_norm(x) on the image rows, camera head on camera_dec_norm(cam rows)
        x_last = s["dec_x_last"]
        dx = ops.layernorm_mod_backward(x_last, d_hooks[depth], f("backbone.dec_norm.weight"), frame_a=fa,
                                        frame_b=fb, frames=Fr, rows_per_frame=rpf, skip_first=True)
        ops.adaln_reduce(fa, fb, f("backbone.dec_norm.weight"), f("backbone.dec_norm.bias"),
                         dgamma=g["backbone.dec_norm.weight"], dbeta=g["backbone.dec_norm.bias"])
        if d_pred is not None:
            d_cam_out = ops.camera_head_backward(s["cam_out"], f("camera_extrinsic_head.1.weight"),
                                                 f("camera_extrinsic_head.1.bias"), B, T, D,
                                                 d_pred.to(F32).contiguous(), g["camera_extrinsic_head.1.weight"],
                                                 g["camera_extrinsic_head.1.bias"])
            ops.layernorm_backward(cam(x_last), d_cam_out, f("backbone.camera_dec_norm.weight"), dx=cam(dx),
                                   dgamma=g["backbone.camera_dec_norm.weight"],
                                   dbeta=g["backbone.camera_dec_norm.bias"])
        for b in ("cam_head", "dec_norms"):
            self.reducer.bucket_ready(self.buckets[b])
        for i in reversed(range(depth)):
            if i + 1 in d_hooks and i + 1 != depth:
                dx.add_(d_hooks[i + 1])
            k = f"backbone.dec_blocks.{i}."
            c = s["dec_blocks"][i]
            m1, m2 = c["mod1"], c["mod2"]
            dmod1, dmod2 = z(Fr, 3 * D), z(Fr, 6 * D)
            # --- mlp_cam (camera rows): cam += fc2(gelu(fc1(cn2)))
            dyc, _ = ops.grad_prep(cam(dx), want_t=False, colsum=g[k + "mlp_cam.fc2.bias"])
            dac = self._lin_bwd(dyc, c["ac"], k + "mlp_cam.fc2", k + "mlp_cam.fc2.weight")
            dzc, _ = ops.grad_prep(dac, z=c["zc"], want_t=False, colsum=g[k + "mlp_cam.fc1.bias"])
            d_cn2 = self._lin_bwd(dzc, c["cn2b"], k + "mlp_cam.fc1", k + "mlp_cam.fc1.weight", dx_dtype=F32)
            # --- mlp (image rows)
            dbr = ops.gate_backward(dx, frames=Fr, rows_per_frame=rpf, branch=c["br3"], gate=m2[:, 5 * D:],
                                    dgate=dmod2[:, 5 * D:], colsum=g[k + "mlp.fc2.bias"], first_row_mode=2)
            da = self._lin_bwd(dbr, c["a3"], k + "mlp.fc2", k + "mlp.fc2.weight")
            dz, _ = ops.grad_prep(da, z=c["z3"], want_t=False, colsum=g[k + "mlp.fc1.bias"])
            dh = self._lin_bwd(dz, c["h3"], k + "mlp.fc1", k + "mlp.fc1.weight")
            fa.zero_(); fb.zero_()
            ops.layernorm_mod_backward(c["x2"], dh, f(k + "norm3.weight"), frame_a=fa, frame_b=fb, frames=Fr,
                                       rows_per_frame=rpf, skip_first=True, scale=m2[:, 3 * D:4 * D], dres=dx, dx=dx)
            ops.adaln_reduce(fa, fb, f(k + "norm3.weight"), f(k + "norm3.bias"), scale=m2[:, 3 * D:4 * D],
                             dscale=dmod2[:, 3 * D:4 * D], dshift=dmod2[:, 4 * D:5 * D],
                             dgamma=g[k + "norm3.weight"], dbeta=g[k + "norm3.bias"])
            # --- neighbour cross attention
            dbr = ops.gate_backward(dx, frames=Fr, rows_per_frame=rpf, branch=c["br2"], gate=m2[:, 2 * D:3 * D],
                                    dgate=dmod2[:, 2 * D:3 * D], colsum=g[k + "cross_attn.proj.bias"],
                                    first_row_mode=2)
            do = self._lin_bwd(dbr, c["o2"], k + "cross_attn.proj", k + "cross_attn.proj.weight")
            qkv = c["qkv2"]
            dqkv = torch.zeros_like(qkv)            # camera rows are not items: their gradient is zero
            ops.attention_backward(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], c["o2"], do, c["lse2"],
                                   dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], heads=Hh, q_start=pl["nb_q"],
                                   q_len=pl["nb_len"], kv_start0=pl["nb_k0"], kv_len0=pl["nb_len"],
                                   kv_start1=pl["nb_k1"], kv_len1=pl["nb_len1"], max_q_len=N, max_kv_len=2 * N,
                                   scale=0.125, dkv_tables=pl["nb_dkv"])
            ops.rope_rows_backward(dqkv, pl["pos_dec"], heads=Hh, q_col=0, k_col=D, base=100.0, cam_theta=theta)
            dh = self._lin_bwd(dqkv, c["h2"], k + "cross_qkv", k + "cross_attn.qkv.weight",
                               bias=k + "cross_attn.qkv.bias")
            fa.zero_(); fb.zero_()
            ops.layernorm_mod_backward(c["x1"], dh, f(k + "norm2.weight"), frame_a=fa, frame_b=fb, frames=Fr,
                                       rows_per_frame=rpf, skip_first=True, scale=m2[:, :D], dres=dx, dx=dx)
            ops.adaln_reduce(fa, fb, f(k + "norm2.weight"), f(k + "norm2.bias"), scale=m2[:, :D],
                             dscale=dmod2[:, :D], dshift=dmod2[:, D:2 * D], dgamma=g[k + "norm2.weight"],
                             dbeta=g[k + "norm2.bias"])
            # --- modulation2 (SiLU -> Linear on cam_norm2(cam)) and cam_norm2
            dmb, _ = ops.grad_prep(dmod2, want_t=False, colsum=g[k + "modulation2.proj.bias"])
            d_sil = self._lin_bwd(dmb, c["sil2"], k + "modulation2.proj", k + "modulation2.proj.weight", dx_dtype=F32)
            ops.silu_backward(c["cn2"], d_sil, d_cn2, accumulate=True)
            ops.layernorm_backward(cam(c["x1"]), d_cn2, f(k + "cam_norm2.weight"), dres=cam(dx), dx=cam(dx),
                                   dgamma=g[k + "cam_norm2.weight"], dbeta=g[k + "cam_norm2.bias"])
            # --- video + camera self attention (shared qkv / proj weights: image and camera rows together)
            dbr = ops.gate_backward(dx, frames=Fr, rows_per_frame=rpf, branch=c["br1"], gate=m1[:, 2 * D:],
                                    dgate=dmod1[:, 2 * D:], colsum=g[k + "attn.proj.bias"], first_row_mode=1)
            do = self._lin_bwd(dbr, c["o1"], k + "attn.proj", k + "attn.proj.weight")
            qkv = c["qkv1"]
            dqkv = torch.empty_like(qkv)
            ops.attention_backward(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], c["o1"], do, c["lse1"],
                                   dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], heads=Hh,
                                   q_start=pl["vid_start"], q_len=pl["vid_len"], kv_start0=pl["vid_start"],
                                   kv_len0=pl["vid_len"], max_q_len=T * rpf, max_kv_len=T * rpf,
                                   causal_block=rpf, scale=0.125)
            ops.rope_rows_backward(dqkv, pl["pos_dec"], heads=Hh, q_col=0, k_col=D, base=100.0, cam_theta=theta)
            dh = self._lin_bwd(dqkv, c["h1"], k + "attn.qkv", k + "attn.qkv.weight", bias=k + "attn.qkv.bias",
                               dx_dtype=F32)
            fa.zero_(); fb.zero_()
            ops.layernorm_mod_backward(c["x0"], dh, f(k + "norm1.weight"), frame_a=fa, frame_b=fb, frames=Fr,
                                       rows_per_frame=rpf, skip_first=True, scale=m1[:, :D], dres=dx, dx=dx)
            ops.adaln_reduce(fa, fb, f(k + "norm1.weight"), f(k + "norm1.bias"), scale=m1[:, :D],
                             dscale=dmod1[:, :D], dshift=dmod1[:, D:2 * D], dgamma=g[k + "norm1.weight"],
                             dbeta=g[k + "norm1.bias"])
            dmb, _ = ops.grad_prep(dmod1, want_t=False, colsum=g[k + "modulation1.proj.bias"])
            d_sil = self._lin_bwd(dmb, c["sil1"], k + "modulation1.proj", k + "modulation1.proj.weight", dx_dtype=F32)
            d_cn1 = cam(dh)                         # the camera row's gradient through the shared qkv projection
            ops.silu_backward(c["cn1"], d_sil, d_cn1, accumulate=True)
            ops.layernorm_backward(cam(c["x0"]), d_cn1, f(k + "cam_norm1.weight"), dres=cam(dx), dx=cam(dx),
                                   dgamma=g[k + "cam_norm1.weight"], dbeta=g[k + "cam_norm1.bias"])
            s["dec_blocks"][i] = None
            self.reducer.bucket_ready(self.buckets[f"dec{i}"])
        # ---- decoder stem: image rows = decoder_embed(enc_norm(x_enc)), camera rows = the two tokens
        dcam = cam(dx)
        g["backbone.camera_intrinsic_token"].add_(dcam.sum(0))
        if T > 1:
            g["backbone.camera_extrinsic_token"].add_(dcam.reshape(B, T, D)[:, 1:].sum((0, 1)))
        dyc, _ = ops.grad_prep(dx.view(Fr, rpf * D)[:, D:], want_t=False)       # compact (Fr, N*D) = (Fr*N, D) rows
        dyc = dyc.view(Fr * N, D)
        d_inter0 = d_hooks[0]
        self._lin_bwd(dyc, s["inter0"], "decoder_embed", "backbone.decoder_embed.weight",
                      bias="backbone.decoder_embed.bias", dx_dtype=F32, res1=d_inter0, out=d_inter0)
        self.reducer.bucket_ready(self.buckets["dec_stem"])
        return d_inter0

    def _encoder_bwd(self, pl, s, d_inter0):
        p, bb, g = self.p, self.bb, self.g
        f = lambda k: p[k].detach()
        E, lay = bb["enc_embed_dim"], pl["lay"]
        Fr, Np, N = pl["Fr"], pl["Np"], pl["N"]
        M = Fr * N
        dx = ops.layernorm_backward(s["enc_x_last"], d_inter0, f("backbone.enc_norm.weight"),
                                    dgamma=g["backbone.enc_norm.weight"], dbeta=g["backbone.enc_norm.bias"])
        for i in reversed(range(bb["enc_depth"])):
            k = f"backbone.enc_blocks.{i}."
            gi = {n: g[k + n] for n in ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias",
                                       "attn.proj.weight", "attn.proj.bias", "norm2.weight", "norm2.bias",
                                       "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")}
            dx = eg.block_backward(dx, self.enc_w[i], gi, lay, s["enc_blocks"][i])
            s["enc_blocks"][i] = None
            self.reducer.bucket_ready(self.buckets[f"enc{i}"])
        # stem: patch embedding (rows 0..Np-1 of a frame) and the intrinsic token (row Np)
        cs = torch.zeros((E,), dtype=F32, device=self.dev)
        dy, _ = ops.grad_prep(dx, want_t=False, colsum=cs)
        cols_full = torch.zeros((Fr, N, s["enc_cols"].shape[1]), dtype=BF16, device=self.dev)
        cols_full[:, :Np] = s["enc_cols"].view(Fr, Np, -1)
        ops.gemm_tn_acc(dy, cols_full.view(M, -1), g["backbone.patch_embed.proj.weight"].view(E, -1))
        d_intr = dx.view(Fr, N * E)[:, Np * E:]
        ci = torch.zeros((E,), dtype=F32, device=self.dev)
        di, _ = ops.grad_prep(d_intr, want_t=False, colsum=ci)
        K16 = torch.zeros((Fr, 16), dtype=BF16, device=self.dev)
        K16[:, :9] = s["K9"]
        dwi = torch.zeros((E, 16), dtype=F32, device=self.dev)
        ops.gemm_tn_acc(di, K16, dwi)
        g["backbone.intrinsic_encoder.weight"].add_(dwi[:, :9])
        g["backbone.intrinsic_encoder.bias"].add_(ci)
        g["backbone.patch_embed.proj.bias"].add_(cs - ci)
        self.reducer.bucket_ready(self.buckets["enc_stem"])


def _conv_omap(omap: dict) -> dict:
    """row-mapping keywords of ops.gemm that ops.conv_gemm also understands"""
    return omap
