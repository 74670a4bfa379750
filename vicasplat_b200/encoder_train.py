"""Training step of the ViT-Large image encoder (patch embedding + intrinsic token,
24 blocks with fused RoPE2D, final LayerNorm; backbone_vica.py:450-480,535-541, croco/blocks.py:81-236)
with hand-written backward, bucketed gradient all-reduce and the fused AdamW.

* Parameters are fp32 masters under the reference's ``state_dict`` names; the bf16 GEMM operand copies
  (W and W^T of every linear layer) are re-derived after each optimizer step by ONE ``vs_grad_prep``
  pass per weight.
* Gradients live in one flat fp32 buffer per BUCKET (a bucket = one block; plus the stem:
  patch_embed / intrinsic_encoder / enc_norm); ``param.grad`` are views.  Small tensors (biases,
  LayerNorm) come first in a bucket: they are accumulated with atomics and need zeroing, the weight
  gradients behind them are accumulated by the wgrad GEMM's in-place residual.
* ``GradReducer``: the one collective of this system (the reference wraps the model in
  DDP, main.py:111).  A bucket is all-reduced (NCCL, asynchronously) as soon as its block's backward
  has been enqueued, i.e. in reverse layer order, overlapping the remaining backward pass.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import encoder_grad as eg, ops

_BLOCK_SMALL = ("norm1.weight", "norm1.bias", "attn.qkv.bias", "attn.proj.bias", "norm2.weight",
                "norm2.bias", "mlp.fc1.bias", "mlp.fc2.bias")
_BLOCK_BIG = ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")
_STEM_SMALL = ("backbone.patch_embed.proj.bias", "backbone.intrinsic_encoder.bias",
               "backbone.enc_norm.weight", "backbone.enc_norm.bias")
_STEM_BIG = ("backbone.intrinsic_encoder.weight", "backbone.patch_embed.proj.weight")


@dataclass(frozen=True)
class ViTEncoderConfig:
    """The values of backbone/vica.yaml that shape the image encoder (any object with these attributes
    works as `cfg`, e.g. the oracle's EncoderConfig in the tests)."""
    enc_embed_dim: int = 1024
    enc_depth: int = 24
    enc_num_heads: int = 16
    patch_size: int = 16
    mlp_ratio: float = 4.0
    ln_eps: float = 1e-6


class GradBucket:
    """One flat fp32 gradient buffer; `views[name]` are the per-parameter gradients inside it."""

    def __init__(self, shapes: Dict[str, torch.Size], small: List[str], big: List[str], device):
        self.names = list(small) + list(big)
        off, offs = 0, {}
        for n in self.names:
            offs[n] = off
            off += (shapes[n].numel() + 3) // 4 * 4          # 16-byte aligned views
            if n == small[-1]:
                self.n_small = off
        self.flat = torch.zeros((off,), dtype=torch.float32, device=device)
        self.views = {n: self.flat[offs[n]: offs[n] + shapes[n].numel()].view(shapes[n]) for n in self.names}

    def zero(self, weights_too: bool = True) -> None:
        (self.flat if weights_too else self.flat[: self.n_small]).zero_()


class GradReducer:
    """Averages gradient buckets over the ranks of the default process group.  ``bucket_ready`` starts
    an asynchronous all-reduce on the bucket's flat buffer (NCCL orders it after the kernels already
    enqueued on the current stream and runs it beside the ones enqueued later); ``finish`` makes the
    current stream wait for all of them.  World size 1 / no process group: no-ops.

    ``compress_bf16=True`` sends bf16 copies of the buckets (half the bytes on the wire, SURVEY §8e:
    1.16 GB instead of 2.31 GB for the whole model) and writes the averaged result back into the fp32
    bucket; the default keeps the reference's fp32 all-reduce."""

    def __init__(self, compress_bf16: bool = False) -> None:
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._nccl = self.world > 1 and dist.get_backend() == "nccl"
        self.compress_bf16 = compress_bf16
        self._pending: list = []
        self.bytes_reduced = 0

    def bucket_ready(self, bucket: GradBucket) -> None:
        if self.world == 1:
            return
        op = dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM
        wire = bucket.flat.to(torch.bfloat16) if self.compress_bf16 else bucket.flat
        self._pending.append((dist.all_reduce(wire, op=op, async_op=True), bucket, wire))
        self.bytes_reduced += wire.numel() * wire.element_size()

    def finish(self) -> None:
        for work, bucket, wire in self._pending:
            work.wait()
            if wire is not bucket.flat:
                bucket.flat.copy_(wire)
            if not self._nccl:                                # gloo has no AVG
                bucket.flat.div_(self.world)
        self._pending.clear()


class VitEncoderTrainer:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg, frames: int, img_hw, device,
                 reducer: Optional[GradReducer] = None):
        self.cfg, self.dev = cfg, device
        self.E, self.depth, self.P = cfg.enc_embed_dim, cfg.enc_depth, cfg.patch_size
        self.frames = frames
        self.gh, self.gw = img_hw[0] // self.P, img_hw[1] // self.P
        self.lay = eg.FrameLayout.make(frames, self.gh, self.gw, cfg.enc_num_heads, device)
        self.reducer = reducer or GradReducer()
        # fp32 masters under the reference's names
        self.params: Dict[str, torch.Tensor] = {}
        want = list(_STEM_SMALL) + list(_STEM_BIG)
        for i in range(self.depth):
            want += [f"backbone.enc_blocks.{i}.{n}" for n in _BLOCK_SMALL + _BLOCK_BIG]
        for n in want:
            self.params[n] = sd[n].detach().to(device=device, dtype=torch.float32).contiguous().clone().requires_grad_(True)
        # gradient buckets; param.grad = view
        self.block_buckets: List[GradBucket] = []
        for i in range(self.depth):
            k = f"backbone.enc_blocks.{i}."
            shapes = {n: self.params[k + n].shape for n in _BLOCK_SMALL + _BLOCK_BIG}
            self.block_buckets.append(GradBucket(shapes, list(_BLOCK_SMALL), list(_BLOCK_BIG), device))
            for n, v in self.block_buckets[-1].views.items():
                self.params[k + n].grad = v
        self.stem_bucket = GradBucket({n: self.params[n].shape for n in _STEM_SMALL + _STEM_BIG},
                                      list(_STEM_SMALL), list(_STEM_BIG), device)
        for n, v in self.stem_bucket.views.items():
            self.params[n].grad = v
        self._dw_intr = torch.zeros((self.E, 16), dtype=torch.float32, device=device)
        self._colsum_all = torch.zeros((self.E,), dtype=torch.float32, device=device)
        self.w: List[Dict[str, torch.Tensor]] = [dict() for _ in range(self.depth)]
        self.repack()
        self._saved: Optional[dict] = None

    # ---- parameters
    def parameters(self) -> List[torch.Tensor]:
        return list(self.params.values())

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return dict(self.params)

    @torch.no_grad()
    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        """Copies reference-named tensors INTO the fp32 masters (gradient views, optimizer state and
        device pointer tables stay valid) and re-derives the bf16 operands."""
        missing = [k for k in self.params if k not in sd]
        if strict and missing:
            raise KeyError(f"VitEncoderTrainer.load_state_dict: missing {missing[:4]}{' ...' if len(missing) > 4 else ''}")
        for k, p in self.params.items():
            if k in sd:
                if tuple(sd[k].shape) != tuple(p.shape):
                    raise ValueError(f"VitEncoderTrainer.load_state_dict: {k} has shape {tuple(sd[k].shape)}, expected {tuple(p.shape)}")
                p.copy_(sd[k])
        self.repack()

    @torch.no_grad()
    def repack(self) -> None:
        """bf16 operand copies of every weight, both orientations, one pass per weight."""
        for i in range(self.depth):
            k, w = f"backbone.enc_blocks.{i}.", self.w[i]
            for name in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                w[name], w[name + ".t"] = ops.grad_prep(self.params[k + name + ".weight"])
                w[name + ".bias"] = self.params[k + name + ".bias"]
            for name in ("norm1", "norm2"):
                w[name + ".weight"], w[name + ".bias"] = self.params[k + name + ".weight"], self.params[k + name + ".bias"]
        self.w_patch, _ = ops.grad_prep(self.params["backbone.patch_embed.proj.weight"].flatten(1), want_t=False)

    # ---- forward
    @torch.no_grad()
    def forward(self, image: torch.Tensor, intrinsics: torch.Tensor) -> torch.Tensor:
        """image (frames, 3, H, W) fp32, intrinsics (frames, 3, 3) -> enc_norm(tokens) bf16 (frames * n, E)."""
        p, lay, E = self.params, self.lay, self.E
        Np, N = self.gh * self.gw, lay.n
        M = self.frames * N
        cols = ops.patchify(image.contiguous(), self.P)
        x = torch.empty((M, E), dtype=torch.float32, device=self.dev)
        ops.gemm(cols, self.w_patch, bias=p["backbone.patch_embed.proj.bias"], out=x,
                 out_gin=Np, out_gout=N, out_off=0)
        K9 = intrinsics.reshape(self.frames, 9).to(torch.float32).contiguous()
        ops.intrinsic_token(K9, p["backbone.intrinsic_encoder.weight"], p["backbone.intrinsic_encoder.bias"],
                            x, self.frames, E, N, Np)
        saved = []
        for i in range(self.depth):
            s = eg.Saved()
            x = eg.block_forward(x, self.w[i], lay, s)
            saved.append(s)
        out, _ = ops.layernorm(x, p["backbone.enc_norm.weight"], p["backbone.enc_norm.bias"], eps=self.cfg.ln_eps)
        self._saved = dict(cols=cols, K9=K9, blocks=saved, x_last=x)
        return out

    # ---- backward
    @torch.no_grad()
    def backward(self, d_out: torch.Tensor) -> None:
        """d_out: gradient w.r.t. forward()'s output, (frames * n, E) fp32 or bf16.  Fills every
        ``param.grad``; buckets are handed to the reducer in reverse layer order as they complete."""
        s, p, lay, E = self._saved, self.params, self.lay, self.E
        assert s is not None, "backward() needs a forward() first"
        Np, N, Fr = self.gh * self.gw, lay.n, self.frames
        M = Fr * N
        for b in self.block_buckets:
            b.zero()
        self.stem_bucket.zero()
        sg = self.stem_bucket.views
        dx = ops.layernorm_backward(s["x_last"], d_out, p["backbone.enc_norm.weight"],
                                    dgamma=sg["backbone.enc_norm.weight"], dbeta=sg["backbone.enc_norm.bias"],
                                    eps=self.cfg.ln_eps)
        for i in reversed(range(self.depth)):
            dx = eg.block_backward(dx, self.w[i], self.block_buckets[i].views, lay, s["blocks"][i])
            s["blocks"][i] = None                                   # activations of block i are dead
            self.reducer.bucket_ready(self.block_buckets[i])
        # stem.  Patch embedding (croco/blocks.py:195-225): dW = dY^T cols over the patch rows; the
        # intrinsic rows are zero rows of the padded im2col matrix.  Intrinsic token (Linear 9 -> E,
        # backbone_vica.py:535-536): the strided rows frame * n + Np.
        self._colsum_all.zero_()
        dy, _ = ops.grad_prep(dx, want_t=False, colsum=self._colsum_all)
        cols_full = torch.zeros((Fr, N, s["cols"].shape[1]), dtype=torch.bfloat16, device=self.dev)
        cols_full[:, :Np] = s["cols"].view(Fr, Np, -1)
        dWp = sg["backbone.patch_embed.proj.weight"].view(E, -1)
        ops.gemm(dy, cols_full.view(M, -1), tn=True, out=dWp, res1=dWp)
        d_intr = dx.view(Fr, N * E)[:, Np * E:]                     # (Fr, E) view, row stride N * E
        di, _ = ops.grad_prep(d_intr, want_t=False, colsum=sg["backbone.intrinsic_encoder.bias"])
        K16 = torch.zeros((Fr, 16), dtype=torch.bfloat16, device=self.dev)
        K16[:, :9] = s["K9"]
        ops.gemm(di, K16, tn=True, out=self._dw_intr)
        sg["backbone.intrinsic_encoder.weight"].copy_(self._dw_intr[:, :9])
        torch.sub(self._colsum_all, sg["backbone.intrinsic_encoder.bias"],
                  out=sg["backbone.patch_embed.proj.bias"])
        self.reducer.bucket_ready(self.stem_bucket)
        self.reducer.finish()
        self._saved = None


# ------------------------------------------------------------------ bucket plan of the WHOLE model (C4)
# Parameters that are constructed but never executed (refinenet4 is called with one input, so its
# resConfUnit1 is dead: dpt_block.py:184,196; dpt_head.py:58; dpt_gs_head.py:143).  Measured on the
# unmodified reference by oracle/make_unused_params_golden.py -> tests/golden/unused_params.json.
# They stay in state_dict (checkpoints) but are left out of every gradient bucket: the reference
# needs DDP's find_unused_parameters=True for them (src/main.py:111), a static plan does not.
UNUSED_PARAMETER_MARK = ".scratch.refinenet4.resConfUnit1."


def plan_buckets(param_names, enc_depth: int, dec_depth: int):
    """Static gradient-bucket plan for all parameters of VicaSplat, in the order their gradients
    become final during the backward pass (reverse execution order: heads, decoder norms, decoder
    blocks N-1..0, decoder stem, encoder blocks N-1..0, encoder stem) -- the order ``GradReducer``
    is fed in.  Returns (buckets: list of (name, [parameter names]), unused: [parameter names])."""
    order = ["gs_head", "pts_head", "cam_head", "dec_norms"]
    order += [f"dec{i}" for i in reversed(range(dec_depth))] + ["dec_stem"]
    order += [f"enc{i}" for i in reversed(range(enc_depth))] + ["enc_stem"]
    members = {n: [] for n in order}
    unused = []

    def bucket_of(k: str) -> str:
        if k.startswith("gaussian_param_head."):
            return "gs_head"
        if k.startswith("downstream_head1."):
            return "pts_head"
        if k.startswith(("camera_extrinsic_head.", "camera_intrinsic_head.")):
            return "cam_head"
        if k.startswith(("backbone.dec_norm.", "backbone.camera_dec_norm.")):
            return "dec_norms"
        if k.startswith("backbone.dec_blocks."):
            return "dec" + k.split(".")[2]
        if k.startswith(("backbone.decoder_embed.", "backbone.camera_extrinsic_token",
                         "backbone.camera_intrinsic_token")):
            return "dec_stem"
        if k.startswith("backbone.enc_blocks."):
            return "enc" + k.split(".")[2]
        if k.startswith(("backbone.patch_embed.", "backbone.intrinsic_encoder.", "backbone.enc_norm.")):
            return "enc_stem"
        raise KeyError(f"plan_buckets: no bucket rule for parameter {k!r}")

    for k in param_names:
        if UNUSED_PARAMETER_MARK in k:
            unused.append(k)
            continue
        b = bucket_of(k)
        if b not in members:
            raise KeyError(f"plan_buckets: {k!r} maps to {b!r}, beyond the configured depth")
        members[b].append(k)
    return [(n, members[n]) for n in order if members[n]], unused
