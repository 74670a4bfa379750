"""Drop-in for the reference's trainable ViT ``Block`` (src/model/encoder/backbone/croco/blocks.py:115-130)
on the hand-written sm_100a training path.

Same constructor arguments, same sub-module / ``state_dict`` names (``norm1``, ``attn.qkv``,
``attn.proj``, ``norm2``, ``mlp.fc1``, ``mlp.fc2``) and the same ``forward(x, xpos)``, so
``VicaNet.enc_blocks`` (backbone_vica.py:395-399) can be built from this class and trained by the
reference's own loop: parameters stay ordinary fp32 ``nn.Parameter``s, gradients arrive through
``torch.autograd`` (one custom Function per block) and DDP / the optimizer see nothing unusual.
Under the Function: ``encoder_grad.block_forward`` / ``block_backward`` (tcgen05 GEMMs and attention,
fused HBM-bound kernels).  CUDA only; dropout / drop-path must be 0 (they are in every shipped
config) and head_dim must be 64.
"""
from __future__ import annotations

import torch
from torch import nn

from . import encoder_grad as eg, ops

_ORDER = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
          "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
          "mlp.fc2.weight", "mlp.fc2.bias")


class _Attention(nn.Module):
    def __init__(self, dim, rope=None, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.rope = rope


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden_features, in_features)


class _BlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, block, lay, *params):
        shape = x.shape
        x2 = x.detach().reshape(-1, shape[-1]).contiguous()
        w = block._packed()
        saved = eg.Saved() if any(ctx.needs_input_grad) else None
        out = eg.block_forward(x2, w, lay, saved)
        ctx.w, ctx.lay, ctx.saved = w, lay, saved
        return out.view(shape)

    @staticmethod
    def backward(ctx, dout):
        g = eg.zero_grads(ctx.w)
        d2 = dout.reshape(-1, dout.shape[-1]).to(torch.float32).contiguous()
        dx = eg.block_backward(d2, ctx.w, g, ctx.lay, ctx.saved)
        ctx.saved = None
        return (dx.view(dout.shape), None, None, *[g[n] for n in _ORDER])


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, rope=None):
        super().__init__()
        if drop or attn_drop or drop_path:
            raise NotImplementedError("vicasplat_b200 Block: dropout / drop-path are not supported")
        if act_layer is not nn.GELU:
            raise NotImplementedError("vicasplat_b200 Block: the activation is exact-erf GELU")
        if dim % num_heads or dim // num_heads != 64:
            raise NotImplementedError("vicasplat_b200 Block: head_dim must be 64")
        if rope is None:
            raise NotImplementedError("vicasplat_b200 Block: RoPE2D is fused into the qkv projection (rope=None is not wired)")
        if not qkv_bias:
            raise NotImplementedError("vicasplat_b200 Block: qkv_bias=False is not wired (every config sets it)")
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, rope=rope, num_heads=num_heads, qkv_bias=qkv_bias)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))
        self._pack, self._versions = None, None

    def _named(self):
        mods = {"norm1": self.norm1, "attn.qkv": self.attn.qkv, "attn.proj": self.attn.proj,
                "norm2": self.norm2, "mlp.fc1": self.mlp.fc1, "mlp.fc2": self.mlp.fc2}
        return {f"{k}.{leaf}": getattr(m, leaf) for k, m in mods.items() for leaf in ("weight", "bias")}

    def _packed(self):
        """bf16 operand copies (W and W^T), rebuilt when a parameter has been updated in place."""
        named = self._named()
        versions = tuple((p.data_ptr(), p._version) for p in named.values())
        if self._pack is None or versions != self._versions:
            w = {}
            with torch.no_grad():
                for k, p in named.items():
                    if k.endswith(".weight") and p.dim() == 2:
                        w[k[:-7]], w[k[:-7] + ".t"] = ops.grad_prep(p.detach())
                    else:
                        w[k] = p.detach()
                w["ln_eps"] = float(self.norm1.eps)
            self._pack, self._versions = w, versions
        return self._pack

    def forward(self, x, xpos):
        if not x.is_cuda:
            raise RuntimeError("vicasplat_b200 Block needs CUDA tensors (there is no CPU fallback)")
        B, N, C = x.shape
        if self.norm1.eps != self.norm2.eps:
            raise NotImplementedError("vicasplat_b200 Block: norm1 / norm2 must share eps")
        base = float(getattr(self.attn.rope, "base", 100.0))
        pos = xpos.reshape(B * N, 2).to(torch.int32).contiguous()
        start = torch.arange(B, dtype=torch.int32, device=x.device) * N
        length = torch.full((B,), N, dtype=torch.int32, device=x.device)
        lay = eg.FrameLayout(B, N, self.attn.num_heads, pos, start, length, rope_base=base)
        named = self._named()
        return _BlockFn.apply(x.to(torch.float32), self, lay, *[named[n] for n in _ORDER])
