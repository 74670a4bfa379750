"""``VicaSplat`` encoder plugin (src/model/encoder/vicasplat.py:59-290) on hand-written sm_100a
kernels.

* Same constructor / ``forward`` contract and the same ``state_dict()`` keys as the reference
  (847 keys at the shipped config), so reference checkpoints load by key.
* fp32 master parameters live in this nn.Module; ``EncoderEngine`` keeps packed bf16 copies laid
  out for the tcgen05 GEMM / implicit-GEMM kernels and a preallocated activation workspace, and
  replays the whole forward (~560 launches) as ONE CUDA graph per input shape.
* Residual streams are fp32 in HBM; GEMM operands bf16; accumulation fp32 in TMEM.

Launch sequence per stage (reference lines in brackets):
  patch embed [croco/blocks.py:195-225] -> 24 x ViT block [croco/blocks.py:81-130]
  -> enc_norm, decoder_embed, camera tokens [backbone_vica.py:482-494]
  -> 12 x MixDecoderBlock [backbone_vica.py:280-335] with the camera token stored as row 0 of every
     frame (258 rows per frame), so image and camera tokens share every GEMM launch
  -> camera head [vicasplat.py:179-199] -> 2 x DPT head [heads/dpt_block.py, dpt_head.py,
     dpt_gs_head.py] as NHWC implicit GEMMs -> Gaussian adapter [common/gaussian_adapter.py:167-212]
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch
from torch import Tensor, nn

from . import ops
from ._lib import VS_ACT_GELU, VS_ACT_NONE, VS_ACT_RELU

BF16 = torch.bfloat16
F32 = torch.float32


# ------------------------------------------------------------------------------------ config
@dataclass
class OpacityMappingCfg:
    initial: float = 0.0
    final: float = 0.0
    warm_up: int = 1


@dataclass
class GaussianAdapterCfg:
    gaussian_scale_min: float = 0.005
    gaussian_scale_max: float = 0.04
    sh_degree: int = 4
    scale_act: str = "softplus"


def default_backbone_cfg() -> dict:
    """config/model/encoder/backbone/vica.yaml with the re10k_8view overrides."""
    return dict(img_size=256, patch_size=16, enc_embed_dim=1024, enc_depth=24, enc_num_heads=16,
                dec_embed_dim=768, dec_depth=12, dec_num_heads=12, mlp_ratio=4.0,
                temporal_rope_theta=30, rope_dim_list=[32, 32], use_blocked_causal_attention=True,
                use_framewise_modulation=True, use_cross_neighbor_attention=True,
                use_intrinsic_embedding=True)


@dataclass
class VicaSplatCfg:
    name: str = "vicasplat"
    backbone: dict = field(default_factory=default_backbone_cfg)
    visualizer: object = None
    gaussian_adapter: GaussianAdapterCfg = field(default_factory=GaussianAdapterCfg)
    apply_bounds_shim: bool = True
    opacity_mapping: OpacityMappingCfg = field(default_factory=OpacityMappingCfg)
    predict_opacity: bool = False
    input_mean: tuple = (0.5, 0.5, 0.5)
    input_std: tuple = (0.5, 0.5, 0.5)
    pretrained_weights: str = ""
    gs_center_head_type: str = "dpt"
    gs_param_head_type: str = "dpt_gs"
    predict_conf: bool = False
    camera_type: str = "dq"


@dataclass
class Gaussians:
    """src/model/types.py:7-12 (+ scales / rotations as the adapter returns them)."""
    means: Tensor
    covariances: Tensor
    harmonics: Tensor
    opacities: Tensor
    scales: Optional[Tensor] = None
    rotations: Optional[Tensor] = None
    cov6: Optional[Tensor] = None      # packed upper triangle (xx,xy,xz,yy,yz,zz): what the rasterizer reads


class _AttrDict(dict):
    __getattr__ = dict.__getitem__


LAYER_DIMS = (96, 192, 384, 768)
FEAT = 256
GSP_XYZ, GSP_LD = 96, 100   # inference head-output rows: 83 Gaussian parameters padded to 96, then the centre


def _param_shapes(bb: dict, raw_gs_dim: int) -> Dict[str, tuple]:
    """key -> shape of every state_dict entry, in the reference's naming (Appendix B)."""
    E, D, P = bb["enc_embed_dim"], bb["dec_embed_dim"], bb["patch_size"]
    he, hd = int(E * bb["mlp_ratio"]), int(D * bb["mlp_ratio"])
    s: Dict[str, tuple] = {}

    def lin(k, o, i):
        s[k + ".weight"], s[k + ".bias"] = (o, i), (o,)

    def ln(k, c):
        s[k + ".weight"], s[k + ".bias"] = (c,), (c,)

    def cv(k, o, i, kk, bias=True):
        s[k + ".weight"] = (o, i, kk, kk)
        if bias:
            s[k + ".bias"] = (o,)

    s["backbone.camera_extrinsic_token"] = (D,)
    s["backbone.camera_intrinsic_token"] = (D,)
    cv("backbone.patch_embed.proj", E, 3, P)
    for i in range(bb["enc_depth"]):
        k = f"backbone.enc_blocks.{i}"
        ln(k + ".norm1", E); lin(k + ".attn.qkv", 3 * E, E); lin(k + ".attn.proj", E, E)
        ln(k + ".norm2", E); lin(k + ".mlp.fc1", he, E); lin(k + ".mlp.fc2", E, he)
    ln("backbone.enc_norm", E)
    lin("backbone.decoder_embed", D, E)
    for i in range(bb["dec_depth"]):
        k = f"backbone.dec_blocks.{i}"
        ln(k + ".cam_norm1", D); lin(k + ".modulation1.proj", 3 * D, D); ln(k + ".norm1", D)
        lin(k + ".attn.qkv", 3 * D, D); lin(k + ".attn.proj", D, D)
        ln(k + ".cam_norm2", D); lin(k + ".modulation2.proj", 6 * D, D); ln(k + ".norm2", D)
        for n in ("projq", "projk", "projv", "proj"):
            lin(f"{k}.cross_attn.{n}", D, D)
        ln(k + ".norm3", D)
        lin(k + ".mlp.fc1", hd, D); lin(k + ".mlp.fc2", D, hd)
        lin(k + ".mlp_cam.fc1", hd, D); lin(k + ".mlp_cam.fc2", D, hd)
    ln("backbone.dec_norm", D)
    ln("backbone.camera_dec_norm", D)
    lin("backbone.intrinsic_encoder", E, 9)
    for head in ("downstream_head1", "gaussian_param_head"):
        k = head + ".dpt"
        for idx in range(4):
            s[f"{k}.scratch.layer_rn.{idx}.weight"] = (FEAT, LAYER_DIMS[idx], 3, 3)
        for r in (1, 2, 3, 4):
            rk = f"{k}.scratch.refinenet{r}"
            cv(rk + ".out_conv", FEAT, FEAT, 1)
            for u in ("resConfUnit1", "resConfUnit2"):
                cv(f"{rk}.{u}.conv1", FEAT, FEAT, 3)
                cv(f"{rk}.{u}.conv2", FEAT, FEAT, 3)
        if head == "downstream_head1":
            cv(k + ".head.0", FEAT // 2, FEAT, 3); cv(k + ".head.2", FEAT // 2, FEAT // 2, 3)
            cv(k + ".head.4", 3, FEAT // 2, 1)
        else:
            cv(k + ".head.0", FEAT, FEAT, 3, bias=False); cv(k + ".head.4", raw_gs_dim, FEAT, 1)
            cv(k + ".input_merger.0", FEAT, 3, 7)
        dims = [E, D, D, D]
        for idx in range(4):
            cv(f"{k}.act_postprocess.{idx}.0", LAYER_DIMS[idx], dims[idx], 1)
        c0, c1, c3 = LAYER_DIMS[0], LAYER_DIMS[1], LAYER_DIMS[3]
        s[f"{k}.act_postprocess.0.1.weight"], s[f"{k}.act_postprocess.0.1.bias"] = (c0, c0, 4, 4), (c0,)
        s[f"{k}.act_postprocess.1.1.weight"], s[f"{k}.act_postprocess.1.1.bias"] = (c1, c1, 2, 2), (c1,)
        cv(f"{k}.act_postprocess.3.1", c3, c3, 3)
    lin("camera_extrinsic_head.1", 8, D)
    return s


class _Node(nn.Module):
    """Anonymous container: only there to give parameters the reference's dotted names."""


def _register(root: nn.Module, dotted: str, p: nn.Parameter) -> None:
    *path, leaf = dotted.split(".")
    m = root
    for name in path:
        if name not in m._modules:
            m.add_module(name, _Node())
        m = m._modules[name]
    m.register_parameter(leaf, p)


# ------------------------------------------------------------------------------------ module
class VicaSplat(nn.Module):
    """Drop-in for the reference encoder plugin (``ENCODERS['vicasplat']``)."""
    patch_size: int = 16

    def __init__(self, cfg: Optional[VicaSplatCfg] = None, weight_dtype=None, device=None,
                 precision: str = "bf16") -> None:
        """precision: the 16-bit operand format of the forward-only engine --
          "bf16"  speed mode (default): bf16 GEMM / attention operands, fp32 accumulation and residual streams;
          "fp16"  parity mode: fp16 operands, i.e. the 10-bit mantissa the reference's TF32 matmuls round
                  their fp32 operands to (backbone_vica.py:9) at the same tensor-core rate as bf16 (twice
                  kind::tf32's) -- for callers who want reference-grade numerics.  Activations must stay
                  below fp16's 65 504 (they do after LayerNorm; the residual streams are fp32 either way)."""
        super().__init__()
        if precision not in ("bf16", "fp16"):
            raise ValueError("precision must be 'bf16' or 'fp16'")
        self.precision = precision
        self.cfg = cfg if cfg is not None else VicaSplatCfg()
        bb = dict(default_backbone_cfg(), **dict(self.cfg.backbone))
        if not bb["use_intrinsic_embedding"]:
            raise NotImplementedError("use_intrinsic_embedding=False (camera_intrinsic_head) is "
                                      "not on the shipped 8-view path")
        if self.cfg.camera_type != "dq" or self.cfg.predict_conf:
            raise NotImplementedError("only camera_type='dq', predict_conf=False are implemented")
        if self.cfg.gaussian_adapter.scale_act != "softplus":
            raise NotImplementedError("only scale_act='softplus' is implemented")
        assert bb["enc_embed_dim"] // bb["enc_num_heads"] == 64
        assert bb["dec_embed_dim"] // bb["dec_num_heads"] == 64
        assert bb["dec_depth"] > 9
        self._bb = bb
        self.d_sh = (self.cfg.gaussian_adapter.sh_degree + 1) ** 2
        self.raw_gs_dim = 1 + 7 + 3 * self.d_sh
        self.camera_extrinsic_channels = 8
        for key, shape in _param_shapes(bb, self.raw_gs_dim).items():
            _register(self, key, nn.Parameter(torch.empty(shape, dtype=F32, device=device)))
        # the reference registers layer_rn[i] under two names (dpt_block.py:33-75): same tensor
        for head in ("downstream_head1", "gaussian_param_head"):
            scratch = self._modules[head]._modules["dpt"]._modules["scratch"]
            for idx in range(4):
                node = _Node()
                node.register_parameter("weight", scratch._modules["layer_rn"]._modules[str(idx)].weight)
                scratch.add_module(f"layer{idx + 1}_rn", node)
        self.backbone.config = _AttrDict(bb)
        mask = torch.ones((self.d_sh,), dtype=F32)
        for deg in range(1, self.cfg.gaussian_adapter.sh_degree + 1):
            mask[deg * deg:(deg + 1) ** 2] = 0.1 * 0.25 ** deg
        self.register_buffer("sh_mask", mask, persistent=False)
        self.reset_parameters()
        self._engine: Optional[EncoderEngine] = None

    # ---- initialisation (scheme of backbone_vica.py:431-448, vicasplat.py:118-127)
    @torch.no_grad()
    def reset_parameters(self) -> None:
        for name, p in self.named_parameters():
            if p.dim() == 1:
                if name.endswith("token"):
                    nn.init.normal_(p, std=0.02)
                elif name.endswith("weight"):
                    nn.init.ones_(p)
                else:
                    nn.init.zeros_(p)
            elif "modulation" in name or name.startswith("camera_extrinsic_head"):
                nn.init.zeros_(p)
            elif p.dim() == 2:
                nn.init.xavier_uniform_(p)
            else:
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))

    def enable_gradient_checkpointing(self) -> None:
        """Accepted for interface compatibility (vicasplat.py:140): the training path keeps bf16
        activations (~6 GB per scene at 8 views) and runs micro-batches instead of recomputing."""

    def get_data_shim(self):
        mean, std = self.cfg.input_mean, self.cfg.input_std

        def data_shim(batch):
            for view in ("context", "target"):
                if view in batch and "image" in batch[view]:
                    img = batch[view]["image"]
                    m = torch.tensor(mean, dtype=img.dtype, device=img.device).view(1, 1, 3, 1, 1)
                    s = torch.tensor(std, dtype=img.dtype, device=img.device).view(1, 1, 3, 1, 1)
                    batch[view] = dict(batch[view], image=(img - m) / s)
            return batch
        return data_shim

    def invalidate(self) -> None:
        """Call after changing parameters (load_state_dict does it) so bf16 copies are re-packed."""
        self._engine = None
        self._train_engine = None

    def train_engine(self):
        """The training-path engine behind the differentiable forward (vicasplat_b200.train.TrainEngine,
        gradients returned through torch.autograd); its bf16 operand copies follow in-place parameter
        updates (optimizer steps) through the parameters' version counters."""
        from .train import TrainEngine
        versions = tuple(p._version for p in self.parameters())
        eng = getattr(self, "_train_engine", None)
        if eng is None:
            eng = self._train_engine = TrainEngine(self, attach_grads=False)
        elif versions != self._train_versions:
            eng.repack()
        self._train_versions = versions
        return eng

    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self.invalidate()
        return r

    def _apply(self, fn, *a, **k):
        self._engine = None
        self._train_engine = None
        return super()._apply(fn, *a, **k)

    def set_precision(self, precision: str) -> None:
        if precision not in ("bf16", "fp16"):
            raise ValueError("precision must be 'bf16' or 'fp16'")
        self.precision = precision
        self._engine = None

    def engine(self) -> "EncoderEngine":
        if self._engine is None:
            self._engine = EncoderEngine(self, precision=self.precision)
        return self._engine

    def forward(self, context: dict, global_step: int = 0, visualization_dump: Optional[dict] = None,
                distill: bool = False, compute_viewspace_depth: bool = True,
                clone_outputs: bool = True, **kwargs) -> dict:
        """In training mode with gradients enabled the call is DIFFERENTIABLE (one autograd.Function over
        the hand-written backward pass, ``_forward_train``): every output the reference's training step
        consumes carries gradients to all parameters.  Otherwise (``eval()`` or ``torch.no_grad()``) the
        forward-only engine replays its CUDA graph.

        ``clone_outputs=False`` (forward-only path) returns views of the engine's static output buffers
        (valid until the next forward on this module): what a consumer on the same stream needs,
        without the ~3 GB of copies per 8-scene batch."""
        image = context["image"]
        if not image.is_cuda:
            raise RuntimeError("vicasplat_b200.VicaSplat runs on CUDA only (no CPU fallback)")
        intr = context.get("intrinsics", None)
        assert intr is not None, "use_intrinsic_embedding=True needs context['intrinsics']"
        if self.training and torch.is_grad_enabled() and not distill:
            return self._forward_train(context, image, intr, compute_viewspace_depth, visualization_dump)
        with torch.no_grad():
            return self._forward_infer(context, image, intr, distill, compute_viewspace_depth,
                                       clone_outputs, visualization_dump)

    def _forward_train(self, context, image, intr, compute_viewspace_depth, visualization_dump):
        names = [n for n, _ in self.named_parameters()]
        params = [p for _, p in self.named_parameters()]
        raw, cov, cov6, sh, opac, pred, c2w, scales, rot = _TrainFn.apply(self, names, image, intr, *params)
        centers = raw[..., :3]                      # a view, as in the reference (vicasplat.py:256-259)
        depth = None
        if compute_viewspace_depth:
            ext = context["extrinsics"]
            Rinv = torch.linalg.inv(ext[:, :, :3, :3])
            rel = centers - ext[:, :, None, None, :3, 3]
            depth = (rel * Rinv[:, :, None, None, 2, :]).sum(-1)
        gaussians = Gaussians(means=centers, covariances=cov, harmonics=sh, opacities=opac[..., None],
                              scales=scales, rotations=rot, cov6=cov6)
        if visualization_dump is not None:
            visualization_dump["depth"] = gaussians.means[..., -1:]
        return dict(pred_extrins=pred, pred_intrins=None, gaussian_camera_extrins=c2w,
                    gaussian_camera_intrins=None, gaussian_centers=centers, confidence=None,
                    context_view_depths=depth, gaussians=gaussians, raw_gaussians=raw, cov6=cov6)

    def _forward_infer(self, context, image, intr, distill, compute_viewspace_depth, clone_outputs,
                       visualization_dump):
        B, T, _, H, W = image.shape
        out = self.engine().run(image, intr, heads=not distill, gs=not distill,
                                clone_outputs=clone_outputs)
        centers = out["raw"][..., :3] if not distill else out["centers"]
        depth = None
        if compute_viewspace_depth:
            ext = context["extrinsics"]
            Rinv = torch.linalg.inv(ext[:, :, :3, :3])
            rel = centers - ext[:, :, None, None, :3, 3]
            depth = (rel * Rinv[:, :, None, None, 2, :]).sum(-1)
        res = dict(pred_extrins=out["pred_extrins"], pred_intrins=None,
                   gaussian_camera_extrins=out["c2w"], gaussian_camera_intrins=None,
                   gaussian_centers=centers, confidence=None, context_view_depths=depth)
        if distill:
            return res
        g = out["gaussians"]
        gaussians = Gaussians(means=centers, covariances=g["cov"], harmonics=g["sh"],
                              opacities=g["opac"][..., None], scales=g["scales"], rotations=g["rot"],
                              cov6=g["cov6"])
        if visualization_dump is not None:
            visualization_dump["depth"] = gaussians.means[..., -1:]
        res.update(gaussians=gaussians, raw_gaussians=out["raw"], cov6=g["cov6"])
        return res


class _TrainFn(torch.autograd.Function):
    """VicaSplat.forward as ONE autograd node: forward = TrainEngine.forward (keeps activations), backward
    = the hand-written backward pass; parameter gradients are handed back to autograd (so DDP hooks and
    any optimizer see ordinary ``.grad`` accumulation)."""

    @staticmethod
    def forward(ctx, model, names, image, intr, *params):
        eng = model.train_engine()
        out = eng.forward(image, intr)
        B, T, _, H, W = image.shape
        shp = (B, T, H, W)
        ctx.eng, ctx.names, ctx.G = eng, names, B * T * H * W
        ctx.set_materialize_grads(False)
        c2w, scales, rot = out["c2w"], out["scales"].view(*shp, 3), out["rot"].view(*shp, 4)
        ctx.mark_non_differentiable(c2w, scales, rot)
        return (out["raw"].view(*shp, -1), out["cov"].view(*shp, 3, 3), out["cov6"].view(*shp, 6),
                out["sh"].view(*shp, 3, -1), out["opac"].view(*shp), out["pred_extrins"], c2w, scales, rot)

    @staticmethod
    def backward(ctx, d_raw, d_cov, d_cov6, d_sh, d_opac, d_pred, *_):
        G = ctx.G
        flat = lambda t, *s: None if t is None else t.reshape(G, *s).to(torch.float32).contiguous()
        eng = ctx.eng
        eng.backward(d_raw=flat(d_raw, -1), d_cov=flat(d_cov, 3, 3), d_cov6=flat(d_cov6, 6),
                     d_sh=None if d_sh is None else flat(d_sh, 3, d_sh.shape[-1]), d_opac=flat(d_opac),
                     d_pred=d_pred)
        return (None, None, None, None, *[eng.g.get(n) for n in ctx.names])


# ------------------------------------------------------------------------------------ engine
def _bf(t: Tensor, dtype=BF16) -> Tensor:
    return t.detach().to(dtype).contiguous()


def _pack_conv(w: Tensor, dtype=BF16) -> Tensor:
    """[N, Cin, kh, kw] -> 16-bit [N, kh*kw*cin_pad] (tap-major, channel-minor, zero padded)."""
    n, cin, kh, kw = w.shape
    cp = (cin + 63) // 64 * 64
    p = torch.zeros((n, kh * kw, cp), dtype=F32, device=w.device)
    p[:, :, :cin] = w.detach().permute(0, 2, 3, 1).reshape(n, kh * kw, cin)
    return _bf(p.reshape(n, -1), dtype)


class EncoderEngine:
    """Packed weights + activation workspace + CUDA-graph replay for one VicaSplat module."""

    def __init__(self, model: VicaSplat, use_graph: bool = True, precision: str = "bf16"):
        self.m = model
        self.bb = model._bb
        self.use_graph = use_graph
        self.half = torch.float16 if precision == "fp16" else BF16     # 16-bit operand format (see VicaSplat)
        # K = 448 stem: the bilinear-x2 residual gathered in the GEMM epilogue costs ~4x the main
        # loop in issue slots; a stand-alone upsample + plain bf16 residual is HBM-bound instead
        self.fuse_stem_upsample = os.environ.get("VS_FUSE_STEM_UP", "1") == "1"
        self.dev = next(model.parameters()).device
        assert self.dev.type == "cuda", "move the module to CUDA first"
        self.sd = {k: v.detach() for k, v in model.state_dict().items()}
        self.w: Dict[str, Tensor] = {}
        self._pack()
        self._plans: Dict[tuple, dict] = {}

    # ---- weights
    def _pack(self) -> None:
        sd, w, bb = self.sd, self.w, self.bb
        f = lambda k: sd[k].to(F32).contiguous()
        half = self.half
        _bf = lambda t: t.detach().to(half).contiguous()                # noqa: F811 (shadows the module helper)
        _pack_conv = lambda t: globals()["_pack_conv"](t, half)         # noqa: F811
        for k, v in sd.items():
            if v.dim() == 1:
                w[k] = f(k)
        w["patch"] = _bf(sd["backbone.patch_embed.proj.weight"].flatten(1))
        for i in range(bb["enc_depth"]):
            k = f"backbone.enc_blocks.{i}"
            for n in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"):
                w[f"{k}.{n}"] = _bf(sd[f"{k}.{n}.weight"])
        w["decoder_embed"] = _bf(sd["backbone.decoder_embed.weight"])
        for i in range(bb["dec_depth"]):
            k = f"backbone.dec_blocks.{i}"
            for n in ("modulation1.proj", "modulation2.proj", "attn.qkv", "attn.proj",
                      "cross_attn.proj", "mlp.fc1", "mlp.fc2", "mlp_cam.fc1", "mlp_cam.fc2"):
                w[f"{k}.{n}"] = _bf(sd[f"{k}.{n}.weight"])
            w[f"{k}.cross_qkv"] = _bf(torch.cat([sd[f"{k}.cross_attn.{n}.weight"]
                                                 for n in ("projq", "projk", "projv")], 0))
            w[f"{k}.cross_qkv.bias"] = torch.cat([sd[f"{k}.cross_attn.{n}.bias"]
                                                  for n in ("projq", "projk", "projv")], 0).to(F32).contiguous()
        w["intr.w"] = f("backbone.intrinsic_encoder.weight")
        w["cam_head.w"] = f("camera_extrinsic_head.1.weight")
        for head in ("downstream_head1", "gaussian_param_head"):
            k = head + ".dpt"
            for idx in range(4):
                w[f"{k}.ap{idx}.0"] = _bf(sd[f"{k}.act_postprocess.{idx}.0.weight"].flatten(1))
                w[f"{k}.rn{idx}"] = _pack_conv(sd[f"{k}.scratch.layer_rn.{idx}.weight"])
            for idx, kk in ((0, 4), (1, 2)):   # ConvTranspose(k == stride) as a GEMM to (tap, cout)
                wt = sd[f"{k}.act_postprocess.{idx}.1.weight"]               # [in, out, k, k]
                w[f"{k}.ap{idx}.1"] = _bf(wt.permute(2, 3, 1, 0).reshape(kk * kk * wt.shape[1], wt.shape[0]))
                w[f"{k}.ap{idx}.1.bias"] = sd[f"{k}.act_postprocess.{idx}.1.bias"].to(F32).repeat(kk * kk).contiguous()
            w3 = sd[f"{k}.act_postprocess.3.1.weight"]                       # 3x3 stride 2: im2col GEMM
            w[f"{k}.ap3.1"] = _bf(w3.permute(0, 2, 3, 1).reshape(w3.shape[0], -1))
            for r in (1, 2, 3, 4):
                rk = f"{k}.scratch.refinenet{r}"
                w[rk + ".out"] = _bf(sd[rk + ".out_conv.weight"].flatten(1))
                for u in ("resConfUnit1", "resConfUnit2"):
                    for c in ("conv1", "conv2"):
                        w[f"{rk}.{u}.{c}"] = _pack_conv(sd[f"{rk}.{u}.{c}.weight"])
            w[k + ".head.0"] = _pack_conv(sd[k + ".head.0.weight"])
            if head == "downstream_head1":
                w[k + ".head.2"] = _pack_conv(sd[k + ".head.2.weight"])
                w[k + ".head.4.w"] = sd[k + ".head.4.weight"].flatten(1).to(F32).contiguous()
            else:
                # 1x1 head padded to 96 outputs: every 32-column chunk of its epilogue is a full one
                # (the 19-column chunk of N = 83 took the row-by-row store path: 1.10 -> 0.6 ms per 8 scenes)
                w4 = torch.zeros((GSP_XYZ, FEAT), dtype=F32, device=self.dev)
                w4[: self.m.raw_gs_dim] = sd[k + ".head.4.weight"].flatten(1)
                w[k + ".head.4"] = _bf(w4)
                b4 = torch.zeros((GSP_XYZ,), dtype=F32, device=self.dev)
                b4[: self.m.raw_gs_dim] = sd[k + ".head.4.bias"]
                w[k + ".head.4.bias96"] = b4
                # 7x7 stem as kh = 7, kw = 1 over windows of 8 pixels x 8 (zero-padded) channels:
                # K index = dy * 64 + dx * 8 + c
                w7 = sd[k + ".input_merger.0.weight"]                        # [256,3,7,7]
                p = torch.zeros((w7.shape[0], 7, 8, 8), dtype=F32, device=w7.device)
                p[:, :, :7, :3] = w7.permute(0, 2, 3, 1)
                w[k + ".merger"] = _bf(p.reshape(w7.shape[0], -1))

    # ---- per-shape plan: buffers + item tables
    def _plan(self, B, T, H, W) -> dict:
        key = (B, T, H, W)
        if key in self._plans:
            return self._plans[key]
        bb, dev = self.bb, self.dev
        P, E, D = bb["patch_size"], bb["enc_embed_dim"], bb["dec_embed_dim"]
        gh, gw = H // P, W // P
        assert H % P == 0 and W % P == 0 and gh % 2 == 0 and gw % 2 == 0
        Fr, Np = B * T, gh * gw
        N, rpf = Np + 1, Np + 2
        he, hd = int(E * bb["mlp_ratio"]), int(D * bb["mlp_ratio"])
        i32 = dict(dtype=torch.int32, device=dev)
        z = lambda *s, dt=self.half: torch.zeros(s, dtype=dt, device=dev)
        pl = dict(B=B, T=T, H=H, W=W, gh=gh, gw=gw, Fr=Fr, Np=Np, N=N, rpf=rpf)
        pl["image"] = z(Fr, 3, H, W, dt=F32)
        pl["K9"] = z(Fr, 9, dt=F32)
        # encoder
        pl["x_enc"] = z(Fr * N, E, dt=F32)
        pl["h_enc"] = z(Fr * N, E)
        pl["qkv_enc"] = z(Fr * N, 3 * E)
        pl["att_enc"] = z(Fr * N, E)
        pl["mlp_enc"] = z(Fr * N, he)
        pl["inter0"] = z(Fr * N, E)
        ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
        pos = torch.cat([torch.stack([ys, xs], -1).reshape(Np, 2), torch.tensor([[gh, 0]])], 0)
        pl["pos_enc"] = pos[None].expand(Fr, N, 2).reshape(Fr * N, 2).to(**i32).contiguous()
        fr = torch.arange(Fr, **i32)
        pl["enc_start"], pl["enc_len"] = fr * N, torch.full((Fr,), N, **i32)
        # decoder: row 0 of each frame is the camera token (position y = -1 - t: temporal rope)
        pos_d = torch.zeros((B, T, rpf, 2), dtype=torch.int32)
        pos_d[:, :, 1:] = pos.to(torch.int32)
        pos_d[:, :, 0, 0] = -1 - torch.arange(T, dtype=torch.int32)[None]
        pl["pos_dec"] = pos_d.reshape(Fr * rpf, 2).to(dev).contiguous()
        pl["x_dec"] = z(Fr * rpf, D, dt=F32)
        pl["h_dec"] = z(Fr * rpf, D)
        pl["qkv_dec"] = z(Fr * rpf, 3 * D)
        pl["att_dec"] = z(Fr * rpf, D)
        pl["mlp_dec"] = z(Fr * rpf, hd)
        pl["cam_n"] = z(Fr, D, dt=F32)
        pl["cam_nb"] = z(Fr, D)
        pl["cam_h"] = z(Fr, hd)
        pl["mod1"] = z(Fr, 3 * D, dt=F32)
        pl["mod2"] = z(Fr, 6 * D, dt=F32)
        sc = torch.arange(B, **i32)
        pl["vid_start"], pl["vid_len"] = sc * (T * rpf), torch.full((B,), T * rpf, **i32)
        t_idx = torch.arange(T)
        prev = torch.where(t_idx > 0, t_idx - 1, t_idx + 1)
        nxt = torch.where(t_idx < T - 1, t_idx + 1, t_idx - 1)
        base = (torch.arange(B)[:, None] * T)
        pl["nb_q"] = ((base + t_idx[None]) * rpf + 1).reshape(-1).to(**i32)
        pl["nb_k0"] = ((base + prev[None]) * rpf + 1).reshape(-1).to(**i32)
        pl["nb_k1"] = ((base + nxt[None]) * rpf + 1).reshape(-1).to(**i32)
        pl["nb_len"] = torch.full((Fr,), N, **i32)
        pl["nb_len1"] = ((prev != nxt)[None].expand(B, T).reshape(-1) * N).to(**i32)
        hooks = [0, bb["dec_depth"] * 2 // 4, bb["dec_depth"] * 3 // 4, bb["dec_depth"]]
        pl["hooks"] = hooks
        pl["hook_buf"] = {h: z(Fr * rpf, D) for h in hooks[1:]}
        pl["cam_out"] = z(Fr, D, dt=F32)
        pl["pred"] = None
        # outputs
        pl["raw"] = z(Fr * H * W, 3 + self.m.raw_gs_dim, dt=F32)
        # head outputs, 16-byte aligned rows: 83 Gaussian parameters at columns 0.., xyz at 84..86
        pl["gsp"] = z(Fr * H * W, GSP_LD, dt=F32)   # head outputs: params [0, 83), centre xyz at GSP_XYZ
        self._plans[key] = pl
        return pl

    # ---- stages
    def _encoder(self, pl, taps):
        w, bb = self.w, self.bb
        E, H = bb["enc_embed_dim"], bb["enc_num_heads"]
        Fr, Np, N = pl["Fr"], pl["Np"], pl["N"]
        x = pl["x_enc"]
        cols = ops.patchify(pl["image"], bb["patch_size"], half=self.half)
        ops.gemm(cols, w["patch"], bias=w["backbone.patch_embed.proj.bias"], out=x,
                 out_gin=Np, out_gout=N, out_off=0)
        ops.intrinsic_token(pl["K9"], w["intr.w"], w["backbone.intrinsic_encoder.bias"], x, Fr, E, N, Np)
        for i in range(bb["enc_depth"]):
            k = f"backbone.enc_blocks.{i}"
            ops.layernorm(x, w[k + ".norm1.weight"], w[k + ".norm1.bias"], out_bf16=pl["h_enc"])
            # RoPE2D of q and k (croco/blocks.py:101-103) is applied in the GEMM epilogue
            qkv = ops.gemm(pl["h_enc"], w[k + ".attn.qkv"], bias=w[k + ".attn.qkv.bias"], out=pl["qkv_enc"],
                           rope=(pl["pos_enc"], 0, E, H, 100.0, 30.0))
            ops.attention(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], pl["att_enc"], heads=H,
                          q_start=pl["enc_start"], q_len=pl["enc_len"], kv_start0=pl["enc_start"],
                          kv_len0=pl["enc_len"], max_q_len=N, max_kv_len=N, scale=0.125)
            ops.gemm(pl["att_enc"], w[k + ".attn.proj"], bias=w[k + ".attn.proj.bias"], res1=x, out=x)
            ops.layernorm(x, w[k + ".norm2.weight"], w[k + ".norm2.bias"], out_bf16=pl["h_enc"])
            ops.gemm(pl["h_enc"], w[k + ".mlp.fc1"], bias=w[k + ".mlp.fc1.bias"], act=VS_ACT_GELU,
                     out=pl["mlp_enc"])
            ops.gemm(pl["mlp_enc"], w[k + ".mlp.fc2"], bias=w[k + ".mlp.fc2.bias"], res1=x, out=x)
            if taps is not None:
                taps[f"enc{i}"] = x.clone()
        ops.layernorm(x, w["backbone.enc_norm.weight"], w["backbone.enc_norm.bias"], out_bf16=pl["inter0"])
        if taps is not None:
            taps["inter0"] = pl["inter0"].clone()

    def _decoder(self, pl, taps):
        w, bb = self.w, self.bb
        D, H = bb["dec_embed_dim"], bb["dec_num_heads"]
        theta = float(bb["temporal_rope_theta"])
        Fr, T, N, rpf = pl["Fr"], pl["T"], pl["N"], pl["rpf"]
        x = pl["x_dec"]
        ops.gemm(pl["inter0"], w["decoder_embed"], bias=w["backbone.decoder_embed.bias"], out=x,
                 out_gin=N, out_gout=rpf, out_off=1)
        ops.camera_tokens(w["backbone.camera_intrinsic_token"], w["backbone.camera_extrinsic_token"],
                          x, Fr, T, D, rpf)
        cam_rows = x.view(Fr, rpf * D)[:, :D]          # strided (Fr, D) view of the camera rows
        for i in range(bb["dec_depth"]):
            k = f"backbone.dec_blocks.{i}"
            # --- video + camera self attention
            ops.layernorm(cam_rows, w[k + ".cam_norm1.weight"], w[k + ".cam_norm1.bias"],
                          out_f32=pl["cam_n"], want_bf16=False)
            sil = ops.silu_bf16(pl["cam_n"], Fr, D, half=self.half)
            ops.gemm(sil, w[k + ".modulation1.proj"], bias=w[k + ".modulation1.proj.bias"], out=pl["mod1"])
            m1 = pl["mod1"]
            ops.layernorm(x, w[k + ".norm1.weight"], w[k + ".norm1.bias"],
                          w0=w[k + ".cam_norm1.weight"], b0=w[k + ".cam_norm1.bias"],
                          scale=m1[:, :D], shift=m1[:, D:2 * D], rows_per_frame=rpf, out_bf16=pl["h_dec"])
            qkv = ops.gemm(pl["h_dec"], w[k + ".attn.qkv"], bias=w[k + ".attn.qkv.bias"], out=pl["qkv_dec"],
                           rope=(pl["pos_dec"], 0, D, H, 100.0, theta))
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], pl["att_dec"], heads=H,
                          q_start=pl["vid_start"], q_len=pl["vid_len"], kv_start0=pl["vid_start"],
                          kv_len0=pl["vid_len"], max_q_len=T * rpf, max_kv_len=T * rpf,
                          causal_block=rpf, scale=0.125)
            ops.gemm(pl["att_dec"], w[k + ".attn.proj"], bias=w[k + ".attn.proj.bias"],
                     gate=m1[:, 2 * D:], gate_rows=rpf, first_row_mode=1, res1=x, out=x)
            # --- neighbour cross attention (image rows only)
            ops.layernorm(cam_rows, w[k + ".cam_norm2.weight"], w[k + ".cam_norm2.bias"],
                          out_f32=pl["cam_n"], out_bf16=pl["cam_nb"])
            sil = ops.silu_bf16(pl["cam_n"], Fr, D, half=self.half)
            ops.gemm(sil, w[k + ".modulation2.proj"], bias=w[k + ".modulation2.proj.bias"], out=pl["mod2"])
            m2 = pl["mod2"]
            ops.layernorm(x, w[k + ".norm2.weight"], w[k + ".norm2.bias"], scale=m2[:, :D],
                          shift=m2[:, D:2 * D], rows_per_frame=rpf, out_bf16=pl["h_dec"])
            qkv = ops.gemm(pl["h_dec"], w[k + ".cross_qkv"], bias=w[k + ".cross_qkv.bias"], out=pl["qkv_dec"],
                           rope=(pl["pos_dec"], 0, D, H, 100.0, theta))
            ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], pl["att_dec"], heads=H,
                          q_start=pl["nb_q"], q_len=pl["nb_len"], kv_start0=pl["nb_k0"],
                          kv_len0=pl["nb_len"], kv_start1=pl["nb_k1"], kv_len1=pl["nb_len1"],
                          max_q_len=N, max_kv_len=2 * N, scale=0.125)
            ops.gemm(pl["att_dec"], w[k + ".cross_attn.proj"], bias=w[k + ".cross_attn.proj.bias"],
                     gate=m2[:, 2 * D:3 * D], gate_rows=rpf, first_row_mode=2, res1=x, out=x)
            # --- MLPs
            ops.layernorm(x, w[k + ".norm3.weight"], w[k + ".norm3.bias"], scale=m2[:, 3 * D:4 * D],
                          shift=m2[:, 4 * D:5 * D], rows_per_frame=rpf, out_bf16=pl["h_dec"])
            ops.gemm(pl["h_dec"], w[k + ".mlp.fc1"], bias=w[k + ".mlp.fc1.bias"], act=VS_ACT_GELU,
                     out=pl["mlp_dec"])
            ops.gemm(pl["mlp_dec"], w[k + ".mlp.fc2"], bias=w[k + ".mlp.fc2.bias"],
                     gate=m2[:, 5 * D:], gate_rows=rpf, first_row_mode=2, res1=x, out=x)
            ops.gemm(pl["cam_nb"], w[k + ".mlp_cam.fc1"], bias=w[k + ".mlp_cam.fc1.bias"],
                     act=VS_ACT_GELU, out=pl["cam_h"])
            ops.gemm(pl["cam_h"], w[k + ".mlp_cam.fc2"], bias=w[k + ".mlp_cam.fc2.bias"], res1=x,
                     out=x, out_gin=1, out_gout=rpf, out_off=0)
            layer = i + 1
            if layer in pl["hook_buf"]:
                if layer == bb["dec_depth"]:
                    ops.layernorm(x, w["backbone.dec_norm.weight"], w["backbone.dec_norm.bias"],
                                  out_bf16=pl["hook_buf"][layer])
                else:
                    ops.layernorm(x, normalize=False, out_bf16=pl["hook_buf"][layer])
            if taps is not None:
                taps[f"dec{layer}"] = x.clone()
        ops.layernorm(cam_rows, w["backbone.camera_dec_norm.weight"], w["backbone.camera_dec_norm.bias"],
                      out_f32=pl["cam_out"], want_bf16=False)
        pred, c2w = ops.camera_head(pl["cam_out"], D, w["cam_head.w"], w["camera_extrinsic_head.1.bias"],
                                    pl["B"], T, D)
        pl["pred"], pl["c2w"] = pred, c2w
        if taps is not None:
            taps["cam_out"] = pl["cam_out"].clone()

    def _conv(self, x, key, *, k=3, N=FEAT, bias=None, act=VS_ACT_NONE, res1=None, res2=None,
              relu_copy=False, out_dtype=BF16):
        out2 = torch.empty(x.shape[:3] + (N,), dtype=self.half, device=x.device) if relu_copy else None
        out = ops.conv_gemm(x, self.w[key], kh=k, kw=k, pad=k // 2, N=N, bias=bias, act=act,
                            res1=res1, res2=res2, out2=out2, out_dtype=out_dtype)
        return (out, out2) if relu_copy else out

    def _rcu(self, rk, x, x_relu, extra=None, relu_copy=False):
        """ResidualConvUnit: conv2(relu(conv1(relu(x)))) + x (+ extra)."""
        b = lambda c: self.w[f"{rk}.{c}.bias"]
        y = self._conv(x_relu, f"{rk}.conv1", bias=b("conv1"), act=VS_ACT_RELU)
        return self._conv(y, f"{rk}.conv2", bias=b("conv2"), res1=x, res2=extra, relu_copy=relu_copy)

    def _fusion(self, rk, x, x_relu):
        """resConfUnit2 -> bilinear x2 -> 1x1 out_conv (dpt_block.py:196-205).  A 1x1 convolution
        (+ bias) commutes with bilinear interpolation (the weights of every output pixel sum to 1),
        so the out_conv runs BEFORE the upsampling, on a quarter of the pixels."""
        y = self._rcu(rk + ".resConfUnit2", x, x_relu)
        n, h, w_, c = y.shape
        out = ops.gemm(y.view(-1, c), self.w[rk + ".out"], bias=self.w[rk + ".out_conv.bias"])
        return ops.upsample2x(out.view(n, h, w_, FEAT))

    def _trunk(self, pl, head, taps):
        w, bb = self.w, self.bb
        k = head + ".dpt"
        Fr, Np, N, rpf, gh, gw = pl["Fr"], pl["Np"], pl["N"], pl["rpf"], pl["gh"], pl["gw"]
        E, D = bb["enc_embed_dim"], bb["dec_embed_dim"]
        layers = []
        for idx, hook in enumerate(pl["hooks"]):
            if idx == 0:
                A, C, gs = pl["inter0"], E, N * E
            else:
                A, C, gs = pl["hook_buf"][hook][1:], D, rpf * D    # skip the camera row
            t = ops.gemm(A, w[f"{k}.ap{idx}.0"], K=C, a_rows=Np, a_groups=Fr, a_row_stride=C,
                         a_group_stride=gs, bias=w[f"{k}.act_postprocess.{idx}.0.bias"])
            c = LAYER_DIMS[idx]
            if idx in (0, 1):
                kk = 4 if idx == 0 else 2
                t = ops.gemm(t, w[f"{k}.ap{idx}.1"], bias=w[f"{k}.ap{idx}.1.bias"])
                t = ops.pixel_shuffle(t, Fr, gh, gw, c, kk)
            elif idx == 2:
                t = t.view(Fr, gh, gw, c)
            else:
                cols = ops.im2col(t, nchw_f32=False, n=Fr, h=gh, w=gw, c=c, k=3, stride=2, pad=1,
                                  kpad=9 * c)
                t = ops.gemm(cols, w[f"{k}.ap3.1"], bias=w[f"{k}.act_postprocess.3.1.bias"])
                t = t.view(Fr, gh // 2, gw // 2, c)
            layers.append(self._conv(t, f"{k}.rn{idx}", relu_copy=True))
        s = k + ".scratch.refinenet"
        l3, l3r = layers[3]
        path = self._fusion(s + "4", l3, l3r)
        for r, (l, lr) in ((3, layers[2]), (2, layers[1]), (1, layers[0])):
            x, xr = self._rcu(f"{s}{r}.resConfUnit1", l, lr, extra=path, relu_copy=True)
            path = self._fusion(f"{s}{r}", x, xr)
        if taps is not None:
            taps[head + ".path1"] = path.clone()
        return path

    def _heads(self, pl, taps, gs=True):
        w = self.w
        Fr, H, W = pl["Fr"], pl["H"], pl["W"]
        gsp = pl["gsp"]
        # --- Gaussian centres: 'regression' head + exp-depth postprocess
        k = "downstream_head1.dpt"
        p1 = self._trunk(pl, "downstream_head1", taps)
        y = self._conv(p1, k + ".head.0", N=FEAT // 2, bias=w[k + ".head.0.bias"])
        y = self._conv(ops.upsample2x(y), k + ".head.2", N=FEAT // 2, bias=w[k + ".head.2.bias"],
                       act=VS_ACT_RELU)
        ops.pts_tail(y, FEAT // 2, w[k + ".head.4.w"], w[k + ".head.4.bias"], gsp[:, GSP_XYZ:], Fr * H * W)
        if not gs:
            return
        # --- Gaussian parameters: trunk x2 + relu(conv7x7(image)) -> conv3x3 + ReLU -> 1x1
        k = "gaussian_param_head.dpt"
        p1 = self._trunk(pl, "gaussian_param_head", taps)                  # (Fr, H/2, W/2, 256)
        img8 = ops.image_nhwc8(pl["image"], pad=3, half=self.half)        # (Fr, H+6, W+8, 8)
        # relu(conv7x7(image)) + bilinear_x2(p1): the image is addressed through an overlapping TMA
        # view (no im2col buffer), the upsampling happens in the epilogue (no full-res copy of p1)
        if self.fuse_stem_upsample:
            merged = ops.conv_gemm(img8, w[k + ".merger"], kh=7, kw=1, pad=0, N=FEAT,
                                   bias=w[k + ".input_merger.0.bias"], act=VS_ACT_RELU, res1=p1,
                                   res_up2=True,
                                   view=(Fr, H, W, 64, H + 6, 8, (W + 8) * 8, (H + 6) * (W + 8) * 8))
        else:
            merged = ops.conv_gemm(img8, w[k + ".merger"], kh=7, kw=1, pad=0, N=FEAT,
                                   bias=w[k + ".input_merger.0.bias"], act=VS_ACT_RELU,
                                   res1=ops.upsample2x(p1),
                                   view=(Fr, H, W, 64, H + 6, 8, (W + 8) * 8, (H + 6) * (W + 8) * 8))
        y = self._conv(merged, k + ".head.0", act=VS_ACT_RELU)
        ops.gemm(y.view(-1, FEAT), w[k + ".head.4"], bias=w[k + ".head.4.bias96"], out=gsp, N=GSP_XYZ)

    def _forward(self, pl, heads=True, gs=True, taps=None):
        self._encoder(pl, taps)
        self._decoder(pl, taps)
        if heads:
            self._heads(pl, taps, gs)
            if gs:
                pl["gauss"] = ops.gaussian_adapter(pl["gsp"], self.m.d_sh, self.m.sh_mask.to(self.dev),
                                                   center_col=GSP_XYZ, param_col=0, raw_out=pl["raw"])
            else:   # distill: only the centres are produced
                pl["raw"][:, :3].copy_(pl["gsp"][:, GSP_XYZ:GSP_XYZ + 3])

    # ---- public
    @torch.no_grad()
    def run(self, image: Tensor, intrinsics: Tensor, heads: bool = True, gs: bool = True,
            taps: Optional[dict] = None, clone_outputs: bool = True) -> dict:
        B, T, _, H, W = image.shape
        pl = self._plan(B, T, H, W)
        pl["image"].copy_(image.reshape(B * T, 3, H, W))
        pl["K9"].copy_(intrinsics.reshape(B * T, 9))
        mode = (heads, gs)
        if self.use_graph and taps is None:
            graphs = pl.setdefault("graphs", {})
            if mode not in graphs:
                # warm-up on a side stream (sets kernel attributes, fills allocator pools), then capture
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._forward(pl, heads, gs)
                torch.cuda.current_stream().wait_stream(s)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._forward(pl, heads, gs)
                graphs[mode] = (g, dict(pred=pl["pred"], c2w=pl["c2w"], gauss=pl.get("gauss")))
            g, outs = graphs[mode]
            g.replay()
        else:
            self._forward(pl, heads, gs, taps)
            outs = dict(pred=pl["pred"], c2w=pl["c2w"], gauss=pl.get("gauss"))
        cl = (lambda t: t.clone()) if clone_outputs else (lambda t: t)
        Cr = pl["raw"].shape[1]
        raw = cl(pl["raw"]).view(B, T, H, W, Cr)
        res = dict(pred_extrins=cl(outs["pred"]), c2w=cl(outs["c2w"]), raw=raw, centers=raw[..., :3])
        if heads and gs:
            gq = outs["gauss"]
            shp = (B, T, H, W)
            res["gaussians"] = dict(
                cov=cl(gq["cov"]).view(*shp, 3, 3), cov6=cl(gq["cov6"]).view(*shp, 6),
                sh=cl(gq["sh"]).view(*shp, 3, self.m.d_sh), opac=cl(gq["opac"]).view(*shp),
                scales=cl(gq["scales"]).view(*shp, 3), rot=cl(gq["rot"]).view(*shp, 4))
        return res
