"""Generate golden vectors for the encoder path by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python oracle/make_encoder_golden.py            # writes tests/golden/encoder_*.npz

The reference's third-party imports that are absent from the image are stubbed (SURVEY.md
Appendix E); only ``pypose.SO3`` needs behaviour (xyzw Hamilton algebra).  The weights are NOT the
reference's own initialisation: they come from ``oracle.encoder_ref.synth_state_dict`` (a seeded,
machine-independent generator) and are loaded with ``load_state_dict(strict=True)``, which also
pins the 847-key state_dict contract.  Inputs are seeded too, so the tests can regenerate both.
"""
from __future__ import annotations

import functools
import importlib.abc
import importlib.machinery
import inspect
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")

ABSENT = {"matplotlib", "omegaconf", "e3nn", "pytorch3d", "pyquaternion", "dacite", "lightning",
          "pytorch_lightning", "hydra", "lpips", "diff_gaussian_rasterization", "gsplat",
          "skimage", "evo", "plyfile", "moviepy", "plotly", "imageio", "colorama", "svg",
          "colorspacious", "timm", "trimesh", "gradio", "viser", "nerfview", "skvideo"}


def install_stubs():
    class Anything:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return Anything()
        def __getattr__(self, k): return Anything()
        def __class_getitem__(cls, k): return cls
        def __mro_entries__(self, bases): return (Anything,)
        def __or__(self, o): return self
        def __ror__(self, o): return self

    class HollowModule(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (Anything,), {})

    class Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, name, path, target=None):
            if name.split(".")[0] in ABSENT:
                return importlib.machinery.ModuleSpec(name, self, is_package=True)
        def create_module(self, spec):
            m = HollowModule(spec.name)
            m.__path__ = []
            return m
        def exec_module(self, m): pass

    sys.meta_path.insert(0, Finder())

    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    class AttrDict(dict):
        __getattr__ = dict.__getitem__

    class CfgFallback:
        def __getattr__(self, k):
            d = self.__dict__
            if "_cfg" in d and k in d["_cfg"]:
                return d["_cfg"][k]
            return nn.Module.__getattr__(self, k)

    class ConfigMixin(CfgFallback):
        @property
        def config(self):
            return self.__dict__["_cfg"]

    class ModelMixin(CfgFallback, nn.Module):
        pass

    def register_to_config(init):
        sig = inspect.signature(init)

        @functools.wraps(init)
        def wrapped(self, *a, **k):
            ba = sig.bind(self, *a, **k)
            ba.apply_defaults()
            self.__dict__["_cfg"] = AttrDict({n: v for n, v in ba.arguments.items() if n != "self"})
            init(self, *a, **k)
        return wrapped

    module("diffusers")
    module("diffusers.models", ModelMixin=ModelMixin)
    module("diffusers.models.normalization", RMSNorm=nn.Identity)
    module("diffusers.configuration_utils", ConfigMixin=ConfigMixin,
           register_to_config=register_to_config)

    class SO3:
        """xyzw unit-quaternion algebra: the subset of pypose.SO3 that misc/dq.py touches."""
        def __init__(self, t): self.t = t.t if isinstance(t, SO3) else t
        def tensor(self): return self.t
        lshape = property(lambda s: s.t.shape[:-1])
        device = property(lambda s: s.t.device)
        dtype = property(lambda s: s.t.dtype)
        def Inv(self): return SO3(torch.cat([-self.t[..., :3], self.t[..., 3:]], -1))
        def __mul__(self, o):
            if not isinstance(o, SO3):
                return self.t * o
            x1, y1, z1, w1 = self.t.unbind(-1)
            x2, y2, z2, w2 = o.t.unbind(-1)
            return SO3(torch.stack([w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                                    w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                                    w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2,
                                    w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2], -1))
        def __rmul__(self, o): return o * self.t
        def __truediv__(self, o): return self.t / o
        def __neg__(self): return -self.t
        def norm(self, *a, **k): return self.t.norm(*a, **k)
        def __getitem__(self, i): return self.t[i]
        def matrix(self):
            x, y, z, w = self.t.unbind(-1)
            return torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                                2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                                2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
                               -1).reshape(*x.shape, 3, 3)

    def identity_SO3(*shape):
        t = torch.zeros(*shape, 4)
        t[..., 3] = 1
        return SO3(t)

    module("pypose", SO3=SO3, identity_SO3=identity_SO3, LieTensor=SO3)
    module("pypose.lietensor")
    module("pypose.lietensor.lietensor", LieType=object, SO3Type=object)


def build_reference(cfg):
    """cfg: oracle.encoder_ref.EncoderConfig -> the reference's VicaSplat module (eval)."""
    from src.model.encoder.vicasplat import VicaSplat, VicaSplatCfg, OpacityMappingCfg
    from src.model.encoder.common.gaussian_adapter import GaussianAdapterCfg
    bb = dict(img_size=cfg.img_size, patch_size=cfg.patch_size, enc_embed_dim=cfg.enc_embed_dim,
              enc_depth=cfg.enc_depth, enc_num_heads=cfg.enc_num_heads,
              dec_embed_dim=cfg.dec_embed_dim, dec_depth=cfg.dec_depth,
              dec_num_heads=cfg.dec_num_heads, mlp_ratio=cfg.mlp_ratio,
              temporal_rope_theta=cfg.temporal_rope_theta, rope_dim_list=[32, 32],
              use_blocked_causal_attention=True, use_framewise_modulation=True,
              use_cross_neighbor_attention=True, use_intrinsic_embedding=True)
    vc = VicaSplatCfg(name="vicasplat", backbone=bb, visualizer=None,
                      gaussian_adapter=GaussianAdapterCfg(0.005, 0.04, cfg.sh_degree, "softplus"),
                      apply_bounds_shim=True, opacity_mapping=OpacityMappingCfg(0.0, 0.0, 1),
                      predict_opacity=False)
    return VicaSplat(vc).eval()


def synth_inputs(B, T, size, seed=250307):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand((B, T, 3, size, size), generator=g) * 2 - 1
    K = torch.tensor([[0.86, 0, 0.5], [0, 0.86, 0.5], [0, 0, 1.0]]).expand(B, T, 3, 3).clone()
    return image, K


CASES = {
    # name: (EncoderConfig kwargs, B, T, pixel stride of the stored raw_gaussians sample)
    "small": (dict(img_size=64, enc_depth=2, dec_depth=10), 1, 3, 4),
    "full2v": (dict(), 1, 2, 16),
    # the headline shape of BASELINE.json configs[1] (8 views, full depth): the 2 064-key video attention,
    # the 8-row blocked-causal camera mask and the two-segment neighbour attention at full size
    "full8v": (dict(), 1, 8, 16),
    # two DIFFERENT clips in one batch (cross-scene isolation of every attention / modulation table)
    "full4v_b2": (dict(), 2, 4, 32),
}


def main():
    os.chdir("/tmp")                      # keep cwd free of a `curope/` folder (namespace shadowing)
    sys.path.insert(0, str(REF))
    sys.path.insert(0, str(ROOT))
    install_stubs()
    from oracle import encoder_ref as er
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    only = set(sys.argv[1:])                # optional: names of the cases to (re)generate
    for name, (kw, B, T, stride) in CASES.items():
        if only and name not in only:
            continue
        cfg = er.EncoderConfig(**kw)
        sd = er.synth_state_dict(cfg, seed=0)
        model = build_reference(cfg)
        ref_keys = set(model.state_dict().keys())
        assert ref_keys == set(sd.keys()), (sorted(ref_keys - set(sd))[:5], sorted(set(sd) - ref_keys)[:5])
        model.load_state_dict(sd, strict=True)
        image, K = synth_inputs(B, T, cfg.img_size)
        with torch.no_grad():
            vid = image.permute(0, 2, 1, 3, 4)
            _, cam_ext, _, inter = model.backbone(vid, K)
            out = model({"image": image, "intrinsics": K}, compute_viewspace_depth=False)
        raw = out["raw_gaussians"]
        g = out["gaussians"]
        sub = (slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
        data = dict(
            n_keys=np.int64(len(ref_keys)),
            pred_extrins=out["pred_extrins"].numpy(),
            gaussian_camera_extrins=out["gaussian_camera_extrins"].numpy(),
            camera_tokens=cam_ext.numpy(),
            raw_sub=raw[sub].numpy(),
            raw_mean=raw.mean(dim=(0, 1, 2, 3)).numpy(),
            raw_std=raw.std(dim=(0, 1, 2, 3)).numpy(),
            cov_sub=g.covariances[sub].numpy(),
            sh_sub=g.harmonics[sub].numpy(),
            opac_sub=g.opacities[sub].numpy(),
            inter_mean=np.array([t.mean().item() for t in inter]),
            inter_std=np.array([t.std().item() for t in inter]),
            inter_last_row=np.stack([t[0, -1, -1, :64].numpy() for t in inter]),
        )
        np.savez_compressed(out_dir / f"encoder_{name}.npz", **data)
        print(name, "keys", len(ref_keys), "raw", tuple(raw.shape),
              "bytes", (out_dir / f"encoder_{name}.npz").stat().st_size)


if __name__ == "__main__":
    main()
