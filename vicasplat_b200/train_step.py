"""One optimisation step of the reference's training loop on the sm_100a kernels (BASELINE.json configs[2]
/ [3]): ``ModelWrapper.training_step`` (src/model/model_wrapper.py:184-321: encoder -> decoder.forward ->
losses) + backward + DDP gradient all-reduce (src/main.py:110-115) + nan_to_num / clip 0.5 / AdamW
(src/main.py:40-45, config/main.yaml:70, model_wrapper.py:884-951), without the autograd graph:

    per micro-batch of scenes
        TrainEngine.forward            encoder: clips -> per-pixel Gaussians + poses (activations kept)
        per scene   render_forward     V target views of the scene's Gaussians       (scenes on SceneStreams)
        vs_mse_loss per scene, LPIPS   loss values + dL/dcolour (LossMse, loss_mse.py:23-31; LossLpips as one
                                       sweep over the micro-batch's images); camera loss on the predicted poses
        per scene   render_backward    -> this scene's slice of d means / d cov6 / d SH / d opacity
        TrainEngine.backward           -> parameter gradients (accumulated over the micro-batches); on the
                                       LAST micro-batch every bucket is all-reduced as soon as it is final
    FusedAdamW.step, TrainEngine.repack

Gradient accumulation over micro-batches is exact (the wgrad GEMMs accumulate in fp32), so a per-GPU
batch larger than what the kept activations allow (~6 GB per scene at 8 views) costs no accuracy.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import decoder as dec, ops
from .encoder import VicaSplat
from .encoder_train import GradReducer
from .optim import FusedAdamW
from .rasterizer import SceneStreams, render_backward, render_forward, take_deferred, verify_deferred, _deferred
from .train import TrainEngine


class _NoReduce:
    """stands in for the reducer on all but the last micro-batch"""
    world = 1

    def bucket_ready(self, bucket):
        pass

    def finish(self):
        pass


class TrainStep:
    def __init__(self, model: VicaSplat, *, lr: float = 4e-5, backbone_lr_multiplier: float = 0.25,
                 new_param_keywords=("gaussian_param_head", "intrinsic_encoder"), weight_decay: float = 0.05,
                 betas=(0.9, 0.95), max_grad_norm: float = 0.5, micro_batch: int = 8, mse_weight: float = 1.0,
                 camera_weight: float = 0.1, lpips=None, lpips_weight: float = 0.05,
                 reducer: Optional[GradReducer] = None, background=(0.0, 0.0, 0.0)):
        """lr / backbone_lr_multiplier / new_param_keywords: the reference's two parameter groups
        (model_wrapper.py:884-927 with config/experiment/re10k_8view.yaml:48-55: parameters whose name
        contains a keyword train at lr, the pretrained rest at lr * multiplier).  Losses: MSE on the renders
        (config/loss/mse.yaml, weight 1) and -- when the context carries ground-truth extrinsics -- the
        dual-quaternion camera loss (config/loss/camera.yaml, weight 0.1); LPIPS (config/loss/lpips.yaml,
        weight 0.05) when an ``lpips.LpipsVgg`` is given -- the `lpips` package's VGG16 + linear-layer weights
        are not in this image, ``LpipsVgg.stand_in`` is a random network of the same cost."""
        self.model = model
        self.reducer = reducer or GradReducer()
        self.eng = TrainEngine(model, reducer=self.reducer)
        self.micro_batch = micro_batch
        self.mse_weight = mse_weight
        self.camera_weight = camera_weight
        self.lpips, self.lpips_weight = lpips, lpips_weight
        named = [(n, p) for n, p in model.named_parameters() if n not in set(self.eng.unused)]
        is_new = lambda n: any(k in n for k in new_param_keywords)
        new = [p for n, p in named if is_new(n)]
        old = [p for n, p in named if not is_new(n)]
        self.opt = FusedAdamW([{"params": new, "lr": lr}, {"params": old, "lr": lr * backbone_lr_multiplier}],
                              lr=lr, betas=betas, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        self.bg = torch.tensor(background, dtype=torch.float32, device=self.eng.dev)
        self.last_loss = None

    @torch.no_grad()
    def step(self, context: dict, target: dict, override_gaussians=None, check_overflow=True) -> torch.Tensor:
        """Gradients of the batch (``accumulate``), then nan_to_num / clip / AdamW and the operand re-pack.
        Returns the mean loss (device scalar)."""
        loss = self.accumulate(context, target, override_gaussians, check_overflow)
        self.opt.step()
        self.eng.repack()
        return loss

    def _render(self, b, j, out, override_gaussians, view_t, full_t, campos, tanfov, check_overflow, V, Gs, H, W):
        """render scene b (slot j of the micro-batch): (colour (V,3,H,W), forward state)"""
        g = slice(j * Gs, (j + 1) * Gs)
        gauss = dict(means=out["means"][g], cov6=out["cov6"][g], sh=out["sh"][g], opac=out["opac"][g])
        if override_gaussians is not None:
            gauss = override_gaussians(b, gauss)
        cam = slice(b * V, (b + 1) * V)
        color, _depth, _alpha, st = render_forward(
            gauss["means"], gauss["cov6"], gauss["opac"], gauss["sh"], sh_degree=4, sh_layout="chan_major",
            viewmatrix=view_t[cam], projmatrix=full_t[cam], campos=campos[cam], tanfov=tanfov[cam],
            bg=self.bg, H=H, W=W)
        if check_overflow:
            _deferred.append((st.num_pairs, st.max_pairs, st.max_tile, (V, Gs, H, W)))
        return color, st

    @torch.no_grad()
    def accumulate(self, context: dict, target: dict, override_gaussians=None, check_overflow=True) -> torch.Tensor:
        """Forward + backward of one batch: every ``param.grad`` holds the gradient of the mean loss
        (averaged over the ranks when a process group is up).
        context: image (B,T,3,H,W) in [-1,1], intrinsics (B,T,3,3); target: extrinsics (B,V,4,4),
        intrinsics (B,V,3,3), near / far (B,V), image (B,V,3,H,W).
        override_gaussians(scene_index, dict of this scene's (G, ...) views) -> dict: lets a synthetic
        benchmark place the splats (see bench.py); the gradient still flows to the encoder's outputs."""
        eng = self.eng
        image, intr = context["image"], context["intrinsics"]
        B, T, _, H, W = image.shape
        V = target["extrinsics"].shape[1]
        Gs = T * H * W
        mb = min(self.micro_batch, B)
        assert B % mb == 0, "the batch must be a multiple of the micro-batch"
        tanfov, view_t, full_t, campos = dec._cameras(target["extrinsics"].flatten(0, 1), target["intrinsics"].flatten(0, 1),
                                                      target["near"].flatten(), target["far"].flatten())
        losses = []
        n_micro = B // mb
        real = eng.reducer
        for mi in range(n_micro):
            sl = slice(mi * mb, (mi + 1) * mb)
            out = eng.forward(image[sl], intr[sl])
            G = mb * Gs
            dev = out["raw"].device
            z = lambda *s: torch.zeros((G, *s), dtype=torch.float32, device=dev)
            d_means, d_cov6, d_sh, d_opac = z(3), z(6), z(3, self.model.d_sh), z()
            # the scenes' render and render-backward chains are independent: round-robin over SceneStreams;
            # the losses sit between them on the current stream, LPIPS as ONE sweep over the micro-batch's
            # mb * V images (its deep layers have too few rows per scene to fill the GPU)
            colors, states = [None] * mb, [None] * mb
            with SceneStreams(dev) as ss:
                for j in range(mb):
                    with ss.scene(j):
                        colors[j], states[j] = self._render(mi * mb + j, j, out, override_gaussians, view_t, full_t,
                                                            campos, tanfov, check_overflow, V, Gs, H, W)
                    ss.keep(colors[j])
            g_colors = []
            for j in range(mb):
                loss, g_color = ops.mse_loss(colors[j], target["image"][mi * mb + j], self.mse_weight / B)
                losses.append(loss)
                g_colors.append(g_color)
            if self.lpips is not None and self.lpips_weight > 0:
                ll, gl = self.lpips.loss_and_grad(torch.cat(colors), target["image"][sl].flatten(0, 1),
                                                  self.lpips_weight * mb / B)
                losses.append(ll)
                for j in range(mb):
                    g_colors[j].add_(gl[j * V:(j + 1) * V])
            with SceneStreams(dev) as ss:
                for j in range(mb):
                    g = slice(j * Gs, (j + 1) * Gs)
                    with ss.scene(j):
                        render_backward(states[j], g_colors[j],
                                        out=dict(d_means=d_means[g], d_cov6=d_cov6[g], d_opac=d_opac[g],
                                                 d_sh=d_sh[g].view(Gs, -1)), want_tau=False)
            del colors, states, g_colors
            d_pred = None
            if self.camera_weight > 0 and "extrinsics" in context:
                from .loss import camera_loss
                with torch.enable_grad():      # (mb, T-1, 8) numbers: host-side torch arithmetic (loss.py)
                    pred = out["pred_extrins"].detach().requires_grad_(True)
                    lc = camera_loss(pred, context["extrinsics"][sl], self.camera_weight * mb / B)
                    lc.backward()
                d_pred = pred.grad
                losses.append(lc.detach())
            last = mi == n_micro - 1
            eng.reducer = real if last else _NoReduce()
            eng.backward(d_means=d_means, d_cov6=d_cov6, d_sh=d_sh, d_opac=d_opac, d_pred=d_pred, zero=(mi == 0))
        eng.reducer = real
        if check_overflow:      # the counters travel with the loss: one synchronisation per step
            recs = take_deferred()
            counts = torch.stack([r[0] for r in recs]).cpu()
            verify_deferred(counts, recs)
        self.last_loss = torch.stack(losses).sum()
        return self.last_loss
