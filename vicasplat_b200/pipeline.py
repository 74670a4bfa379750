"""Steady-state scene pipeline: host clips in, host renders + poses out.

The reference evaluates one batch at a time (``ModelWrapper.test_step``, src/model/model_wrapper.py:
207-333: ``encoder(context)`` -> ``decoder.forward(gaussians, target cameras)`` -> ``.cpu()``), so the
host<->device copies of a batch sit in series with its compute.  ``ScenePipeline`` keeps the same two
plugin calls but runs consecutive batches through ``depth`` slots on three CUDA streams:

    copy-in stream   pinned host clip  -> slot image / intrinsics           (H2D of batch i+1)
    compute stream   VicaSplat.forward -> DecoderSplattingCUDA.forward      (batch i)
    copy-out stream  slot colour / depth / poses -> pinned host buffers     (D2H of batch i-1)

Ordering is by events only; the single host synchronisation is ``Ticket.result()``, which waits for
the copy-out event of that batch.  Nothing here touches the numerics: the same kernels run in the
same order per batch as in the sequential calls (tests/test_gpu_pipeline.py compares them bit for bit).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

import torch
from torch import Tensor

from .rasterizer import take_deferred


@dataclass
class Ticket:
    """Handle of one submitted batch; ``result()`` blocks until its outputs are in host memory."""
    slot: "_Slot"
    seq: int

    def result(self) -> Dict[str, Tensor]:
        if self.slot.seq != self.seq:
            raise RuntimeError("ScenePipeline: this ticket's slot was reused; call result() before "
                               "submitting `depth` more batches")
        self.slot.out_done.synchronize()
        if self.slot.records:
            from .rasterizer import verify_deferred
            verify_deferred(self.slot.counts_h[:len(self.slot.records)], self.slot.records)
        return dict(color=self.slot.color_h, depth=self.slot.depth_h, pred_extrins=self.slot.pose_h)


class _Slot:
    def __init__(self):
        self.seq = -1
        self.shape = None
        self.records = []
        self.in_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.out_done = torch.cuda.Event()

    def ensure(self, dev, B, T, H, W, V, h, w):
        shape = (B, T, H, W, V, h, w)
        if self.shape == shape:
            return
        f32 = dict(dtype=torch.float32, device=dev)
        self.image_d = torch.empty((B, T, 3, H, W), **f32)
        self.K_d = torch.empty((B, T, 3, 3), **f32)
        self.color_d = torch.empty((B, V, 3, h, w), **f32)
        self.depth_d = torch.empty((B, V, h, w), **f32)
        self.pose_d = torch.empty((B, T - 1, 8), **f32)
        self.color_h = torch.empty((B, V, 3, h, w), dtype=torch.float32).pin_memory()
        self.depth_h = torch.empty((B, V, h, w), dtype=torch.float32).pin_memory()
        self.pose_h = torch.empty((B, T - 1, 8), dtype=torch.float32).pin_memory()
        self.counts_d = torch.zeros((B, 2), dtype=torch.int64, device=dev)
        self.counts_h = torch.zeros((B, 2), dtype=torch.int64).pin_memory()
        self.shape = shape


class ScenePipeline:
    """encoder -> decoder over a stream of host batches with copy / compute overlap.

    ``submit(context, target)``:
      context  {"image": (B,T,3,H,W) float32 host tensor in [-1,1] (pinned for async copies),
                "intrinsics": (B,T,3,3)}  -- the encoder plugin's ``context`` dict
      target   {"extrinsics": (B,V,4,4), "intrinsics": (B,V,3,3), "near": (B,V), "far": (B,V),
                "image_shape": (h, w)} device tensors -- the decoder plugin's camera arguments;
                optional "gaussians": render these instead of the encoder's prediction
    """

    def __init__(self, model, decoder, depth: int = 2):
        assert depth >= 2, "at least two slots are needed to overlap anything"
        self.model, self.decoder = model, decoder
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("ScenePipeline runs on CUDA only (no CPU fallback)")
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_compute = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self.slots: List[_Slot] = [_Slot() for _ in range(depth)]
        self.seq = 0
        # the first batches of a shape calibrate the binning capacity hint with the checked
        # (synchronising) render; later ones run sync-free and are verified in Ticket.result()
        self._calibrated = set()
        self._check = True

    @torch.no_grad()
    def submit(self, context: Dict[str, Tensor], target: dict) -> Ticket:
        image, K = context["image"], context["intrinsics"]
        B, T, _, H, W = image.shape
        V = target["extrinsics"].shape[1]
        h, w = target["image_shape"]
        slot = self.slots[self.seq % len(self.slots)]
        # the slot's previous batch must have left the device before its buffers are overwritten
        slot.out_done.synchronize()
        slot.ensure(self.dev, B, T, H, W, V, h, w)
        slot.seq = self.seq
        key = (B, V, h, w, T * H * W if "gaussians" not in target else int(target["gaussians"].means[0].numel()) // 3)
        self._check = True if key not in self._calibrated else "deferred"
        self._calibrated.add(key)
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_stream(cur)                 # whatever produced `target` on the caller's stream
            slot.image_d.copy_(image, non_blocking=True)
            slot.K_d.copy_(K, non_blocking=True)
            slot.in_done.record(self.s_in)
        with torch.cuda.stream(self.s_compute):
            self.s_compute.wait_event(slot.in_done)
            enc = self.model({"image": slot.image_d, "intrinsics": slot.K_d},
                             compute_viewspace_depth=False, clone_outputs=False)
            g = target.get("gaussians") or enc["gaussians"]
            take_deferred()                            # drop records of renders that were not ours
            out = self.decoder.forward(g, target["extrinsics"], target["intrinsics"], target["near"],
                                       target["far"], (h, w), check_overflow=self._check)
            slot.records = take_deferred()
            if slot.records:
                torch.stack([r[0] for r in slot.records], out=slot.counts_d[:len(slot.records)])
            slot.color_d.copy_(out.color)
            slot.depth_d.copy_(out.depth)
            slot.pose_d.copy_(enc["pred_extrins"])
            slot.compute_done.record(self.s_compute)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(slot.compute_done)
            slot.color_h.copy_(slot.color_d, non_blocking=True)
            slot.depth_h.copy_(slot.depth_d, non_blocking=True)
            slot.pose_h.copy_(slot.pose_d, non_blocking=True)
            slot.counts_h.copy_(slot.counts_d, non_blocking=True)
            slot.out_done.record(self.s_out)
        self.seq += 1
        return Ticket(slot, slot.seq)

    def drain(self) -> None:
        for s in self.slots:
            s.out_done.synchronize()
