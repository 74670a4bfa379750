"""Fused multi-tensor AdamW with global-norm clipping and a non-finite scan.

Mirrors what the reference's training step does around ``optimizer.step()``:
``torch.optim.AdamW(param_dicts, lr, weight_decay=0.05, betas=(0.9, 0.95))`` with per-group learning
rates (src/model/model_wrapper.py:884-951) and Lightning's ``gradient_clip_val: 0.5`` (global L2 norm,
config/main.yaml:70) -- 847 small launches and a host-visible ``.any()`` per tensor there, two launches
and no host synchronisation here (``vs_adamw_step``).  State and parameters stay ordinary torch
tensors, so ``state_dict`` round-trips with ``torch.optim.AdamW``'s layout.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List

import torch

from . import _lib
from ._lib import AdamWParams, check, ptr, stream_ptr

CHUNK = 16384    # VS_ADAMW_CHUNK


class FusedAdamW:
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: float = 0.0, skip_nonfinite: bool = True):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups: List[dict] = []
        for g in groups:
            g = dict(g)
            g["params"] = [p for p in g["params"] if p.requires_grad]
            g.setdefault("lr", lr)
            self.param_groups.append(g)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.max_grad_norm, self.skip_nonfinite = max_grad_norm, skip_nonfinite
        self.step_count = 0
        self._params = [p for g in self.param_groups for p in g["params"]]
        for p in self._params:
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters (no CPU fallback)")
        self.dev = self._params[0].device
        self.state = {p: dict(exp_avg=torch.zeros_like(p), exp_avg_sq=torch.zeros_like(p)) for p in self._params}
        i64 = dict(dtype=torch.int64, device=self.dev)
        self._t_params = torch.tensor([p.data_ptr() for p in self._params], **i64)
        self._t_m = torch.tensor([self.state[p]["exp_avg"].data_ptr() for p in self._params], **i64)
        self._t_v = torch.tensor([self.state[p]["exp_avg_sq"].data_ptr() for p in self._params], **i64)
        self._t_sizes = torch.tensor([p.numel() for p in self._params], **i64)
        ct, cs = [], []
        for i, p in enumerate(self._params):
            for s in range(0, p.numel(), CHUNK):
                ct.append(i); cs.append(s)
        self._t_ct = torch.tensor(ct, dtype=torch.int32, device=self.dev)
        self._t_cs = torch.tensor(cs, **i64)
        self._partials = torch.zeros((len(ct) + 2,), dtype=torch.float32, device=self.dev)
        self._counter = torch.zeros((1,), dtype=torch.int32, device=self.dev)
        self.grad_norm = torch.zeros((), dtype=torch.float32, device=self.dev)
        self.found_inf = torch.zeros((), dtype=torch.int32, device=self.dev)
        self._t_grads = torch.zeros((len(self._params),), **i64)
        self._grad_ptrs = None
        self._lrs_host = None
        self._t_lrs = torch.zeros((len(self._params),), dtype=torch.float32, device=self.dev)

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p in self._params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    @torch.no_grad()
    def step(self) -> None:
        """One AdamW step on every parameter (all must have a gradient).  ``self.grad_norm`` (before
        clipping) and ``self.found_inf`` are device scalars: reading them is the caller's sync."""
        lib = _lib.load()
        gp = []
        for p in self._params:
            if p.grad is None:
                raise RuntimeError("FusedAdamW.step: a parameter has no gradient (feed zeros for unused ones)")
            g = p.grad
            if not (g.is_contiguous() and g.dtype == torch.float32):
                raise RuntimeError("FusedAdamW.step: gradients must be contiguous fp32")
            gp.append(g.data_ptr())
        if gp != self._grad_ptrs:                       # gradient buffers normally persist across steps
            self._t_grads.copy_(torch.tensor(gp, dtype=torch.int64), non_blocking=True)
            self._grad_ptrs = gp
        lrs = [float(g["lr"]) for g in self.param_groups for _ in g["params"]]
        if lrs != self._lrs_host:                       # schedulers change group["lr"] between steps
            self._t_lrs.copy_(torch.tensor(lrs, dtype=torch.float32), non_blocking=True)
            self._lrs_host = lrs
        self.step_count += 1
        q = AdamWParams()
        q.n_tensors, q.n_chunks = len(self._params), self._t_ct.numel()
        q.params, q.grads, q.exp_avg, q.exp_avg_sq = ptr(self._t_params), ptr(self._t_grads), ptr(self._t_m), ptr(self._t_v)
        q.sizes, q.lrs, q.chunk_tensor, q.chunk_start = ptr(self._t_sizes), ptr(self._t_lrs), ptr(self._t_ct), ptr(self._t_cs)
        q.beta1, q.beta2, q.eps, q.weight_decay = self.betas[0], self.betas[1], self.eps, self.weight_decay
        q.step, q.max_grad_norm, q.skip_nonfinite = self.step_count, self.max_grad_norm, int(self.skip_nonfinite)
        q.partials, q.counter = ptr(self._partials), ptr(self._counter)
        q.grad_norm_out, q.found_inf_out = ptr(self.grad_norm), ptr(self.found_inf)
        check(lib.vs_adamw_step(C.byref(q), C.c_void_p(stream_ptr())), "vs_adamw_step")
