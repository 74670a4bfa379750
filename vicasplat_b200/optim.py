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


_NONFINITE = {"keep": 0, "skip": 1, "sanitize": 2}


class FusedAdamW:
    """nonfinite: what a non-finite gradient element does to the step --
      "sanitize" (default, the reference: GradientNanCheckCallback applies torch.nan_to_num_ to the
                 gradients and the optimizer still steps, src/main.py:40-45);
      "skip"     drop the whole step (``found_inf`` tells; the step counter does not advance);
      "keep"     use the gradients as they are (torch.optim.AdamW alone)."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: float = 0.0, nonfinite: str = "sanitize"):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        if nonfinite not in _NONFINITE:
            raise ValueError(f"FusedAdamW: nonfinite must be one of {sorted(_NONFINITE)}")
        self.param_groups: List[dict] = []
        for g in groups:
            g = dict(g)
            g["params"] = [p for p in g["params"] if p.requires_grad]
            g.setdefault("lr", lr)
            # one weight decay / betas / eps for all tensors goes to the kernel: a group that asks for
            # different ones (e.g. a no-decay group) must not be silently decayed with the global value
            for key, glob in (("weight_decay", weight_decay), ("betas", tuple(betas)), ("eps", eps)):
                if key in g and (tuple(g[key]) if key == "betas" else g[key]) != glob:
                    raise ValueError(f"FusedAdamW: per-group {key}={g[key]!r} differs from the optimizer's "
                                     f"{glob!r}; only 'lr' may vary between groups")
            self.param_groups.append(g)
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.max_grad_norm, self.nonfinite = max_grad_norm, nonfinite
        self._params = [p for g in self.param_groups for p in g["params"]]
        if not self._params:
            raise ValueError("FusedAdamW: no trainable parameters")
        for p in self._params:
            if not (p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("FusedAdamW needs contiguous fp32 parameters")
        self.dev = self._params[0].device       # CPU parameters can hold / exchange state; step() needs CUDA
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
        # the step counter lives on the device and advances only when an update is applied
        self._step = torch.zeros((), dtype=torch.int32, device=self.dev)
        self._t_grads = torch.zeros((len(self._params),), **i64)
        self._grad_ptrs = None
        self._lrs_host = None
        self._t_lrs = torch.zeros((len(self._params),), dtype=torch.float32, device=self.dev)

    @property
    def step_count(self) -> int:
        """Number of APPLIED updates (reads the device counter: a host synchronisation)."""
        return int(self._step.item())

    @step_count.setter
    def step_count(self, v: int) -> None:
        self._step.fill_(int(v))

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
        if self.dev.type != "cuda":
            raise RuntimeError("FusedAdamW.step needs CUDA parameters (there is no CPU fallback)")
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
        q = AdamWParams()
        q.n_tensors, q.n_chunks = len(self._params), self._t_ct.numel()
        q.params, q.grads, q.exp_avg, q.exp_avg_sq = ptr(self._t_params), ptr(self._t_grads), ptr(self._t_m), ptr(self._t_v)
        q.sizes, q.lrs, q.chunk_tensor, q.chunk_start = ptr(self._t_sizes), ptr(self._t_lrs), ptr(self._t_ct), ptr(self._t_cs)
        q.beta1, q.beta2, q.eps, q.weight_decay = self.betas[0], self.betas[1], self.eps, self.weight_decay
        q.step, q.max_grad_norm, q.skip_nonfinite = 0, self.max_grad_norm, _NONFINITE[self.nonfinite]
        q.partials, q.counter = ptr(self._partials), ptr(self._counter)
        q.grad_norm_out, q.found_inf_out = ptr(self.grad_norm), ptr(self.found_inf)
        q.step_counter = ptr(self._step)
        check(lib.vs_adamw_step(C.byref(q), C.c_void_p(stream_ptr())), "vs_adamw_step")

    # ---- checkpoint / resume in torch.optim.AdamW's layout (Lightning saves ``optimizer.state_dict()``
    # with every checkpoint of the reference: src/main.py:80-99, resumed through ``ckpt_path``)
    def state_dict(self) -> dict:
        state, groups, idx = {}, [], 0
        step = float(self.step_count)
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                st = self.state[p]
                state[idx] = {"step": torch.tensor(step),
                              "exp_avg": st["exp_avg"], "exp_avg_sq": st["exp_avg_sq"]}
                ids.append(idx)
                idx += 1
            meta = {k: v for k, v in g.items() if k != "params"}
            meta.setdefault("betas", tuple(self.betas))
            meta.setdefault("eps", self.eps)
            meta.setdefault("weight_decay", self.weight_decay)
            groups.append({**meta, "params": ids})
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, sd: dict) -> None:
        """Accepts ``torch.optim.AdamW.state_dict()`` (or this class's).  Moments are copied INTO the
        existing buffers (the device tables keep pointing at them); learning rates follow the saved
        groups; an empty state (an optimizer that never stepped) resets the step counter."""
        groups = sd["param_groups"]
        if len(groups) != len(self.param_groups) or any(
                len(a["params"]) != len(b["params"]) for a, b in zip(groups, self.param_groups)):
            raise ValueError("FusedAdamW.load_state_dict: parameter groups do not match")
        steps = set()
        for saved, mine in zip(groups, self.param_groups):
            for key, glob in (("weight_decay", self.weight_decay), ("betas", tuple(self.betas)), ("eps", self.eps)):
                if key in saved and (tuple(saved[key]) if key == "betas" else saved[key]) != glob:
                    raise ValueError(f"FusedAdamW.load_state_dict: saved group has {key}={saved[key]!r}, this "
                                     f"optimizer runs {glob!r} for all tensors")
            if "lr" in saved:
                mine["lr"] = saved["lr"]
            for i, p in zip(saved["params"], mine["params"]):
                st = sd["state"].get(i)
                if st is None:
                    self.state[p]["exp_avg"].zero_()
                    self.state[p]["exp_avg_sq"].zero_()
                    steps.add(0)
                    continue
                if st["exp_avg"].shape != p.shape:
                    raise ValueError(f"FusedAdamW.load_state_dict: state {i} has shape {tuple(st['exp_avg'].shape)}, "
                                     f"parameter has {tuple(p.shape)}")
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                steps.add(int(round(float(st["step"]))))
        if len(steps) > 1:
            raise ValueError(f"FusedAdamW.load_state_dict: parameters disagree on the step count {sorted(steps)} "
                             "(one bias correction is shared by all tensors)")
        self.step_count = steps.pop() if steps else 0
        self._lrs_host = None
