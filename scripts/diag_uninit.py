"""Diagnostic: poison the caching allocator's free blocks with NaN before every stage of a training step;
any read of uninitialised memory then shows up as a non-finite gradient."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tests.test_gpu_train_step import _setup
from vicasplat_b200 import rasterizer, train_step as tsm
from vicasplat_b200.rasterizer import RasterOverflow
from vicasplat_b200.train_step import TrainStep

cuda = torch.device("cuda:0")
model, context, target, scenes = _setup(cuda)
model.train()


def poison():
    junk = [torch.full((1 << 28,), float("nan"), device=cuda) for _ in range(8)]          # 8 x 1 GiB
    junk += [torch.full((s,), float("nan"), device=cuda) for s in (1 << 10, 1 << 14, 1 << 17) for _ in range(256)]
    torch.cuda.synchronize()
    del junk


def override(b, gz):
    s = scenes[b]
    return dict(means=s["means"] + gz["means"], cov6=s["cov6"] + gz["cov6"], sh=s["harmonics"] + gz["sh"],
                opac=s["opacities"] + (gz["opac"] - 0.5))


where = sys.argv[1] if len(sys.argv) > 1 else "all"
ts = TrainStep(model, micro_batch=2, camera_weight=0.0)
for _ in range(2):
    try:
        ts.accumulate(context, target, override_gaussians=override)
    except RasterOverflow:
        pass
ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

_rb = tsm.render_backward
_bw = ts.eng.backward
_fw = ts.eng.forward


def rb(*a, **k):
    if where in ("all", "render_bwd"):
        poison()
    return _rb(*a, **k)


def bw(*a, **k):
    if where in ("all", "enc_bwd"):
        poison()
    return _bw(*a, **k)


def fw(*a, **k):
    if where in ("all", "enc_fwd"):
        poison()
    return _fw(*a, **k)


tsm.render_backward = rb
ts.eng.backward = bw
ts.eng.forward = fw
ts.accumulate(context, target, override_gaussians=override)
bad, worst = [], (0.0, "")
for n, p in model.named_parameters():
    if p.grad is None:
        continue
    if not torch.isfinite(p.grad).all():
        bad.append(n)
    else:
        e = ((p.grad - ref[n]).norm() / ref[n].norm().clamp_min(1e-20)).item()
        worst = max(worst, (e, n))
print(f"poison at {where}: {len(bad)} tensors non-finite; first: {bad[:8]}; worst finite difference {worst}")
