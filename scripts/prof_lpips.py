"""LPIPS value + gradient sweep alone: time per call at 12 images (one scene) and 96 images (a micro-batch),
with the per-family CUDA-event breakdown of vicasplat_b200.ops.TIMERS."""
import os
import sys
from collections import defaultdict
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops
from vicasplat_b200.lpips import LpipsVgg

dev = torch.device("cuda:0")
net = LpipsVgg.stand_in(dev)
g = torch.Generator().manual_seed(0)
for n in (12, 96):
    pred = torch.rand((n, 3, 256, 256), generator=g).to(dev)
    tgt = torch.rand((n, 3, 256, 256), generator=g).to(dev)
    for _ in range(3):
        net.loss_and_grad(pred, tgt, 0.05)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        net.loss_and_grad(pred, tgt, 0.05)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if os.environ.get("VS_PROFILE_STEP") == "1" and n == 96:
        torch.cuda.cudart().cudaProfilerStart()
        net.loss_and_grad(pred, tgt, 0.05)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    ops.TIMERS = {}
    net.loss_and_grad(pred, tgt, 0.05)
    fam = ops.family_ms(ops.TIMERS)
    gem = defaultdict(lambda: [0.0, 0])
    for a, b, meta in ops.TIMERS.get("gemm", []):
        gem[meta][0] += a.elapsed_time(b)
        gem[meta][1] += 1
    ops.TIMERS = None
    print(f"images {n}: {ms:.3f} ms per call = {ms / n * 12:.3f} ms per 12-view scene; families {fam}")
    for m, v in sorted(gem.items(), key=lambda kv: -kv[1][0])[:14]:
        tf = 2.0 * m[1] * m[2] * m[3] * v[1] / (v[0] * 1e-3) / 1e12
        print(f"   {v[0]:7.3f} ms n={v[1]} {tf:7.1f} TF/s {m}")
