"""Diagnostic: is LpipsVgg.loss_and_grad bit-reproducible?  Compares repeated calls on the same inputs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops
from vicasplat_b200.lpips import LpipsVgg, TAPS

dev = torch.device("cuda:0")
net = LpipsVgg.stand_in(dev, seed=3)
g = torch.Generator().manual_seed(0)
for n, s in ((4, 64), (2, 64), (12, 256)):
    pred = torch.rand((n, 3, s, s), generator=g).to(dev)
    tgt = torch.rand((n, 3, s, s), generator=g).to(dev)
    ref_l, ref_g = net.loss_and_grad(pred, tgt, 0.5)
    diffs = []
    for _ in range(6):
        junk = torch.full((64 << 20,), float("nan"), device=dev)      # dirty the allocator's free blocks
        del junk
        l, gr = net.loss_and_grad(pred, tgt, 0.5)
        diffs.append(((gr - ref_g).norm() / ref_g.norm()).item())
    print(f"n={n} s={s}: loss {ref_l.item():.6f}; grad rel diff over repeats {['%.1e' % d for d in diffs]}; "
          f"finite {torch.isfinite(ref_g).all().item()}")
    # per stage: features twice
    a1, p1 = net._features(pred, keep=True)
    a2, p2 = net._features(pred, keep=True)
    bad = [i for i, (x, y) in enumerate(zip(a1, a2)) if not torch.equal(x, y)]
    print("   feature maps differing between two forward passes:", bad)
