"""Diagnostic: gradients of a 2-scene batch as one micro-batch vs two (tests/test_gpu_train_step.py), per tensor."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tests.test_gpu_train_step import _setup
from vicasplat_b200.lpips import LpipsVgg
from vicasplat_b200.rasterizer import RasterOverflow
from vicasplat_b200.train_step import TrainStep

cuda = torch.device("cuda:0")
model, context, target, scenes = _setup(cuda)
model.train()
B, T = context["image"].shape[:2]
ext = torch.eye(4, device=cuda).repeat(B, T, 1, 1)
ext[:, 1:, :3, 3] = 0.1 * torch.randn((B, T - 1, 3), generator=torch.Generator().manual_seed(2)).to(cuda)
context = dict(context, extrinsics=ext)
use_lpips = "nolpips" not in sys.argv
net = LpipsVgg.stand_in(cuda, seed=3) if use_lpips else None


def override(b, gz):
    s = scenes[b]
    return dict(means=s["means"] + gz["means"], cov6=s["cov6"] + gz["cov6"], sh=s["harmonics"] + gz["sh"],
                opac=s["opacities"] + (gz["opac"] - 0.5))


LOG = []
if net is not None:
    _orig = net.loss_and_grad

    def _logged(pred, tgt, w, want_grad=True):
        l, gr = _orig(pred, tgt, w, want_grad)
        LOG.append((pred.double().sum().item(), pred.double().pow(2).sum().item(), tgt.double().sum().item(),
                    gr.double().sum().item(), gr.double().pow(2).sum().item(), l.item()))
        return l, gr
    net.loss_and_grad = _logged


BLOG = []


def grads(mb):
    ts = TrainStep(model, micro_batch=mb, lpips=net, lpips_weight=0.5, camera_weight=0.1)
    _bw = ts.eng.backward

    def bw(**k):
        BLOG.append(tuple(k[n].double().pow(2).sum().item() if k.get(n) is not None else 0.0
                          for n in ("d_means", "d_cov6", "d_sh", "d_opac", "d_pred")))
        return _bw(**k)
    ts.eng.backward = bw
    tries = 0
    for _ in range(3):
        try:
            loss = ts.accumulate(context, target, override_gaussians=override)
            break
        except RasterOverflow:
            tries += 1
    return loss.item(), tries, {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}


for rep in range(4):
    l2, t2, g2 = grads(2)
    l1, t1, g1 = grads(1)
    l2b, t2b, g2b = grads(2)
    err = sorted(((((g1[n] - g2[n]).norm() / g2[n].norm().clamp_min(1e-20)).item(), n) for n in g2), reverse=True)
    err_same = max(((g2b[n] - g2[n]).norm() / g2[n].norm().clamp_min(1e-20)).item() for n in g2)
    if err_same > 1e-4:
        import re
        from collections import defaultdict
        grp = defaultdict(list)
        for n in g2:
            e = ((g2b[n] - g2[n]).norm() / g2[n].norm().clamp_min(1e-20)).item()
            m = re.match(r"(backbone\.(?:enc|dec)_blocks\.\d+|[a-z_0-9]+\.dpt\.[a-z_0-9]+(?:\.[a-z_0-9]+)?|[a-z_]+)", n)
            grp[m.group(1) if m else n].append((e, n))
        def order(n):
            m = re.match(r"backbone\.(enc|dec)_blocks\.(\d+)\.", n)
            if m:
                return (1 if m.group(1) == "dec" else 2, -int(m.group(2)), n)
            return (0 if not n.startswith("backbone") else 3, 0, n)
        rows = sorted(((order(n), ((g2b[n] - g2[n]).norm() / g2[n].norm().clamp_min(1e-20)).item()) for n in g2))
        first = next(i for i, (k, e) in enumerate(rows) if e > 1e-6 and k[0] in (1, 2))
        print("   first tensors (backward order) around the point where two identical runs start to differ:")
        for k, e in rows[max(0, first - 12):first + 40]:
            print(f"        {k[2]:60s} {e:.1e}")
        print("   per-group MIN error between two identical mb=2 runs (backward order is heads -> dec 11..0 -> enc 23..0):")
        for k in grp:
            v = sorted(grp[k])
            print(f"      {k:55s} min {v[0][0]:.1e} ({v[0][1].split('.')[-2]}.{v[0][1].split('.')[-1]})  median {v[len(v)//2][0]:.1e}  max {v[-1][0]:.1e}")
    print("   lpips calls (sum pred, sum pred^2, sum tgt, sum grad, sum grad^2, loss):")
    for rec in LOG:
        print("    ", " ".join(f"{x:.10e}" for x in rec))
    LOG.clear()
    print("   eng.backward inputs (sum of squares of d_means, d_cov6, d_sh, d_opac, d_pred):")
    for rec in BLOG:
        print("    ", " ".join(f"{x:.10e}" for x in rec))
    BLOG.clear()
    print(f"rep {rep}: retries {t2}/{t1}/{t2b}; mb2 vs mb1 worst {err[0][0]:.2e} ({err[0][1]}), 2nd {err[1][0]:.2e} ({err[1][1]}), "
          f"median {err[len(err) // 2][0]:.2e}; mb2 vs mb2 again {err_same:.2e}")
