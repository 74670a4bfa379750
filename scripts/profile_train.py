#!/usr/bin/env python
"""Where the training step's time goes: CUDA events around the sections of TrainStep.step and around
every launch family (vicasplat_b200.ops.TIMERS), plus the GEMM launches ranked by time.
usage: profile_train.py [scenes_per_micro_batch]"""
import json
import os
import sys
from collections import defaultdict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from vicasplat_b200 import decoder as dec, ops as vops, synthetic  # noqa: E402
from vicasplat_b200.encoder import VicaSplat  # noqa: E402
from vicasplat_b200.rasterizer import RasterOverflow  # noqa: E402
from vicasplat_b200.train import TrainEngine  # noqa: E402
from vicasplat_b200.train_step import TrainStep  # noqa: E402

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T, V, S = 8, 12, 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VicaSplat().to(dev)
with torch.no_grad():
    for n, p in model.named_parameters():
        if "modulation" in n or n.startswith("camera_extrinsic_head"):
            p.normal_(0, 0.02)
ts = TrainStep(model, micro_batch=NB)
image, K = synthetic.clip(NB, T, S)
ctx = dict(image=image.to(dev), intrinsics=K.to(dev))
scenes = []
for i in range(NB):
    sc = synthetic.gaussian_scene(T, S, S, V, seed=1 + i, device=dev)
    sc["cov6"] = dec._cov6(sc["covariances"]).contiguous()
    scenes.append(sc)
stk = lambda k: torch.stack([s[k] for s in scenes])
target = dict(extrinsics=stk("extrinsics"), intrinsics=stk("intrinsics"), near=stk("near"), far=stk("far"),
              image=torch.rand((NB, V, 3, S, S), device=dev))


def override(b, g):
    s = scenes[b]
    return dict(means=s["means"] + g["means"], cov6=s["cov6"] + g["cov6"], sh=s["harmonics"] + g["sh"],
                opac=s["opacities"] + (g["opac"] - 0.5))


for _ in range(3):
    try:
        ts.step(ctx, target, override_gaussians=override)
    except RasterOverflow:
        pass
torch.cuda.synchronize()

if os.environ.get("VS_PROFILE_STEP") == "1":
    # one training step between cudaProfilerStart/Stop for
    #   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none ...
    torch.cuda.cudart().cudaProfilerStart()
    ts.step(ctx, target, override_gaussians=override)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("profiled one training step")
    sys.exit(0)

# ---- sections (monkey-patched event marks around the engine's stages)
marks = []


def mark(name):
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    marks.append((name, e))


eng = ts.eng
for name in ("_encoder_fwd", "_decoder_fwd", "_heads_fwd", "_heads_bwd", "_decoder_bwd", "_encoder_bwd"):
    fn = getattr(eng, name)

    def wrap(*a, _fn=fn, _n=name, **k):
        mark("(before) " + _n)
        r = _fn(*a, **k)
        mark(_n)
        return r
    setattr(eng, name, wrap)
_opt_step, _repack = ts.opt.step, eng.repack
ts.opt.step = lambda: (mark("(before) opt"), _opt_step(), mark("adamw"))[1]
eng.repack = lambda: (_repack(), mark("repack"))[0]
mark("start")
ts.step(ctx, target, override_gaussians=override)
mark("end")
torch.cuda.synchronize()
sections = defaultdict(float)
for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
    key = n1 if not n1.startswith("(before)") else "render fwd+bwd / glue before " + n1[9:]
    sections[key] += e0.elapsed_time(e1)
total = marks[0][1].elapsed_time(marks[-1][1])

# ---- families
for name in ("_encoder_fwd", "_decoder_fwd", "_heads_fwd", "_heads_bwd", "_decoder_bwd", "_encoder_bwd"):
    delattr(eng, name)
ts.opt.step, eng.repack = _opt_step, _repack
vops.TIMERS = {}
ts.step(ctx, target, override_gaussians=override)
fam = vops.family_ms(vops.TIMERS)
gem = defaultdict(lambda: [0.0, 0])
for e0, e1, meta in vops.TIMERS.get("gemm", []):
    gem[meta][0] += e0.elapsed_time(e1)
    gem[meta][1] += 1
vops.TIMERS = None
out = dict(scenes=NB, total_ms=total, sections=dict(sections), families=fam,
           gemm_top=[dict(meta=list(map(str, m)), ms=v[0], n=v[1],
                          tflops=(2.0 * m[1] * m[2] * m[3] * v[1] / (v[0] * 1e-3) / 1e12) if v[0] > 0 else None)
                     for m, v in sorted(gem.items(), key=lambda kv: -kv[1][0])[:40]])
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"profile_train_b{NB}.json").write_text(json.dumps(out, indent=1))
print(json.dumps(dict(total_ms=total, sections=dict(sections), families=fam), indent=1))
for g in out["gemm_top"][:25]:
    print(f"{g['ms']:8.3f} ms  n={g['n']:3d}  {g['tflops'] or 0:7.1f} TF/s  {g['meta']}")
