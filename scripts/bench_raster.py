#!/usr/bin/env python
"""Raster-side workloads of BASELINE.json beyond the headline bench line:

  --workload train  configs[2] raster part: 8-view 256x256 training step of the render path -- per
                    scene: 12-view forward render, MSE loss against target images (fused value +
                    gradient), rasterizer backward to means / cov6 / SH / opacity -- batch of
                    --batch scenes (default 24) per step.
  --workload stress configs[4]: 16-view 512x512 raster-only forward, G = 2^21 Gaussians per scene.

Prints one JSON line per run (same key conventions as bench.py).  Synthetic, seeded scenes."""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from bench import ClockSampler, load_peaks, raster_bytes  # noqa: E402
from vicasplat_b200 import decoder as dec, synthetic  # noqa: E402
from vicasplat_b200.loss import mse  # noqa: E402
from vicasplat_b200.rasterizer import rasterize_views  # noqa: E402
import vicasplat_b200.rasterizer as rmod  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", choices=["train", "stress"], default="train")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    if a.workload == "train":
        T, V, S, G, NB = 8, 12, 256, 8 * 256 * 256, a.batch or 24
        sc = synthetic.gaussian_scene(T, S, S, V, seed=1)
    else:
        T, V, S, G, NB = 16, 16, 512, 1 << 21, a.batch or 1
        sc = synthetic.gaussian_scene(T, S, S, V, seed=1, n_gauss=G)
    sc = {k: v.to(dev) for k, v in sc.items()}
    tanfov, view_t, full_t, campos = dec._cameras(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
    cov6 = dec._cov6(sc["covariances"]).contiguous()
    kw = dict(sh_degree=4, sh_layout="chan_major", viewmatrix=view_t, projmatrix=full_t, campos=campos,
              tanfov=tanfov, bg=torch.zeros((V, 3), device=dev), H=S, W=S, want_n_touched=False)
    for _ in range(2):   # calibrate the binning capacity hint (checked calls)
        rasterize_views(sc["means"], cov6, sc["opacities"], shs=sc["harmonics"], **kw)
    n_pairs = int((rmod._capacity_hint[(V, G, S, S)][0] - 4096) / 1.25)
    kw.update(check_overflow=False)
    target = torch.rand((V, 3, S, S), device=dev)
    leaves = [sc["means"], cov6, sc["opacities"], sc["harmonics"]]

    def step():
        for _ in range(NB):
            if a.workload == "stress":
                with torch.no_grad():
                    out = rasterize_views(sc["means"], cov6, sc["opacities"], shs=sc["harmonics"], **kw)[0]
            else:
                m, c, o, s = (t.detach().requires_grad_(True) for t in leaves)
                color = rasterize_views(m, c, o, shs=s, **kw)[0]
                loss = mse(color, target, 1.0)
                loss.backward()
                out = m.grad
        return out

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clk:
        e0.record()
        for _ in range(a.steps):
            out = step()
        e1.record()
        torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    ms = e0.elapsed_time(e1) / a.steps
    peaks = load_peaks()
    scenes = NB * 1e3 / ms
    fwd_bytes = NB * raster_bytes(V, G, S, S)
    # backward: one more read of the parameters, one write of their gradients, one read of dL/dRGB
    alg = fwd_bytes if a.workload == "stress" else NB * V * (3 * G * 232.0 + 2 * S * S * 16.0)
    gbs = alg / (ms * 1e-3) / 1e9
    line = dict(
        metric=("raster_only_scenes_per_sec_16view_512x512" if a.workload == "stress"
                else "raster_train_scenes_per_sec_8view_256x256"),
        value=scenes, unit="scenes/s", n_gpus=1, steps=a.steps, warmup=max(a.warmup, 3), ms_per_step=ms,
        higher_is_better=True, dtype="f32", data="synthetic",
        mpix_per_sec=scenes * V * S * S / 1e6,
        config=dict(workload=("configs[4]: 16-view 512x512 raster-only forward, G=2^21 Gaussians/scene"
                              if a.workload == "stress" else
                              "configs[2] render path: 12-view 256x256 forward + fused MSE (value+grad) + "
                              "rasterizer backward, G=524288 Gaussians/scene"),
                    scenes_per_step=NB, views=V, image=f"{S}x{S}", gaussians=G, raster_pairs=n_pairs,
                    l2="Gaussian parameters + binning workspace exceed L2; no explicit flush"),
        clocks=clk.result,
        roofline=dict(bound="hbm", kernel="rasterizer chain" + ("" if a.workload == "stress" else " fwd+bwd"),
                      achieved=gbs, peak=peaks["hbm"], unit="GB/s", frac=gbs / peaks["hbm"], traffic=None,
                      bytes_per_step=alg, peak_source=peaks["which"] + " copy"))
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
