"""Conv weight gradient (a_mode 3) vs the plain token wgrad (a_mode 2) of vs_gemm, for ncu / event timing.
usage: prof_wgrad.py [frames]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vicasplat_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((n, 256, 256, 256), generator=g, device=dev).to(torch.bfloat16)
dy = torch.randn((n, 256, 256, 256), generator=g, device=dev).to(torch.bfloat16)
dW3 = torch.zeros((256, 9 * 256), device=dev)
dW1 = torch.zeros((256, 256), device=dev)
dWt = torch.zeros((256, 256), device=dev)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


px = n * 65536
cases = {
    "conv3x3 wgrad (a_mode 3)": (lambda: ops.conv_wgrad(dy, x, dW3, kh=3, kw=3, pad=1), 2.0 * 256 * 2304 * px),
    "conv1x1 wgrad (a_mode 3)": (lambda: ops.conv_wgrad(dy, x, dW1, kh=1, kw=1, pad=0), 2.0 * 256 * 256 * px),
    "same 1x1 as token wgrad (a_mode 2)": (lambda: ops.gemm_tn_acc(dy.view(-1, 256), x.view(-1, 256), dWt), 2.0 * 256 * 256 * px),
}
for s in (4, 8, 16, 32, 64):
    cases[f"conv3x3 wgrad split_k={s}"] = (lambda s=s: ops.conv_wgrad(dy, x, dW3, kh=3, kw=3, pad=1, split_k=s), 2.0 * 256 * 2304 * px)
if os.environ.get("VS_PROFILE_STEP") == "1":
    cases["conv3x3 wgrad (a_mode 3)"][0]()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    cases["conv3x3 wgrad (a_mode 3)"][0]()
    cases["same 1x1 as token wgrad (a_mode 2)"][0]()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
for name, (fn, fl) in cases.items():
    ms = timed(fn)
    print(f"{name:40s} {ms:8.3f} ms  {fl / (ms * 1e-3) / 1e12:7.1f} TF/s")
