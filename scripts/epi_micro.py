"""Epilogue-only GEMM microbenchmark: K = 64 (one k-block) so the kernel time is the epilogue's."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops
from vicasplat_b200._lib import VS_ACT_GELU, VS_ACT_NONE
dev = torch.device("cuda:0")
bf = torch.bfloat16
M, K = 16448 * 4, 64
for N in (1024, 256):
    A = torch.randn((M, K), device=dev).to(bf)
    W = torch.randn((N, K), device=dev).to(bf)
    bias = torch.randn((N,), device=dev)
    xf = torch.randn((M, N), device=dev)
    xb = torch.randn((M, N), device=dev).to(bf)
    ob = torch.empty((M, N), device=dev, dtype=bf)
    of = torch.empty((M, N), device=dev)
    cases = {
        "bf16 out": lambda: ops.gemm(A, W, bias=bias, out=ob),
        "bf16 out, no bias": lambda: ops.gemm(A, W, out=ob),
        "f32 out": lambda: ops.gemm(A, W, bias=bias, out=of),
        "gelu bf16 out": lambda: ops.gemm(A, W, bias=bias, act=VS_ACT_GELU, out=ob),
        "f32 res + f32 out (in place)": lambda: ops.gemm(A, W, bias=bias, res1=xf, out=xf),
        "bf16 res + bf16 out": lambda: ops.gemm(A, W, bias=bias, res1=xb, out=ob),
    }
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        tiles = (M / 128) * (N / 256) / 148          # 128 x 256 CTA tiles per CTA
        print(f"N={N:5d} {name:32s} {us:8.1f} us  {us / tiles * 1.9:8.0f} clk/tile(1.9GHz)  "
              f"{M * N * 1e-3 / us:7.1f} Gelem/s")
