"""Times vs_gemm on the model's GEMM shapes against torch.matmul (cuBLAS) -- tuning aid (GPU box)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops, _lib

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
M = 2056 * B
shapes = [("qkv", M, 1024, 3072, None), ("proj+res", M, 1024, 1024, "res"), ("fc1+gelu", M, 1024, 4096, "gelu"),
          ("fc2+res", M, 4096, 1024, "res"), ("dec qkv", 2064 * B, 768, 2304, None),
          ("dec fc1", 2064 * B, 768, 3072, "gelu"), ("dec proj", 2064 * B, 768, 768, "res")]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n * 1e3


for name, m, k, n, epi in shapes:
    A = torch.randn((m, k), device=dev).to(torch.bfloat16)
    W = (torch.randn((n, k), device=dev) / k ** 0.5).to(torch.bfloat16)
    bias = torch.randn((n,), device=dev)
    x = torch.randn((m, n), device=dev)
    out16 = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    kw = dict(bias=bias)
    if epi == "res":
        kw.update(res1=x, out=x)
    elif epi == "gelu":
        kw.update(act=_lib.VS_ACT_GELU, out=out16)
    else:
        kw.update(out=out16)
    res = {}
    for bn in (0, 128, 256):
        res[bn] = timeit(lambda: ops.gemm(A, W, block_n=bn, **kw))
    t_cublas = timeit(lambda: torch.matmul(A, W.T, out=out16))
    fl = 2.0 * m * k * n
    print(f"{name:10s} M={m} K={k} N={n}: ours auto {res[0]:7.1f} us ({fl/res[0]/1e6:6.0f} TF/s)  bn128 {res[128]:7.1f}  "
          f"bn256 {res[256]:7.1f} | cuBLAS {t_cublas:7.1f} us ({fl/t_cublas/1e6:6.0f} TF/s)", flush=True)
