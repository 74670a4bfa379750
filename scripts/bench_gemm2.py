import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops, _lib
dev = torch.device("cuda:0")

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n): fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n * 1e3

for (m, k, n) in [(8192, 8192, 4096), (8192, 1024, 4096), (8192, 2048, 4096), (8192, 4096, 4096), (16384, 1024, 4096), (8192 * 4, 1024, 1024)]:
    A = torch.randn((m, k), device=dev).to(torch.bfloat16)
    W = (torch.randn((n, k), device=dev) / k ** 0.5).to(torch.bfloat16)
    out = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
    fl = 2.0 * m * k * n
    t1 = timeit(lambda: ops.gemm(A, W, out=out))
    t2 = timeit(lambda: ops.gemm(A, W, out=out, block_n=128))
    tc = timeit(lambda: torch.matmul(A, W.T, out=out))
    print(f"M={m} K={k} N={n}: ours {t1:8.1f} us {fl/t1/1e6:6.0f} TF/s | bn128 {t2:8.1f} {fl/t2/1e6:6.0f} | cuBLAS {tc:8.1f} us {fl/tc/1e6:6.0f} TF/s", flush=True)
