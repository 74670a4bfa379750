"""block_n sweep for the epilogue-bound GEMM shapes (which tile width should the cost model pick?)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops
from vicasplat_b200._lib import VS_ACT_GELU, VS_ACT_NONE
dev = torch.device("cuda:0"); bf = torch.bfloat16
def case(m, n, k, act=VS_ACT_NONE, res=False):
    A = torch.randn((m, k), device=dev).to(bf); W = (torch.randn((n, k), device=dev) / k ** 0.5).to(bf)
    bias = torch.randn((n,), device=dev)
    x = torch.randn((m, n), device=dev) if res else None
    out = x if res else torch.empty((m, n), device=dev, dtype=bf)
    res_t = []
    for bn in (0, 64, 128, 256):
        fn = lambda: ops.gemm(A, W, bias=bias, act=act, res1=x, out=out, block_n=bn)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        res_t.append(f"bn={bn}: {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us")
    print(f"M={m} N={n} K={k} act={act} res={res}:  " + "  ".join(res_t))
case(16448, 1024, 1024, res=True)
case(16512, 768, 768, res=True)
case(16448, 4096, 1024, act=VS_ACT_GELU)
case(16512, 3072, 768, act=VS_ACT_GELU)
case(16448, 3072, 1024)
case(16512, 2304, 768)
case(16448, 1024, 4096, res=True)
case(16512, 768, 3072, res=True)
case(2056, 1024, 1024, res=True)
case(2056, 3072, 1024)
case(2056, 4096, 1024, act=VS_ACT_GELU)
case(2056, 1024, 4096, res=True)
case(2064, 768, 768, res=True)
case(2064, 2304, 768)
