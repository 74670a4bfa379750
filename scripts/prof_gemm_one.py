import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops
dev = torch.device("cuda:0")
m, k, n = 8224, 1024, 3072
A = torch.randn((m, k), device=dev).to(torch.bfloat16)
W = (torch.randn((n, k), device=dev) / k ** 0.5).to(torch.bfloat16)
bias = torch.randn((n,), device=dev)
out = torch.empty((m, n), device=dev, dtype=torch.bfloat16)
for _ in range(5):
    ops.gemm(A, W, bias=bias, out=out)
torch.cuda.synchronize()
