import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops
dev = torch.device("cuda:0")
bf = torch.bfloat16
M, K, N = 16448 * 4, 64, 1024
A = torch.randn((M, K), device=dev).to(bf)
W = torch.randn((N, K), device=dev).to(bf)
ob = torch.empty((M, N), device=dev, dtype=bf)
for _ in range(4):
    ops.gemm(A, W, out=ob)
torch.cuda.synchronize()
