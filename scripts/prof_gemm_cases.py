"""A few representative GEMM launches of the encoder (8 scenes) for ncu:
   ncu --set full --import-source on -k regex:gemm_tc05 ... python scripts/prof_gemm_cases.py [case]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops
from vicasplat_b200._lib import VS_ACT_GELU, VS_ACT_RELU, VS_ACT_NONE
dev = torch.device("cuda:0")
case = sys.argv[1] if len(sys.argv) > 1 else "all"
bf = torch.bfloat16
torch.manual_seed(0)

def lin(m, n, k, act=VS_ACT_NONE, res=False, f32=False):
    A = torch.randn((m, k), device=dev).to(bf)
    W = (torch.randn((n, k), device=dev) / k ** 0.5).to(bf)
    bias = torch.randn((n,), device=dev)
    x = torch.randn((m, n), device=dev) if res else None
    out = x if res else torch.empty((m, n), device=dev, dtype=torch.float32 if f32 else bf)
    def run():
        ops.gemm(A, W, bias=bias, act=act, res1=x, out=out)
    return run

def stem(fr=16, H=256, W=256):
    img8 = torch.randn((fr, H + 6, W + 8, 8), device=dev).to(bf)
    w = (torch.randn((256, 448), device=dev) / 20).to(bf)
    bias = torch.randn((256,), device=dev)
    p1 = torch.randn((fr, H // 2, W // 2, 256), device=dev).to(bf)
    out = torch.empty((fr, H, W, 256), device=dev, dtype=bf)
    def run():
        ops.conv_gemm(img8, w, kh=7, kw=1, pad=0, N=256, bias=bias, act=VS_ACT_RELU, res1=p1, res_up2=True,
                      out=out, view=(fr, H, W, 64, H + 6, 8, (W + 8) * 8, (H + 6) * (W + 8) * 8))
    return run

cases = dict(
    proj=lin(16448, 1024, 1024, res=True),
    fc1=lin(16448, 4096, 1024, act=VS_ACT_GELU),
    qkv=lin(16448, 3072, 1024),
    dproj=lin(16512, 768, 768, res=True),
    dfc1=lin(16512, 3072, 768, act=VS_ACT_GELU),
    dqkv=lin(16512, 2304, 768),
    stem=stem(),
)
for name, fn in cases.items():
    if case != "all" and case != name:
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")
