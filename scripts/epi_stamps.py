"""Debug: in-kernel clock64 stamps of one GEMM epilogue chunk.  Needs the instrumented library:
    python -c "import __graft_entry__ as g; g.build_debug_library()"     (nvcc ... -DVS_EPI_TIMING)
then run this script on the GPU box."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from vicasplat_b200 import _lib
_lib.LIB_PATH = ROOT / "vicasplat_b200" / "lib" / "libvs_dbg.so"
from vicasplat_b200 import ops
from vicasplat_b200._lib import VS_ACT_GELU
dev = torch.device("cuda:0")
bf = torch.bfloat16
lib = _lib.load()
names = ["tile start", "acc ready", "chunk start", "tmem ld done", "STS issued", "syncwarp1", "stores issued", "syncwarp2", "tile end"]
def run(tag, M, N, K, **kw):
    A = torch.randn((M, K), device=dev).to(bf); W = torch.randn((N, K), device=dev).to(bf)
    for _ in range(3):
        ops.gemm(A, W, **kw)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    assert lib.vs_debug_epi_stamps(buf) == 0
    t = list(buf)[:9]
    print(tag, " ".join(f"{names[i]}=+{t[i] - t[0]}" for i in range(9)))
M = 16448 * 4
bias = torch.randn((1024,), device=dev)
run("K=64 bf16 nobias ", M, 1024, 64, out=torch.empty((M, 1024), device=dev, dtype=bf))
run("K=64 bf16 bias   ", M, 1024, 64, bias=bias, out=torch.empty((M, 1024), device=dev, dtype=bf))
x = torch.randn((M, 1024), device=dev)
run("K=64 f32 res     ", M, 1024, 64, bias=bias, res1=x, out=x)
run("K=1024 f32 res   ", 16448, 1024, 1024, bias=bias, res1=x[:16448], out=x[:16448])
xb = torch.randn((M, 256), device=dev).to(bf)
run("K=448 N=256 bf16 res", M, 256, 448, bias=bias[:256], res1=xb, out=torch.empty((M, 256), device=dev, dtype=bf))
run("K=448 N=256 bf16    ", M, 256, 448, bias=bias[:256], out=torch.empty((M, 256), device=dev, dtype=bf))
b4 = torch.randn((4096,), device=dev)
run("K=1024 N=4096 gelu  ", 16448, 4096, 1024, bias=b4, act=VS_ACT_GELU, out=torch.empty((16448, 4096), device=dev, dtype=bf))
run("K=1024 N=3072 bf16  ", 16448, 3072, 1024, bias=b4[:3072], out=torch.empty((16448, 3072), device=dev, dtype=bf))
run("K=256 N=96 f32      ", M, 96, 256, bias=bias[:96], out=torch.empty((M, 100), device=dev)[:, :96], out_dtype=torch.float32)
