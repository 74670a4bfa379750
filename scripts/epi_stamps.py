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
