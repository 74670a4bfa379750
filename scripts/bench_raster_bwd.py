"""render_backward alone on the bench scene (12 views, 256x256, 524 288 Gaussians): ms per call."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import decoder as dec, synthetic
from vicasplat_b200.rasterizer import render_forward, render_backward

dev = torch.device("cuda:0")
T, V, S = 8, 12, 256
sc = synthetic.gaussian_scene(T, S, S, V, seed=1, device=dev)
cov6 = dec._cov6(sc["covariances"]).contiguous()
tanfov, view_t, full_t, campos = dec._cameras(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
kw = dict(sh_degree=4, sh_layout="chan_major", viewmatrix=view_t, projmatrix=full_t, campos=campos, tanfov=tanfov,
          bg=torch.zeros(3, device=dev), H=S, W=S)
for _ in range(2):
    color, depth, alpha, st = render_forward(sc["means"], cov6, sc["opacities"], sc["harmonics"], **kw)
g = torch.randn_like(color)
G = sc["means"].shape[0]
out = dict(d_means=torch.zeros((G, 3), device=dev), d_cov6=torch.zeros((G, 6), device=dev),
           d_opac=torch.zeros((G,), device=dev), d_sh=torch.zeros((G, 75), device=dev))
for want_tau in (False, True):
    for _ in range(3):
        render_backward(st, g, out=out, want_tau=want_tau)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        render_backward(st, g, out=out, want_tau=want_tau)
    e1.record()
    torch.cuda.synchronize()
    print(f"render_backward want_tau={want_tau}: {e0.elapsed_time(e1) / 10:.3f} ms")
