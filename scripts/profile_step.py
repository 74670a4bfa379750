"""One eager (un-graphed) bench step between cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off ... python scripts/profile_step.py
The kernels are exactly those the CUDA graph of bench.py replays."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import synthetic, decoder as dec
from vicasplat_b200.encoder import VicaSplat, EncoderEngine
from vicasplat_b200.rasterizer import rasterize_views

T, V, S = 8, 12, 256
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VicaSplat().to(dev).eval()
with torch.no_grad():
    for n, p in model.named_parameters():
        if "modulation" in n or n.startswith("camera_extrinsic_head"):
            p.normal_(0, 0.02)
eng = EncoderEngine(model, use_graph=False)
image, K = synthetic.clip(NB, T, S)
image, K = image.to(dev), K.to(dev)
sc = {k: v.to(dev) for k, v in synthetic.gaussian_scene(T, S, S, V, seed=1).items()}
tanfov, view_t, full_t, campos = dec._cameras(sc["extrinsics"], sc["intrinsics"], sc["near"], sc["far"])
cov6 = dec._cov6(sc["covariances"]).contiguous()
kw = dict(shs=sc["harmonics"], sh_degree=4, sh_layout="chan_major", viewmatrix=view_t, projmatrix=full_t,
          campos=campos, tanfov=tanfov, bg=torch.zeros((V, 3), device=dev), H=S, W=S, want_n_touched=False)
rasterize_views(sc["means"], cov6, sc["opacities"], **kw)          # calibrates the pair capacity hint
for _ in range(2):
    eng.run(image, K, clone_outputs=False)
    rasterize_views(sc["means"], cov6, sc["opacities"], check_overflow=False, **kw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
eng.run(image, K, clone_outputs=False)
rasterize_views(sc["means"], cov6, sc["opacities"], check_overflow=False, **kw)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one step")
