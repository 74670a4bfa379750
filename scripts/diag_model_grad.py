"""Per-tensor gradient errors of TrainEngine against autograd over the fp32 oracle (diagnostic)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
import test_gpu_model_grad as T  # noqa: E402
from oracle import encoder_ref as er  # noqa: E402
from vicasplat_b200.train import TrainEngine  # noqa: E402

dev = torch.device("cuda:0")
group = sys.argv[1] if len(sys.argv) > 1 else "pose"
cfg, model, image, K = T._build(dev)
eng = TrainEngine(model)
out = eng.forward(image, K)
G = out["raw"].shape[0]
gen = lambda i, s: torch.randn(s, generator=torch.Generator().manual_seed(700 + i)).to(dev)
d_pred = gen(1, out["pred_extrins"].shape)
d_raw = gen(0, (G, 86)); d_raw[:, :3] = 0


d_c = gen(5, (G, 3))
hold = {}


def make_loss(o):
    if group == "pose":
        return (o["pred_extrins"] * d_pred).sum()
    if group == "centers":
        c_ref = o["raw_gaussians"][..., :3].detach().reshape(G, 3)
        hold["wc"] = (d_c / (1 + (c_ref * c_ref).sum(-1, keepdim=True))).contiguous()
        return (o["gaussians"]["means"].reshape(G, 3) * hold["wc"]).sum()
    return (o["raw_gaussians"].reshape(G, 86) * d_raw).sum()


_, grad_of = T._oracle_grads(cfg, image, K, make_loss, dev, masks=eng.relu_masks())
if group == "pose":
    eng.backward(d_pred=d_pred)
elif group == "centers":
    eng.backward(d_means=hold["wc"])
else:
    eng.backward(d_raw=d_raw)
rows = []
for k, p in model.named_parameters():
    ref = grad_of(k)
    if ref is None or ref.norm() == 0:
        continue
    err = ((p.grad - ref).norm() / ref.norm()).item()
    cos = torch.nn.functional.cosine_similarity(p.grad.flatten().double(), ref.flatten().double(), dim=0).item()
    rows.append((k, err, cos, p.grad.norm().item() / ref.norm().item()))
out_dir = ROOT / "gpurun_out"
out_dir.mkdir(exist_ok=True)
with open(out_dir / f"diag_model_grad_{group}.txt", "w") as f:
    for k, e, c, r in rows:
        f.write(f"{e:9.3e} cos={c:.5f} ratio={r:.4f} {k}\n")
print("done", len(rows))
