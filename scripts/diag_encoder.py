"""Per-stage error report of the GPU encoder against the fp32 oracle (debug aid, GPU box)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from oracle import encoder_ref as er
from tests.test_gpu_encoder import _build, _rel

name = sys.argv[1] if len(sys.argv) > 1 else "small"
dev = torch.device("cuda:0")
cfg, model, sd, image, K, _ = _build(name, dev)
B, T = image.shape[:2]
taps = {}
out = model.engine().run(image, K, taps=taps)
with torch.no_grad():
    ref = er.forward(sd, image, K, cfg, stages=True)
    gh = cfg.img_size // 16
    flat = [t.flatten(0, 1) for t in ref["intermediates"]]
    for head in ("downstream_head1", "gaussian_param_head"):
        p1 = er.dpt_trunk(sd, head + ".dpt", flat, gh, gh, cfg).permute(0, 2, 3, 1)
        print(head, "path1 rel", _rel(taps[head + ".path1"], p1), "absmax", p1.abs().max().item())
    raw, rr_ = out["raw"], ref["raw_gaussians"]
    print("centers rel", _rel(raw[..., :3], rr_[..., :3]), "max|c|", rr_[..., :3].abs().max().item())
    d_ref = torch.log1p(rr_[..., :3].norm(dim=-1))
    d_got = torch.log1p(raw[..., :3].norm(dim=-1))
    print("pre-exp distance abs err: mean", (d_ref - d_got).abs().mean().item(), "max", (d_ref - d_got).abs().max().item(), "mean d", d_ref.mean().item())
    print("params rel", _rel(raw[..., 3:], rr_[..., 3:]))
    for i in range(1, cfg.dec_depth):
        rpf = taps[f"dec{i}"].shape[0] // (B * T)
        print("dec", i, _rel(taps[f"dec{i}"].view(B, T, rpf, -1)[:, :, 1:-1], ref["intermediates"][i]))
