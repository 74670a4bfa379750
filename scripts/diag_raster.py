import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tests.test_gpu_raster import _scene, _oracle, _ours
dev = torch.device("cuda:0")
for hw, n_ctx, n_tgt, seed in [(64, 2, 3, 1), (48, 1, 2, 2), (80, 2, 1, 3), (48, 1, 1, 9)]:
    sc = _scene(n_ctx, hw, n_tgt, seed)
    c, d = _ours(sc, hw, dev)
    for dt in (torch.float32, torch.float64):
        rc, rd = _oracle(sc, hw, dt)
        ec = (c.cpu().double() - rc).abs()
        ed = (d.cpu().double() - rd).abs() / rd.abs().clamp_min(1)
        print(hw, seed, dt, "color max %.2e ok %.5f | depth max %.2e ok %.5f" % (ec.max(), (ec <= 1e-4).double().mean(), ed.max(), (ed <= 1e-4).double().mean()))
