import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tests.test_gpu_raster_backward import _setup, _oracle_grads, _our_grads, _rel
dev = torch.device("cuda:0")
for (hw, n_ctx, n_tgt, seed) in [(32, 2, 2, 3), (48, 1, 3, 5), (48, 1, 1, 5), (64, 2, 2, 7)]:
    sc, wc, wd = _setup(hw, n_ctx, n_tgt, seed)
    for pose in (False, True):
        ref = _oracle_grads(sc, wc, wd, hw, pose)
        got = _our_grads(sc, wc, wd, hw, dev, pose)
        print(hw, n_ctx, n_tgt, seed, "pose", pose, "loss", got["loss"].item(), ref["loss"].item(),
              {k: round(_rel(got[k], ref[k]), 5) for k in ("means", "cov6", "sh", "op") + (("theta", "rho") if pose else ())})
        if pose:
            print("   theta", got["theta"].cpu().tolist(), ref["theta"].tolist())
            print("   rho  ", got["rho"].cpu().tolist(), ref["rho"].tolist())
