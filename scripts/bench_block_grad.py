"""Times the training path of ONE ViT encoder block at the bench size of BASELINE
configs[2]: `scenes` scenes x 8 frames x 257 tokens.  Prints one JSON line with forward-train /
backward times, executed FLOP (2*MAC) and the per-kernel-family split of the backward pass.

    python scripts/bench_block_grad.py [scenes=8] [iters=10]
"""
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from vicasplat_b200 import encoder_grad as eg, ops, synthetic
from vicasplat_b200.encoder_train import ViTEncoderConfig


def main():
    scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dev = torch.device("cuda:0")
    cfg = ViTEncoderConfig(enc_depth=1)
    sd = {k: v.to(dev) for k, v in synthetic.vit_encoder_state_dict(depth=1, seed=0).items()
          if k.startswith("backbone.enc_blocks.0.")}
    w = eg.pack_block(sd, "backbone.enc_blocks.0", dev)
    g = eg.zero_grads(w)
    Fr = scenes * 8
    lay = eg.FrameLayout.make(Fr, 16, 16, cfg.enc_num_heads, dev)
    M, E = Fr * lay.n, cfg.enc_embed_dim
    x = torch.randn((M, E), device=dev)
    dout = torch.randn((M, E), device=dev)

    def fwd():
        s = eg.Saved()
        return eg.block_forward(x, w, lay, s), s

    def time(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(3):
        out, saved = fwd()
        eg.block_backward(dout, w, g, lay, saved)
    if os.environ.get("VS_PROFILE"):                 # ncu --profile-from-start off: one forward + backward
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        out, saved = fwd()
        eg.block_backward(dout, w, g, lay, saved)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    t_fwd = time(fwd, iters)
    out, saved = fwd()
    t_bwd = time(lambda: eg.block_backward(dout, w, g, lay, saved), iters)

    ops.TIMERS = {}
    eg.block_backward(dout, w, g, lay, saved)
    fam = ops.family_ms(ops.TIMERS)
    ops.TIMERS = None

    lin = 2.0 * M * 12 * E * E                       # qkv + proj + fc1 + fc2
    att = 4.0 * lay.n * lay.n * 64 * cfg.enc_num_heads * Fr
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    line = {
        "what": "ViT encoder block (croco/blocks.py:81-130) training path", "scenes": scenes, "rows": M,
        "fwd_train_ms": round(t_fwd, 3), "bwd_ms": round(t_bwd, 3),
        "fwd_tflops": round((lin + att) / t_fwd / 1e9, 1),
        "bwd_tflops_algorithmic": round((2 * lin + 2.5 * att) / t_bwd / 1e9, 1),
        "bwd_family_ms": {k: round(v, 3) for k, v in fam.items()},
        "peaks": peaks,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
