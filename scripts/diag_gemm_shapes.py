"""Per-shape GEMM efficiency of one eager encoder forward (CUDA events around every launch).
usage: python scripts/diag_gemm_shapes.py [scenes]"""
import sys
from collections import defaultdict
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops, synthetic
from vicasplat_b200.encoder import VicaSplat, EncoderEngine

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VicaSplat().to(dev).eval()
eng = EncoderEngine(model, use_graph=False)
image, K = synthetic.clip(NB, 8, 256)
image, K = image.to(dev), K.to(dev)
for _ in range(2):
    eng.run(image, K, clone_outputs=False)
acc = defaultdict(lambda: [0, 0.0])
REP = 3
for _ in range(REP):
    ops.TIMERS = {}
    eng.run(image, K, clone_outputs=False)
    torch.cuda.synchronize()
    for fam, v in ops.TIMERS.items():
        for e0, e1, meta in v:
            key = (fam, meta)
            acc[key][0] += 1
            acc[key][1] += e0.elapsed_time(e1)
    ops.TIMERS = None
rows = []
tot = 0.0
for (fam, meta), (n, ms) in acc.items():
    ms /= REP; n //= REP
    tf = None
    if meta is not None and meta[0] == "attn":
        _, items, heads, ql, kl, cb = meta
        tf = 4.0 * items * heads * ql * kl * 64 * n / (ms * 1e-3) / 1e12
    elif meta is not None:
        kind, M, N, Kd = meta[:4]
        tf = 2.0 * M * N * Kd * n / (ms * 1e-3) / 1e12
    rows.append((ms, fam, meta, n, tf))
    tot += ms
rows.sort(key=lambda r: -r[0])
print(f"# scenes={NB} total event-timed ms={tot:.2f}")
for ms, fam, meta, n, tf in rows:
    print(f"{ms:8.3f} ms  n={n:3d}  {fam:10s} {meta}  {'' if tf is None else f'{tf:7.1f} TF/s'}")
