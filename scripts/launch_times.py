"""Per-launch CUDA-event times of one un-graphed encoder forward (the roofline leg of bench.py, itemised):
   python scripts/launch_times.py OUT.json [reps]
Median over `reps` passes per launch index; used to A/B two builds of the library on one box."""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from vicasplat_b200 import ops, synthetic
from vicasplat_b200.encoder import VicaSplat, EncoderEngine

out = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = VicaSplat().to(dev).eval()
eng = EncoderEngine(model, use_graph=False)
image, K = synthetic.clip(8, 8, 256)
image, K = image.to(dev), K.to(dev)
for _ in range(2):
    eng.run(image, K, clone_outputs=False)
runs = []
for _ in range(reps):
    ops.TIMERS = {}
    eng.run(image, K, clone_outputs=False)
    torch.cuda.synchronize()
    runs.append({k: [(e[0].elapsed_time(e[1]), e[2]) for e in v] for k, v in ops.TIMERS.items()})
ops.TIMERS = None
res = {}
for fam in runs[0]:
    rows = []
    for i in range(len(runs[0][fam])):
        ts = sorted(r[fam][i][0] for r in runs)
        rows.append(dict(ms=ts[len(ts) // 2], meta=str(runs[0][fam][i][1])))
    res[fam] = rows
json.dump(res, open(out, "w"))
print({k: round(sum(r["ms"] for r in v), 3) for k, v in res.items()})
