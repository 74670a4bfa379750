"""Per-shape table of the GEMM / implicit-GEMM launches of one encoder forward from scripts/launch_times.py's
json:  python scripts/gemm_shapes.py gpurun_out/lt.json > profiles/rN_gemm_shapes_b8.txt"""
import ast
import json
import sys
from collections import defaultdict

d = json.load(open(sys.argv[1]))
agg = defaultdict(lambda: [0.0, 0])
for a in d["gemm"]:
    agg[a["meta"]][0] += a["ms"]
    agg[a["meta"]][1] += 1
tot = sum(v[0] for v in agg.values())
print(f"# GEMM family: {tot:.3f} ms per 8-scene encoder pass (CUDA events per launch, un-graphed pass, median of the reps)")
print("# (kind, M, N, K, activation, has residual, fp32 output) | launches | ms | TF/s executed | share")
fl_tot = 0.0
for k, (ms, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    m = ast.literal_eval(k)
    fl = 2.0 * m[1] * m[2] * m[3] * n
    fl_tot += fl
    print(f"{k:62s} {n:4d} {ms:8.3f} {fl / (ms * 1e-3) / 1e12:8.1f} {100 * ms / tot:5.1f}%")
print(f"# total executed {fl_tot / 1e12:.2f} TFLOP -> {fl_tot / (tot * 1e-3) / 1e12:.1f} TF/s")
for fam in ("attention", "layernorm", "upsample2x"):
    if fam in d:
        print(f"# {fam}: {sum(a['ms'] for a in d[fam]):.3f} ms over {len(d[fam])} launches")
