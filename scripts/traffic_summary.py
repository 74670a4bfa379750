"""Per-kernel time and DRAM traffic from an ncu csv with gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum (scripts/gpu_round.sh traffic)."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [set(), 0.0, 0.0, 0.0])
for row in csv.DictReader(lines):
    n = re.sub(r'\(.*', '', row['Kernel Name'])[:72]
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']; m = row['Metric Name']
    a = agg[n]; a[0].add(row['ID'])
    if m == 'gpu__time_duration.sum':
        a[1] += v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    else:
        mb = v / 1e6 if u == 'byte' else (v / 1e3 if u == 'Kbyte' else (v if u == 'Mbyte' else v * 1e3))
        a[2 if 'read' in m else 3] += mb
print("# kernel | launches | time us | dram read MB | dram write MB")
gemm = 0.0
for k, (ids, t, r, w) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:72s} {len(ids):5d} {t:10.1f} {r:10.1f} {w:10.1f}")
    if 'gemm_tc05' in k: gemm += (r + w) * 1e6
print(f"# GEMM family total DRAM bytes: {gemm:.0f}")
