"""Summarise an ncu gpu__time_duration launch list (csv) by kernel; optional per-grid GEMM table."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0; L = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    n = re.sub(r'\(.*', '', row['Kernel Name']); agg[n][0] += 1; agg[n][1] += v; tot += v
    L.append((v, int(row['ID']), n[-34:], row['Grid Size']))
print(f"# total {tot:.1f} us over {len(L)} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d} avg={t/n:8.1f}  {k[:100]}")
if len(sys.argv) > 2:
    print("# top launches")
    for v, i, n, g in sorted(L, reverse=True)[:int(sys.argv[2])]:
        print(f"{v:9.1f} id={i:4d} {n} {g}")
