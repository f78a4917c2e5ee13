"""Summarise one decode step of an `ncu --metrics gpu__time_duration.sum --csv` launch list (probe, not a test).
usage: python tests/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import re
import sys

with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
rows = [(x["Kernel Name"], x["Grid Size"], float(x["Metric Value"]) / 1000) for x in csv.DictReader(lines)]
idx = [i for i, (n, g, t) in enumerate(rows) if "embed_kernel" in n]
print(len(rows), "launches; embed at", idx)
a, b = idx[0], idx[1]
step = rows[a:b]
total = sum(t for _, _, t in step)
print(f"one step: {len(step)} kernels, {total:.1f} us")
agg = collections.OrderedDict()
for n, g, t in step:
    n = re.sub(r"\(.*", "", n).replace("void ", "").replace("scv::", "").replace("<unnamed>::", "").replace("unnamed>::", "")
    agg.setdefault((n[:64], g), [0, 0.0])
    agg[(n[:64], g)][0] += 1
    agg[(n[:64], g)][1] += t
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100 * t / total:5.1f}%  n={c:3d} avg={t / c:7.2f}  {k[0]}  {k[1]}")
