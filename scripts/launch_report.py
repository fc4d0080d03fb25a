"""Summaries of an `ncu --metrics gpu__time_duration.sum --csv` launch list of scripts/net_once.py / step_once.py (last step only).
  python scripts/launch_report.py gpurun_out/net_launches.csv [min_us]"""
import collections
import csv
import re
import sys

lines = open(sys.argv[1]).readlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.DictReader(lines[start:]))
names = [re.sub(r"\(.*", "", r["Kernel Name"])[:60] for r in rows]
n = len(rows)
period = None
for L in range(20, n // 2 + 1):                      # the step repeats: find its period from the tail
    if names[n - L:] == names[n - 2 * L:n - L]:
        period = L
        break
step = rows[n - period:] if period else rows
dur = [float(r["Metric Value"].replace(",", "")) / 1000 for r in step]
print(f"{len(step)} kernels per step, {sum(dur):.1f} us summed (cold-cache, serialised)")
agg = collections.defaultdict(lambda: [0, 0.0])
for r, d in zip(step, dur):
    k = re.sub(r"void |at::|<unnamed>::|native::", "", re.sub(r"\(.*", "", r["Kernel Name"]))[:64]
    agg[k][0] += 1
    agg[k][1] += d
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:66s} {v[0]:4d} {v[1]:9.1f} us")
for lo, hi in ((0, 5), (5, 10), (10, 20), (20, 40), (40, 80), (80, 160), (160, 1e9)):
    sel = [x for x in dur if lo <= x < hi]
    print(f"{lo:5.0f}-{hi:<10.0f} us: {len(sel):4d} launches {sum(sel):9.1f} us")
if len(sys.argv) > 2:
    t = 0.0
    for i, (r, d) in enumerate(zip(step, dur)):
        t += d
        if d >= float(sys.argv[2]):
            k = re.sub(r"void |at::|<unnamed>::|native::", "", re.sub(r"\(.*", "", r["Kernel Name"]))[:58]
            print(f"{i:4d} t={t / 1000:6.2f}ms {d:7.1f} us {r['Grid Size']:>16s} {k}")
