"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
Usage: summarize_launches.py launches.csv [first_id last_id]  > profiles/xxx.md"""
import csv
import re
import sys
from collections import OrderedDict

path = sys.argv[1]
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    i = int(r["ID"])
    if lo <= i <= hi:
        rows.append((i, r["Kernel Name"], float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name[:70]


agg = OrderedDict()
for _, name, us, grid, block in rows:
    d = agg.setdefault(short(name), [0, 0.0])
    d[0] += 1
    d[1] += us
total = sum(v[1] for v in agg.values())
print(f"launches {len(rows)} (ids {rows[0][0]}..{rows[-1][0]}), total device time {total / 1e3:.3f} ms "
      f"(ncu: cold-cache, serialised - compare SHARES)\n")
print("| kernel | launches | total us | share |")
print("|---|---:|---:|---:|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f}% |")
