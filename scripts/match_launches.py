"""Joins the C-ABI call sequence of one UNet step (gpurun_out/step_calls.json, written by scripts/unet_step.py)
with the ncu launch list of the SAME step, giving real device time per call signature.
Usage: match_launches.py launches.csv step_calls.json [step_index_from_end=1]"""
import csv
import json
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
kern = [(r["Kernel Name"], float(r["Metric Value"]) / 1e3) for r in rows]
calls = json.load(open(sys.argv[2]))
need = sum(c[2] for c in calls)
# scripts/unet_step.py records its FIRST call (which also projects the per-prompt K/V cache)
seg = kern[:need]
assert len(seg) == need, (len(seg), need)
agg = OrderedDict()
cls = OrderedDict()
i = 0
for name, sig, nk, fl, by in calls:
    us = sum(k[1] for k in seg[i:i + nk])
    i += nk
    for table, key in ((agg, f"{name[3:]} {sig}"), (cls, name)):
        d = table.setdefault(key, [0, 0.0, 0.0, 0.0])
        d[0] += 1; d[1] += us; d[2] += fl; d[3] += by
tot = sum(v[1] for v in cls.values())
print(f"one UNet step: {need} kernels, {tot / 1e3:.3f} ms device time (ncu, serialised, cold-ish caches)\n")
print("| entry point | calls | us | share | TFLOP/s | GB/s (algorithmic) |\n|---|---:|---:|---:|---:|---:|")
for k, (n, us, fl, by) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% | {fl / us / 1e6:.1f} | {by / us / 1e3:.1f} |")
print("\n| call signature | calls | us total | us each | TFLOP/s | GB/s |\n|---|---:|---:|---:|---:|---:|")
for k, (n, us, fl, by) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"| {k} | {n} | {us:.1f} | {us / n:.1f} | {fl / us / 1e6:.1f} | {by / us / 1e3:.1f} |")
