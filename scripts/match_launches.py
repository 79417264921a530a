"""Joins the C-ABI call sequence of one UNet step (gpurun_out/step_calls.json, written by scripts/unet_step.py)
with the ncu launch list of the SAME step, giving real device time per call signature.
Usage: match_launches.py launches.csv step_calls.json [traffic.json]"""
import csv
import json
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
allrows = list(csv.DictReader(lines))
rows = [r for r in allrows if r["Metric Name"] == "gpu__time_duration.sum"]
dram = {}
for r in allrows:                      # optional: dram__bytes_read.sum / dram__bytes_write.sum collected in the same pass
    if r["Metric Name"].startswith("dram__bytes_"):
        v = float(r["Metric Value"]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1)
        dram[r["ID"]] = dram.get(r["ID"], 0.0) + v
kern = [(r["Kernel Name"], float(r["Metric Value"]) / 1e3, dram.get(r["ID"])) for r in rows]
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
    db = sum((k[2] or 0.0) for k in seg[i:i + nk])
    i += nk
    for table, key in ((agg, f"{name[3:]} {sig}"), (cls, name)):
        d = table.setdefault(key, [0, 0.0, 0.0, 0.0, 0.0])
        d[0] += 1; d[1] += us; d[2] += fl; d[3] += by; d[4] += db
tot = sum(v[1] for v in cls.values())
print(f"one UNet step: {need} kernels, {tot / 1e3:.3f} ms device time (ncu, serialised, cold-ish caches)\n")
print("| entry point | calls | us | share | TFLOP/s | GB/s (algorithmic) |\n|---|---:|---:|---:|---:|---:|")
for k, (n, us, fl, by, db) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% | {fl / us / 1e6:.1f} | {by / us / 1e3:.1f} |")
if dram:
    print("\nDRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, same ncu pass) vs algorithmic bytes, per launch:\n")
    print("| entry point | launches | algorithmic MB / launch | DRAM MB / launch | ratio |\n|---|---:|---:|---:|---:|")
    for k, (n, us, fl, by, db) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
        if by > 0:
            print(f"| `{k}` | {n} | {by / n / 1e6:.2f} | {db / n / 1e6:.2f} | {db / by:.2f} |")
    if len(sys.argv) > 3 and sys.argv[3].endswith(".json"):
        json.dump({k: {"launches": n, "us": us, "algorithmic_bytes": by, "dram_bytes": db, "flops": fl}
                   for k, (n, us, fl, by, db) in cls.items()}, open(sys.argv[3], "w"), indent=1)
print("\n| call signature | calls | us total | us each | TFLOP/s | GB/s |\n|---|---:|---:|---:|---:|---:|")
for k, (n, us, fl, by, db) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"| {k} | {n} | {us:.1f} | {us / n:.1f} | {fl / us / 1e6:.1f} | {by / us / 1e3:.1f} |")
