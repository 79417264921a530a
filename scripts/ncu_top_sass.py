"""Top SASS lines by warp-stall samples from `ncu -i rep --page source --csv`.  Usage: ncu_top_sass.py src.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
body = rows[2:]
i_s = hdr.index("Warp Stall Sampling (All Samples)")
i_e = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(float(r[i_s] or 0) for r in body)
print("total samples", tot, "lines", len(body))
idx = sorted(range(len(body)), key=lambda i: -float(body[i][i_s] or 0))[:n]
for i in idx:
    r = body[i]
    st = sorted(((float(r[c] or 0), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print(f"{float(r[i_s] or 0):8.0f} {100*float(r[i_s] or 0)/tot:5.1f}%  line {i:5d} exec {r[i_e]:>8s}  {r[1][:90]:90s} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
