"""Device timeline of the tile attention kernel (N=4096, d=40 or N=1024, d=80). Usage: attn_tile_trace.py [d]"""
import sys, torch, statistics as st
sys.path.insert(0, ".")
from adaprompt_b200 import ops, _lib
d = int(sys.argv[1]) if len(sys.argv) > 1 else 40
lib = _lib.load()
B, heads = 16, 8
N = 4096 if d == 40 else 1024
dp = 48 if d == 40 else d
q = torch.randn(B * N, 2 * heads * dp, device="cuda").to(torch.bfloat16) * 0.3
k = q[:, heads * dp:]
vt = torch.randn(heads * d, B * N, device="cuda").to(torch.bfloat16)
o = torch.empty(B * N, heads * d, device="cuda", dtype=torch.bfloat16)
f = lambda: ops.attention(q, k, vt, o, B=B, heads=heads, Nq=N, d=d, ldq=2 * heads * dp, Nk=N, ldk=2 * heads * dp, ldvt=B * N, kv_stride=N)
f(); torch.cuda.synchronize()
NC = 512
buf = torch.zeros(NC * 64 * 8 + NC + 2 * 64 * 8, dtype=torch.int64, device="cuda")
rc = lib.af_attention_bf16_trace(q.data_ptr(), 2 * heads * dp, k.data_ptr(), 2 * heads * dp, vt.data_ptr(), B * N, N,
                                 o.data_ptr(), buf.data_ptr(), B, heads, N, N, d, torch.cuda.current_stream().cuda_stream)
_lib.check(rc, "af_attention_bf16_trace")
torch.cuda.synchronize()
h = buf.cpu()
t = h[:NC * 64 * 8].view(NC, 64, 8)
smid = h[NC * 64 * 8:NC * 64 * 8 + NC].tolist()
t2 = h[NC * 64 * 8 + NC:].view(2, 64, 8)
nb = N // (128 if d == 40 else 64)
lo, hi = 4, min(nb - 2, 28)
def phases(rows_t, names, nev):
    rows = rows_t[lo:hi].tolist(); nxt = rows_t[lo + 1:hi + 1, 0].tolist()
    durs = [[r[i + 1] - r[i] for r in rows] for i in range(nev - 1)] + [[n - r[nev - 1] for r, n in zip(rows, nxt)]]
    per = st.mean([n - r[0] for r, n in zip(rows, nxt)])
    return ", ".join(f"{nm} {st.mean(dd):.0f}" for nm, dd in zip(names, durs)) + f" | period {per:.0f}"
sm = ["wait S", "ld S", "max", "wait P free", "exp+st issue", "st drain+arrive", "loop"]
NEV, NEV_I, EXP0, EXP1 = 7, 7, 4, 5
inames = ["wait K", "wait S free", "issue S", "wait V", "wait P", "issue PV", "loop"]
print(f"d={d} N={N}: cycles per key block, blocks {lo}..{hi}")
print("softmax CTA0:", phases(t[0], sm, NEV))
print("issuer      :", phases(t2[0], inames, NEV_I))
print("producer    :", phases(t2[1], ["wait K slot", "issue K", "wait V slot", "issue V+loop"], 4))
# co-resident partner of CTA 0 in the first wave: same SM, start within 20k cycles
part = [c for c in range(1, NC) if smid[c] == smid[0] and abs(int(t[c, 0, 0]) - int(t[0, 0, 0])) < 20000]
print("CTA 0 on SM", smid[0], "first-wave partners:", part)
if part:
    c = part[0]
    print(f"softmax CTA{c}:", phases(t[c], sm, NEV))
    t0 = int(t[0, 0, 0])
    print("exp phase [start, end) of both CTAs relative to CTA 0's first stamp, blocks 8..15; overlap = both exponentiating")
    for j in range(8, 16):
        a0, a1 = int(t[0, j, EXP0]) - t0, int(t[0, j, EXP1]) - t0
        # partner block whose exp phase is nearest in time
        best = min(range(min(nb, 64)), key=lambda jj: abs(int(t[c, jj, EXP0]) - t0 - a0))
        b0, b1 = int(t[c, best, EXP0]) - t0, int(t[c, best, EXP1]) - t0
        ov = max(0, min(a1, b1) - max(a0, b0))
        print(f"  j={j}: CTA0 exp [{a0},{a1})  CTA{c} j={best} exp [{b0},{b1})  overlap {ov}")
    # exp-phase duration of CTA 0 against how much of it the partner was exponentiating too
    ivs = [(int(t[c, jj, EXP0]) - t0, int(t[c, jj, EXP1]) - t0) for jj in range(min(nb, 64))]
    rows = []
    for j in range(2, min(nb, 64) - 1):
        a0, a1 = int(t[0, j, EXP0]) - t0, int(t[0, j, EXP1]) - t0
        ov = sum(max(0, min(a1, b1) - max(a0, b0)) for b0, b1 in ivs)
        per = int(t[0, j + 1, 0]) - int(t[0, j, 0])
        rows.append((j, a1 - a0, ov, per))
    print("block: exp duration, overlap with the partner's exp phases, block period")
    print("  " + "  ".join(f"{j}:{dur}/{ov}/{per}" for j, dur, ov, per in rows))
    lo_ov = [r for r in rows if r[2] < 0.2 * r[1]]
    hi_ov = [r for r in rows if r[2] > 0.8 * r[1]]
    if lo_ov: print(f"  blocks with < 20 % overlap: n={len(lo_ov)} mean exp {st.mean(r[1] for r in lo_ov):.0f} period {st.mean(r[3] for r in lo_ov):.0f}")
    if hi_ov: print(f"  blocks with > 80 % overlap: n={len(hi_ov)} mean exp {st.mean(r[1] for r in hi_ov):.0f} period {st.mean(r[3] for r in hi_ov):.0f}")
