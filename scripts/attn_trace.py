"""Device timeline of one CTA of the pair attention kernel (N=4096, d=40). Usage: attn_trace.py [variant]"""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops, _lib
var = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lib = _lib.load()
lib.af_attention_set_pair_variant(var)
B, heads, N, d, dp = 16, 8, 4096, 40, 48
q = torch.randn(B * N, 2 * heads * dp, device="cuda").to(torch.bfloat16) * 0.3
k = q[:, heads * dp:]
vt = torch.randn(heads * d, B * N, device="cuda").to(torch.bfloat16)
o = torch.empty(B * N, heads * d, device="cuda", dtype=torch.bfloat16)
f = lambda: ops.attention(q, k, vt, o, B=B, heads=heads, Nq=N, d=d, ldq=2 * heads * dp, Nk=N, ldk=2 * heads * dp, ldvt=B * N, kv_stride=N)
f(); torch.cuda.synchronize()
buf = torch.zeros(4 * 64 * 8, dtype=torch.int64, device="cuda")
lib.af_attention_set_trace(buf.data_ptr())
f(); torch.cuda.synchronize()
lib.af_attention_set_trace(None)
t = buf.cpu().view(4, 64, 8)
t0 = int(t[t > 0].min())
pieces = 2 if var & 1 else 1
print(f"variant {var}; clock64 cycles relative to first stamp; softmax events: top, S ready, S in regs, max done, " + ", ".join(f"P{h} free, P{h} out" for h in range(pieces)))
for j in list(range(0, 6)) + list(range(14, 20)):
    for a, name in enumerate(("sm0", "sm1", "mma0", "mma1")):
        ev = [int(x) - t0 if x > 0 else -1 for x in t[a, j]]
        n = 4 + 2 * pieces if a < 2 else 2 + 2 * pieces
        print(f"j={j:2d} {name:5s} " + " ".join(f"{e:7d}" for e in ev[:n]))
if var & 8:
    import statistics as st
    names = ["wait S", "ld S", "max+xchg", "wait P free", "wait turn", "exp", "loop"]
    for a in (0, 1):
        rows = t[a, 4:28].tolist(); nxt = t[a, 5:29, 0].tolist()
        durs = [[r[i + 1] - r[i] for r in rows] for i in range(6)] + [[n - r[6] for r, n in zip(rows, nxt)]]
        print(f"row-split softmax warp tile {a}: " + ", ".join(f"{nm} {st.mean(dd):.0f}" for nm, dd in zip(names, durs)) + f"  | period {st.mean([n - r[0] for r, n in zip(rows, nxt)]):.0f}")
    sys.exit(0)
# per-phase averages over blocks 4..27
import statistics as st
for a in (0, 1):
    rows = t[a, 4:28].tolist()
    nxt = t[a, 5:29, 0].tolist()
    names = ["wait S", "ld S", "max", "wait P0 free", "exp0"] + (["wait P1 free", "exp1"] if pieces == 2 else []) + ["loop"]
    durs = [[r[i + 1] - r[i] for r in rows] for i in range(3 + 2 * pieces)] + [[n - r[3 + 2 * pieces] for r, n in zip(rows, nxt)]]
    print(f"softmax warp tile {a}: " + ", ".join(f"{nm} {st.mean(dd):.0f}" for nm, dd in zip(names, durs)) + f"  | period {st.mean([n - r[0] for r, n in zip(rows, nxt)]):.0f}")
