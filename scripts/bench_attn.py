"""Micro-benchmark of af_attention_bf16 at the UNet's shapes (batch 16). Usage: bench_attn.py [d] [reps]"""
import sys, math, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops
only = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, heads = 16, 8
def run(N, Nk, d, cross):
    dp = 48 if d == 40 else d
    C = heads * d
    q = torch.randn(B * N, 2 * heads * dp, device="cuda").to(torch.bfloat16) * 0.3
    if cross:
        nkp = 80
        k = torch.randn(B * nkp, heads * dp, device="cuda").to(torch.bfloat16)
        vt = torch.randn(C, B * nkp, device="cuda").to(torch.bfloat16)
        kw = dict(Nk=Nk, ldk=heads * dp, ldvt=B * nkp, kv_stride=nkp)
    else:
        k = q[:, heads * dp:]
        vt = torch.randn(C, B * N, device="cuda").to(torch.bfloat16)
        kw = dict(Nk=N, ldk=2 * heads * dp, ldvt=B * N, kv_stride=N)
    o = torch.empty(B * N, C, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.attention(q, k, vt, o, B=B, heads=heads, Nq=N, d=d, ldq=2 * heads * dp, **kw)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 4.0 * B * heads * N * kw["Nk"] * d
    print(f"N={N:5d} Nk={kw['Nk']:5d} d={d:3d} {'cross' if cross else 'self '}: {ms:8.3f} ms  {fl/ms/1e9:8.1f} TFLOP/s", flush=True)
for (N, d) in ((4096, 40), (1024, 80), (256, 160), (64, 160)):
    if only and d != only: continue
    run(N, N, d, False)
    run(N, 77, d, True)
