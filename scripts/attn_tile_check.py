"""Self-attention kernels: parity on a few shapes (incl. ragged / masked), then timing at the UNet's batch-16 shapes.
Usage: attn_tile_check.py [reps]"""
import sys, math, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from adaprompt_b200 import ops
import test_kernels_gpu as T

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B, heads = 16, 8

def bench(N, d):
    dp = 48 if d == 40 else d
    C = heads * d
    q = torch.randn(B * N, 2 * heads * dp, device="cuda").to(torch.bfloat16) * 0.3
    k = q[:, heads * dp:]
    vt = torch.randn(C, B * N, device="cuda").to(torch.bfloat16)
    o = torch.empty(B * N, C, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.attention(q, k, vt, o, B=B, heads=heads, Nq=N, d=d, ldq=2 * heads * dp, Nk=N, ldk=2 * heads * dp,
                              ldvt=B * N, kv_stride=N)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 4.0 * B * heads * N * N * d / ms / 1e9

errs = [("40/4096", T._attn_case(1, 4096, 4096, 40)), ("40/1000m", T._attn_case(2, 1000, 1000, 40, seed=3, mask=True)),
        ("40/296", T._attn_case(3, 296, 296, 40, seed=4)), ("80/1024", T._attn_case(2, 1024, 1024, 80)),
        ("80/1000m", T._attn_case(2, 1000, 1000, 80, seed=7, mask=True))]
torch.cuda.synchronize()
print("parity " + ("ok " if all(e < 6e-3 for _, e in errs) else "BAD ") + " ".join(f"{n}={e:.2e}" for n, e in errs))
for N, d in ((4096, 40), (9216, 40), (1024, 80), (2304, 80), (256, 160), (64, 160)):
    ms, tf = bench(N, d)
    print(f"N={N:5d} d={d:3d}: {ms:8.3f} ms  {tf:7.1f} TFLOP/s", flush=True)
