"""A/B of the self-attention schedules: parity of each variant on a few shapes (incl. ragged / masked), then timing at
the UNet's batch-16 shapes.  Usage: attn_tile_check.py v1,v2,... [reps]
variant < 16: round-1 pair / row-split kernels; 16 + x: attention_tile.cu (bit 0 two threads per row, bits 1-2 FMA-pipe
share of the exponentials); + 32: d = 80 through the tile kernel as well."""
import sys, math, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from adaprompt_b200 import ops, _lib
import test_kernels_gpu as T

variants = [int(v) for v in sys.argv[1].split(",")] if len(sys.argv) > 1 else [8, 16, 17]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = _lib.load()
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
    fl = 4.0 * B * heads * N * N * d
    return ms, fl / ms / 1e9

for var in variants:
    lib.af_attention_set_pair_variant(var)
    errs = []
    try:
        errs.append(("40/4096", T._attn_case(1, 4096, 4096, 40)))
        errs.append(("40/1000m", T._attn_case(2, 1000, 1000, 40, seed=3, mask=True)))
        errs.append(("40/296", T._attn_case(3, 296, 296, 40, seed=4)))
        errs.append(("80/1024", T._attn_case(2, 1024, 1024, 80)))
        errs.append(("80/1000m", T._attn_case(2, 1000, 1000, 80, seed=7, mask=True)))
    except Exception as e:  # noqa
        print(f"variant {var}: FAILED {e!r}", flush=True)
        continue
    torch.cuda.synchronize()
    ok = all(e < 6e-3 for _, e in errs)
    t40 = bench(4096, 40)
    t80 = bench(1024, 80)
    print(f"variant {var:3d}: parity {'ok ' if ok else 'BAD'} " + " ".join(f"{n}={e:.2e}" for n, e in errs) +
          f" | N4096 d40 {t40[0]:.3f} ms {t40[1]:.0f} TF/s | N1024 d80 {t80[0]:.3f} ms {t80[1]:.0f} TF/s", flush=True)
