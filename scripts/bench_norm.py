"""Micro-benchmark of af_groupnorm_apply / af_layernorm at the UNet's shapes (batch 16). Usage: bench_norm.py [reps]"""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
def timeit(f, nbytes, name):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(reps):
        flush.zero_()                      # evict the 126 MB L2 between repetitions
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    print(f"{name:40s} {ms*1e3:8.1f} us  {nbytes/ms/1e6:8.1f} GB/s", flush=True)
B = 16
for HW, C in ((4096, 320), (4096, 640), (4096, 960), (1024, 640), (1024, 1280), (1024, 1920), (256, 1280), (256, 2560), (64, 1280), (64, 2560)):
    side = int(HW ** 0.5)
    x = torch.randn(B, side, side, C, device="cuda")
    slots = HW // 32                                  # what a GEMM / conv epilogue produces (one slot per 32 rows)
    xs = x.view(B, slots, 32, C)
    st = ops.GNStats(torch.stack((xs.sum(2), (xs * xs).sum(2)), dim=-1).contiguous().view(-1), slots, C)
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    y = torch.empty(B, side, side, C, device="cuda", dtype=torch.bfloat16)
    timeit(lambda: ops.groupnorm_apply(x, st, g, b, 1e-5, True, y), B * HW * C * 6, f"groupnorm finalize + apply B{B} HW{HW} C{C}")
for rows, C in ((65536, 320), (16384, 640), (4096, 1280)):
    x = torch.randn(rows, C, device="cuda")
    g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
    y = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16)
    timeit(lambda: ops.layernorm(x, g, b, 1e-5, y), rows * C * 6, f"layernorm rows{rows} C{C}")
