"""CTA-pair (cta_group::2) vs single-CTA tiles for the conv / GEMM shapes of the UNet at batch 16: microseconds per launch."""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops
from adaprompt_b200.packing import pack_conv3x3
dev = "cuda"
def timeit(f, reps=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3
print("shape".ljust(44) + "   auto   never  always")
for B, H, C0, C1, Cout in ((16, 64, 320, 0, 320), (16, 64, 640, 0, 640), (16, 64, 640, 320, 320), (16, 32, 640, 0, 640), (16, 32, 1280, 0, 1280),
                           (16, 32, 1280, 640, 640), (16, 16, 1280, 0, 1280), (16, 16, 1280, 1280, 1280), (16, 8, 1280, 0, 1280)):
    x0 = torch.randn(B, H, H, C0, device=dev).bfloat16()
    x1 = torch.randn(B, H, H, C1, device=dev).bfloat16() if C1 else None
    w = pack_conv3x3((torch.randn(Cout, C0 + C1, 3, 3, device=dev) * 0.01).bfloat16())
    out = torch.empty(B, H, H, Cout, device=dev)
    res = torch.randn(B, H, H, Cout, device=dev)
    row = f"conv B{B} {H}x{H} C{C0 + C1}->{Cout}".ljust(44)
    for mode in (0, 1, 2):
        with ops.launch_options(pair_mode=mode):
            row += f"{timeit(lambda: ops.conv3x3(x0, w, out, x1=x1, residual=res)):8.1f}"
    print(row, flush=True)
for M, N, K, res, geglu in ((65536, 320, 320, True, False), (65536, 768, 320, False, False), (16384, 640, 640, True, False), (4096, 1280, 1280, True, False),
                            (4096, 1280, 5120, True, False), (16384, 640, 2560, True, False), (65536, 320, 1280, True, False), (4096, 2560, 1280, False, False)):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if res else torch.bfloat16)
    r = torch.randn(M, N, device=dev) if res else None
    row = f"gemm M{M} N{N} K{K}{' res' if res else ''}".ljust(44)
    for mode in (0, 1, 2):
        with ops.launch_options(pair_mode=mode):
            row += f"{timeit(lambda: ops.gemm(a, w, out, residual=r)):8.1f}"
    print(row, flush=True)
