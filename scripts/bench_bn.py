"""N-tile width (bn hint) x split-K for the conv / GEMM shapes of the UNet at batch 16: microseconds per launch."""
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
bns = (0, 128, 160, 256)
print("shape".ljust(40) + "".join(f"{'bn=%d' % b if b else 'auto':>9s}" for b in bns) + "   (split auto) |" + "".join(f"{'bn=%d' % b if b else 'auto':>9s}" for b in bns) + "   (split off)")
for B, H, C0, C1, Cout, stride in ((16, 8, 1280, 0, 1280, 1), (16, 8, 1280, 1280, 1280, 1), (16, 16, 1280, 0, 1280, 1), (16, 16, 1280, 1280, 1280, 1),
                                   (16, 16, 1280, 640, 1280, 1), (16, 16, 640, 0, 1280, 1), (16, 32, 640, 0, 640, 1), (16, 32, 1280, 0, 1280, 1),
                                   (16, 32, 1280, 640, 640, 1), (16, 32, 640, 320, 640, 1), (16, 64, 320, 0, 320, 1), (16, 64, 640, 0, 640, 1),
                                   (16, 64, 320, 320, 320, 1), (16, 64, 640, 320, 320, 1), (16, 64, 320, 0, 320, 2), (16, 32, 640, 0, 640, 2)):
    x0 = torch.randn(B, H, H, C0, device=dev).bfloat16()
    x1 = torch.randn(B, H, H, C1, device=dev).bfloat16() if C1 else None
    w = pack_conv3x3((torch.randn(Cout, C0 + C1, 3, 3, device=dev) * 0.01).bfloat16())
    Ho = H // stride
    out = torch.empty(B, Ho, Ho, Cout, device=dev)
    res = torch.randn(B, Ho, Ho, Cout, device=dev)
    row = f"conv B{B} {H}x{H} C{C0 + C1}->{Cout} s{stride}".ljust(40)
    for sk in (0, 1):
        for bn in bns:
            with ops.launch_options(split_k=sk):
                row += f"{timeit(lambda: ops.conv3x3(x0, w, out, x1=x1, stride=stride, residual=res, bn=bn)):9.1f}"
        row += "                  "
    print(row, flush=True)
for M, N, K, res in ((4096, 1280, 1280, True), (4096, 1280, 5120, True), (4096, 2560, 1280, False), (16384, 640, 640, True),
                     (16384, 640, 2560, True), (16384, 1280, 640, False), (65536, 320, 320, True), (65536, 320, 1280, True), (65536, 768, 320, False)):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if res else torch.bfloat16)
    r = torch.randn(M, N, device=dev) if res else None
    row = f"gemm M{M} N{N} K{K}{' res' if res else ''}".ljust(40)
    for sk in (0, 1):
        for bn in bns:
            with ops.launch_options(split_k=sk):
                row += f"{timeit(lambda: ops.gemm(a, w, out, residual=r, bn=bn)):9.1f}"
        row += "                  "
    print(row, flush=True)
