"""pair (cta_group::2) vs single-CTA schedule across N-tile sizes for conv / GEMM shapes of the UNet."""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import _lib, ops
lib = _lib.load()
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def conv(B, H, C, Co, bns):
    x = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(Co, 3, 3, C, device="cuda") * (9 * C) ** -0.5).to(torch.bfloat16)
    out = torch.empty(B, H, H, Co, device="cuda")
    fl = 2.0 * B * H * H * Co * 9 * C
    for bn in bns:
        row = f"conv B{B} {H}x{H} C{C}->{Co} bn{bn:3d}:"
        for mode in (0, 1):
            lib.af_gemm_set_pair_mode(mode)
            ms = timeit(lambda: ops.conv3x3(x, w, out, bn=bn))
            row += f"  pair={mode} {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s"
        print(row, flush=True)
def gemm(M, N, K, bns, res=False):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if res else torch.bfloat16)
    r = torch.randn(M, N, device="cuda") if res else None
    fl = 2.0 * M * N * K
    for bn in bns:
        row = f"gemm M{M} N{N} K{K}{' res' if res else ''} bn{bn:3d}:"
        for mode in (0, 1):
            lib.af_gemm_set_pair_mode(mode)
            ms = timeit(lambda: ops.gemm(a, w, out, residual=r, bn=bn))
            row += f"  pair={mode} {ms*1e3:7.1f} us {fl/ms/1e9:7.1f} TF/s"
        print(row, flush=True)
conv(16, 64, 320, 320, (64, 128, 160))
conv(16, 32, 640, 640, (128, 160))
conv(16, 16, 1280, 1280, (128, 160, 256))
conv(16, 8, 1280, 1280, (64, 128, 160, 256))
conv(16, 16, 2560, 1280, (160, 256))
gemm(65536, 320, 1280, (64, 160))
gemm(16384, 640, 2560, (128, 160))
gemm(4096, 1280, 5120, (128, 160, 256))
gemm(65536, 320, 320, (64, 160), res=True)
gemm(8192, 8192, 8192, (128, 160, 256))
