"""Micro-benchmark of af_gemm_bf16 / af_conv3x3_bf16 at UNet shapes. Usage: bench_gemm.py [case] [reps]"""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops
case = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
def timeit(f, flops, name):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms*1e3:9.1f} us  {flops/ms/1e9:8.1f} TFLOP/s", flush=True)
def gemm_case(M, N, K, res, f32, geglu=False, bn=0):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    No = N // 2 if geglu else N
    out = torch.empty(M, No, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    r = torch.randn(M, No, device="cuda") if res else None
    f = lambda: ops.gemm(a, w, out, bias=bias, residual=r, geglu=geglu, bn=bn)
    timeit(f, 2.0 * M * N * K, f"gemm M{M} N{N} K{K}{' res' if res else ''}{' f32' if f32 else ''}{' geglu' if geglu else ''} bn{bn}")
def conv_case(B, H, C, Co):
    x = torch.randn(B, H, H, C, device="cuda").to(torch.bfloat16)
    w = (torch.randn(Co, 3, 3, C, device="cuda") * (9 * C) ** -0.5).to(torch.bfloat16)
    bias = torch.randn(Co, device="cuda")
    out = torch.empty(B, H, H, Co, device="cuda")
    f = lambda: ops.conv3x3(x, w, out, bias=bias)
    timeit(f, 2.0 * B * H * H * Co * 9 * C, f"conv3x3 B{B} {H}x{H} C{C}->{Co}")
if case in ("all", "res"):
    gemm_case(65536, 320, 320, True, True)
if case in ("all", "geglu"):
    gemm_case(65536, 2560, 320, False, False, geglu=True)
if case == "f32":
    gemm_case(65536, 320, 320, False, True)
if case == "qk":
    gemm_case(65536, 768, 320, False, False)
if case == "conv8":
    conv_case(16, 8, 1280, 1280)
if case == "conv64":
    conv_case(16, 64, 320, 320)
if case == "all":
    gemm_case(65536, 320, 320, False, False)
    gemm_case(65536, 320, 320, False, True)
    gemm_case(16384, 640, 640, True, True)
    gemm_case(4096, 1280, 1280, True, True)
    gemm_case(65536, 768, 320, False, False)
    gemm_case(8192, 8192, 8192, False, False, bn=256)
    gemm_case(8192, 8192, 8192, False, False, bn=128)
    conv_case(16, 64, 320, 320)
    conv_case(16, 8, 1280, 1280)
if case == "bn":
    for bn in (0, 128):
        gemm_case(65536, 320, 320, True, True, bn=bn)
        gemm_case(65536, 320, 320, False, True, bn=bn)
        gemm_case(65536, 320, 320, False, False, bn=bn)
        gemm_case(16384, 640, 640, True, True, bn=bn)
        gemm_case(16384, 640, 640, False, False, bn=bn)
        gemm_case(4096, 1280, 1280, True, True, bn=bn)
        gemm_case(65536, 768, 320, False, False, bn=bn)
        gemm_case(65536, 384, 320, False, False, bn=bn)
        gemm_case(16384, 1280, 640, False, False, bn=bn)
        gemm_case(4096, 2560, 1280, False, False, bn=bn)
        gemm_case(320, 65536, 320, False, False, bn=bn)
if case == "vt":
    for bn in (128, 160, 256):
        gemm_case(320, 65536, 320, False, False, bn=bn)
        gemm_case(640, 16384, 640, False, False, bn=bn)
        gemm_case(1280, 4096, 1280, False, False, bn=bn)
