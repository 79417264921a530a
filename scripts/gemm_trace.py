"""Device timeline of CTA 0 of af_gemm_bf16. Usage: gemm_trace.py M N K [res] [f32]"""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops, _lib
M, N, K = (int(v) for v in sys.argv[1:4])
res, f32 = "res" in sys.argv, "f32" in sys.argv
lib = _lib.load()
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
r = torch.randn(M, N, device="cuda") if res else None
f = lambda: ops.gemm(a, w, out, bias=bias, residual=r)
for _ in range(3): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print(f"M{M} N{N} K{K} res={res} f32={f32}: {e0.elapsed_time(e1) * 100:.1f} us per launch")
buf = torch.zeros(4 * 32 * 8, dtype=torch.int64, device="cuda")
with ops.launch_options(trace=buf):
    f()
torch.cuda.synchronize()
t = buf.cpu().view(4, 32, 8)
t0 = int(t[t > 0].min())
names = (("prod", 3), ("mma", 4), ("epi", 6))
for tile in range(8):
    for a_, (nm, n) in enumerate(names):
        ev = [int(x) - t0 if x > 0 else -1 for x in t[a_, tile, :n]]
        print(f"tile {tile} {nm:4s} " + " ".join(f"{e:7d}" for e in ev))
ep = t[2]
valid = [i for i in range(1, 31) if ep[i + 1, 0] > 0]
if valid:
    import statistics as st
    per = st.mean([int(ep[i + 1, 0] - ep[i, 0]) for i in valid])
    print("epilogue warp 0, mean cycles: tile period", per,
          "| wait acc", st.mean([int(ep[i, 1] - ep[i, 0]) for i in valid]),
          "| first chunk: ld", st.mean([int(ep[i, 2] - ep[i, 1]) for i in valid]),
          "math+sts", st.mean([int(ep[i, 3] - ep[i, 2]) for i in valid]),
          "store+next", st.mean([int(ep[i, 4] - ep[i, 3]) for i in valid]),
          "(ld split: issue+bias", st.mean([int(ep[i, 6] - ep[i, 1]) for i in valid]),
          "residual wait", st.mean([int(ep[i, 7] - ep[i, 6]) for i in valid]) if res else 0,
          "tmem wait", st.mean([int(ep[i, 2] - (ep[i, 7] if res else ep[i, 6])) for i in valid]), ")",
          "| rest of tile", st.mean([int(ep[i, 5] - ep[i, 4]) for i in valid]))
    x = t[3]
    print("first chunk detail: before ld issue", st.mean([int(x[i, 0] - ep[i, 1]) for i in valid]),
          "| ld issue", st.mean([int(x[i, 1] - x[i, 0]) for i in valid]),
          "| bias loads", st.mean([int(ep[i, 6] - x[i, 1]) for i in valid]),
          "| math (tmem wait -> o[] ready)", st.mean([int(x[i, 2] - ep[i, 2]) for i in valid]),
          "| st.shared", st.mean([int(x[i, 3] - x[i, 2]) for i in valid]),
          "| fence.proxy.async", st.mean([int(x[i, 4] - x[i, 3]) for i in valid]),
          "| syncwarp(+stats)", st.mean([int(ep[i, 3] - x[i, 4]) for i in valid]))
    mm = t[1]
    print("MMA issuer, mean cycles: wait acc free", st.mean([int(mm[i, 1] - mm[i, 0]) for i in valid]),
          "| wait first stage", st.mean([int(mm[i, 2] - mm[i, 1]) for i in valid]),
          "| k loop", st.mean([int(mm[i, 3] - mm[i, 2]) for i in valid]),
          "of which waiting for operands", st.mean([int(mm[i, 4]) for i in valid]))
