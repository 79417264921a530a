import sys, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from adaprompt_b200 import ops, _lib
lib = _lib.load()
var = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
lib.af_attention_set_pair_variant(var)
DEV = "cuda"; heads, d, dp = 8, 40, 48
g = torch.Generator().manual_seed(1)
q = torch.randn(B, N, heads, d, generator=g).to(DEV); k = torch.randn(B, N, heads, d, generator=g).to(DEV); v = torch.randn(B, N, heads, d, generator=g).to(DEV)
qs = (q * (d ** -0.5 * math.log2(math.e))).to(torch.bfloat16); kb, vb = k.to(torch.bfloat16), v.to(torch.bfloat16)
qbuf = torch.zeros(B, N, heads, dp, device=DEV, dtype=torch.bfloat16); kbuf = torch.zeros_like(qbuf)
qbuf[..., :d] = qs; kbuf[..., :d] = kb
C = heads * d
vt = torch.zeros(C, B * N, device=DEV, dtype=torch.bfloat16)
vt.view(C, B, N)[:] = vb.permute(2, 3, 0, 1).reshape(C, B, N)
s = torch.einsum("bihd,bjhd->bhij", qs.float(), kb.float())
p = torch.softmax(s * math.log(2.0), dim=-1)
ref = torch.einsum("bhij,bjhd->bihd", p, vb.float())
for rep in range(3):
    out = torch.empty(B, N, C, device=DEV, dtype=torch.bfloat16)
    ops.attention(qbuf, kbuf, vt, out, B=B, heads=heads, Nq=N, Nk=N, d=d, ldq=heads * dp, ldk=heads * dp, ldvt=B * N, kv_stride=N)
    torch.cuda.synchronize()
    o = out.float().view(B, N, heads, d)
    err = (o - ref).pow(2).sum(-1).sqrt() / ref.pow(2).sum(-1).sqrt().clamp_min(1e-6)   # [B, N, heads]
    bad = err > 0.02
    print(f"rep {rep}: rel {((o-ref).norm()/ref.norm()).item():.3e}  bad rows {int(bad.sum())} of {bad.numel()}  nan {int(torch.isnan(o).sum())}")
    if bad.any():
        idx = bad.nonzero()
        print("  bad per head:", torch.bincount(idx[:, 2], minlength=heads).tolist())
        print("  bad per q-tile:", torch.bincount(idx[:, 1] // 128, minlength=N // 128).tolist())
        print("  bad per lane quarter:", torch.bincount((idx[:, 1] % 128) // 32, minlength=4).tolist())
        print("  first few:", idx[:8].tolist(), err[bad][:8].tolist())
