import sys, math, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops
B, N, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
heads, C = 8, 8 * d
dp = 48 if d == 40 else d
g = torch.Generator().manual_seed(0)
q = torch.randn(B, N, heads, d, generator=g).cuda(); k = torch.randn(B, N, heads, d, generator=g).cuda(); v = torch.randn(B, N, heads, d, generator=g).cuda()
qs = (q * (d ** -0.5 * math.log2(math.e))).to(torch.bfloat16); kb = k.to(torch.bfloat16); vb = v.to(torch.bfloat16)
qk = torch.zeros(B * N, 2, heads, dp, device="cuda", dtype=torch.bfloat16)
qk[:, 0, :, :d] = qs.reshape(B * N, heads, d); qk[:, 1, :, :d] = kb.reshape(B * N, heads, d)
qk = qk.reshape(B * N, 2 * heads * dp)
ldvt = max(64, B * N)
vt = torch.zeros(C, ldvt, device="cuda", dtype=torch.bfloat16)
vt[:, :B * N] = vb.permute(2, 3, 0, 1).reshape(C, B * N)
out = torch.empty(B * N, C, device="cuda", dtype=torch.bfloat16)
ops.attention(qk, qk[:, heads * dp:], vt, out, B=B, heads=heads, Nq=N, Nk=N, d=d, ldq=2 * heads * dp, ldk=2 * heads * dp, ldvt=ldvt, kv_stride=N)
torch.cuda.synchronize()
s = torch.einsum("bihd,bjhd->bhij", qs.float(), kb.float())
p = torch.softmax(s * math.log(2.0), dim=-1)
ref = torch.einsum("bhij,bjhd->bihd", p, vb.float()).reshape(B * N, C)
print("OK", B, N, d, ((out.float() - ref).norm() / ref.norm()).item())
