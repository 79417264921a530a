"""One short-context cross-attention launch at the headline shape (B16, N4096, 77 keys, d40) for ncu."""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops
B, heads, N, d, dp, Nk, nkp = 16, 8, 4096, 40, 48, 77, 80
q = torch.randn(B * N, heads * dp, device="cuda").to(torch.bfloat16) * 0.3
k = torch.randn(B * nkp, heads * dp, device="cuda").to(torch.bfloat16)
vt = torch.randn(heads * d, B * nkp, device="cuda").to(torch.bfloat16)
o = torch.empty(B * N, heads * d, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(q, k, vt, o, B=B, heads=heads, Nq=N, Nk=Nk, d=d, ldq=heads * dp, ldk=heads * dp, ldvt=B * nkp, kv_stride=nkp)
torch.cuda.synchronize()
print("ok", o.float().abs().mean().item())
