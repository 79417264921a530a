"""One self-attention launch at the headline shape (B16, N4096, d40) for ncu."""
import sys, torch
sys.path.insert(0, ".")
from adaprompt_b200 import ops
B, heads, N, d, dp = 16, 8, 4096, 40, 48
q = torch.randn(B * N, 2 * heads * dp, device="cuda").to(torch.bfloat16) * 0.3
vt = torch.randn(heads * d, B * N, device="cuda").to(torch.bfloat16)
o = torch.empty(B * N, heads * d, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(q, q[:, heads * dp:], vt, o, B=B, heads=heads, Nq=N, Nk=N, d=d, ldq=2 * heads * dp, ldk=2 * heads * dp, ldvt=B * N, kv_stride=N)
torch.cuda.synchronize()
print("ok", o.float().abs().mean().item())
