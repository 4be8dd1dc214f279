import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
B, H, Nq, Nk, D = 4, 8, 4096, 4096, 72
q = torch.randn(B, Nq, H, D, device="cuda").bfloat16(); k = torch.randn(B, Nk, H, D, device="cuda").bfloat16(); v = torch.randn(B, Nk, H, D, device="cuda").bfloat16()
for _ in range(2): ops.attention(q, k, v)
torch.cuda.synchronize()
