import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
M, N, K = 131072, 864, 288
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16(); bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(2): ops.gemm(a, w, bias=bias, out=out)
torch.cuda.synchronize()
