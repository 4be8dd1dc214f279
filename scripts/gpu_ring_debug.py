import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
import torch.nn.functional as F
dev = "cuda"
def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
for (N, K, glu, af32) in [(3072, 3072, False, False), (9216, 3072, False, True), (16384, 3072, True, True), (3072, 8192, False, False), (32064, 3072, False, True)]:
    x = torch.randn(1, K, device=dev) * 3
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    if af32:
        xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5); a = x; kw = dict(a_rms_eps=1e-5)
    else:
        a = x.bfloat16(); xn = a.float(); kw = {}
    ref = xn @ w.float().T
    if glu:
        r = ref.reshape(1, -1, 2, 16); ref = (F.silu(r[:, :, 0]) * r[:, :, 1]).reshape(1, -1)
    y = ops.gemm(a, w, glu=glu, act="silu" if glu else None, out_dtype=torch.float32, **kw)
    torch.cuda.synchronize()
    print(N, K, glu, af32, "rel", rel(y, ref), flush=True)
