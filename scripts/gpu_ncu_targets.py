import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
def run(M, N, K, **kw):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    extra = {}
    if kw.get("bias"): extra["bias"] = torch.randn(N, device=dev)
    if kw.get("act"): extra["act"] = kw["act"]
    if kw.get("glu"): extra["glu"] = True
    n_out = N // 2 if kw.get("glu") else N
    od = kw.get("od", torch.bfloat16)
    if kw.get("resid"): extra["resid"] = torch.randn(M, n_out, device=dev)
    out = torch.empty(M, n_out, device=dev, dtype=od)
    ops.gemm(a, w, out=out, **extra)
    torch.cuda.synchronize()
run(32768, 576, 576, bias=True, resid=True, od=torch.float32)
run(32768, 2304, 576, bias=True, act="gelu")
run(1, 16384, 3072, glu=True, act="silu")
run(1, 3072, 8192, resid=True, od=torch.float32)
# memory attention shape: 1 head x 256, 4096 queries, 28k keys
q = torch.randn(1, 4096, 1, 256, device=dev).bfloat16(); k = torch.randn(1, 28700, 1, 256, device=dev).bfloat16(); v = torch.randn_like(k)
ops.attention(q, k, v)
torch.cuda.synchronize()
