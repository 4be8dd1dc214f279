import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
def bench(M, N, K, iters=20, **kw):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    extra = {}
    if kw.get("bias"): extra["bias"] = torch.randn(N, device=dev)
    if kw.get("act"): extra["act"] = kw["act"]
    od = kw.get("od", torch.bfloat16)
    if kw.get("resid"): extra["resid"] = torch.randn(M, N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=od)
    bn = kw.get("bn", 0)
    for _ in range(3): ops.gemm(a, w, out=out, force_block_n=bn, **extra)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(a, w, out=out, force_block_n=bn, **extra)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"M={M} N={N} K={K} {kw}: {ms*1000:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
for act in (None, "gelu", "relu", "quick_gelu"):
    bench(32768, 2304, 576, bias=True, act=act)
for bn in (64, 128, 256):
    bench(32768, 2304, 576, bias=True, act="gelu", bn=bn)
    bench(32768, 2304, 576, bn=bn)
bench(32768, 1728, 576, bias=True)
bench(32768, 576, 576, bias=True, resid=True, od=torch.float32)
bench(32768, 576, 2304, bias=True, resid=True, od=torch.float32)
bench(131072, 1152, 288, bias=True, act="gelu")
bench(131072, 1152, 288, bias=True)
bench(524288, 576, 144, bias=True, act="gelu")
bench(524288, 576, 144, bias=True)
bench(524288, 432, 144, bias=True)
for M, N, K in [(1, 3072, 8192), (1, 16384, 3072), (1, 9216, 3072), (1, 3072, 3072), (1, 32064, 3072)]:
    bench(M, N, K, iters=50, od=torch.float32)
