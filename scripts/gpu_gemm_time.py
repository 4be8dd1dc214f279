"""Times the hot GEMM shapes (CUDA events, 20 reps, L2 flushed by the working set itself for the big ones)."""
import os, sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
def t(M, N, K, **kw):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    extra = {}
    if kw.get("bias"): extra["bias"] = torch.randn(N, device=dev)
    if kw.get("act"): extra["act"] = kw["act"]
    if kw.get("glu"): extra["glu"] = True
    n_out = N // 2 if kw.get("glu") else N
    out = torch.empty(M, n_out, device=dev, dtype=kw.get("od", torch.bfloat16))
    if kw.get("resid"): out.normal_(); extra["resid"] = out
    for _ in range(3): ops.gemm(a, w, out=out, **extra)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gemm(a, w, out=out, **extra)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"{M}x{N}x{K} {kw}: {us:.1f} us  {2.0*M*N*K/us/1e6:.0f} TF/s", flush=True)
print("VG_GEMM_BACKOFF =", os.environ.get("VG_GEMM_BACKOFF", "0"))
t(32768, 2304, 576, bias=True, act="gelu")
t(32768, 576, 2304, bias=True, resid=True, od=torch.float32)
t(32768, 1728, 576, bias=True)
t(131072, 1152, 288, bias=True, act="gelu")
t(1696, 16384, 3072, glu=True, act="silu")
t(2050, 1408, 6144, bias=True, resid=True, od=torch.float32)
t(2050, 6144, 1408, bias=True, act="gelu")
t(4096, 256, 256, bias=True)
t(131072, 864, 288, bias=True)
t(524288, 576, 144, bias=True, act="gelu")
t(524288, 432, 144, bias=True)
t(131072, 288, 1152, bias=True, resid=True, od=torch.float32)
t(32768, 576, 576, bias=True, resid=True, od=torch.float32)
