"""GPU bring-up check for the tcgen05 GEMM: compares against torch fp32 matmul of the same bf16 inputs."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops

torch.manual_seed(0)
dev = "cuda"
res = []

def ref(a, w, bias=None, act=None, glu=False, col_scale=None, resid=None, alpha=1.0):
    y = alpha * (a.float() @ w.float().transpose(-1, -2))
    if bias is not None: y = y + bias
    f = {None: lambda x: x, "gelu": torch.nn.functional.gelu, "quick_gelu": lambda x: x * torch.sigmoid(1.702 * x),
         "relu": torch.relu, "silu": torch.nn.functional.silu, "sigmoid": torch.sigmoid}[act]
    if glu:
        yb = y.reshape(*y.shape[:-1], -1, 2, 16)
        y = (f(yb[..., 0, :]) * yb[..., 1, :]).reshape(*y.shape[:-1], -1)
    else:
        y = f(y)
    if col_scale is not None: y = y * col_scale
    if resid is not None: y = y + resid.float()
    return y

def run(name, M, N, K, bn=0, batch=0, out_dtype=torch.float32, **kw):
    shp = (batch,) if batch else ()
    a = torch.randn(*shp, M, K, device=dev).bfloat16()
    w = (torch.randn(*shp, N, K, device=dev) / K ** 0.5).bfloat16()
    args = {}
    if kw.get("bias"): args["bias"] = torch.randn(N, device=dev)
    if kw.get("act"): args["act"] = kw["act"]
    if kw.get("glu"): args["glu"] = True
    n_out = N // 2 if kw.get("glu") else N
    if kw.get("col_scale"): args["col_scale"] = torch.randn(n_out, device=dev)
    if kw.get("resid"): args["resid"] = torch.randn(*shp, M, n_out, device=dev).to(kw["resid"])
    if kw.get("alpha"): args["alpha"] = kw["alpha"]
    y = ops.gemm(a, w, out_dtype=out_dtype, force_block_n=bn, **args)
    torch.cuda.synchronize()
    r = ref(a, w, **args)
    err = ((y.float() - r).norm() / r.norm()).item()
    mx = (y.float() - r).abs().max().item()
    tol = 2e-5 if out_dtype == torch.float32 else 5e-3
    ok = err < tol
    res.append((name, ok, err, mx))
    print(f"{'OK ' if ok else 'BAD'} {name}: M={M} N={N} K={K} bn={bn} batch={batch} rel={err:.3e} max={mx:.3e}", flush=True)
    return ok

first = run("basic128", 128, 128, 64, bn=128)
if not first:
    # dump a small diagnostic
    a = torch.eye(128, 64, device=dev).bfloat16(); w = torch.eye(128, 64, device=dev).bfloat16()
    y = ops.gemm(a, w, out_dtype=torch.float32, force_block_n=128)
    print(y[:8, :8]); print(y.abs().sum())
run("basic64", 128, 64, 64, bn=64)
run("basic256", 128, 256, 64, bn=256)
run("k256", 128, 128, 256, bn=128)
run("multi_tile", 512, 512, 512, bn=128)
run("multi_tile256", 1024, 1024, 1024, bn=256)
run("ragged", 300, 200, 136, bn=128)
run("ragged64", 77, 40, 72, bn=64)
run("ragged256", 1000, 1000, 1000, bn=256)
run("persist_many", 4096, 4096, 512, bn=256)
run("auto_small", 9, 256, 256)
run("auto_big", 1664, 9216, 3072)
run("bias_gelu", 256, 384, 256, bias=True, act="gelu")
run("bias_quick", 256, 384, 256, bias=True, act="quick_gelu")
run("glu_silu", 256, 512, 256, glu=True, act="silu")
run("scale_resid_f32", 256, 384, 256, bias=True, col_scale=True, resid=torch.float32)
run("resid_bf16_out_bf16", 256, 384, 256, resid=torch.bfloat16, out_dtype=torch.bfloat16)
run("out_bf16", 384, 1024, 512, out_dtype=torch.bfloat16)
run("alpha", 128, 128, 128, alpha=0.125)
run("batched", 200, 96, 64, batch=5)
run("batched_big", 4096, 32, 32, batch=4)
# naive kernel agreement
a = torch.randn(333, 200, device=dev).bfloat16(); w = torch.randn(129, 200, device=dev).bfloat16()
y1 = ops.gemm(a, w, out_dtype=torch.float32); y2 = ops.gemm(a, w, out_dtype=torch.float32, naive=True)
print("naive vs tcgen05 max diff", (y1 - y2).abs().max().item())

# timing
def bench(M, N, K, bn=0, iters=20):
    a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, w, out=out, force_block_n=bn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.gemm(a, w, out=out, force_block_n=bn)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2 * M * N * K / ms / 1e9
    print(f"bench M={M} N={N} K={K} bn={bn}: {ms:.3f} ms  {tf:.1f} TFLOP/s", flush=True)
    return tf
for shp in [(8192, 8192, 8192, 256), (8192, 8192, 8192, 128), (1664, 9216, 3072, 0), (1664, 16384, 3072, 0), (1664, 3072, 8192, 0),
            (65536, 432, 144, 0), (4096, 1728, 576, 0), (4616, 4096, 1024, 0)]:
    bench(*shp)
    # torch reference timing
M, N, K = 8192, 8192, 8192
a = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
for _ in range(3): a @ w.T
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): a @ w.T
e1.record(); torch.cuda.synchronize()
print("cublas 8192^3 TFLOP/s", 2 * M * N * K / (e0.elapsed_time(e1) / 20) / 1e9)
bad = [r for r in res if not r[1]]
print("FAILED:" if bad else "ALL OK", bad)
