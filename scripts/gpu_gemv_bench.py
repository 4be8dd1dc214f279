"""Kernel-only timing of the decode GEMV shapes (CUDA-graph replay of 20 launches removes the Python launch cost)."""
import os, sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
M = int(os.environ.get("VG_BENCH_M", "1"))
REPS = int(os.environ.get("VG_BENCH_REPS", "5"))
shapes = [(9216, 3072, False), (16384, 3072, True), (3072, 3072, False), (3072, 8192, False), (32064, 3072, False)]
if os.environ.get('VG_BENCH_SHAPE'):
    shapes = [shapes[int(os.environ['VG_BENCH_SHAPE'])]]
tot = 0.0
for N, K, glu in shapes:
    ws = [(torch.randn(N, K, device=dev) / K ** 0.5).bfloat16() for _ in range(4)]  # rotate weights: no L2 reuse (4 x >= 19 MB... mostly > L2 share)
    x = torch.randn(M, K, device=dev)
    out = torch.empty(M, N // 2 if glu else N, device=dev, dtype=torch.bfloat16)
    f = lambda: [ops.gemm(x, w, glu=glu, act="silu" if glu else None, out=out, a_rms_eps=1e-5) for w in ws for _ in range(REPS)]
    f(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (8 * REPS) * 1000
    gb = N * K * 2 / us / 1e3
    tot += us
    print(f"M={M} N={N} K={K} glu={glu}: {us:.1f} us  {gb:.0f} GB/s")
print("sum per layer-set us", tot)
