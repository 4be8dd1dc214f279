"""Timing of the tcgen05 attention at the config-2 shapes (VG_ATTN_TC2=0: one query tile per block, two blocks per SM; =1: two query tiles per block)."""
import os, sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
def bench(B, H, Nq, Nk, D, causal=False, window=0, iters=10):
    q = torch.randn(B, Nq, H, D, device=dev).bfloat16(); k = torch.randn(B, Nk, H, D, device=dev).bfloat16(); v = torch.randn(B, Nk, H, D, device=dev).bfloat16()
    for _ in range(2): ops.attention(q, k, v, causal=causal, window=window)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): ops.attention(q, k, v, causal=causal, window=window)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = 4.0 * B * H * Nq * Nk * D * (0.5 if causal else 1.0)
    print(f"TC2={os.environ.get('VG_ATTN_TC2', '1')} B={B} H={H} Nq={Nq} Nk={Nk} D={D} causal={causal} window={window}: {ms*1000:.0f} us {fl/ms/1e9:.0f} TFLOP/s", flush=True)
bench(8, 8, 4096, 4096, 72); bench(1, 32, 1696, 1696, 96, True); bench(1, 32, 3360, 3360, 96, True, 2047)
bench(2, 16, 1025, 1025, 88); bench(8, 16, 577, 577, 64); bench(128, 8, 256, 256, 72)
