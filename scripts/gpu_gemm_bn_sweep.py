"""Tile-width sweep (force_block_n) over the small-K Hiera GEMM shapes against the heuristic's choice (0)."""
import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
def t(M, N, K, bn, act=None, resid=False):
    a = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16(); bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.float32 if resid else torch.bfloat16)
    kw = dict(bias=bias, out=out)
    if act: kw["act"] = act
    if resid: out.normal_(); kw["resid"] = out
    if bn: kw["force_block_n"] = bn
    for _ in range(3): ops.gemm(a, w, **kw)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gemm(a, w, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1e3
shapes = [(131072, 864, 288, None, False), (131072, 1152, 288, "gelu", False), (131072, 288, 1152, None, True), (131072, 288, 288, None, True),
          (524288, 432, 144, None, False), (524288, 576, 144, "gelu", False), (524288, 144, 576, None, True), (524288, 144, 144, None, True),
          (32768, 576, 576, None, True), (32768, 576, 2304, None, True), (32768, 1728, 576, None, False), (32768, 2304, 576, "gelu", False),
          (8192, 1152, 4608, None, True), (8192, 4608, 1152, "gelu", False), (8192, 3456, 1152, None, False)]
for M, N, K, act, resid in shapes:
    row = []
    for bn in (0, 64, 96, 128, 192, 256):
        try:
            row.append(f"{bn}:{t(M, N, K, bn, act, resid):.1f}")
        except Exception as e:
            row.append(f"{bn}:err")
    print(f"{M}x{N}x{K} act={act} resid={resid}: " + "  ".join(row), flush=True)
