import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
torch.manual_seed(0)
dev = "cuda"
def ref(q, k, v, scale, causal):
    q, k, v = [t.float().permute(0, 2, 1, 3) for t in (q, k, v)]
    s = q @ k.transpose(-1, -2) * scale
    if causal:
        Nq, Nk = s.shape[-2:]
        i = torch.arange(Nq, device=s.device)[:, None] + (Nk - Nq); j = torch.arange(Nk, device=s.device)[None]
        s = s.masked_fill(j > i, float("-inf"))
    return (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3)
def run(B, H, Nq, Nk, D, causal=False):
    q = torch.randn(B, Nq, H, D, device=dev).bfloat16(); k = torch.randn(B, Nk, H, D, device=dev).bfloat16(); v = torch.randn(B, Nk, H, D, device=dev).bfloat16()
    o = ops.attention(q, k, v, causal=causal); torch.cuda.synchronize()
    r = ref(q, k, v, D ** -0.5, causal)
    err = ((o.float() - r).norm() / r.norm()).item()
    ol = ops.attention(q, k, v, causal=causal, legacy=True); torch.cuda.synchronize()
    errl = ((ol.float() - r).norm() / r.norm()).item()
    print(f"B={B} H={H} Nq={Nq} Nk={Nk} D={D} causal={causal}: tc rel={err:.2e} legacy rel={errl:.2e} {'OK' if err < 6e-3 else 'BAD'}", flush=True)
    return q, k, v
cases = [(1, 1, 128, 128, 64), (1, 1, 128, 128, 128), (1, 2, 256, 384, 64), (1, 1, 128, 128, 256), (1, 1, 256, 512, 256), (2, 4, 300, 700, 96), (1, 8, 300, 300, 72),
         (1, 16, 1025, 1025, 88), (1, 16, 577, 577, 64)]
only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for i, c in enumerate(cases):
    if only is None or only == i: run(*c)
if only is None:
    run(1, 32, 333, 333, 96, True); run(1, 4, 1700, 1700, 96, True); run(1, 1, 4096, 8300, 256); run(2, 2, 300, 5000, 64)
    def bench(B, H, Nq, Nk, D, causal=False, legacy=False, iters=10):
        q = torch.randn(B, Nq, H, D, device=dev).bfloat16(); k = torch.randn(B, Nk, H, D, device=dev).bfloat16(); v = torch.randn(B, Nk, H, D, device=dev).bfloat16()
        for _ in range(2): ops.attention(q, k, v, causal=causal, legacy=legacy)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): ops.attention(q, k, v, causal=causal, legacy=legacy)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = 4.0 * B * H * Nq * Nk * D * (0.5 if causal else 1.0)
        print(f"bench B={B} H={H} Nq={Nq} Nk={Nk} D={D} causal={causal} legacy={legacy}: {ms*1000:.0f} us {fl/ms/1e9:.0f} TFLOP/s", flush=True)
    for leg in (False, True):
        bench(1, 1, 4096, 28700, 256, legacy=leg); bench(8, 8, 4096, 4096, 72, legacy=leg); bench(1, 32, 1696, 1696, 96, True, legacy=leg)
        bench(2, 16, 1025, 1025, 88, legacy=leg); bench(8, 16, 577, 577, 64, legacy=leg)
