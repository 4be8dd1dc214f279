"""Launches ONE instance of each hot kernel at its config-2 shape, in a fixed order, for `ncu --set full` (one GPU, no bench number
is taken from this run). Order (= ncu launch ids):
  0 gemm  Hiera stage-3 fc1    32768 x 2304 x 576   bias+GELU, bf16 out
  1 gemm  Hiera stage-3 fc2    32768 x 576 x 2304   bias + in-place fp32 residual
  2 gemm  Phi-3 prefill gu     1696 x 16384 x 3072  SwiGLU epilogue
  3 gemm  Phi-3 prefill down   1696 x 3072 x 8192   in-place fp32 residual
  4 gemv  Phi-3 decode gu      1 x 16384 x 3072     fused RMSNorm + SwiGLU
  5 gemv  Phi-3 decode down    1 x 3072 x 8192      residual
  6 attn  Phi-3 prefill        1696 x 1696, 32 heads x 96, causal (packed qkv cache layout)
  7 attn  SAM2 memory attn     4096 q x 28700 keys, 1 head x 256 (split-K + merge = 2 launches)
  9 attn  Hiera global         8 imgs x 4096 tokens x 8 heads x 72
 10 attn  Hiera stage-3 window 128 windows x 256 tokens x 8 heads x 72
 11 attn  Hiera stage-2 window 8192 windows x 16 tokens x 4 heads x 72 (window_attn_small_kernel)
 12 attn  Phi-3 decode         1 query x 1709 keys, 32 heads x 96
 13 gemv  batched decode gu    4 x 16384 x 3072     fused RMSNorm + SwiGLU (gemv_ring_mma_kernel)
 14 gemv  batched decode down  4 x 3072 x 8192      residual (gemv_ring_mma_kernel)
 15, 16   fused SAM2 mask decoder HW = 4096, B = 1 (split-bf16 tensor-core variant, then fp32-FMA variant)
Run under `ncu -k regex:"gemm_tcgen05|gemv_|flash_attn|window_attn|attn_decode|mask_decoder"` so that only these launches are profiled.
"""
import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import ops
dev = "cuda"
BF = torch.bfloat16
def run(M, N, K, **kw):
    a = torch.randn(M, K, device=dev).to(torch.float32 if kw.get("rms") else BF); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    extra = {}
    if kw.get("bias"): extra["bias"] = torch.randn(N, device=dev)
    if kw.get("act"): extra["act"] = kw["act"]
    if kw.get("glu"): extra["glu"] = True
    if kw.get("rms"): extra["a_rms_eps"] = 1e-5
    n_out = N // 2 if kw.get("glu") else N
    od = kw.get("od", BF)
    out = torch.empty(M, n_out, device=dev, dtype=od)
    if kw.get("resid"): out.normal_(); extra["resid"] = out
    ops.gemm(a, w, out=out, **extra)
    torch.cuda.synchronize()
run(32768, 2304, 576, bias=True, act="gelu")
run(32768, 576, 2304, bias=True, resid=True, od=torch.float32)
run(1696, 16384, 3072, glu=True, act="silu")
run(1696, 3072, 8192, resid=True, od=torch.float32)
run(1, 16384, 3072, glu=True, act="silu", rms=True)
run(1, 3072, 8192, resid=True, od=torch.float32)
# Phi-3 prefill attention on the packed [L, 3*hidden] cache rows
L, H, hd, D = 1696, 32, 96, 3072
cache = torch.randn(2048, 3 * D, device=dev).bfloat16()
q = cache[:L, :D].unflatten(1, (H, hd))[None]; k = cache[:L, D:2 * D].unflatten(1, (H, hd))[None]; v = cache[:L, 2 * D:].unflatten(1, (H, hd))[None]
ops.attention(q, k, v, scale=hd ** -0.5, causal=True, window=2047)
torch.cuda.synchronize()
q = torch.randn(1, 4096, 1, 256, device=dev).bfloat16(); k = torch.randn(1, 28700, 1, 256, device=dev).bfloat16(); v = torch.randn_like(k)
ops.attention(q, k, v)
torch.cuda.synchronize()
qkv = torch.randn(8 * 4096, 3 * 576, device=dev).bfloat16()
q, k, v = (qkv[:, j * 576:(j + 1) * 576].unflatten(1, (8, 72)) for j in range(3))
ops.attention(q.unflatten(0, (8, 4096)), k.unflatten(0, (8, 4096)), v.unflatten(0, (8, 4096)), scale=72 ** -0.5)
torch.cuda.synchronize()
ops.attention(q[None], k[None], v[None], scale=72 ** -0.5, q_win=(16, 64, 4, 4), k_win=(16, 64, 4, 4))
torch.cuda.synchronize()
# Hiera stage-2 windows: 8 imgs x 128x128 tokens, 4x4 windows, 4 heads x 72 (one block per window, all heads)
qkv2 = torch.randn(8 * 128 * 128, 3 * 288, device=dev).bfloat16()
q2, k2, v2 = (qkv2[:, j * 288:(j + 1) * 288].unflatten(1, (4, 72)) for j in range(3))
ops.attention(q2[None], k2[None], v2[None], scale=72 ** -0.5, q_win=(4, 128, 32, 32), k_win=(4, 128, 32, 32))
torch.cuda.synchronize()
nk = torch.tensor([1709], device=dev, dtype=torch.int32)
ws = ops.decode_workspace(1, H, hd, dev)
q1 = torch.randn(1, 3 * D, device=dev).bfloat16()[:, :D].unflatten(1, (H, hd))[None]
ops.attention(q1, cache[:2048, D:2 * D].unflatten(1, (H, hd))[None], cache[:2048, 2 * D:].unflatten(1, (H, hd))[None], scale=hd ** -0.5, causal=True, window=2047,
              nk_dev=nk, decode_ws=ws)
torch.cuda.synchronize()
def run_m(M, N, K, **kw):
    a = torch.randn(M, K, device=dev).to(torch.float32 if kw.get("rms") else BF); w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    extra = {"a_rms_eps": 1e-5} if kw.get("rms") else {}
    if kw.get("glu"): extra.update(glu=True, act="silu")
    out = torch.randn(M, N // 2 if kw.get("glu") else N, device=dev).to(kw.get("od", BF))
    if kw.get("resid"): extra["resid"] = out
    ops.gemm(a, w, out=out, **extra)
    torch.cuda.synchronize()
run_m(4, 16384, 3072, glu=True, rms=True)
run_m(4, 3072, 8192, resid=True, od=torch.float32)
# the fused SAM2 mask decoder at full size (HW = 4096 -> 128 row blocks, B = 1 object, 9 tokens): tensor-core variant, then fp32-FMA variant
from videoglamm_b200.config import Sam2Config
from videoglamm_b200.sam2_engine import Sam2Engine
from videoglamm_b200.statedict import random_state_dict, sam2_param_shapes
cfg = Sam2Config()
eng = Sam2Engine(random_state_dict(sam2_param_shapes(cfg), dev, seed=1), cfg, dev)
HW = eng.HW
pix = torch.randn(HW, 256, device=dev); s0 = torch.randn(16 * HW, 32, device=dev); s1 = torch.randn(4 * HW, 64, device=dev); sp = torch.randn(1, 3, 256, device=dev)
torch.cuda.synchronize()
eng.mask_decoder(pix, s0, s1, sp)
torch.cuda.synchronize()
eng.fused_exact = True
eng.mask_decoder(pix, s0, s1, sp)
torch.cuda.synchronize()
