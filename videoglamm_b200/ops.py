"""Tensor-level wrappers over the C ABI. torch is used for device memory and streams only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p | None:
    return None if t is None else C.c_void_p(t.data_ptr())


def _dt(t: torch.dtype) -> int:
    if t == torch.bfloat16:
        return L.BF16
    if t == torch.float32:
        return L.F32
    raise TypeError(f"unsupported dtype {t}")


def _require_cuda(*ts: torch.Tensor | None) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("videoglamm_b200 ops run on CUDA tensors only (there is no CPU path)")


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: torch.Tensor | None = None, act: str | None = None, glu: bool = False,
         col_scale: torch.Tensor | None = None, resid: torch.Tensor | None = None, out: torch.Tensor | None = None,
         out_dtype: torch.dtype = torch.bfloat16, alpha: float = 1.0, force_block_n: int = 0, naive: bool = False) -> torch.Tensor:
    """out = epilogue(alpha * a @ w^T). a: bf16 [M,K] or [B,M,K]; w: bf16 [N,K] or [B,N,K] (nn.Linear layout).
    Rows of `a`/`w` may be strided (last dim contiguous)."""
    _require_cuda(a, w, bias, col_scale, resid, out)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16, (a.dtype, w.dtype)
    assert a.stride(-1) == 1 and w.stride(-1) == 1
    batched = a.dim() == 3
    if batched:
        Bt, M, K = a.shape
        assert w.dim() == 3 and w.shape[0] == Bt
        N = w.shape[1]
    else:
        assert a.dim() == 2 and w.dim() == 2
        Bt, (M, K), N = 1, a.shape, w.shape[0]
    assert w.shape[-1] == K, (a.shape, w.shape)
    n_out = N // 2 if glu else N
    if out is None:
        out = torch.empty((Bt, M, n_out) if batched else (M, n_out), device=a.device, dtype=out_dtype)
    assert out.stride(-1) == 1 and out.shape[-1] == n_out and out.shape[-2] == M
    g = L.VgGemm()
    g.A, g.lda, g.strideA = a.data_ptr(), a.stride(-2), a.stride(0) if batched else 0
    g.B, g.ldb, g.strideB = w.data_ptr(), w.stride(-2), w.stride(0) if batched else 0
    g.C, g.ldc, g.strideC = out.data_ptr(), out.stride(-2), out.stride(0) if batched else 0
    g.M, g.N, g.K, g.batch = M, N, K, Bt
    g.out_dtype = _dt(out.dtype)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
        g.bias = bias.data_ptr()
    g.act = L.ACT[act]
    g.glu = int(glu)
    if col_scale is not None:
        assert col_scale.dtype == torch.float32 and col_scale.numel() == n_out and col_scale.is_contiguous()
        g.col_scale = col_scale.data_ptr()
    if resid is not None:
        assert resid.stride(-1) == 1 and resid.shape[-1] == n_out and resid.shape[-2] == M
        g.resid, g.ldr, g.strideR, g.resid_dtype = resid.data_ptr(), resid.stride(-2), resid.stride(0) if batched else 0, _dt(resid.dtype)
    g.alpha = alpha
    g.force_block_n = force_block_n
    fn = L.lib().vg_gemm_naive if naive else L.lib().vg_gemm
    L.check(fn(C.byref(g), _stream()), "vg_gemm")
    return out


def launch_count() -> int:
    return int(L.lib().vg_launch_count())


def launch_count_reset() -> None:
    L.lib().vg_launch_count_reset()
