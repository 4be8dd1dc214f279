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


class W3:
    """B-side operand of a split-bf16 ("x3") GEMM: w fp32 [N,K] (or [B,N,K]) stored as bf16 [.., N, 3*Kp] = [hi | hi | lo] (include/vglamm.h,
    vg_split3). Built once at load time for weights; `gemm` accepts it in place of a bf16 weight and then takes an fp32 A."""

    def __init__(self, w: torch.Tensor):
        _require_cuda(w)
        assert w.dtype == torch.float32 and w.dim() in (2, 3)
        self.shape = tuple(w.shape)
        self.K = w.shape[-1]
        self.w3 = split3(w.reshape(-1, self.K), b_side=True).view(*w.shape[:-1], -1)
        self.Kp = self.w3.shape[-1] // 3

    def dim(self):
        return len(self.shape)


def split3(x: torch.Tensor, b_side: bool = False) -> torch.Tensor:
    """fp32 [rows, K] (row stride arbitrary) -> bf16 [rows, 3*Kp], Kp = K rounded up to 8: [hi | lo | hi] (A side) or [hi | hi | lo] (B side)."""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    rows, K = x.shape
    kp = (K + 7) // 8 * 8
    out = torch.empty((rows, 3 * kp), device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().vg_split3(_ptr(x), C.c_longlong(x.stride(0)), _ptr(out), C.c_longlong(3 * kp), C.c_longlong(rows), K, kp, int(b_side), _stream()),
            "vg_split3")
    return out


def gemm3(a: torch.Tensor, w: "W3 | torch.Tensor", **kw) -> torch.Tensor:
    """fp32-grade GEMM on the bf16 tensor cores: a fp32 [M,K] / [B,M,K]; w a W3 (weights) or an fp32 activation [N,K] / [B,N,K] (split on the fly)."""
    assert a.dtype == torch.float32 and a.stride(-1) == 1
    if not isinstance(w, W3):
        w = W3(w.contiguous())
    assert a.shape[-1] == w.K, (a.shape, w.shape)
    a2 = a.reshape(-1, w.K) if a.is_contiguous() else a
    assert a2.dim() == 2
    a3 = split3(a2).view(*a.shape[:-1], -1)
    return gemm(a3, w.w3, **kw)


def gemm(a: torch.Tensor, w: torch.Tensor, *, bias: torch.Tensor | None = None, act: str | None = None, glu: bool = False,
         col_scale: torch.Tensor | None = None, resid: torch.Tensor | None = None, out: torch.Tensor | None = None,
         out_dtype: torch.dtype = torch.bfloat16, alpha: float = 1.0, force_block_n: int = 0, naive: bool = False,
         a_rms_eps: float | None = None) -> torch.Tensor:
    """out = epilogue(alpha * a @ w^T). a: bf16 [M,K] or [B,M,K]; w: bf16 [N,K] or [B,N,K] (nn.Linear layout).
    Rows of `a`/`w` may be strided (last dim contiguous)."""
    if isinstance(w, W3):
        assert a_rms_eps is None and not naive
        return gemm3(a, w, bias=bias, act=act, glu=glu, col_scale=col_scale, resid=resid, out=out, out_dtype=out_dtype, alpha=alpha)
    _require_cuda(a, w, bias, col_scale, resid, out)
    if a_rms_eps is not None:  # fp32 A rows, RMS-normalised on the fly (skinny path only)
        assert a.dtype == torch.float32 and a.dim() == 2 and a.shape[0] <= 8 and not naive
    else:
        assert a.dtype == torch.bfloat16, a.dtype
    assert w.dtype == torch.bfloat16, w.dtype
    assert a.stride(-1) == 1 and w.stride(-1) == 1
    batched = a.dim() == 3
    if batched:
        Bt, M, K = a.shape
        assert (w.dim() == 3 and w.shape[0] == Bt) or w.dim() == 2  # 2D w: shared by every batch entry
        N = w.shape[-2]
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
    g.B, g.ldb, g.strideB = w.data_ptr(), w.stride(-2), w.stride(0) if (batched and w.dim() == 3) else 0
    g.C, g.ldc, g.strideC = out.data_ptr(), out.stride(-2), out.stride(0) if batched else 0
    g.M, g.N, g.K, g.batch = M, N, K, Bt
    g.out_dtype = _dt(out.dtype)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() in (N, Bt * N)
        g.bias, g.strideBias = bias.data_ptr(), (N if bias.numel() == Bt * N and Bt > 1 else 0)
    g.act = L.ACT[act]
    g.glu = int(glu)
    if col_scale is not None:
        assert col_scale.dtype == torch.float32 and col_scale.numel() == n_out and col_scale.is_contiguous()
        g.col_scale = col_scale.data_ptr()
    if resid is not None:
        assert resid.stride(-1) == 1 and resid.shape[-1] == n_out and resid.shape[-2] == M
        g.resid, g.ldr, g.strideR, g.resid_dtype = resid.data_ptr(), resid.stride(-2), resid.stride(0) if (batched and resid.dim() == 3) else 0, _dt(resid.dtype)
    g.alpha = alpha
    g.force_block_n = force_block_n
    if a_rms_eps is not None:
        g.a_dtype, g.a_rms_eps = L.F32, a_rms_eps
    fn = L.lib().vg_gemm_naive if naive else L.lib().vg_gemm
    if _GEMM_PROF is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(fn(C.byref(g), _stream()), "vg_gemm")
        e1.record()
        _GEMM_PROF.append((e0, e1, 2.0 * M * N * K * Bt, (M, N, K, Bt)))
        return out
    L.check(fn(C.byref(g), _stream()), "vg_gemm")
    return out


_GEMM_PROF: list | None = None


def gemm_profile(enable: bool):
    """Per-launch CUDA-event timing of every GEMM on the current stream (bench.py roofline). enable=False returns the totals."""
    global _GEMM_PROF
    if enable:
        _GEMM_PROF = []
        return None
    torch.cuda.synchronize()
    rec, _GEMM_PROF = _GEMM_PROF or [], None
    by_shape: dict = {}
    for e0, e1, fl, shp in rec:
        ms = e0.elapsed_time(e1)
        a = by_shape.setdefault(shp, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ms
        a[2] += fl
    top = sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:40]
    tensor = [(k, v) for k, v in by_shape.items() if k[0] > 8]      # tcgen05 kernel; M <= 8 takes the weight-streaming (HBM-bound) kernel
    big = [(k, v) for k, v in tensor if 2.0 * k[0] * k[1] * k[2] * k[3] >= 1e9]   # launches large enough for the tensor pipe to be the limiter
    skinny = [(k, v) for k, v in by_shape.items() if k[0] <= 8]
    return {"n": len(rec), "ms": sum(v[1] for v in by_shape.values()), "flops": sum(v[2] for v in by_shape.values()),
            "tensor": {"n": sum(v[0] for _, v in tensor), "ms": sum(v[1] for _, v in tensor), "flops": sum(v[2] for _, v in tensor)},
            "tensor_big": {"n": sum(v[0] for _, v in big), "ms": sum(v[1] for _, v in big), "flops": sum(v[2] for _, v in big)},
            "skinny": {"n": sum(v[0] for _, v in skinny), "ms": sum(v[1] for _, v in skinny),
                       "weight_bytes": sum(v[0] * 2.0 * k[1] * k[2] * k[3] for k, v in skinny)},
            "top_shapes": [{"MNKB": list(k), "launches": v[0], "ms": round(v[1], 4), "tflops": round(v[2] / max(v[1], 1e-9) / 1e9, 1)} for k, v in top]}


def launch_count() -> int:
    return int(L.lib().vg_launch_count())


def launch_count_reset() -> None:
    L.lib().vg_launch_count_reset()


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, scale: float | None = None, causal: bool = False, window: int = 0,
              out: torch.Tensor | None = None, q_win: tuple | None = None, k_win: tuple | None = None,
              nk_dev: torch.Tensor | None = None, decode_ws: torch.Tensor | None = None, legacy: bool = False,
              fused_rope_theta: float = 0.0) -> torch.Tensor:
    """q [B,Nq,H,D], k/v [B,Nk,H,D] bf16 views (any strides with unit last stride) -> out [B,Nq,H,D] (contiguous unless given).
    q_win/k_win = (win, gridw, nwx, nwy): Hiera 2D window addressing; then B = images*nwx*nwy and the tensors passed are the
    un-partitioned [images*gridh*gridw, H, D] views (dim 0 of size 1 is added by the caller: shape [1, rows, H, D])."""
    _require_cuda(q, k, v, out)
    assert q.dtype == k.dtype == v.dtype and q.dtype in (torch.bfloat16, torch.float32)
    exact = q.dtype == torch.float32                    # fp32 operands -> the exact fp32 kernel (vg_attention_f32)
    a = L.VgAttention()
    if q_win is None:
        B, Nq, H, D = q.shape
        Nk = k.shape[1]
    else:
        win, gridw, nwx, nwy = q_win
        kw = k_win or q_win
        H, D = q.shape[-2:]
        n_img = q.shape[1] // (nwy * win * gridw)
        B, Nq, Nk = n_img * nwx * nwy, win * win, kw[0] * kw[0]
        a.q_win, a.q_gridw, a.q_nwx, a.q_nwy = q_win
        a.k_win, a.k_gridw, a.k_nwx, a.k_nwy = kw
    if out is None:
        out = torch.empty(q.shape, device=q.device, dtype=q.dtype)
    assert out.dtype == q.dtype
    for t in (q, k, v, out):
        assert t.stride(-1) == 1
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.q_bs, a.q_rs, a.q_hs = q.stride(0), q.stride(1), q.stride(2)
    a.k_bs, a.k_rs, a.k_hs = k.stride(0), k.stride(1), k.stride(2)
    a.v_bs, a.v_rs, a.v_hs = v.stride(0), v.stride(1), v.stride(2)
    a.o_bs, a.o_rs, a.o_hs = out.stride(0), out.stride(1), out.stride(2)
    a.B, a.H, a.Nq, a.Nk, a.D = B, H, Nq, Nk, D
    a.scale = float(scale if scale is not None else D ** -0.5)
    a.causal, a.window = int(causal), int(window)
    a.force_legacy = int(legacy)
    if exact:
        assert nk_dev is None and decode_ws is None and not fused_rope_theta, "decode-graph features exist on the bf16 path only"
        if q_win is None and not causal and not window and Nk >= 256 and B * H * ((Nq + 15) // 16) < 148:
            ws = torch.empty((16 * B * H * Nq * (D + 2),), device=q.device, dtype=torch.float32)
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
        L.check(L.lib().vg_attention_f32(C.byref(a), _stream()), "vg_attention_f32")
        return out
    if nk_dev is not None:
        assert nk_dev.dtype == torch.int32 and nk_dev.is_cuda and Nq == 1 and nk_dev.numel() in (1, B)
        a.nk_dev = nk_dev.data_ptr()
        a.nk_per_batch = int(nk_dev.numel() == B and B > 1)
    if fused_rope_theta:
        assert Nq == 1 and q.stride(2) == D, "fused RoPE/append: q must be the packed [q|k|v] row of the newest token (one row per batch entry)"
        a.fused_rope_theta = float(fused_rope_theta)
    if q_win is None and not causal and Nq > 1 and Nk >= 2048 and B * H * ((Nq + 63) // 64) < 148:
        # few query tiles, many keys (SAM2 memory attention): split-K scratch for up to 8 partials
        dp = 32 if D <= 32 else 64 if D <= 64 else 80 if D <= 80 else 96 if D <= 96 else 128 if D <= 128 else 256
        ws = torch.empty((8 * B * H * Nq * (max(dp, 128 if D <= 128 else 256) + 2),), device=q.device, dtype=torch.float32)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    if decode_ws is not None:  # persistent ZEROED fp32 scratch (see decode_workspace): split-K partials + in-kernel merge tickets
        assert Nq == 1 and decode_ws.dtype == torch.float32 and decode_ws.is_cuda
        a.workspace, a.workspace_bytes, a.ws_has_counters = decode_ws.data_ptr(), decode_ws.numel() * 4, 1
    elif Nq == 1 and q_win is None and Nk >= 768:  # decode: split-K scratch
        ws = torch.empty((B * H * 32 * (max(D, 128) + 130),), device=q.device, dtype=torch.float32)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
    L.check(L.lib().vg_attention(C.byref(a), _stream()), "vg_attention")
    return out


def norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None = None, *, eps: float, rms: bool = False,
         out_dtype: torch.dtype = torch.bfloat16, out: torch.Tensor | None = None) -> torch.Tensor:
    """LayerNorm / RMSNorm over the last dim. x [..., C] with unit last stride and uniform row stride (2D after flatten)."""
    _require_cuda(x, w, b, out)
    cols = x.shape[-1]
    x2 = x.reshape(-1, cols) if x.is_contiguous() else x
    assert x2.dim() == 2 and x2.stride(1) == 1
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    o2 = out.reshape(-1, cols) if out.is_contiguous() else out
    assert o2.dim() == 2 and o2.stride(1) == 1
    assert w.dtype == torch.float32 and (b is None or b.dtype == torch.float32)
    L.check(L.lib().vg_norm(_ptr(x2), _dt(x2.dtype), C.c_longlong(x2.stride(0)), _ptr(o2), _dt(o2.dtype), C.c_longlong(o2.stride(0)), _ptr(w), _ptr(b),
                            x2.shape[0], cols, C.c_float(eps), int(rms), _stream()), "vg_norm")
    return out


def add(a: torch.Tensor, b: torch.Tensor | None = None, *, alpha: float = 1.0, out_dtype: torch.dtype | None = None,
        out: torch.Tensor | None = None, rows: int | None = None) -> torch.Tensor:
    """out[r] = a[r % a_rows] + alpha*b[r % b_rows] (2D after flattening leading dims). b=None casts. `rows` > a rows tiles a."""
    _require_cuda(a, b, out)
    cols = a.shape[-1]
    a2 = a.reshape(-1, cols) if a.is_contiguous() else a
    assert a2.dim() == 2 and a2.stride(1) == 1
    n_rows = rows or a2.shape[0]
    assert n_rows % a2.shape[0] == 0
    if out is None:
        out = torch.empty(a.shape if n_rows == a2.shape[0] else (n_rows, cols), device=a.device, dtype=out_dtype or a.dtype)
    o2 = out.reshape(-1, cols) if out.is_contiguous() else out
    if b is not None:
        b2 = b.reshape(-1, cols) if b.is_contiguous() else b
        assert b2.dim() == 2 and b2.stride(1) == 1 and n_rows % b2.shape[0] == 0
        bargs = (_ptr(b2), _dt(b2.dtype), C.c_longlong(b2.stride(0)), b2.shape[0])
    else:
        bargs = (None, 0, C.c_longlong(0), 1)
    L.check(L.lib().vg_add(_ptr(a2), _dt(a2.dtype), C.c_longlong(a2.stride(0)), a2.shape[0], *bargs, C.c_float(alpha), _ptr(o2), _dt(o2.dtype),
                           C.c_longlong(o2.stride(0)), n_rows, cols, _stream()), "vg_add")
    return out


def cast(a: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return add(a, None, out_dtype=dtype)


def pool2x2(x: torch.Tensor, n: int, h: int, w: int, *, is_max: bool, out_dtype: torch.dtype | None = None) -> torch.Tensor:
    """x: [n*h*w, C] view (row stride arbitrary) -> [n*(h/2)*(w/2), C]"""
    _require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1 and x.shape[0] == n * h * w
    c = x.shape[1]
    y = torch.empty((n * (h // 2) * (w // 2), c), device=x.device, dtype=out_dtype or x.dtype)
    L.check(L.lib().vg_pool2x2(_ptr(x), _dt(x.dtype), C.c_longlong(x.stride(0)), _ptr(y), _dt(y.dtype), C.c_longlong(c), n, h, w, c, int(is_max),
                               _stream()), "vg_pool2x2")
    return y


def upsample2x_add(lat: torch.Tensor, top: torch.Tensor, n: int, h: int, w: int) -> torch.Tensor:
    """lat [n*2h*2w, C] fp32 + nearest-upsampled top [n*h*w, C] fp32"""
    _require_cuda(lat, top)
    assert lat.dtype == top.dtype == torch.float32 and lat.is_contiguous() and top.is_contiguous()
    out = torch.empty_like(lat)
    L.check(L.lib().vg_upsample2x_add(_ptr(lat), _ptr(top), _ptr(out), n, h, w, lat.shape[-1], _stream()), "vg_upsample2x_add")
    return out


def convt2x2_epilogue(g: torch.Tensor, bias: torch.Tensor, feat: torch.Tensor | None, ln_w, ln_b, eps: float, n: int, h: int, w: int,
                      cout: int, out_dtype=torch.bfloat16, feat_bcast: bool = False) -> torch.Tensor:
    _require_cuda(g, bias, feat)
    assert g.is_contiguous() and g.shape[-1] == 4 * cout and (feat is None or feat.is_contiguous())
    out = torch.empty((n * 4 * h * w, cout), device=g.device, dtype=out_dtype)
    L.check(L.lib().vg_convt2x2_epilogue(_ptr(g), _dt(g.dtype), _ptr(bias), _ptr(feat), _dt(feat.dtype) if feat is not None else 0, _ptr(ln_w),
                                         _ptr(ln_b), C.c_float(eps), _ptr(out), _dt(out.dtype), n, h, w, cout, int(feat_bcast), _stream()), "vg_convt2x2_epilogue")
    return out


def im2col(x: torch.Tensor, k: int, s: int, p: int, kpad: int | None = None, out_dtype: torch.dtype = torch.bfloat16) -> tuple[torch.Tensor, int, int]:
    """x NCHW (fp32/bf16, contiguous) -> bf16 (or fp32) [n*oh*ow, kpad]"""
    _require_cuda(x)
    assert x.is_contiguous() and x.dim() == 4
    n, cin, h, w = x.shape
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    kk = cin * k * k
    kpad = kpad or (kk + 7) // 8 * 8
    out = torch.empty((n * oh * ow, kpad), device=x.device, dtype=out_dtype)
    L.check(L.lib().vg_im2col(_ptr(x), _dt(x.dtype), _ptr(out), _dt(out_dtype), n, cin, h, w, k, s, p, oh, ow, kpad, _stream()), "vg_im2col")
    return out, oh, ow


def dwconv(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, n: int, h: int, wd: int, k: int, out_dtype=None) -> torch.Tensor:
    """x [n*h*wd, C] channels-last; w [C,k,k] fp32"""
    _require_cuda(x, w, bias)
    assert x.is_contiguous() and w.is_contiguous() and w.dtype == torch.float32
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    L.check(L.lib().vg_dwconv(_ptr(x), _dt(x.dtype), _ptr(w), _ptr(bias), _ptr(y), _dt(y.dtype), n, h, wd, x.shape[-1], k, _stream()), "vg_dwconv")
    return y


def conv3x3s2_ln_gelu(x: torch.Tensor, w: torch.Tensor, bias, ln_w, ln_b, eps: float, n: int, h: int, wd: int, cin: int, cout: int,
                      out_dtype=torch.float32) -> torch.Tensor:
    """x [n*h*wd, cin] channels-last; w [(ky,kx,ci), cout] fp32"""
    _require_cuda(x, w)
    assert x.is_contiguous() and w.is_contiguous() and tuple(w.shape) == (9 * cin, cout)
    y = torch.empty((n * (h // 2) * (wd // 2), cout), device=x.device, dtype=out_dtype)
    L.check(L.lib().vg_conv3x3s2_ln_gelu(_ptr(x), _dt(x.dtype), _ptr(w), _ptr(bias), _ptr(ln_w), _ptr(ln_b), C.c_float(eps), _ptr(y), _dt(y.dtype),
                                         n, h, wd, cin, cout, _stream()), "vg_conv3x3s2_ln_gelu")
    return y


def bilinear(x: torch.Tensor, oh: int, ow: int, *, mode: int = 0, scale: float = 1.0, bias: float = 0.0) -> torch.Tensor:
    """x [n, ih, iw] fp32 -> [n, oh, ow] fp32 (mode 0/1/2), uint8 (mode 3: x > 0) or bit-packed uint8 [n, ceil(oh*ow/8)] (mode 4: x > 0, numpy.packbits order)"""
    _require_cuda(x)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 3
    n, ih, iw = x.shape
    if mode == 4:
        y = torch.empty((n, (oh * ow + 7) // 8), device=x.device, dtype=torch.uint8)
    else:
        y = torch.empty((n, oh, ow), device=x.device, dtype=torch.uint8 if mode == 3 else torch.float32)
    L.check(L.lib().vg_bilinear(_ptr(x), _ptr(y), n, ih, iw, oh, ow, mode, C.c_float(scale), C.c_float(bias), _stream()), "vg_bilinear")
    return y


def rope2d_table(dim: int, gw: int, gh: int, device) -> torch.Tensor:
    """(cos, sin) of every (grid position, rotation pair) for rope2d_(..., table=): fp32 [gw*gh, dim/2, 2]"""
    t = torch.empty((gw * gh, dim // 2, 2), device=device, dtype=torch.float32)
    L.check(L.lib().vg_rope2d_table(_ptr(t), dim, gw, gh, _stream()), "vg_rope2d_table")
    return t


def rope2d_(x: torch.Tensor, dim: int, gw: int, gh: int, table: torch.Tensor | None = None, nblk: int = 1) -> torch.Tensor:
    """in place on bf16 / fp32 [rows, >=dim] (row stride arbitrary). With `table` (rope2d_table): table lookup instead of powf / sincosf per
    element, and `nblk` adjacent column blocks of width dim rotated in one launch (q and k of packed qkv rows)."""
    _require_cuda(x, table)
    assert x.dim() == 2 and x.stride(1) == 1
    if table is not None:
        assert x.shape[1] >= nblk * dim and tuple(table.shape) == (gw * gh, dim // 2, 2)
        L.check(L.lib().vg_rope2d_cached(_ptr(x), _dt(x.dtype), C.c_longlong(x.stride(0)), x.shape[0], dim, nblk, _ptr(table), gw * gh, _stream()), "vg_rope2d_cached")
        return x
    assert nblk == 1
    L.check(L.lib().vg_rope2d(_ptr(x), _dt(x.dtype), C.c_longlong(x.stride(0)), x.shape[0], dim, gw, gh, _stream()), "vg_rope2d")
    return x


def rope_neox_(qkv: torch.Tensor, pos: torch.Tensor, n_heads_total: int, head_dim: int, theta: float) -> torch.Tensor:
    _require_cuda(qkv, pos)
    assert qkv.dim() == 2 and qkv.stride(1) == 1 and pos.dtype == torch.int32
    L.check(L.lib().vg_rope_neox(_ptr(qkv), _dt(qkv.dtype), C.c_longlong(qkv.stride(0)), _ptr(pos), qkv.shape[0], n_heads_total, head_dim, C.c_float(theta),
                                 _stream()), "vg_rope_neox")
    return qkv


def embedding(table: torch.Tensor, ids: torch.Tensor, out_dtype=torch.float32) -> torch.Tensor:
    _require_cuda(table, ids)
    assert table.dtype == torch.bfloat16 and ids.dtype == torch.int64 and ids.is_contiguous() and table.stride(1) == 1
    out = torch.empty((ids.numel(), table.shape[1]), device=table.device, dtype=out_dtype)
    L.check(L.lib().vg_embedding(_ptr(table), C.c_longlong(table.stride(0)), _ptr(ids), _ptr(out), _dt(out.dtype), C.c_longlong(table.shape[1]),
                                 ids.numel(), table.shape[1], _stream()), "vg_embedding")
    return out


def argmax(x: torch.Tensor) -> torch.Tensor:
    _require_cuda(x)
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.empty((x.shape[0],), device=x.device, dtype=torch.int64)
    L.check(L.lib().vg_argmax(_ptr(x), _dt(x.dtype), C.c_longlong(x.stride(0)), _ptr(out), x.shape[0], x.shape[1], _stream()), "vg_argmax")
    return out


def row_scatter(src: torch.Tensor, cache: torch.Tensor, pos_dev: torch.Tensor, col0: int, cols: int) -> None:
    """cache[pos_dev[0], col0:col0+cols] = src[0, col0:col0+cols] (bf16), position read on the device."""
    _require_cuda(src, cache, pos_dev)
    assert src.dtype == cache.dtype == torch.bfloat16 and pos_dev.dtype == torch.int32 and cache.stride(1) == 1
    L.check(L.lib().vg_row_scatter(_ptr(src), _ptr(cache), C.c_longlong(cache.stride(0)), _ptr(pos_dev), col0, cols, _stream()), "vg_row_scatter")


def launch_count_add(n: int) -> None:
    L.lib().vg_launch_count_add(int(n))


def decode_workspace(B: int, H: int, D: int, device) -> torch.Tensor:
    """Zero-initialised scratch for single-query attention with in-kernel split-K merge (up to 32 partials per head + B*H ticket counters)."""
    return torch.zeros((B * H * 32 * (max(D, 128) + 130) + B * H,), device=device, dtype=torch.float32)


def rope_kv_append_(qkv: torch.Tensor, cache: torch.Tensor, pos_dev: torch.Tensor, n_heads: int, head_dim: int, theta: float) -> None:
    _require_cuda(qkv, cache, pos_dev)
    assert qkv.dtype == cache.dtype == torch.bfloat16 and pos_dev.dtype == torch.int32 and qkv.is_contiguous() and cache.stride(1) == 1
    L.check(L.lib().vg_rope_kv_append(_ptr(qkv), _ptr(cache), C.c_longlong(cache.stride(0)), _ptr(pos_dev), n_heads, head_dim, C.c_float(theta),
                                      _stream()), "vg_rope_kv_append")


def multimask_select(masks: torch.Tensor, iou: torch.Tensor, delta: float = 0.05, thresh: float = 0.98) -> torch.Tensor:
    """masks fp32 [B,4,hw], iou fp32 [B,4] -> int64 [B] selected mask index (dynamic multimask via stability)."""
    _require_cuda(masks, iou)
    assert masks.dtype == iou.dtype == torch.float32 and masks.is_contiguous() and iou.is_contiguous() and masks.shape[1] == 4
    out = torch.empty((masks.shape[0],), device=masks.device, dtype=torch.int64)
    L.check(L.lib().vg_multimask_select(_ptr(masks), _ptr(iou), masks.shape[0], masks.shape[2], C.c_float(delta), C.c_float(thresh), _ptr(out),
                                        _stream()), "vg_multimask_select")
    return out


def connected_components(mask: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """sam2._C.get_connected_componnets (sam2/csrc/connected_components.cu:213-289): mask uint8 [N,1,H,W] CUDA contiguous -> (labels int32
    [N,1,H,W], counts int32 [N,1,H,W]); raises RuntimeError on bad input like the reference's AT_ASSERTM."""
    if not mask.is_cuda:
        raise RuntimeError("inputs must be a CUDA tensor")
    if mask.dim() != 4 or mask.shape[1] != 1:
        raise RuntimeError("inputs must be [N, 1, H, W] shape")
    if mask.dtype != torch.uint8:
        raise RuntimeError("inputs must be a uint8 type")
    mask = mask.contiguous()
    n, _, h, w = mask.shape
    labels = torch.empty((n, 1, h, w), device=mask.device, dtype=torch.int32)
    counts = torch.empty_like(labels)
    ws = torch.empty((2 * n * h * w,), device=mask.device, dtype=torch.int32)
    L.check(L.lib().vg_connected_components(_ptr(mask), _ptr(labels), _ptr(counts), n, h, w, _ptr(ws), C.c_longlong(ws.numel() * 4), _stream()),
            "vg_connected_components")
    return labels, counts


def conv3x3s2_ln_gelu_gemm(x: torch.Tensor, w_gemm: torch.Tensor, bias, ln_w, ln_b, eps: float, n: int, h: int, wd: int, cin: int, cout: int,
                           out_dtype=torch.float32) -> torch.Tensor:
    """Conv2d(3x3, s2, p1) + LayerNorm2d + GELU on the tensor cores: channels-last im2col (bf16) -> GEMM (+bias, fp32) -> LayerNorm+GELU.
    x [n*h*wd, cin] channels-last (cin % 8 == 0); w_gemm bf16 [cout, (ky,kx,ci)]."""
    _require_cuda(x, w_gemm)
    assert x.is_contiguous() and tuple(w_gemm.shape) == (cout, 9 * cin) and w_gemm.dtype == torch.bfloat16 and cin % 8 == 0
    rows = n * (h // 2) * (wd // 2)
    cols = torch.empty((rows, 9 * cin), device=x.device, dtype=torch.bfloat16)
    L.check(L.lib().vg_im2col3x3s2_cl(_ptr(x), _dt(x.dtype), _ptr(cols), n, h, wd, cin, _stream()), "vg_im2col3x3s2_cl")
    g = gemm(cols, w_gemm, bias=bias, out_dtype=torch.float32)
    y = torch.empty((rows, cout), device=x.device, dtype=out_dtype)
    L.check(L.lib().vg_layernorm_gelu(_ptr(g), C.c_longlong(g.stride(0)), _ptr(y), _dt(y.dtype), C.c_longlong(y.stride(0)), _ptr(ln_w), _ptr(ln_b), rows, cout,
                                      C.c_float(eps), _stream()), "vg_layernorm_gelu")
    return y


def mask_decoder_fused(weights: "L.VgMaskDecoderWeights", pix: torch.Tensor, dense_pe: torch.Tensor, no_mask: torch.Tensor, s0: torch.Tensor, s1: torch.Tensor,
                       tokens: torch.Tensor, fs: int, pix_shared: bool = False, exact: bool = False):
    """vg_mask_decoder: the whole SAM2 mask decoder for the B objects of one frame in ONE launch (include/vglamm.h). pix fp32 [B*HW, 256] (or [HW, 256]
    shared), s0 fp32 [16*HW, 32], s1 fp32 [4*HW, 64], tokens fp32 [B, Nt, 256] -> (masks [B,4,(4fs)^2], iou [B,4], mask tokens [B,4,256], obj [B,1])."""
    _require_cuda(pix, dense_pe, no_mask, s0, s1, tokens)
    for t in (pix, dense_pe, no_mask, s0, s1, tokens):
        assert t.dtype == torch.float32 and t.is_contiguous()
    B, Nt, Cc = tokens.shape
    HW = fs * fs
    assert Cc == 256 and s0.shape == (16 * HW, 32) and s1.shape == (4 * HW, 64) and pix.shape == ((HW if pix_shared else B * HW), 256)
    dev = pix.device
    masks = torch.empty((B, 4, 16 * HW), device=dev, dtype=torch.float32)
    iou = torch.empty((B, 4), device=dev, dtype=torch.float32)
    mt = torch.empty((B, 4, 256), device=dev, dtype=torch.float32)
    obj = torch.empty((B, 1), device=dev, dtype=torch.float32)
    nbytes = int(L.lib().vg_mask_decoder_workspace_bytes(B, fs))
    ws = torch.empty((nbytes,), device=dev, dtype=torch.uint8)
    d = L.VgMaskDecoder()
    d.pix, d.pix_shared, d.dense_pe, d.no_mask, d.s0, d.s1, d.tokens = pix.data_ptr(), int(pix_shared), dense_pe.data_ptr(), no_mask.data_ptr(), s0.data_ptr(), s1.data_ptr(), tokens.data_ptr()
    d.B, d.Nt, d.fs, d.C, d.heads, d.mlp, d.depth = B, Nt, fs, 256, 8, 2048, 2
    d.weights = C.pointer(weights)
    d.masks, d.iou, d.mask_tokens, d.obj = masks.data_ptr(), iou.data_ptr(), mt.data_ptr(), obj.data_ptr()
    d.workspace, d.workspace_bytes = ws.data_ptr(), nbytes
    d.force_exact = int(exact)
    L.check(L.lib().vg_mask_decoder(C.byref(d), _stream()), "vg_mask_decoder")
    return masks, iou, mt, obj


def resize_u8_pil(frames: torch.Tensor, oh: int, ow: int, tables: tuple) -> torch.Tensor:
    """uint8 [n,h,w,c] cuda -> uint8 [n,oh,ow,c]: PIL.Image.resize(BILINEAR), bit-exact. tables = (kh, bh, kv, bv) int32 cuda tensors
    (preprocess.pil_bilinear_coeffs for (w -> ow) and (h -> oh))."""
    _require_cuda(frames, *tables)
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.dim() == 4
    n, h, w, c = frames.shape
    kh, bh, kv, bv = tables
    assert kh.dtype == bh.dtype == kv.dtype == bv.dtype == torch.int32 and kh.shape[0] == ow and kv.shape[0] == oh
    tmp = torch.empty((n, h, ow, c), device=frames.device, dtype=torch.uint8)
    out = torch.empty((n, oh, ow, c), device=frames.device, dtype=torch.uint8)
    L.check(L.lib().vg_resize_u8_pil(_ptr(frames), _ptr(tmp), _ptr(out), n, h, w, c, oh, ow, _ptr(kh), _ptr(bh), kh.shape[1], _ptr(kv), _ptr(bv), kv.shape[1],
                                     _stream()), "vg_resize_u8_pil")
    return out


def sam_normalize_squash(frames: torch.Tensor, S: int, mean: tuple, std: tuple) -> torch.Tensor:
    """uint8 [n,h,w,3] cuda -> fp32 [n,3,S,S] = F.interpolate((x - mean) / std, (S, S), mode="bilinear")."""
    _require_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.dim() == 4 and frames.shape[3] == 3
    n, h, w, _ = frames.shape
    out = torch.empty((n, 3, S, S), device=frames.device, dtype=torch.float32)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    L.check(L.lib().vg_sam_normalize_squash(_ptr(frames), _ptr(out), n, h, w, S, m3, s3, _stream()), "vg_sam_normalize_squash")
    return out


def resize_u8_cv(frames: torch.Tensor, oh: int, ow: int, tables: tuple) -> torch.Tensor:
    """cv2.resize(x, (ow, oh)) (INTER_LINEAR, bit-exact) on uint8 [n,h,w,c] cuda frames; tables = (xi, xc, yi, yc) from preprocess.cv_linear_tables."""
    _require_cuda(frames, *tables)
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.dim() == 4
    xi, xc, yi, yc = tables
    assert all(t.dtype == torch.int32 and t.is_contiguous() for t in tables) and xi.numel() == ow and yi.numel() == oh
    n, h, w, c = frames.shape
    out = torch.empty((n, oh, ow, c), device=frames.device, dtype=torch.uint8)
    L.check(L.lib().vg_resize_u8_cv(_ptr(frames), _ptr(out), n, h, w, c, oh, ow, _ptr(xi), _ptr(xc), _ptr(yi), _ptr(yc), _stream()), "vg_resize_u8_cv")
    return out


def u8_normalize_chw(frames: torch.Tensor, oh: int, ow: int, top: int, left: int, mean: tuple, std: tuple, mode: int) -> torch.Tensor:
    """uint8 [n,h,w,3] cuda, crop [top, top+oh) x [left, left+ow) -> fp32 [n,3,oh,ow]; mode 0: HF rescale + normalize (float32), mode 1: (x/255 - mean)/std in float64."""
    _require_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.is_contiguous() and frames.dim() == 4 and frames.shape[3] == 3
    n, h, w, _ = frames.shape
    out = torch.empty((n, 3, oh, ow), device=frames.device, dtype=torch.float32)
    m3, s3 = (C.c_double * 3)(*mean), (C.c_double * 3)(*std)
    L.check(L.lib().vg_u8_normalize_chw(_ptr(frames), _ptr(out), n, h, w, oh, ow, top, left, mode, m3, s3, _stream()), "vg_u8_normalize_chw")
    return out


def gemm_ksplit(a: torch.Tensor, w: torch.Tensor, nparts: int, *, resid: torch.Tensor | None = None, out: torch.Tensor | None = None,
                block_n: int = 64) -> torch.Tensor:
    """Skinny projection out = a @ w^T (+ resid) with FEW output columns (a bf16 [M <= 8, K], w bf16 [N, K]) as `nparts` batched tensor-core GEMMs over
    K slices + a partial-sum kernel: N/64 tiles alone would leave most SMs idle while the weights stream (HBM-bound), N/64 * nparts tiles fill them."""
    M, K = a.shape
    N = w.shape[0]
    assert K % nparts == 0 and (K // nparts) % 8 == 0 and a.is_contiguous() and w.is_contiguous()
    kp = K // nparts
    parts = torch.empty((nparts, M, N), device=a.device, dtype=torch.float32)
    gemm(a.view(M, nparts, kp).permute(1, 0, 2), w.view(N, nparts, kp).permute(1, 0, 2), out=parts, force_block_n=block_n)
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous() and (resid is None or (resid.dtype == torch.float32 and resid.is_contiguous()))
    L.check(L.lib().vg_sum_parts(_ptr(parts), C.c_longlong(M * N), nparts, _ptr(resid), _ptr(out), _stream()), "vg_sum_parts")
    return out
