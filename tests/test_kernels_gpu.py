"""Kernel-level parity: every C-ABI op vs a plain torch fp32 reference of the same op on the same (bf16-rounded) inputs.
Tolerances: fp32-output ops are compared at 1e-5 relative L2 (only accumulation order differs); bf16-output ops at 4e-3
(one bf16 rounding of the result, 2^-9 relative)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from videoglamm_b200 import ops  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)
    torch.backends.cudnn.allow_tf32 = False  # the torch reference must be true fp32
    torch.backends.cuda.matmul.allow_tf32 = False


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (128, 64, 64, 64), (128, 256, 64, 256), (300, 200, 136, 0), (77, 40, 72, 64),
                                      (1000, 1000, 1000, 256), (9, 256, 256, 0), (1664, 9216, 3072, 0), (1, 32064, 3072, 0), (4096, 4, 32, 0),
                                      (1000, 576, 136, 192), (1000, 288, 200, 96), (700, 200, 72, 192), (20000, 576, 2304, 0), (20000, 1728, 576, 0),
                                      (40000, 288, 288, 0), (40000, 864, 288, 0)])
def test_gemm_plain(M, N, K, bn):
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    y = ops.gemm(a, w, out_dtype=torch.float32, force_block_n=bn)
    assert rel(y, a.float() @ w.float().T) < 1e-5


@pytest.mark.parametrize("act", ["gelu", "quick_gelu", "relu", "silu", "sigmoid"])
def test_gemm_bias_act(act):
    a = torch.randn(256, 256, device=DEV).bfloat16()
    w = (torch.randn(384, 256, device=DEV) / 16).bfloat16()
    bias = torch.randn(384, device=DEV)
    f = dict(gelu=F.gelu, quick_gelu=lambda x: x * torch.sigmoid(1.702 * x), relu=torch.relu, silu=F.silu, sigmoid=torch.sigmoid)[act]
    y = ops.gemm(a, w, bias=bias, act=act, out_dtype=torch.float32)
    assert rel(y, f(a.float() @ w.float().T + bias)) < 1e-5


def test_gemm_glu_scale_resid_batched():
    a = torch.randn(256, 256, device=DEV).bfloat16()
    w = (torch.randn(512, 256, device=DEV) / 16).bfloat16()
    y = ops.gemm(a, w, glu=True, act="silu", out_dtype=torch.float32)
    r = (a.float() @ w.float().T).reshape(256, -1, 2, 16)
    assert rel(y, (F.silu(r[:, :, 0]) * r[:, :, 1]).reshape(256, -1)) < 1e-5
    bias, cs, res = torch.randn(512, device=DEV), torch.randn(512, device=DEV), torch.randn(256, 512, device=DEV)
    y = ops.gemm(a, w, bias=bias, col_scale=cs, resid=res, out=res.clone())
    assert rel(y, (a.float() @ w.float().T + bias) * cs + res) < 1e-5
    ab = torch.randn(5, 200, 64, device=DEV).bfloat16()
    wb = torch.randn(5, 96, 64, device=DEV).bfloat16()
    assert rel(ops.gemm(ab, wb, out_dtype=torch.float32), ab.float() @ wb.float().transpose(1, 2)) < 1e-5
    yb = ops.gemm(a, w, out_dtype=torch.bfloat16)
    assert rel(yb, a.float() @ w.float().T) < 4e-3


def test_gemm_n_fastest_tiles_with_epilogues():
    """M >> N with several N tiles: N-fastest tile order + 192-column tiles (Hiera stage-3 shapes), in-place residual and GELU epilogues."""
    a = torch.randn(20000, 2304, device=DEV).bfloat16()
    w = (torch.randn(576, 2304, device=DEV) / 48).bfloat16()
    bias, res = torch.randn(576, device=DEV), torch.randn(20000, 576, device=DEV)
    want = a.float() @ w.float().T + bias + res
    assert rel(ops.gemm(a, w, bias=bias, resid=res, out=res), want) < 1e-5
    a2 = torch.randn(20000, 576, device=DEV).bfloat16()
    w2 = (torch.randn(2304, 576, device=DEV) / 24).bfloat16()
    b2 = torch.randn(2304, device=DEV)
    assert rel(ops.gemm(a2, w2, bias=b2, act="gelu"), F.gelu(a2.float() @ w2.float().T + b2)) < 4e-3


def test_gemm_matches_naive_device_reference():
    a = torch.randn(333, 200, device=DEV).bfloat16()
    w = torch.randn(129, 200, device=DEV).bfloat16()
    assert rel(ops.gemm(a, w, out_dtype=torch.float32), ops.gemm(a, w, out_dtype=torch.float32, naive=True)) < 1e-5


def test_gemm_rejects_bad_alignment():
    a = torch.randn(16, 20, device=DEV).bfloat16()[:, :12]  # lda=20 not a multiple of 8
    w = torch.randn(8, 12, device=DEV).bfloat16()
    with pytest.raises(RuntimeError):
        ops.gemm(a, w)


def sdpa_ref(q, k, v, scale, causal=False, window=0):
    q, k, v = [t.float().permute(0, 2, 1, 3) for t in (q, k, v)]
    s = q @ k.transpose(-1, -2) * scale
    Nq, Nk = s.shape[-2:]
    i = torch.arange(Nq, device=s.device)[:, None] + (Nk - Nq)
    j = torch.arange(Nk, device=s.device)[None]
    mask = torch.ones(Nq, Nk, dtype=torch.bool, device=s.device)
    if causal:
        mask &= j <= i
    if window:
        mask &= j >= i - window
    s = s.masked_fill(~mask, float("-inf"))
    return (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3)


@pytest.mark.parametrize("B,H,Nq,Nk,D,causal,window", [
    (2, 4, 64, 64, 64, False, 0), (1, 16, 577, 577, 64, False, 0), (1, 16, 1025, 1025, 88, False, 0), (1, 8, 300, 300, 72, False, 0),
    (1, 32, 333, 333, 96, True, 0), (1, 4, 1, 500, 96, True, 0), (1, 4, 700, 700, 96, True, 128), (2, 1, 256, 1100, 256, False, 0),
    (3, 8, 9, 4096, 16, False, 0), (3, 8, 4096, 9, 16, False, 0), (2, 8, 9, 9, 32, False, 0), (1, 2, 130, 70, 128, False, 0),
    (1, 1, 4096, 8300, 256, False, 0), (2, 2, 300, 5000, 64, False, 0), (1, 1, 70, 2049, 96, False, 0),  # these three: split-K over keys
    (1, 32, 1696, 1696, 96, True, 2047),   # Phi-3 prefill: the sliding window never cuts -> tcgen05 causal path
    (1, 32, 3378, 3378, 96, True, 2047),   # C4/C5 prefill (L > 2047): the band cuts -> tcgen05 kernel with band-edge masking
    (1, 4, 1000, 1000, 96, True, 300), (1, 2, 200, 1500, 64, True, 64), (1, 4, 700, 700, 96, True, 127)])   # narrow bands, Nq != Nk
def test_attention(B, H, Nq, Nk, D, causal, window):
    q = torch.randn(B, Nq, H, D, device=DEV).bfloat16()
    k = torch.randn(B, Nk, H, D, device=DEV).bfloat16()
    v = torch.randn(B, Nk, H, D, device=DEV).bfloat16()
    o = ops.attention(q, k, v, causal=causal, window=window)
    assert rel(o, sdpa_ref(q, k, v, D ** -0.5, causal, window)) < 6e-3  # P is rounded to bf16 before PV


@pytest.mark.parametrize("B,H,Nq,Nk,D,causal,window", [(8, 8, 2048, 640, 72, False, 0), (4, 32, 1024, 1024, 96, True, 0), (4, 32, 1000, 1300, 96, True, 300),
                                                      (16, 8, 1100, 333, 64, False, 0)])
def test_attention_two_query_tiles_per_block(B, H, Nq, Nk, D, causal, window):
    """>= 1024 query tiles x heads x batch and D <= 128: flash_attn_tc2_kernel (two query tiles per block share the K/V ring; causal / banded
    pairs with different key ranges, a ragged last pair, key counts that are not a multiple of the tile)."""
    torch.manual_seed(B + Nq)
    q, k, v = [torch.randn(B, n, H, D, device=DEV).bfloat16() for n in (Nq, Nk, Nk)]
    o = ops.attention(q, k, v, causal=causal, window=window)
    assert rel(o, sdpa_ref(q, k, v, D ** -0.5, causal, window)) < 4e-3


def test_attention_packed_qkv_strides():
    B, N, H, D = 2, 200, 4, 72
    qkv = torch.randn(B, N, 3, H, D, device=DEV).bfloat16()
    q, k, v = qkv.unbind(2)
    assert rel(ops.attention(q, k, v), sdpa_ref(q, k, v, D ** -0.5)) < 6e-3


@pytest.mark.parametrize("legacy", [False, True])
def test_attention_hiera_windows_16(legacy):
    """Hiera stage-3 geometry: 16x16-token windows (256 tokens), 8 heads x 72 on the packed qkv rows -> tcgen05 kernel with windowed TMA
    boxes (legacy=True: the mma.sync kernel on the same call)."""
    n_img, gh, gw, win, H, D = 3, 32, 48, 16, 8, 72
    qkv = torch.randn(n_img * gh * gw, 3 * H * D, device=DEV).bfloat16()
    q, k, v = (qkv[:, j * H * D:(j + 1) * H * D].unflatten(1, (H, D)) for j in range(3))
    nwy, nwx = gh // win, gw // win

    def part(t):
        t = t.reshape(n_img, nwy, win, nwx, win, H, D).permute(0, 1, 3, 2, 4, 5, 6)
        return t.reshape(-1, win * win, H, D)

    o = ops.attention(q[None], k[None], v[None], q_win=(win, gw, nwx, nwy), k_win=(win, gw, nwx, nwy), legacy=legacy)
    assert rel(part(o[0]), sdpa_ref(part(q), part(k), part(v), D ** -0.5)) < 6e-3


@pytest.mark.parametrize("win,H", [(4, 4), (8, 2), (4, 8), (8, 4)])
def test_attention_hiera_small_windows(win, H):
    """Hiera stage-1/2 geometry: 8x8 / 4x4-token windows, 72-wide heads, q/k/v = column blocks of the packed qkv rows -> one block per window
    (window_attn_small_kernel). Checked against SDPA on the partitioned windows and against the generic kernel (legacy=True)."""
    n_img, gh, gw, D = 2, 16, 24, 72
    qkv = torch.randn(n_img * gh * gw, 3 * H * D, device=DEV).bfloat16()
    q, k, v = (qkv[:, j * H * D:(j + 1) * H * D].unflatten(1, (H, D)) for j in range(3))
    nwy, nwx = gh // win, gw // win

    def part(t):
        t = t.reshape(n_img, nwy, win, nwx, win, H, D).permute(0, 1, 3, 2, 4, 5, 6)
        return t.reshape(-1, win * win, H, D)

    o = ops.attention(q[None], k[None], v[None], q_win=(win, gw, nwx, nwy), k_win=(win, gw, nwx, nwy))
    assert rel(part(o[0]), sdpa_ref(part(q), part(k), part(v), D ** -0.5)) < 6e-3
    o2 = ops.attention(q[None], k[None], v[None], q_win=(win, gw, nwx, nwy), k_win=(win, gw, nwx, nwy), legacy=True)
    assert rel(o, o2) < 6e-3


@pytest.mark.parametrize("pooled", [False, True])
def test_attention_hiera_windows(pooled):
    n_img, gh, gw, win, H, D = 2, 16, 24, 8, 2, 72
    qkv = torch.randn(n_img * gh * gw, 3, H, D, device=DEV).bfloat16()
    q, k, v = qkv.unbind(1)
    nwy, nwx = gh // win, gw // win

    def part(t, g_h, g_w, ws):  # window_partition on [n, gh, gw, H, D]
        t = t.reshape(n_img, g_h // ws, ws, g_w // ws, ws, H, D).permute(0, 1, 3, 2, 4, 5, 6)
        return t.reshape(-1, ws * ws, H, D)

    if pooled:
        qp = F.max_pool2d(q.float().reshape(n_img, gh, gw, H * D).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, H, D).bfloat16()
        o = ops.attention(qp[None], k[None], v[None], q_win=(win // 2, gw // 2, nwx, nwy), k_win=(win, gw, nwx, nwy))
        ref = sdpa_ref(part(qp, gh // 2, gw // 2, win // 2), part(k, gh, gw, win), part(v, gh, gw, win), D ** -0.5)
        got = part(o[0], gh // 2, gw // 2, win // 2)
    else:
        o = ops.attention(q[None], k[None], v[None], q_win=(win, gw, nwx, nwy))
        ref = sdpa_ref(part(q, gh, gw, win), part(k, gh, gw, win), part(v, gh, gw, win), D ** -0.5)
        got = part(o[0], gh, gw, win)
    assert rel(got, ref) < 6e-3


@pytest.mark.parametrize("rows,cols,rms,xdt", [(100, 256, False, torch.float32), (77, 1408, True, torch.float32), (5, 3072, True, torch.bfloat16),
                                               (1000, 144, False, torch.float32), (64, 64, False, torch.bfloat16), (10, 32, False, torch.float32),
                                               (300, 1408, True, torch.bfloat16), (50, 576, False, torch.bfloat16)])
def test_norm(rows, cols, rms, xdt):
    x = (torch.randn(rows, cols, device=DEV) * 3 + 1).to(xdt)
    w, b = torch.randn(cols, device=DEV), torch.randn(cols, device=DEV)
    y = ops.norm(x, w, None if rms else b, eps=1e-6, rms=rms, out_dtype=torch.float32)
    xf = x.float()
    ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w if rms else F.layer_norm(xf, (cols,), w, b, 1e-6)
    assert rel(y, ref) < 1e-5


def test_add_pool_upsample():
    a = torch.randn(12, 40, device=DEV)
    b = torch.randn(3, 40, device=DEV).bfloat16()
    assert rel(ops.add(a, b, alpha=0.1), a + 0.1 * b.float().repeat(4, 1)) < 1e-6
    assert ops.cast(a, torch.bfloat16).dtype == torch.bfloat16
    a4, p4 = torch.randn(6 * 50, 144, device=DEV), torch.randn(50, 144, device=DEV)       # fp32 + row-broadcast fp32: 16-byte vector path
    assert torch.equal(ops.add(a4, p4), (a4.view(6, 50, 144) + p4).view(-1, 144))
    assert rel(ops.add(p4, a4[:50], alpha=0.1, rows=150), (p4 + 0.1 * a4[:50]).repeat(3, 1)) < 1e-7
    x = torch.randn(2 * 8 * 6, 3 * 16, device=DEV).bfloat16()
    q = x[:, :16]
    ref = F.max_pool2d(q.float().reshape(2, 8, 6, 16).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, 16)
    assert rel(ops.pool2x2(q, 2, 8, 6, is_max=True), ref) < 1e-6
    ref = F.avg_pool2d(q.float().reshape(2, 8, 6, 16).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, 16)
    assert rel(ops.pool2x2(q, 2, 8, 6, is_max=False, out_dtype=torch.float32), ref) < 1e-6
    xf = torch.randn(3 * 6 * 10, 36, device=DEV)                          # fp32 -> fp32 (projector pooling): 4-channel vector path; 36 % 8 != 0
    for is_max, fn in ((True, F.max_pool2d), (False, F.avg_pool2d)):
        ref = fn(xf.reshape(3, 6, 10, 36).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, 36)
        assert rel(ops.pool2x2(xf, 3, 6, 10, is_max=is_max), ref) < 1e-6
        xb = xf.bfloat16()                                                 # bf16, 36 channels: scalar path
        refb = fn(xb.float().reshape(3, 6, 10, 36).permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(-1, 36)
        assert rel(ops.pool2x2(xb, 3, 6, 10, is_max=is_max), refb) < 4e-3
    lat, top = torch.randn(2 * 8 * 8, 32, device=DEV), torch.randn(2 * 4 * 4, 32, device=DEV)
    ref = lat.reshape(2, 8, 8, 32) + F.interpolate(top.reshape(2, 4, 4, 32).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert rel(ops.upsample2x_add(lat, top, 2, 4, 4), ref.reshape(-1, 32)) < 1e-6


@pytest.mark.parametrize("cout,with_ln", [(64, True), (32, False)])
def test_convtranspose_as_gemm(cout, with_ln):
    n, h, w, cin = 2, 8, 8, 256 if cout == 64 else 64
    x = torch.randn(n, cin, h, w, device=DEV).bfloat16()
    W = (torch.randn(cin, cout, 2, 2, device=DEV) / cin ** 0.5).bfloat16()
    bias, lw, lb = torch.randn(cout, device=DEV), torch.randn(cout, device=DEV), torch.randn(cout, device=DEV)
    feat = torch.randn(n, cout, 2 * h, 2 * w, device=DEV)
    ref = F.conv_transpose2d(x.float(), W.float(), bias, stride=2) + feat
    if with_ln:
        u = ref.mean(1, keepdim=True)
        s = (ref - u).pow(2).mean(1, keepdim=True)
        ref = lw[:, None, None] * ((ref - u) / torch.sqrt(s + 1e-6)) + lb[:, None, None]
    ref = F.gelu(ref).permute(0, 2, 3, 1).reshape(-1, cout)
    wg = W.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous()  # rows (dy,dx,co), cols ci
    g = ops.gemm(x.permute(0, 2, 3, 1).reshape(-1, cin).contiguous(), wg, out_dtype=torch.float32)
    y = ops.convt2x2_epilogue(g, bias, feat.permute(0, 2, 3, 1).reshape(-1, cout).contiguous(), lw if with_ln else None, lb if with_ln else None,
                              1e-6, n, h, w, cout, out_dtype=torch.float32)
    assert rel(y, ref) < 1e-5


@pytest.mark.parametrize("cin,k,s,p,hw", [(3, 7, 4, 3, 64), (3, 14, 14, 0, 56)])
def test_im2col_conv(cin, k, s, p, hw):
    x = torch.randn(2, cin, hw, hw, device=DEV)
    W = torch.randn(48, cin, k, k, device=DEV).bfloat16()
    bias = torch.randn(48, device=DEV)
    cols, oh, ow = ops.im2col(x, k, s, p)
    wk = torch.zeros(48, cols.shape[1], device=DEV, dtype=torch.bfloat16)
    wk[:, : cin * k * k] = W.reshape(48, -1)
    y = ops.gemm(cols, wk, bias=bias, out_dtype=torch.float32)
    ref = F.conv2d(x.bfloat16().float(), W.float(), bias, stride=s, padding=p).permute(0, 2, 3, 1).reshape(-1, 48)
    assert rel(y, ref) < 1e-5


def test_dwconv_conv3x3_bilinear():
    n, h, w, c = 2, 16, 16, 64
    x = torch.randn(n, c, h, w, device=DEV)
    W, b = torch.randn(c, 1, 7, 7, device=DEV), torch.randn(c, device=DEV)
    ref = F.conv2d(x, W, b, padding=3, groups=c).permute(0, 2, 3, 1).reshape(-1, c)
    y = ops.dwconv(x.permute(0, 2, 3, 1).reshape(-1, c).contiguous(), W.reshape(c, 7, 7).contiguous(), b, n, h, w, 7)
    assert rel(y, ref) < 1e-5
    for hh, ww, cc, od in ((5, 12, 256, torch.bfloat16), (64, 64, 256, torch.float32), (7, 6, 8, torch.float32)):   # vector path (w % 4 == 0) / scalar fallback
        xx = torch.randn(3, cc, hh, ww, device=DEV)
        WW, bb = torch.randn(cc, 1, 7, 7, device=DEV), torch.randn(cc, device=DEV)
        rr = F.conv2d(xx, WW, bb, padding=3, groups=cc).permute(0, 2, 3, 1).reshape(-1, cc)
        yy = ops.dwconv(xx.permute(0, 2, 3, 1).reshape(-1, cc).contiguous(), WW.reshape(cc, 7, 7).contiguous(), bb, 3, hh, ww, 7, out_dtype=od)
        assert yy.dtype == od and rel(yy, rr) < (4e-3 if od == torch.bfloat16 else 1e-5)
    for cin in (1, 4, 16, 64):
        cout = cin * 4
        x = torch.randn(n, cin, h, w, device=DEV)
        W, b = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5), torch.randn(cout, device=DEV)
        lw, lb = torch.randn(cout, device=DEV), torch.randn(cout, device=DEV)
        r = F.conv2d(x, W, b, stride=2, padding=1)
        u = r.mean(1, keepdim=True)
        s = (r - u).pow(2).mean(1, keepdim=True)
        ref = F.gelu(lw[:, None, None] * ((r - u) / torch.sqrt(s + 1e-6)) + lb[:, None, None]).permute(0, 2, 3, 1).reshape(-1, cout)
        y = ops.conv3x3s2_ln_gelu(x.permute(0, 2, 3, 1).reshape(-1, cin).contiguous(), W.permute(2, 3, 1, 0).reshape(9 * cin, cout).contiguous(), b,
                                  lw, lb, 1e-6, n, h, w, cin, cout)
        assert rel(y, ref) < 2e-5
    m = torch.randn(3, 64, 64, device=DEV)
    for oh, ow in ((256, 256), (120, 214), (31, 47)):
        ref = F.interpolate(m[:, None], size=(oh, ow), mode="bilinear", align_corners=False)[:, 0]
        assert rel(ops.bilinear(m, oh, ow), ref) < 1e-5
        assert torch.equal(ops.bilinear(m, oh, ow, mode=3).bool(), ops.bilinear(m, oh, ow) > 0)
        import numpy as np
        packed = ops.bilinear(m, oh, ow, mode=4).cpu().numpy()           # bit-packed masks: numpy.packbits order, planes padded to bytes
        bits = np.unpackbits(packed, axis=-1, count=oh * ow).reshape(-1, oh, ow).astype(bool)
        assert np.array_equal(bits, (ops.bilinear(m, oh, ow) > 0).cpu().numpy())
    ref = torch.sigmoid(F.interpolate(m[:, None], size=(256, 256), mode="bilinear", align_corners=False)[:, 0]) * 20 - 10
    assert rel(ops.bilinear(m, 256, 256, mode=1, scale=20.0, bias=-10.0), ref) < 1e-5


def test_rope_embedding_argmax():
    # SAM2 axial RoPE (position_encoding.py:174-216), 2 repeats along keys
    gw = gh = 8
    dim = 64
    x = torch.randn(2 * gw * gh, dim, device=DEV).bfloat16()
    fr = 1.0 / (10000.0 ** (torch.arange(0, dim, 4, device=DEV)[: dim // 4].float() / dim))
    t = torch.arange(gw * gh, device=DEV, dtype=torch.float32)
    fx, fy = torch.outer(t % gw, fr), torch.outer(torch.div(t, gw, rounding_mode="floor"), fr)
    cis = torch.cat([torch.polar(torch.ones_like(fx), fx), torch.polar(torch.ones_like(fy), fy)], -1).repeat(2, 1)
    ref = torch.view_as_real(torch.view_as_complex(x.float().reshape(-1, dim // 2, 2)) * cis).flatten(1)
    assert rel(ops.rope2d_(x.clone(), dim, gw, gh), ref) < 4e-3
    # Phi-3 rotate-half RoPE on q,k heads of packed qkv
    rows, nh, hd = 37, 4, 96
    qkv = torch.randn(rows, 3 * nh * hd, device=DEV).bfloat16()
    pos = torch.arange(5, 5 + rows, device=DEV, dtype=torch.int32)
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, device=DEV).float() / hd))
    ang = pos.float()[:, None] * inv[None]
    cos, sin = torch.cat([ang.cos(), ang.cos()], -1)[:, None], torch.cat([ang.sin(), ang.sin()], -1)[:, None]
    qk = qkv[:, : 2 * nh * hd].float().reshape(rows, 2 * nh, hd)
    rot = torch.cat([-qk[..., hd // 2:], qk[..., : hd // 2]], -1)
    ref = torch.cat([(qk * cos + rot * sin).reshape(rows, -1), qkv[:, 2 * nh * hd:].float()], 1)
    assert rel(ops.rope_neox_(qkv.clone(), pos, 2 * nh, hd, 10000.0), ref) < 4e-3
    table = torch.randn(100, 48, device=DEV).bfloat16()
    ids = torch.randint(0, 100, (17,), device=DEV)
    assert torch.equal(ops.embedding(table, ids), table[ids].float())
    lg = torch.randn(6, 32064, device=DEV)
    lg[2, 77] = lg[2, 9000] = 50.0
    assert torch.equal(ops.argmax(lg), lg.argmax(-1))


@pytest.mark.parametrize("M", [1, 3, 8])
def test_skinny_gemm_decode_path(M):
    """M <= 8 takes the HBM-streaming warp-per-column kernel: same epilogue contract as the tcgen05 path."""
    K, N = 3072, 1024
    a = torch.randn(M, K, device=DEV).bfloat16()
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias, cs, res = torch.randn(N, device=DEV), torch.randn(N, device=DEV), torch.randn(M, N, device=DEV)
    y = ops.gemm(a, w, bias=bias, act="gelu", col_scale=cs, resid=res, out_dtype=torch.float32, alpha=0.5)
    assert rel(y, F.gelu(0.5 * (a.float() @ w.float().T) + bias) * cs + res) < 1e-5
    y = ops.gemm(a, w, glu=True, act="silu", out_dtype=torch.float32)
    r = (a.float() @ w.float().T).reshape(M, -1, 2, 16)
    assert rel(y, (F.silu(r[:, :, 0]) * r[:, :, 1]).reshape(M, -1)) < 1e-5
    assert rel(ops.gemm(a, w, out_dtype=torch.float32), ops.gemm(a, w, out_dtype=torch.float32, force_block_n=64)) < 1e-5
    big = (torch.randn(16384, 3072, device=DEV) / 55).bfloat16()
    assert rel(ops.gemm(a, big, out_dtype=torch.bfloat16), a.float() @ big.float().T) < 4e-3


@pytest.mark.parametrize("H,Nk,D,window", [(32, 1709, 96, 0), (32, 3000, 96, 2047), (4, 256, 64, 0), (1, 5000, 256, 0), (8, 999, 72, 100)])
def test_attention_single_query_decode(H, Nk, D, window):
    q = torch.randn(1, 1, H, D, device=DEV).bfloat16()
    k = torch.randn(1, Nk, H, D, device=DEV).bfloat16()
    v = torch.randn(1, Nk, H, D, device=DEV).bfloat16()
    o = ops.attention(q, k, v, causal=True, window=window)
    assert rel(o, sdpa_ref(q, k, v, D ** -0.5, True, window)) < 4e-3


def test_norm_bf16_in_place_on_packed_rows():
    """IV2 q/k RMSNorm: bf16 rows that are column slices of the packed qkv matrix, normalised in place (vectorised bf16 path)."""
    qkv = (torch.randn(257, 3 * 1408, device=DEV) * 2).bfloat16()
    w = torch.randn(1408, device=DEV)
    x = qkv[:, 1408:2816]
    xf = x.float()
    ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w
    before = qkv.clone()
    ops.norm(x, w, eps=1e-6, rms=True, out=x)
    assert rel(qkv[:, 1408:2816], ref) < 4e-3
    assert torch.equal(qkv[:, :1408], before[:, :1408]) and torch.equal(qkv[:, 2816:], before[:, 2816:])


@pytest.mark.parametrize("rows,cols", [(1, 3072), (3, 1408), (100, 1024), (2047, 600)])
def test_norm_block_per_row_path(rows, cols):
    x = torch.randn(rows, cols, device=DEV) * 2 + 0.5
    w, b = torch.randn(cols, device=DEV), torch.randn(cols, device=DEV)
    assert rel(ops.norm(x, w, b, eps=1e-5, out_dtype=torch.float32), F.layer_norm(x, (cols,), w, b, 1e-5)) < 1e-5
    assert rel(ops.norm(x, w, eps=1e-5, rms=True, out_dtype=torch.float32), x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5) * w) < 1e-5


@pytest.mark.parametrize("M,N,K,glu", [(1, 9216, 3072, False), (1, 16384, 3072, True), (1, 3072, 8192, False), (3, 1024, 512, False), (1, 32064, 3072, False),
                                       (2, 9216, 3072, False), (4, 16384, 3072, True), (3, 16384, 3072, True), (4, 32064, 3072, False), (2, 1200, 1024, False),
                                       (8, 9216, 3072, False), (7, 16384, 3072, True), (5, 2056, 1024, False)])
def test_skinny_gemm_fused_rmsnorm(M, N, K, glu):
    """fp32 A rows RMS-normalised inside the weight-streaming kernel (norm weight folded into W by the caller)."""
    x = torch.randn(M, K, device=DEV) * 3
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-5)
    ref = xn @ w.float().T
    if glu:
        r = ref.reshape(M, -1, 2, 16)
        ref = (F.silu(r[:, :, 0]) * r[:, :, 1]).reshape(M, -1)
    y = ops.gemm(x, w, glu=glu, act="silu" if glu else None, out_dtype=torch.float32, a_rms_eps=1e-5)
    assert rel(y, ref) < 1e-5


@pytest.mark.parametrize("M", [2, 3, 4, 6, 8])
def test_multi_row_ring_gemv_batched_decode_shapes(M):
    """M = 2..8 rows through the multi-row TMA-ring kernel (csrc/gemv_ring.cu: every weight row streamed once for all rows, mma.sync), at the batched decode's
    o_proj / down_proj shapes: bf16 A, fp32 residual updated in place, row strides larger than the row length; and row m of the M-row result agrees
    with the single-row kernel on row m alone."""
    for N, K in ((3072, 3072), (3072, 8192)):
        abuf = torch.randn(M, K + 64, device=DEV).bfloat16()
        a = abuf[:, :K]
        w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
        x = torch.randn(M, N, device=DEV)
        ref = x + a.float() @ w.float().T
        singles = torch.cat([ops.gemm(a[m:m + 1].contiguous(), w, resid=x[m:m + 1].contiguous(), out_dtype=torch.float32) for m in range(M)])
        y = ops.gemm(a, w, resid=x, out=x)
        assert y.data_ptr() == x.data_ptr() and rel(x, ref) < 1e-5 and rel(x, singles) < 1e-5
    # bf16 output into a column slice (the qkv scratch rows), bias + activation
    K, N = 3072, 2048
    xin = torch.randn(M, K, device=DEV)
    w = (torch.randn(N, K, device=DEV) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device=DEV)
    buf = torch.zeros(M, N + 128, device=DEV, dtype=torch.bfloat16)
    ops.gemm(xin, w, bias=bias, act="gelu", out=buf[:, :N], a_rms_eps=1e-6)
    xn = xin * torch.rsqrt(xin.pow(2).mean(-1, keepdim=True) + 1e-6)
    assert rel(buf[:, :N], F.gelu(xn @ w.float().T + bias)) < 4e-3 and float(buf[:, N:].abs().max()) == 0.0


def test_rope2d_table_path_equals_direct_kernel():
    """vg_rope2d_cached (table of (cos, sin) filled once, q and k column blocks in one launch, rows wrapping around the grid period) produces the
    bits of vg_rope2d."""
    gw = gh = 8
    C = 64
    table = ops.rope2d_table(C, gw, gh, DEV)
    for dt in (torch.bfloat16, torch.float32):
        x = torch.randn(3 * gw * gh + 5, 3 * C + 8, device=DEV).to(dt)
        a, b = x.clone(), x.clone()
        ops.rope2d_(a[:, :C], C, gw, gh)
        ops.rope2d_(a[:, C:2 * C], C, gw, gh)
        ops.rope2d_(b[:, :2 * C], C, gw, gh, table=table, nblk=2)
        assert torch.equal(a, b) and not torch.equal(a[:, :2 * C], x[:, :2 * C]) and torch.equal(a[:, 2 * C:], x[:, 2 * C:])


def test_rope_kv_append_and_fused_merge_decode_attention():
    H, hd, max_len = 8, 96, 2100
    D = H * hd
    cache = torch.randn(max_len, 3 * D, device=DEV).bfloat16()
    row = torch.randn(1, 3 * D, device=DEV).bfloat16()
    pos = torch.tensor([1500], device=DEV, dtype=torch.int32)
    ref_row = ops.rope_neox_(row.clone(), pos, 2 * H, hd, 10000.0)
    q_in = row.clone()
    ops.rope_kv_append_(q_in, cache, pos, H, hd, 10000.0)
    assert torch.equal(q_in[:, :D], ref_row[:, :D])
    assert torch.equal(cache[1500, D:], ref_row[0, D:])
    nk = torch.tensor([1501], device=DEV, dtype=torch.int32)
    ws = ops.decode_workspace(1, H, hd, DEV)
    q = q_in[:, :D].unflatten(1, (H, hd))[None]
    k, v = cache[:2048, D:2 * D].unflatten(1, (H, hd))[None], cache[:2048, 2 * D:].unflatten(1, (H, hd))[None]
    for _ in range(3):  # ticket counters must self-clean between launches
        o = ops.attention(q, k, v, causal=True, nk_dev=nk, decode_ws=ws)
        assert rel(o, sdpa_ref(q, k[:, :1501], v[:, :1501], hd ** -0.5, True, 0)) < 4e-3
    assert int(ws[-H:].abs().sum()) == 0


@pytest.mark.parametrize("pos_val", [1500, 1535, 1536, 70])
def test_decode_attention_with_fused_rope_and_append(pos_val):
    """fused_rope_theta: RoPE(q), RoPE(k_new) + KV-cache append inside the decode attention kernel must be BIT-identical to
    vg_rope_kv_append followed by the plain decode attention (outputs and the appended cache row)."""
    H, hd, max_len = 8, 96, 2100
    D = H * hd
    cache0 = torch.randn(max_len, 3 * D, device=DEV).bfloat16()
    row = torch.randn(1, 3 * D, device=DEV).bfloat16()
    pos = torch.tensor([pos_val], device=DEV, dtype=torch.int32)
    nk = torch.tensor([pos_val + 1], device=DEV, dtype=torch.int32)
    ws = ops.decode_workspace(1, H, hd, DEV)
    views = lambda c: (c[:2048, D:2 * D].unflatten(1, (H, hd))[None], c[:2048, 2 * D:].unflatten(1, (H, hd))[None])
    # un-fused reference pipeline
    c1, r1 = cache0.clone(), row.clone()
    ops.rope_kv_append_(r1, c1, pos, H, hd, 10000.0)
    k1, v1 = views(c1)
    want = ops.attention(r1[:, :D].unflatten(1, (H, hd))[None], k1, v1, causal=True, nk_dev=nk, decode_ws=ws).clone()
    # fused
    c2, r2 = cache0.clone(), row.clone()
    k2, v2 = views(c2)
    for _ in range(2):
        got = ops.attention(r2[:, :D].unflatten(1, (H, hd))[None], k2, v2, causal=True, nk_dev=nk, decode_ws=ws, fused_rope_theta=10000.0)
        assert torch.equal(got, want)
    assert torch.equal(c2, c1) and torch.equal(r2, row)          # cache row appended identically; the packed row itself is left un-rotated
    assert int(ws[-H:].abs().sum()) == 0


@pytest.mark.parametrize("cin,cout,hw", [(16, 64, 32), (64, 256, 16), (8, 24, 8)])
def test_conv3x3s2_ln_gelu_gemm_path(cin, cout, hw):
    """MaskDownSampler wide stages through im2col (channels-last) + tcgen05 GEMM + LayerNorm/GELU kernel vs torch conv2d (bf16 operands)."""
    n = 2
    x = torch.randn(n, cin, hw, hw, device=DEV)
    W, b = torch.randn(cout, cin, 3, 3, device=DEV) / (3 * cin ** 0.5), torch.randn(cout, device=DEV)
    lw, lb = torch.randn(cout, device=DEV), torch.randn(cout, device=DEV)
    r = F.conv2d(x.bfloat16().float(), W.bfloat16().float(), b, stride=2, padding=1)
    u = r.mean(1, keepdim=True)
    s = (r - u).pow(2).mean(1, keepdim=True)
    ref = F.gelu(lw[:, None, None] * ((r - u) / torch.sqrt(s + 1e-6)) + lb[:, None, None]).permute(0, 2, 3, 1).reshape(-1, cout)
    xcl = x.permute(0, 2, 3, 1).reshape(-1, cin).contiguous()
    y = ops.conv3x3s2_ln_gelu_gemm(xcl, W.permute(0, 2, 3, 1).reshape(cout, 9 * cin).bfloat16().contiguous(), b, lw, lb, 1e-6, n, hw, hw, cin, cout)
    assert rel(y, ref) < 1e-4
    yb = ops.conv3x3s2_ln_gelu_gemm(xcl, W.permute(0, 2, 3, 1).reshape(cout, 9 * cin).bfloat16().contiguous(), b, lw, lb, 1e-6, n, hw, hw, cin, cout,
                                    out_dtype=torch.bfloat16)
    assert rel(yb, ref) < 4e-3


# ---------------------------------------------------------------------------------------------------- fp32-grade ops (csrc/precise.cu)
@pytest.mark.parametrize("M,N,K", [(9, 256, 256), (4096, 128, 256), (4096, 256, 256), (1, 256, 256), (300, 2048, 256), (16384, 128, 64), (77, 40, 52)])
def test_split_bf16_gemm_is_fp32_grade(M, N, K):
    """vg_split3 + vg_gemm over K' = 3K: A_hi B_hi + A_lo B_hi + A_hi B_lo with fp32 accumulation vs an fp64 reference (error ~2^-17 relative)."""
    a = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    bias = torch.randn(N, device=DEV)
    y = ops.gemm(a, ops.W3(w), bias=bias, out_dtype=torch.float32)
    ref = (a.double() @ w.double().T + bias.double()).float()
    assert rel(y, ref) < 2e-5, rel(y, ref)
    if K % 8 == 0:
        assert rel(ops.gemm(a.bfloat16(), w.bfloat16(), bias=bias, out_dtype=torch.float32), ref) > 1e-3   # what the split buys


def test_split_bf16_gemm_batched_activation_operand_and_epilogues():
    B, M, N, K = 3, 4, 5000, 32
    a, w = torch.randn(B, M, K, device=DEV), torch.randn(B, N, K, device=DEV)
    y = ops.gemm3(a, w, out_dtype=torch.float32)                       # hyper_in @ upscaled: both operands are activations
    assert rel(y, torch.bmm(a.double(), w.double().transpose(1, 2)).float()) < 2e-5
    x = torch.randn(200, 256, device=DEV)
    w1, w2 = torch.randn(512, 256, device=DEV) / 16, torch.randn(256, 512, device=DEV) / 22
    h = ops.gemm(x, ops.W3(w1), act="relu", out_dtype=torch.float32)
    y = ops.gemm(h, ops.W3(w2), resid=x, out=x.clone(), out_dtype=torch.float32)
    ref = (x.double() + torch.relu(x.double() @ w1.double().T) @ w2.double().T).float()
    assert rel(y, ref) < 2e-5


@pytest.mark.parametrize("B,H,Nq,Nk,D,causal,window", [
    (2, 8, 9, 4096, 16, False, 0),      # mask decoder token -> image (split over keys)
    (2, 8, 4096, 9, 16, False, 0),      # image -> token
    (3, 8, 9, 9, 32, False, 0),         # token self-attention
    (1, 1, 256, 1856, 256, False, 0),   # memory attention (tiny test geometry)
    (1, 4, 300, 300, 96, True, 0), (1, 4, 300, 300, 32, True, 100), (1, 4, 1, 333, 96, True, 0), (1, 2, 50, 210, 72, True, 64)])
def test_attention_exact_fp32(B, H, Nq, Nk, D, causal, window):
    q, k, v = torch.randn(B, Nq, H, D, device=DEV), torch.randn(B, Nk, H, D, device=DEV), torch.randn(B, Nk, H, D, device=DEV)
    o = ops.attention(q, k, v, causal=causal, window=window)
    assert o.dtype == torch.float32
    assert rel(o, sdpa_ref(q, k, v, D ** -0.5, causal, window)) < 5e-6


@pytest.mark.parametrize("pooled", [False, True])
def test_attention_exact_fp32_hiera_windows(pooled):
    n_img, gh, gw, win, H, D = 2, 16, 24, 8, 2, 72
    qkv = torch.randn(n_img * gh * gw, 3, H, D, device=DEV)
    q, k, v = qkv.unbind(1)
    nwy, nwx = gh // win, gw // win

    def part(t, g_h, g_w, ws):
        t = t.reshape(n_img, g_h // ws, ws, g_w // ws, ws, H, D).permute(0, 1, 3, 2, 4, 5, 6)
        return t.reshape(-1, ws * ws, H, D)

    if pooled:
        qp = F.max_pool2d(q.reshape(n_img, gh, gw, H * D).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).reshape(-1, H, D).contiguous()
        o = ops.attention(qp[None], k[None], v[None], q_win=(win // 2, gw // 2, nwx, nwy), k_win=(win, gw, nwx, nwy))
        ref = sdpa_ref(part(qp, gh // 2, gw // 2, win // 2), part(k, gh, gw, win), part(v, gh, gw, win), D ** -0.5)
        assert rel(part(o[0], gh // 2, gw // 2, win // 2), ref) < 5e-6
    else:
        o = ops.attention(q[None], k[None], v[None], q_win=(win, gw, nwx, nwy), k_win=(win, gw, nwx, nwy))
        assert rel(part(o[0], gh, gw, win), sdpa_ref(part(q, gh, gw, win), part(k, gh, gw, win), part(v, gh, gw, win), D ** -0.5)) < 5e-6


def test_rope_and_im2col_fp32_variants():
    x = torch.randn(64 * 3, 256, device=DEV)
    xb = x.bfloat16().float()
    a = ops.rope2d_(xb.clone(), 256, 8, 8)
    b = ops.rope2d_(xb.bfloat16(), 256, 8, 8)
    assert rel(b, a) < 4e-3 and a.dtype == torch.float32
    pos = torch.arange(5, 5 + 10, device=DEV, dtype=torch.int32)
    qkv = torch.randn(10, 3 * 4 * 32, device=DEV).bfloat16().float()
    assert rel(ops.rope_neox_(qkv.bfloat16(), pos, 8, 32, 10000.0), ops.rope_neox_(qkv.clone(), pos, 8, 32, 10000.0)) < 4e-3
    img = torch.randn(2, 3, 28, 28, device=DEV)
    c32, oh, ow = ops.im2col(img, 14, 14, 0, out_dtype=torch.float32)
    assert torch.equal(c32[:, :588].view(2, 2, 2, 3, 14, 14), img.view(2, 3, 2, 14, 2, 14).permute(0, 2, 4, 1, 3, 5))
