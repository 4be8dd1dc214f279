"""ctypes binding of libvglamm.so (C ABI declared in include/vglamm.h).

The product path has NO fallback: if the library is missing or a call fails, a RuntimeError is raised."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_LIB_PATH = Path(__file__).resolve().parent / "lib" / "libvglamm.so"
_lib = None

BF16, F32 = 0, 1
ACT_NONE, ACT_GELU, ACT_QUICKGELU, ACT_RELU, ACT_SILU, ACT_SIGMOID = range(6)
ACT = {None: ACT_NONE, "none": ACT_NONE, "gelu": ACT_GELU, "quick_gelu": ACT_QUICKGELU, "relu": ACT_RELU,
       "silu": ACT_SILU, "sigmoid": ACT_SIGMOID}


class VgGemm(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_longlong), ("strideA", C.c_longlong),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("strideB", C.c_longlong),
        ("C", C.c_void_p), ("ldc", C.c_longlong), ("strideC", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("batch", C.c_int),
        ("out_dtype", C.c_int),
        ("bias", C.c_void_p), ("strideBias", C.c_longlong),
        ("act", C.c_int), ("glu", C.c_int),
        ("col_scale", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_longlong), ("strideR", C.c_longlong), ("resid_dtype", C.c_int),
        ("alpha", C.c_float),
        ("force_block_n", C.c_int),
        ("a_dtype", C.c_int), ("a_rms_eps", C.c_float),
    ]


def lib() -> C.CDLL:
    """Load libvglamm.so (built in-tree by videoglamm_b200.build). Raises if it is absent."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(f"{_LIB_PATH} not built: run `python -m videoglamm_b200.build` (no CPU fallback exists)")
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.vg_last_error.restype = C.c_char_p
        _lib.vg_launch_count.restype = C.c_longlong
        _lib.vg_launch_count_reset.restype = None
        _lib.vg_launch_count_add.restype = None
        _lib.vg_launch_count_add.argtypes = [C.c_longlong]
        _lib.vg_mask_decoder_workspace_bytes.restype = C.c_longlong
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().vg_last_error().decode()}")


class VgAttention(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        *[(n, C.c_longlong) for n in ("q_bs", "q_hs", "q_rs", "k_bs", "k_hs", "k_rs", "v_bs", "v_hs", "v_rs", "o_bs", "o_hs", "o_rs")],
        ("B", C.c_int), ("H", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("D", C.c_int),
        ("scale", C.c_float),
        ("causal", C.c_int), ("window", C.c_int),
        ("q_win", C.c_int), ("q_gridw", C.c_int), ("q_nwx", C.c_int), ("q_nwy", C.c_int),
        ("k_win", C.c_int), ("k_gridw", C.c_int), ("k_nwx", C.c_int), ("k_nwy", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong), ("ws_has_counters", C.c_int), ("force_legacy", C.c_int),
        ("nk_dev", C.c_void_p),
        ("fused_rope_theta", C.c_float), ("nk_per_batch", C.c_int),
    ]


# every symbol include/vglamm.h declares (tests check that the built library exports all of them)
SYMBOLS = [
    "vg_last_error", "vg_version", "vg_launch_count", "vg_launch_count_reset", "vg_gemm", "vg_gemm_naive", "vg_attention", "vg_norm",
    "vg_add", "vg_pool2x2", "vg_upsample2x_add", "vg_convt2x2_epilogue", "vg_im2col", "vg_dwconv", "vg_conv3x3s2_ln_gelu", "vg_bilinear",
    "vg_rope2d", "vg_rope2d_table", "vg_rope2d_cached", "vg_rope_neox", "vg_embedding", "vg_argmax", "vg_row_scatter", "vg_launch_count_add", "vg_rope_kv_append", "vg_multimask_select",
    "vg_connected_components", "vg_im2col3x3s2_cl", "vg_layernorm_gelu", "vg_split3", "vg_attention_f32",
    "vg_mask_decoder", "vg_mask_decoder_workspace_bytes", "vg_mask_decoder_set_timeline",
    "vg_resize_u8_pil", "vg_sam_normalize_squash", "vg_resize_u8_cv", "vg_u8_normalize_chw", "vg_sum_parts",
]


class VgLinear(C.Structure):
    _fields_ = [("w", C.c_void_p), ("b", C.c_void_p)]


class VgMdAttn(C.Structure):
    _fields_ = [("q", VgLinear), ("k", VgLinear), ("v", VgLinear), ("out", VgLinear)]


class VgMdLayer(C.Structure):
    _fields_ = [("sa", VgMdAttn), ("t2i", VgMdAttn), ("i2t", VgMdAttn), ("n1", VgLinear), ("n2", VgLinear), ("n3", VgLinear), ("n4", VgLinear),
                ("fc1", VgLinear), ("fc2", VgLinear)]


class VgMaskDecoderWeights(C.Structure):
    _fields_ = [("layer", VgMdLayer * 2), ("final_t2i", VgMdAttn), ("norm_final", VgLinear), ("up1", VgLinear), ("up_ln", VgLinear), ("up2", VgLinear),
                ("hyper", (VgLinear * 3) * 4), ("iou", VgLinear * 3), ("obj", VgLinear * 3), ("pek", C.c_void_p * 3), ("peq", C.c_void_p * 2),
                ("img_hi", C.c_void_p * 12), ("img_lo", C.c_void_p * 12)]


class VgMaskDecoder(C.Structure):
    _fields_ = [("pix", C.c_void_p), ("pix_shared", C.c_int), ("dense_pe", C.c_void_p), ("no_mask", C.c_void_p), ("s0", C.c_void_p), ("s1", C.c_void_p),
                ("tokens", C.c_void_p), ("B", C.c_int), ("Nt", C.c_int), ("fs", C.c_int), ("C", C.c_int), ("heads", C.c_int), ("mlp", C.c_int),
                ("depth", C.c_int), ("weights", C.POINTER(VgMaskDecoderWeights)), ("masks", C.c_void_p), ("iou", C.c_void_p), ("mask_tokens", C.c_void_p),
                ("obj", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong), ("force_exact", C.c_int)]

