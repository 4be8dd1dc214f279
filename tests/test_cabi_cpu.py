"""No-GPU checks of the boundary: the C-ABI library builds for sm_100a, loads on a CPU-only box, exports every symbol
include/vglamm.h declares, and the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from videoglamm_b200 import _lib as L
    lib = L.lib()
    declared = set(re.findall(r"\b(vg_[a-z0-9_]+)\s*\(", (ROOT / "include" / "vglamm.h").read_text()))
    assert declared == set(L.SYMBOLS), declared ^ set(L.SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None
    assert lib.vg_version() == 1


def test_struct_layouts_match_header_sizes():
    from videoglamm_b200 import _lib as L
    # pointers/long long are 8 bytes, ints/floats 4: sizes computed by ctypes must be stable multiples of 8
    assert ctypes.sizeof(L.VgGemm) % 8 == 0 and ctypes.sizeof(L.VgAttention) % 8 == 0


def test_bad_arguments_return_error_codes_not_crashes():
    from videoglamm_b200 import _lib as L
    lib = L.lib()
    g = L.VgGemm()
    assert lib.vg_gemm(ctypes.byref(g), None) == -1  # VG_EINVAL: null pointers
    assert b"null" in lib.vg_last_error()
    a = L.VgAttention()
    assert lib.vg_attention(ctypes.byref(a), None) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only check")
def test_no_cpu_fallback():
    from videoglamm_b200 import ops
    with pytest.raises(RuntimeError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    from videoglamm_b200.sam2_engine import Sam2Engine
    with pytest.raises(RuntimeError):
        Sam2Engine({}, device="cpu")
