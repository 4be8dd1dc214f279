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


def _header_struct(name: str):
    """Parses `typedef struct <name> { ... } <name>;` of include/vglamm.h into [(field, size, align)] (LP64: pointers / long long 8, int / float 4)."""
    text = (ROOT / "include" / "vglamm.h").read_text()
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const )?(void|float|int|long long|unsigned char)\s*(\*?)\s*(.*)", decl)
        assert m, decl
        base, ptr, names = m.group(2), m.group(3), m.group(4)
        for nm in names.split(","):
            nm = nm.strip()
            is_ptr = bool(ptr) or nm.startswith("*")
            nm = nm.lstrip("*").strip()
            size = 8 if (is_ptr or base == "long long") else 4
            fields.append((nm, size, size))
    return fields


@pytest.mark.parametrize("name", ["VgGemm", "VgAttention"])
def test_ctypes_structs_match_the_header_field_by_field(name):
    """Every field of the C struct, in order, with the offset the C compiler gives it, equals the ctypes mirror (a stale mirror makes the
    library read past the caller's struct)."""
    from videoglamm_b200 import _lib as L
    cls = getattr(L, name)
    want = _header_struct(name)
    assert [f[0] for f in cls._fields_] == [w[0] for w in want]
    off = 0
    for (fname, size, align), (cname, ctype) in zip(want, cls._fields_):
        off = (off + align - 1) // align * align
        assert getattr(cls, cname).offset == off and ctypes.sizeof(ctype) == size, (fname, off, getattr(cls, cname).offset)
        off += size
    assert ctypes.sizeof(cls) == (off + 7) // 8 * 8


def test_integration_md_binding_matches_the_header():
    """The ctypes stub INTEGRATION.md shows a maintainer is executable and mirrors the header exactly."""
    from videoglamm_b200 import _lib as L
    md = (ROOT / "INTEGRATION.md").read_text()
    block = re.search(r"class VgGemm\(ctypes.Structure\):.*?\n\n", md, re.S).group(0)
    ns = {"ctypes": ctypes}
    exec(block, ns)
    doc = ns["VgGemm"]
    assert [f[0] for f in doc._fields_] == [f[0] for f in L.VgGemm._fields_]
    assert ctypes.sizeof(doc) == ctypes.sizeof(L.VgGemm)
    assert all(getattr(doc, n).offset == getattr(L.VgGemm, n).offset for n, _ in doc._fields_)


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
