"""Builds libvglamm.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc. No torch extension machinery:
the library has a plain C ABI (include/vglamm.h) and is loaded with ctypes."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
LIB = ROOT / "lib" / "libvglamm.so"
STAMP = ROOT / "lib" / ".build_stamp"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT.parent / "include" / "vglamm.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every .cu under csrc/ into lib/libvglamm.so (skipped when sources are unchanged)."""
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    LIB.parent.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    objdir = ROOT / "lib" / "obj"
    objdir.mkdir(exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    for src in _sources():
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)] + (["-Xptxas", "-v"] if verbose else [])
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(f"--- nvcc {src.name}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building libvglamm.so")
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
