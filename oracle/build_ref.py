"""TEST INFRASTRUCTURE — builds the reference's own native extension (sam2._C, connected components) from the UNMODIFIED source where it lies
under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box). No reference source is copied into the repo.

    python oracle/build_ref.py        # needs /root/reference, nvcc and the torch headers; ~2 minutes

The result (`oracle/_ref/vg_ref_sam2_C.so`) is a torch C++ extension compiled for sm_100a; tests/test_cc_gpu.py loads it when present and
checks libvglamm's `vg_connected_components` against it bit for bit."""
from __future__ import annotations

import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
OUT = ROOT / "_ref"
SRC = Path(os.environ.get("VG_REFERENCE_ROOT", "/root/reference")) / "VideoGLaMM/model/segment_anything_2/sam2/csrc/connected_components.cu"
NAME = "vg_ref_sam2_C"


def built() -> Path | None:
    p = OUT / f"{NAME}.so"
    return p if p.exists() else None


def build(verbose: bool = False) -> Path | None:
    if built():
        return built()
    if not SRC.exists():
        return None
    from torch.utils.cpp_extension import load

    OUT.mkdir(exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    load(name=NAME, sources=[str(SRC)], build_directory=str(OUT), extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-O3"],
         verbose=verbose, is_python_module=False)
    return built()


def load_ref():
    """Returns the reference's get_connected_componnets(uint8 [N,1,H,W] cuda) -> [labels, counts], or None when the extension is not built."""
    p = built()
    if p is None:
        return None
    import importlib.util

    import torch  # noqa: F401  (libtorch must be loaded first)

    spec = importlib.util.spec_from_file_location(NAME, str(p))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.get_connected_componnets


def stage_reference() -> Path | None:
    """Stages the UNMODIFIED reference python package (VideoGLaMM/{model,utils}: .py + .yaml, ~2 MB) under baseline/_ref/ (git-ignored, travels
    to the GPU box like oracle/_ref) so that `bench.py --impl reference` can execute the reference's own VideoGLaMMForCausalLM.inference there
    through oracle/ref_harness.py (VG_REFERENCE_ROOT). Nothing is copied into tracked paths; returns the staged root or None without /root/reference."""
    import shutil
    src = Path(os.environ.get("VG_REFERENCE_SRC", "/root/reference")) / "VideoGLaMM"
    dst = ROOT.parent / "baseline" / "_ref" / "VideoGLaMM"
    if (dst / "model" / "VideoGLaMM.py").exists():
        return dst.parent
    if not (src / "model" / "VideoGLaMM.py").exists():
        return None
    keep = lambda d, names: [n for n in names if (Path(d) / n).is_file() and not n.endswith((".py", ".yaml", ".yml", ".json", ".txt"))]
    for sub in ("model", "utils"):
        shutil.copytree(src / sub, dst / sub, ignore=keep, dirs_exist_ok=True)
    return dst.parent


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    print(stage_reference())
