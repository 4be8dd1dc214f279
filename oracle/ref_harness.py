"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference (mbzuai-oryx/VideoGLaMM) from /root/reference on CPU.

Only used (a) by tests/golden/make_golden.py to freeze known-answer vectors and (b) by `-m "not gpu"` tests that
validate oracle/*.py against the reference's own modules. /root/reference does not exist on the GPU box, so
nothing reachable from `-m gpu` tests, smoke() or bench.py's own arm may import this module.

Shims: the reference needs hydra/omegaconf/timm (absent here) and calls `.cuda()`
literally; we stub the former and make the latter the identity. No reference source is copied.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import os
import re
import sys
import types
from pathlib import Path

import torch
import yaml

REF_ROOT = Path(os.environ.get("VG_REFERENCE_ROOT", "/root/reference"))
VG = REF_ROOT / "VideoGLaMM"
_installed = False
_FLOAT_RE = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")


def available() -> bool:
    return (VG / "model" / "VideoGLaMM.py").exists()


def _coerce(v):
    if isinstance(v, str) and _FLOAT_RE.match(v):
        return float(v)  # PyYAML reads "1e-6" as str
    return v


def _instantiate(node, **_):
    if isinstance(node, dict):
        kw = {k: _instantiate(v) for k, v in node.items() if k != "_target_"}
        if "_target_" in node:
            mod, _, name = node["_target_"].rpartition(".")
            return getattr(importlib.import_module(mod), name)(**kw)
        return kw
    if isinstance(node, list):
        return [_instantiate(v) for v in node]
    return _coerce(node)


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _compose(config_name: str, overrides=()):
    path = VG / "model" / "segment_anything_2" / "sam2_configs" / config_name
    cfg = yaml.safe_load(path.read_text())
    for ov in overrides:
        key, _, val = ov.lstrip("+").partition("=")
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = yaml.safe_load(val)
    return _Cfg(cfg)


def install_shims() -> None:
    """Make `import model.…` resolve to the reference tree and stub its missing third-party imports."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__spec__ = importlib.machinery.ModuleSpec(name, None)  # transformers probes find_spec() on optional deps
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("hydra", compose=_compose, initialize_config_module=lambda *a, **k: None)
    mod("hydra.utils", instantiate=_instantiate)
    sys.modules["hydra"].utils = sys.modules["hydra.utils"]
    mod("omegaconf", OmegaConf=types.SimpleNamespace(resolve=lambda cfg: None))
    try:
        import timm  # noqa: F401
    except Exception:
        layers = mod("timm.models.layers", DropPath=torch.nn.Identity, to_2tuple=lambda x: (x, x) if not isinstance(x, tuple) else x,
                     trunc_normal_=torch.nn.init.trunc_normal_)
        models = mod("timm.models", layers=layers)
        mod("timm", models=models)
    for name in ("decord", "matplotlib", "matplotlib.pyplot", "skimage", "deepspeed", "peft"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                mod(name)
    # literal .cuda() calls on the reference path -> identity (CPU oracle)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    if str(VG) not in sys.path:
        sys.path.insert(0, str(VG))
    _installed = True


def build_sam2_predictor(seed: int = 1234, dtype=torch.float32):
    """The reference's SAM2VideoPredictor (Hiera-L config) with seeded random weights, on CPU."""
    install_shims()
    from model.segment_anything_2.sam2.build_sam import build_sam2_video_predictor

    torch.manual_seed(seed)
    model = build_sam2_video_predictor("sam2_hiera_l.yaml", ckpt_path=None, device=None)
    return model.to(dtype).eval()


def patch_predictor_cpu(predictor):
    """init_state_from_tensor hard-codes cuda devices; force CPU."""
    orig = predictor.init_state_from_tensor

    def wrapped(*a, **k):
        st = orig(*a, **k)
        st["device"] = torch.device("cpu")
        st["storage_device"] = torch.device("cpu")
        return st

    predictor.init_state_from_tensor = wrapped
    return predictor
