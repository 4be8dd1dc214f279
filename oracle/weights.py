"""TEST INFRASTRUCTURE — deterministic synthetic weights with the reference's state_dict keys and shapes.

There are no checkpoints on the box, so parity runs on seeded random weights. The shape tables
below restate the parameter layout of the reference modules; tests/test_oracle_vs_reference.py
checks that they load with strict=True into the reference classes. Values are drawn on CPU from a seeded
torch.Generator (identical on every machine), with NON-ZERO biases / norm offsets so that parity tests see them.
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import torch

from .sam2 import Sam2Cfg


from videoglamm_b200.statedict import lmm_param_shapes, sam2_param_shapes  # noqa: E402,F401  (the key/shape contract lives with the loader)


_EMBED_RE = re.compile(r"(^|\.)(iou_token|mask_tokens|obj_score_token|not_a_point_embed|no_mask_embed|point_embeddings\.\d+|embed_tokens|"
                       r"position_embedding|class_embedding|cls_token|\w*pos_embed\w*)(\.weight)?$")


def _fill(name: str, shape: tuple, g: torch.Generator) -> torch.Tensor:
    """Deterministic values by parameter role."""
    leaf = name.rsplit(".", 1)[-1]
    parent = name.rsplit(".", 2)[-2] if name.count(".") >= 1 else ""
    is_norm = any(k in parent for k in ("norm", "layernorm", "ln_", "ln1", "ln2", "layer_norm", "layrnorm")) or \
        parent in ("1", "4") and ("mask_downscaling" in name or "output_upscaling" in name) or \
        ("mask_downsampler.encoder" in name and len(shape) == 1 and int(parent) % 3 == 1)
    if "positional_encoding_gaussian_matrix" in name:
        return torch.randn(shape, generator=g)
    if is_norm and leaf == "weight":
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if is_norm and leaf == "bias":
        return 0.1 * torch.randn(shape, generator=g)
    if leaf in ("gamma",) or (leaf == "weight" and len(shape) == 1):  # LayerScale-style vectors
        return 0.1 + 0.05 * torch.randn(shape, generator=g)
    if leaf == "bias" or (len(shape) == 1):
        return 0.1 * torch.randn(shape, generator=g)
    if len(shape) >= 2 and leaf == "weight" and not _EMBED_RE.search(name):
        fan_in = math.prod(shape[1:]) if "output_upscaling" not in name else shape[0]
        bound = math.sqrt(3.0 / fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound
    return 0.5 * torch.randn(shape, generator=g)  # embeddings, tokens, pos tables


def make_state_dict(shapes: "OrderedDict[str, tuple]", seed: int) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return OrderedDict((k, _fill(k, tuple(v), g).float()) for k, v in shapes.items())


def make_sam2_state_dict(cfg: Sam2Cfg, seed: int = 1234, obj_score_bias: float = 4.0):
    """`obj_score_bias` shifts the object-score head so that, with random weights, objects are predicted present
    (positive logit, masks kept) — pass a negative value to exercise the NO_OBJ_SCORE branch instead."""
    sd = make_state_dict(sam2_param_shapes(cfg), seed)
    sd["sam_mask_decoder.pred_obj_score_head.layers.2.bias"] += obj_score_bias
    return sd


def make_lmm_state_dict(cfg, seed: int = 1234):
    return make_state_dict(lmm_param_shapes(cfg), seed)


def make_videoglamm_state_dict(lmm_cfg, sam_cfg: Sam2Cfg, seed: int = 1234, obj_score_bias: float = 4.0):
    """Full VideoGLaMMForCausalLM state dict: LMM keys + SAM2 under `model.visual_model.` (VideoGLaMM.py:107-145)."""
    sd = make_lmm_state_dict(lmm_cfg, seed)
    for k, v in make_sam2_state_dict(sam_cfg, seed + 1, obj_score_bias).items():
        sd["model.visual_model." + k] = v
    return sd
