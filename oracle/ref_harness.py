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
EXTRA_SAM2_OVERRIDES: list = []   # appended to every hydra compose() while build_videoglamm runs (tiny SAM2 for the e2e pin)
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
    for ov in list(overrides) + list(EXTRA_SAM2_OVERRIDES):
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


# ------------------------------------------------------------------------------------------------ full VideoGLaMMForCausalLM
def build_videoglamm(workdir: str, *, phi: dict, clip: dict, iv2_depth: int, sam2_overrides: list, seg_token_idx: int):
    """The reference's VideoGLaMMForCausalLM (VideoGLaMM.py:879-903) built the way chat.py:280-352 builds it, on CPU, at REDUCED
    depth/width where the reference lets a config decide (Phi-3 config, CLIP config dir, SAM2 hydra overrides) and at reduced IV2
    DEPTH only (width 1408 / 224 px / 16x16 grid are hard-coded in arch.py:133-136 and multimodal_projector/builder.py:17-27).
    Nothing under /root/reference is modified: `workdir` holds the CLIP config dir, the dummy IV2 checkpoint and a `model` symlink for
    the relative IV2 config path."""
    global EXTRA_SAM2_OVERRIDES
    install_shims()
    EXTRA_SAM2_OVERRIDES = list(sam2_overrides)
    wd = Path(workdir)
    wd.mkdir(parents=True, exist_ok=True)
    for name in ("model", "utils"):
        if not (wd / name).exists():
            (wd / name).symlink_to(VG / name)
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModel
    clip_dir = wd / "openai" / "clip-vit-large-patch14-336"
    if not (clip_dir / "config.json").exists():
        clip_dir.mkdir(parents=True, exist_ok=True)
        ccfg = CLIPVisionConfig(hidden_size=clip["width"], intermediate_size=clip["mlp"], num_hidden_layers=clip["layers"],
                                num_attention_heads=clip["heads"], image_size=clip["image_size"], patch_size=clip["patch"], hidden_act="quick_gelu")
        CLIPVisionModel(ccfg).save_pretrained(clip_dir)
        CLIPImageProcessor(size={"shortest_edge": clip["image_size"]}, crop_size={"height": clip["image_size"], "width": clip["image_size"]}).save_pretrained(clip_dir)
    iv2_dir = wd / "OpenGVLab" / "InternVideo2-Stage2_1B-224p-f4"
    iv2_pt = iv2_dir / "InternVideo2-stage2_1b-224p-f4.pt"
    if not iv2_pt.exists():
        iv2_dir.mkdir(parents=True, exist_ok=True)
        torch.save({"vision_encoder.pos_embed": torch.zeros(1, 1025, 1408), "vision_encoder.clip_pos_embed": torch.zeros(1, 1025, 1408)}, iv2_pt)
    import model.videogpt_plus.model.internvideo.internvideo2 as iv
    orig_cls = getattr(iv, "_vg_orig_PretrainInternVideo2", iv.PretrainInternVideo2)
    iv._vg_orig_PretrainInternVideo2 = orig_cls

    def shallow(*a, **k):  # same class, same code; only the depth argument of the 1B factory (internvideo2.py:698-723) is overridden
        k["depth"] = iv2_depth
        k["checkpoint_num"] = 0
        k["use_checkpoint"] = False
        return orig_cls(*a, **k)

    iv.PretrainInternVideo2 = shallow
    cwd = os.getcwd()
    os.chdir(wd)
    try:
        from model.VideoGLaMM import VideoGLaMMForCausalLM
        from model.videogpt_plus.model.language_model.phi3 import VideoGPTPlusPhi3Config
        cfg = VideoGPTPlusPhi3Config(vocab_size=phi["vocab"], hidden_size=phi["hidden"], intermediate_size=phi["inter"], num_hidden_layers=phi["layers"],
                                     num_attention_heads=phi["heads"], num_key_value_heads=phi["heads"], rms_norm_eps=1e-5, rope_theta=10000.0,
                                     sliding_window=None, max_position_embeddings=4096, original_max_position_embeddings=4096, pad_token_id=0,
                                     bos_token_id=1, eos_token_id=2, attn_implementation="eager")
        cfg.mm_vision_tower = cfg.vision_tower = str(iv2_pt.relative_to(wd))
        cfg.image_mm_vision_tower = cfg.image_vision_tower = "openai/clip-vit-large-patch14-336"
        cfg.mm_projector_type = cfg.image_mm_projector_type = "mlp2x_gelu"
        cfg.mm_vision_select_layer, cfg.mm_vision_select_feature = -2, "patch"
        cfg.pretrain_mm_mlp_adapter = cfg.pretrain_image_mm_mlp_adapter = None
        cfg.mm_use_im_start_end = cfg.mm_use_im_patch_token = False
        cfg.use_sam2, cfg.train_mask_decoder, cfg.out_dim, cfg.mask_decoder_itm, cfg.seg_token_idx = True, True, 256, False, seg_token_idx
        model = VideoGLaMMForCausalLM(cfg, use_sam2_video_branch=True)
        model.get_model().initialize_vision_modules(model.get_model().config)
        model = model.float().eval()
    finally:
        os.chdir(cwd)
        iv.PretrainInternVideo2 = orig_cls
        EXTRA_SAM2_OVERRIDES = []
    patch_predictor_cpu(model.model.visual_model)
    return model


def force_tokens(model, forced: dict):
    """random weights never emit [SEG], so chosen decode steps are overwritten with fixed ids on both sides. Here: a
    LogitsProcessor handed to the reference's own (unmodified) generate() call at VideoGLaMM.py:790-799."""
    from transformers import LogitsProcessor, LogitsProcessorList

    class Forced(LogitsProcessor):
        def __init__(self):
            self.len0 = None

        def __call__(self, input_ids, scores):
            if self.len0 is None:
                self.len0 = input_ids.shape[1]
            step = input_ids.shape[1] - self.len0
            if step in forced:
                scores = torch.full_like(scores, -float("inf"))
                scores[:, forced[step]] = 0.0
            return scores

    orig = model.generate

    def gen(**kw):
        return orig(logits_processor=LogitsProcessorList([Forced()]), do_sample=False, **kw)

    model.generate = gen
    return model


def fix_clip_hidden_states(model):
    """transformers drift shim (reference pins 4.41.0, README.md:76; 5.5.0 is installed): under 5.5 a CLIPVisionModel called from INSIDE
    GenerationMixin.generate records every encoder layer output twice (hidden_states has 2L+1 entries instead of L+1), so the reference's
    `hidden_states[-2]` (clip_encoder.py:34-36) silently selects the LAST layer instead of the penultimate one; outside generate() the same
    call returns L+1 entries. Parity is defined on the pinned behaviour: drop the duplicated entries."""
    vt = model.get_model().get_image_vision_tower().vision_tower
    orig = vt.forward

    def fwd(*a, **k):
        out = orig(*a, **k)
        hs = out.hidden_states
        if hs is not None:
            keep = [hs[0]] + [h for p, h in zip(hs[:-1], hs[1:]) if h is not p and not (h.shape == p.shape and torch.equal(h, p))]
            out.hidden_states = tuple(keep)
        return out

    vt.forward = fwd
    return model
