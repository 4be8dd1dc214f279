"""Drop-in surface of the reference's top model for inference: `VideoGLaMMForCausalLM.from_pretrained(...)` and
`.inference(images, context_images, images_for_sam, input_ids, resize_list, original_size_list, max_new_tokens,
use_sam2_video_branch)` with the same argument meaning, return value and error behaviour as
/root/reference/VideoGLaMM/model/VideoGLaMM.py:560-596 (dispatch) and :770-879 (video branch), plus every attribute / method the
reference's chat.py touches between `from_pretrained` and `inference` (chat.py:280-352): `resize_token_embeddings`, `config.*`,
`get_model().initialize_vision_modules(cfg)`, `get_vision_tower()` / `get_image_vision_tower()` (objects with `.to()`), `.bfloat16()` /
`.float()` / `.half()` / `.cuda()` / `.eval()`, `.model`, and the `forward(**kwargs)` dispatch of VideoGLaMM.py:897-903.
All arithmetic runs in libvglamm kernels through LmmEngine / Sam2Engine; there is no CPU path: weights are held on the host until the
engines are built (`.cuda()` or the first `inference`), which needs a CUDA device.
"""
from __future__ import annotations

import glob
import json
import os
from types import SimpleNamespace

import numpy as np
import torch

from ..config import LmmConfig, Sam2Config, make_lmm_config
from ..statedict import validate_state_dict, videoglamm_param_shapes

SAM_PREFIX = "model.visual_model."
CLIP_PREFIX = "model.image_vision_tower.vision_tower."
IV2_PREFIX = "model.vision_tower."


class _TowerHandle:
    """What `get_vision_tower()` / `get_image_vision_tower()` return: the reference only moves / casts these modules (chat.py:316-319) and reads
    `is_loaded`; the weights themselves live in the engine."""

    def __init__(self, name: str):
        self.name, self.is_loaded = name, True

    def to(self, *_, **__):
        return self

    half = bfloat16 = float = cuda = eval = to

    def __repr__(self):
        return f"<{self.name} tower of the B200 engine>"


def resolve_clip_dir(name_or_path: str | None) -> str | None:
    """`config.image_mm_vision_tower` is a local HF directory or a hub id ("openai/clip-vit-large-patch14-336", what released checkpoints store).
    Hub ids are resolved through the local HF cache only (there is no network path in this engine); None when it cannot be found."""
    if not name_or_path:
        return None
    if os.path.isdir(name_or_path):
        return name_or_path
    try:
        from huggingface_hub import snapshot_download
        return snapshot_download(name_or_path, local_files_only=True)
    except Exception:
        return None


def load_clip_vision_weights(clip_dir: str) -> dict:
    files = sorted(glob.glob(os.path.join(clip_dir, "*.safetensors")))
    if files:
        from safetensors.torch import load_file
        csd = {}
        for f in files:
            csd.update(load_file(f))
    else:
        csd = torch.load(os.path.join(clip_dir, "pytorch_model.bin"), map_location="cpu")
    return {CLIP_PREFIX + k: v for k, v in csd.items() if k.startswith("vision_model.")}


def load_iv2_weights(path: str) -> dict:
    isd = torch.load(path, map_location="cpu")
    isd = isd.get("model", isd.get("module", isd))
    return {IV2_PREFIX + k: v for k, v in isd.items() if k.startswith("vision_encoder.")}


class VideoGLaMMForCausalLM:
    """Inference engine behind the reference's class name. Not an nn.Module: weights are packed once for the kernels."""

    def __init__(self, state_dict: dict, lmm_cfg: LmmConfig | None = None, sam_cfg: Sam2Config | None = None, *, use_sam2_video_branch: bool = True,
                 eos_token_ids=(), device: str = "cuda", validate: bool = True, precise: str = "heads", defer: bool = False, hf_config: dict | None = None):
        """precise: "heads" (default: SAM2 mask decoder in fp32-grade arithmetic, everything else bf16 operands), "none", or "all" (verification
        mode: every GEMM split-bf16, every attention exact fp32 — see Sam2Engine / LmmEngine). defer=True keeps the weights on the host and builds
        the engines at `.cuda()` / the first `inference` (what `from_pretrained` does, so that the reference's chat.py sequence works)."""
        self.lmm_cfg = lmm_cfg or make_lmm_config()
        self.sam_cfg = sam_cfg or Sam2Config()
        self.precise = precise
        self._validate = validate
        self._sd = state_dict
        self.device = torch.device(device)
        self.lmm = self.sam = None
        self.use_sam2 = True
        self.use_sam2_video_branch = use_sam2_video_branch
        cfg = dict(hf_config or {})
        cfg.update(seg_token_idx=self.lmm_cfg.seg_token_idx, use_sam2=True, eos_token_id=list(eos_token_ids), out_dim=self.lmm_cfg.out_dim,
                   mm_use_im_start_end=cfg.get("mm_use_im_start_end", False), mm_use_im_patch_token=cfg.get("mm_use_im_patch_token", False))
        self.config = SimpleNamespace(**cfg)
        self.model = self                                   # chat.py touches `model.model.vision_tower` on its fp16 path
        self._towers = {"iv2": _TowerHandle("InternVideo2"), "clip": _TowerHandle("CLIP")}
        if not defer:
            self._ensure_engines()

    # ---- engines ----------------------------------------------------------------------------------------------------------------------------
    def _ensure_engines(self):
        if self.lmm is not None:
            return
        if not torch.cuda.is_available():
            raise RuntimeError("videoglamm_b200 needs a CUDA device to build its engines (there is no CPU path)")
        from ..lmm_engine import LmmEngine
        from ..sam2_engine import Sam2Engine
        self.lmm_cfg.seg_token_idx = int(getattr(self.config, "seg_token_idx", self.lmm_cfg.seg_token_idx))   # chat.py:303 assigns it after loading
        if self._validate:
            validate_state_dict(self._sd, videoglamm_param_shapes(self.lmm_cfg, self.sam_cfg))
        dev = str(self.device)
        self.lmm = LmmEngine(self._sd, self.lmm_cfg, dev, precise=self.precise == "all")
        self.sam = Sam2Engine(self._sd, self.sam_cfg, dev, prefix=SAM_PREFIX, precise=self.precise)
        self._sd = None                                     # the packed copies are the weights from here on

    # ---- HF-style construction -------------------------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path: str, low_cpu_mem_usage: bool = False, use_sam2_video_branch: bool = True, torch_dtype=None,
                        clip_path: str | None = None, iv2_path: str | None = None, precise: str = "heads", **unused):
        """Reads an exported VideoGLaMM HF directory (`config.json` + `pytorch_model*.bin` / `*.safetensors` [+ `generation_config.json`]; written by
        train_ds_with_videogptplus.py:320-343). The CLIP tower is not in that export (arch.py:53-57): it is read from `clip_path` or
        `config.image_mm_vision_tower` (a local HF directory, or a hub id resolved through the local HF cache); the InternVideo2 `.pt` from `iv2_path`
        or `config.mm_vision_tower`. Unresolvable tower paths raise here (or in `initialize_vision_modules`), naming the path.
        quantization_config / load_in_*bit are not supported (weights are bf16)."""
        if unused.get("load_in_4bit") or unused.get("load_in_8bit") or unused.get("quantization_config") is not None:
            raise NotImplementedError("bitsandbytes quantisation is not part of the B200 engine (weights are bf16)")
        cfg = json.load(open(os.path.join(path, "config.json")))
        sd: dict = {}
        for f in sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))):
            sd.update(torch.load(f, map_location="cpu"))
        for f in sorted(glob.glob(os.path.join(path, "*.safetensors"))):
            from safetensors.torch import load_file
            sd.update(load_file(f))
        if not sd:
            raise FileNotFoundError(f"no pytorch_model*.bin / *.safetensors under {path}")
        lmm_cfg = make_lmm_config(phi=dict(vocab=cfg.get("vocab_size", 32064), hidden=cfg.get("hidden_size", 3072), layers=cfg.get("num_hidden_layers", 32),
                                           heads=cfg.get("num_attention_heads", 32), inter=cfg.get("intermediate_size", 8192),
                                           rope_theta=cfg.get("rope_theta", 10000.0), eps=cfg.get("rms_norm_eps", 1e-5),
                                           sliding_window=cfg.get("sliding_window", 2047)),
                                  seg_token_idx=cfg.get("seg_token_idx", cfg.get("vocab_size", 32064) - 1), out_dim=cfg.get("out_dim", 256),
                                  **cfg.get("vg_arch_overrides", {}))
        sam_cfg = Sam2Config(**cfg.get("vg_sam2_overrides", {}))
        # EOS: HF generate() stops on generation_config.eos_token_id (Phi-3-mini ships [32000, 32001, 32007]); merge it with config.eos_token_id
        eos = cfg.get("eos_token_id", [])
        eos = list(eos) if isinstance(eos, (list, tuple)) else ([] if eos is None else [eos])
        gen_path = os.path.join(path, "generation_config.json")
        gen_eos: list = []
        if os.path.exists(gen_path):
            g = json.load(open(gen_path)).get("eos_token_id", [])
            gen_eos = list(g) if isinstance(g, (list, tuple)) else ([] if g is None else [g])
        model = cls(sd, lmm_cfg, sam_cfg, use_sam2_video_branch=use_sam2_video_branch, eos_token_ids=sorted(set(eos) | set(gen_eos)), precise=precise,
                    defer=True, hf_config=cfg)
        model._generation_eos = gen_eos
        model._tower_paths = dict(clip=clip_path or cfg.get("image_mm_vision_tower") or cfg.get("image_vision_tower"),
                                  iv2=iv2_path or cfg.get("mm_vision_tower") or cfg.get("vision_tower"))
        model._load_towers(strict=False)                      # chat.py calls initialize_vision_modules later; that call is strict
        return model

    def _load_towers(self, strict: bool):
        """Adds the CLIP / InternVideo2 weights to the host state dict when the export does not carry them."""
        if self._sd is None:
            return
        paths = getattr(self, "_tower_paths", {})
        if not any(k.startswith(CLIP_PREFIX) for k in self._sd):
            d = resolve_clip_dir(paths.get("clip"))
            if d is not None:
                self._sd.update(load_clip_vision_weights(d))
            elif strict:
                raise FileNotFoundError(f"CLIP vision tower '{paths.get('clip')}' is neither a local HF directory nor in the local HF cache "
                                        f"(config.image_mm_vision_tower); pass clip_path= to from_pretrained")
        if not any(k.startswith(IV2_PREFIX + "vision_encoder.blocks.") for k in self._sd):
            p = paths.get("iv2")
            if p and os.path.isfile(p):
                self._sd.update(load_iv2_weights(p))
            elif strict:
                raise FileNotFoundError(f"InternVideo2 checkpoint '{p}' not found (config.mm_vision_tower); pass iv2_path= to from_pretrained")

    # ---- module-protocol methods the reference's chat.py calls on the model (chat.py:290-352) -------------------------------------------------
    def eval(self):
        return self

    def cuda(self, device=None, *_, **__):
        if device is not None:
            self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._ensure_engines()
        return self

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (str, torch.device)) and torch.device(a).type == "cuda":
                return self.cuda(a)
        return self

    def bfloat16(self):
        return self

    float = half = bfloat16

    def get_model(self):
        return self

    def initialize_vision_modules(self, model_args=None, fsdp=None):
        """VideoGPTPlusMetaModel.initialize_vision_modules (arch.py:30-86): there it builds both towers + projectors and loads the tower weights
        named by `mm_vision_tower` / `image_mm_vision_tower`; here the towers are part of the engine, so this makes sure their weights are in the
        state dict (loading them from the same config paths) and fails loudly when they cannot be found."""
        if model_args is not None:
            for src, dst in (("image_mm_vision_tower", "clip"), ("mm_vision_tower", "iv2")):
                v = getattr(model_args, src, None) if not isinstance(model_args, dict) else model_args.get(src)
                if v and hasattr(self, "_tower_paths") and not self._tower_paths.get(dst):
                    self._tower_paths[dst] = v
        self._load_towers(strict=True)

    def get_vision_tower(self):
        return self._towers["iv2"]

    def get_image_vision_tower(self):
        return self._towers["clip"]

    def resize_token_embeddings(self, n: int):
        """HF PreTrainedModel.resize_token_embeddings as chat.py:300 uses it (after tokenizer.add_tokens("[SEG]")): rows are truncated or appended
        to `model.embed_tokens` and `lm_head`; appended rows are the mean of the existing ones (transformers' mean-resizing default). Exported
        checkpoints already carry the resized matrices, so this is normally a no-op."""
        ph = self.lmm_cfg.phi
        if n == ph.vocab:
            return self
        if self._sd is None:
            raise RuntimeError("resize_token_embeddings must be called before the engines are built (.cuda() / inference)")
        for k in ("model.embed_tokens.weight", "lm_head.weight"):
            w = self._sd[k]
            if n < w.shape[0]:
                w = w[:n].clone()
            else:
                w = torch.cat([w, w.float().mean(0, keepdim=True).to(w.dtype).expand(n - w.shape[0], -1)], 0)
            self._sd[k] = w
        ph.vocab = n
        self.config.vocab_size = n
        return self

    def _eos_ids(self) -> tuple:
        e = getattr(self.config, "eos_token_id", ())
        e = tuple(e) if isinstance(e, (list, tuple)) else (() if e is None else (int(e),))
        # chat.py overwrites config.eos_token_id with the tokenizer's single id; the ids of generation_config.json still stop generation (HF semantics)
        return tuple(sorted(set(e) | set(getattr(self, "_generation_eos", ()))))

    def forward(self, **kwargs):
        """VideoGLaMM.py:897-903: `past_key_values` in kwargs -> the LMM forward (HF generate's inner call; replaced here by LmmEngine.generate, so
        it is only offered as a full-sequence forward returning logits), else the training / validation `model_forward`."""
        if "past_key_values" in kwargs:
            return self.lmm_forward(**kwargs)
        return self.model_forward(**kwargs)

    __call__ = forward

    @torch.no_grad()
    def lmm_forward(self, input_ids=None, images=None, context_images=None, **_):
        """One no-cache forward of the LMM (VideoGPTPlusPhi3ForCausalLM.forward, language_model/phi3.py:42-115): logits fp32 [1, L, V]."""
        self._ensure_engines()
        from .. import ops
        lmm = self.lmm
        ctx = None if context_images is None else context_images[0]
        vis = lmm.visual_tokens_any(images[0].to(self.device, torch.float32), None if ctx is None else ctx.to(self.device, torch.float32))
        x = lmm.embed_prompt(input_ids.to(self.device), vis)
        lmm._alloc_cache(x.shape[0] + 1)
        xr = lmm._llm(x, 0, raw=True)
        h = ops.norm(xr, lmm.final_norm, eps=lmm.cfg.phi.eps, rms=True, out_dtype=torch.float32)     # output_hidden_states[-1]
        return SimpleNamespace(logits=lmm.logits_all(xr)[None], hidden_states=(h[None],))

    def model_forward(self, **kwargs):
        from ..model_forward import model_forward
        return model_forward(self, **kwargs)

    # ---- inference -------------------------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, images, context_images, images_for_sam, input_ids, resize_list, original_size_list, max_new_tokens: int = 32,
                  use_sam2_video_branch: bool = False, forced_tokens: dict | None = None, return_logits: bool = False):
        """Same contract as the reference (VideoGLaMM.py:560-596): lists of length 1 (batch 1 is asserted by the callers,
        chat.py:570); returns (output_ids int64 [1, L_txt+N], video_segments_batch: [ {frame_idx: {obj_id: bool ndarray [H,W]}} ]).
        `context_images` None (or [None]) is the still-image input of chat.py:458-487: one CLIP-preprocessed image, 576 un-pooled tokens."""
        if not self.use_sam2 and use_sam2_video_branch:
            raise ValueError("Cannot use SAM2 video branch without SAM2")
        assert len(images) == 1 and len(images_for_sam) == 1 and input_ids.shape[0] == 1
        self._ensure_engines()
        self.lmm.cfg.seg_token_idx = int(self.config.seg_token_idx)
        dev = self.device
        ctx = None if (context_images is None or context_images[0] is None) else context_images[0]
        # host-resident SAM2 frames (100 MB per 8-frame clip at 1024 px): start the transfer now on a side stream, it completes under the LMM stage
        sam_in, copy_done = images_for_sam[0], None
        raw_u8 = sam_in.dtype == torch.uint8          # raw frames [T,H,W,3]: pre-processed on the GPU (H*W*3 bytes per frame over PCIe instead of 12.6 MB)
        if raw_u8:
            assert sam_in.dim() == 4 and sam_in.shape[3] == 3, "uint8 images_for_sam must be raw frames [T,H,W,3]"
        if not sam_in.is_cuda:
            if not hasattr(self, "_copy_stream"):
                self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._copy_stream):
                sam_in = sam_in.to(dev, non_blocking=True)
                copy_done = torch.cuda.Event()
                copy_done.record(self._copy_stream)
        enc_images, enc_ctx = self.encoder_inputs(images[0], ctx)
        out_ids, emb = self.lmm.generate(input_ids.to(dev), enc_images, enc_ctx, max_new_tokens, eos_ids=self._eos_ids(), forced=forced_tokens)
        if emb.shape[0] == 0:
            return (out_ids, [{}], None) if return_logits else (out_ids, [{}])       # VideoGLaMM.py:840-842
        H, W = original_size_list[0]
        run = self.sam.run_video if use_sam2_video_branch else self.sam.run_framewise      # VideoGLaMM.py:576-596 dispatch
        if copy_done is not None:
            torch.cuda.current_stream(dev).wait_event(copy_done)
            sam_in.record_stream(torch.cuda.current_stream(dev))
        if raw_u8:
            sam_in = self.sam_preprocessor_gpu().preprocess_video(sam_in)[0]
        if return_logits:   # parity tests: the fp32 video-resolution logits the masks are thresholded from (VideoGLaMM.py:870-875), as a third value
            logits = torch.stack(run(sam_in.to(dev), emb[:, None, :], (int(H), int(W)), binary=False)).cpu()
            stacked = (logits > 0).numpy()
            segs = {t: {o: stacked[t, o] for o in range(emb.shape[0])} for t in range(stacked.shape[0])}
            return out_ids, [segs], logits
        frames = run(sam_in.to(dev), emb[:, None, :], (int(H), int(W)), binary="packed")
        segs, self.last_packed = self.masks_to_host(frames, int(H), int(W))
        return out_ids, [segs]

    def encoder_inputs(self, images: torch.Tensor, ctx: torch.Tensor | None):
        """(IV2 frames, CLIP context images) on the device as fp32. `images` may be the RAW uint8 frames [T,H,W,3] (with `ctx` the same tensor or None):
        they are then pre-processed on the GPU (preprocess.EncPreprocessGPU: cv2-bilinear 224 px + PIL-bicubic / centre-crop 336 px, the
        reference's EncPreprocessor_VideoGPTPlus arithmetic) — H*W*3 bytes per frame over PCIe instead of 1.9 MB of fp32."""
        dev = self.device
        if images.dtype == torch.uint8:
            assert images.dim() == 4 and images.shape[3] == 3 and (ctx is None or ctx.dtype == torch.uint8), "uint8 images must be raw frames [T,H,W,3]"
            if not hasattr(self, "_enc_pre_gpu"):
                from ..preprocess import EncPreprocessGPU
                self._enc_pre_gpu = EncPreprocessGPU(dev, self.lmm.cfg.iv2.image_size, self.lmm.cfg.clip.image_size)
            return self._enc_pre_gpu.preprocess_video(images)
        return images.to(dev, torch.float32, non_blocking=True), None if ctx is None else ctx.to(dev, torch.float32, non_blocking=True)

    def sam_preprocessor_gpu(self):
        if not hasattr(self, "_sam_pre_gpu"):
            from ..preprocess import SAM_v2_PreprocessGPU
            self._sam_pre_gpu = SAM_v2_PreprocessGPU(self.device, self.sam_cfg.image_size if self.sam_cfg.image_size == 1024 else 1024)
        return self._sam_pre_gpu

    def masks_to_host(self, frames: list, H: int, W: int):
        """Per-frame bit-packed device masks [n_obj, ceil(H*W/8)] -> ({frame: {obj: bool ndarray [H,W]}}, PackedMasks still on the device).
        ONE device->host copy per clip, of H*W/8 bytes per mask, into pinned memory; the bool arrays are expanded on the host."""
        from ..parallel import PackedMasks
        packed = torch.stack(frames)                                                  # [T, n_obj, bpp] uint8
        key = tuple(packed.shape)
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        host = self._pinned.get(key)
        if host is None:
            if len(self._pinned) > 16:
                self._pinned.clear()
            host = self._pinned[key] = torch.empty(key, dtype=torch.uint8, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        T, n_obj = key[:2]
        bits = np.unpackbits(host.numpy(), axis=-1, count=H * W).reshape(T, n_obj, H, W).view(np.bool_)
        return {t: {o: bits[t, o] for o in range(n_obj)} for t in range(T)}, PackedMasks(packed, (T, n_obj, H, W))
