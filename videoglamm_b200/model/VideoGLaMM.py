"""Drop-in surface of the reference's top model for inference: `VideoGLaMMForCausalLM.from_pretrained(...)` and
`.inference(images, context_images, images_for_sam, input_ids, resize_list, original_size_list, max_new_tokens,
use_sam2_video_branch)` with the same argument meaning, return value and error behaviour as
/root/reference/VideoGLaMM/model/VideoGLaMM.py:560-596 (dispatch) and :770-879 (video branch). All arithmetic runs in
libvglamm kernels through LmmEngine / Sam2Engine; there is no CPU path.
"""
from __future__ import annotations

import glob
import json
import os
from types import SimpleNamespace

import numpy as np
import torch

from ..config import LmmConfig, Sam2Config, make_lmm_config
from ..lmm_engine import LmmEngine
from ..sam2_engine import Sam2Engine
from ..statedict import validate_state_dict, videoglamm_param_shapes

SAM_PREFIX = "model.visual_model."


class VideoGLaMMForCausalLM:
    """Inference engine behind the reference's class name. Not an nn.Module: weights are packed once for the kernels."""

    def __init__(self, state_dict: dict, lmm_cfg: LmmConfig | None = None, sam_cfg: Sam2Config | None = None, *, use_sam2_video_branch: bool = True,
                 eos_token_ids=(), device: str = "cuda", validate: bool = True, precise: str = "heads"):
        """precise: "heads" (default: SAM2 mask decoder in fp32-grade arithmetic, everything else bf16 operands), "none", or "all" (verification
        mode: every GEMM split-bf16, every attention exact fp32 — see Sam2Engine / LmmEngine)."""
        self.lmm_cfg = lmm_cfg or make_lmm_config()
        self.sam_cfg = sam_cfg or Sam2Config()
        if validate:
            validate_state_dict(state_dict, videoglamm_param_shapes(self.lmm_cfg, self.sam_cfg))
        self.device = torch.device(device)
        self.lmm = LmmEngine(state_dict, self.lmm_cfg, device, precise=precise == "all")
        self.sam = Sam2Engine(state_dict, self.sam_cfg, device, prefix=SAM_PREFIX, precise=precise)
        self.use_sam2 = True
        self.use_sam2_video_branch = use_sam2_video_branch
        self.config = SimpleNamespace(seg_token_idx=self.lmm_cfg.seg_token_idx, use_sam2=True, eos_token_id=list(eos_token_ids),
                                      out_dim=self.lmm_cfg.out_dim, mm_use_im_start_end=False, mm_use_im_patch_token=False)

    # ---- HF-style construction -------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path: str, low_cpu_mem_usage: bool = False, use_sam2_video_branch: bool = True, torch_dtype=None,
                        clip_path: str | None = None, iv2_path: str | None = None, **_):
        """Reads an exported VideoGLaMM HF directory (`config.json` + `pytorch_model*.bin` / `*.safetensors`; written by
        train_ds_with_videogptplus.py:320-343). CLIP keys are not in that export (arch.py:53-57): pass `clip_path` (an HF
        CLIPVisionModel dir); the InternVideo2 `.pt` is read from `iv2_path` or config.mm_vision_tower."""
        cfg = json.load(open(os.path.join(path, "config.json")))
        sd: dict = {}
        for f in sorted(glob.glob(os.path.join(path, "pytorch_model*.bin"))):
            sd.update(torch.load(f, map_location="cpu"))
        for f in sorted(glob.glob(os.path.join(path, "*.safetensors"))):
            from safetensors.torch import load_file
            sd.update(load_file(f))
        if clip_path:
            from safetensors.torch import load_file
            files = glob.glob(os.path.join(clip_path, "*.safetensors"))
            csd = load_file(files[0]) if files else torch.load(os.path.join(clip_path, "pytorch_model.bin"), map_location="cpu")
            sd.update({"model.image_vision_tower.vision_tower." + k: v for k, v in csd.items() if k.startswith("vision_model.")})
        iv2 = iv2_path or cfg.get("mm_vision_tower")
        if iv2 and os.path.isfile(iv2):
            isd = torch.load(iv2, map_location="cpu")
            isd = isd.get("model", isd.get("module", isd))
            sd.update({"model.vision_tower." + k: v for k, v in isd.items() if k.startswith("vision_encoder.")})
        lmm_cfg = make_lmm_config(phi=dict(vocab=cfg.get("vocab_size", 32064), hidden=cfg.get("hidden_size", 3072), layers=cfg.get("num_hidden_layers", 32),
                                           heads=cfg.get("num_attention_heads", 32), inter=cfg.get("intermediate_size", 8192),
                                           rope_theta=cfg.get("rope_theta", 10000.0), eps=cfg.get("rms_norm_eps", 1e-5),
                                           sliding_window=cfg.get("sliding_window", 2047)),
                                  seg_token_idx=cfg["seg_token_idx"], out_dim=cfg.get("out_dim", 256))
        eos = cfg.get("eos_token_id", [])
        return cls(sd, lmm_cfg, Sam2Config(), use_sam2_video_branch=use_sam2_video_branch, eos_token_ids=eos if isinstance(eos, list) else [eos])

    # no-op module-protocol methods the reference's chat.py calls on the model (chat.py:290-352)
    def eval(self):
        return self

    def cuda(self, *_, **__):
        return self

    def bfloat16(self):
        return self

    def float(self):
        return self

    def get_model(self):
        return self

    def resize_token_embeddings(self, n: int):
        """chat.py:300 calls this after adding `[SEG]`; exported checkpoints already carry the resized matrices, so only the size is checked."""
        if n > self.lmm_cfg.phi.vocab:
            raise ValueError(f"tokenizer has {n} tokens, the packed lm_head/embeddings have {self.lmm_cfg.phi.vocab}")
        return self

    def _eos_ids(self) -> tuple:
        e = getattr(self.config, "eos_token_id", ())
        return tuple(e) if isinstance(e, (list, tuple)) else (() if e is None else (int(e),))

    # ---- inference -------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def inference(self, images, context_images, images_for_sam, input_ids, resize_list, original_size_list, max_new_tokens: int = 32,
                  use_sam2_video_branch: bool = False, forced_tokens: dict | None = None, return_logits: bool = False):
        """Same contract as the reference (VideoGLaMM.py:560-596): lists of length 1 (batch 1 is asserted by the callers,
        chat.py:570); returns (output_ids int64 [1, L_txt+N], video_segments_batch: [ {frame_idx: {obj_id: bool ndarray [H,W]}} ])."""
        if not self.use_sam2 and use_sam2_video_branch:
            raise ValueError("Cannot use SAM2 video branch without SAM2")
        assert len(images) == 1 and len(images_for_sam) == 1 and input_ids.shape[0] == 1
        if context_images is None or context_images[0] is None:
            raise NotImplementedError("still-image inputs (context_images=None, arch.py encode_images) are not on the video path")
        dev = self.device
        # host-resident SAM2 frames (100 MB per 8-frame clip at 1024 px): start the transfer now on a side stream, it completes under the LMM stage
        sam_in, copy_done = images_for_sam[0], None
        if not sam_in.is_cuda:
            if not hasattr(self, "_copy_stream"):
                self._copy_stream = torch.cuda.Stream(device=dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._copy_stream):
                sam_in = sam_in.to(dev, non_blocking=True)
                copy_done = torch.cuda.Event()
                copy_done.record(self._copy_stream)
        out_ids, emb = self.lmm.generate(input_ids.to(dev), images[0].to(dev, torch.float32), context_images[0].to(dev, torch.float32),
                                         max_new_tokens, eos_ids=self._eos_ids(), forced=forced_tokens)
        if emb.shape[0] == 0:
            return out_ids, [{}]                                                     # VideoGLaMM.py:840-842
        H, W = original_size_list[0]
        run = self.sam.run_video if use_sam2_video_branch else self.sam.run_framewise      # VideoGLaMM.py:576-596 dispatch
        if copy_done is not None:
            torch.cuda.current_stream(dev).wait_event(copy_done)
            sam_in.record_stream(torch.cuda.current_stream(dev))
        if return_logits:   # parity tests: the fp32 video-resolution logits the masks are thresholded from (VideoGLaMM.py:870-875), as a third value
            logits = torch.stack(run(sam_in.to(dev), emb[:, None, :], (int(H), int(W)), binary=False)).cpu()
            stacked = (logits > 0).numpy()
            segs = {t: {o: stacked[t, o] for o in range(emb.shape[0])} for t in range(stacked.shape[0])}
            return out_ids, [segs], logits
        frames = run(sam_in.to(dev), emb[:, None, :], (int(H), int(W)), binary="packed")
        segs, self.last_packed = self.masks_to_host(frames, int(H), int(W))
        return out_ids, [segs]

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
