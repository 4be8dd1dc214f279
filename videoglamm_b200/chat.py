"""Entry points the reference's eval scripts import from `chat.py` (`eval_gcg_infer.py:18`, `eval_mevis.py`, `eval_referdavis_infer.py`, ...):
`initialize_model_videogptplus`, `preprocess_vision`, `IMAGE_TOKEN_INDEX` — same names, argument order and return tuple
(/root/reference/VideoGLaMM/chat.py:226-367 and :402-456), backed by the B200 engine. A maintainer switches an eval script over with

    import videoglamm_b200.chat as chat          # instead of `import chat` / `from chat import ...`

Differences: weights stay bf16 (the bitsandbytes 4/8-bit paths and the DeepSpeed fp16 wrapper of the reference are not part of this engine:
`load_in_8bit` / `load_in_4bit` raise), `precision` only selects the dtype of the PREPROCESSED tensors handed to `inference` (they are
converted to fp32 / bf16 on the device there)."""
from __future__ import annotations

import os

from .preprocess import (IMAGE_TOKEN_INDEX, ConvGenerator_VideoGPTPlus, EncPreprocessor_VideoGPTPlus, SAM_v2_Preprocess,  # noqa: F401
                         preprocess_vision)

__all__ = ["IMAGE_TOKEN_INDEX", "initialize_model_videogptplus", "preprocess_vision"]


def initialize_model_videogptplus(model_base, precision, local_rank, load_in_8bit, load_in_4bit, use_sam2_video_branch, base_llm_type):
    """chat.py:226-367 -> (model, tokenizer, enc_preprocessor, conv_generator, sam_preprocessor).
    `model_base`: exported VideoGLaMM HF directory (config.json + weights + tokenizer files). CLIP / InternVideo2 weights are looked up
    next to it the way the reference resolves them: `config.image_mm_vision_tower` (an HF CLIP directory or hub name available offline) and
    `config.mm_vision_tower` (the InternVideo2 `.pt`)."""
    import json

    import torch
    from transformers import AutoTokenizer

    from .model import VideoGLaMMForCausalLM

    if base_llm_type != "phi3":
        raise ValueError("Invalid base_llm_type")                     # chat.py:266-267
    if load_in_8bit or load_in_4bit:
        raise NotImplementedError("bitsandbytes quantisation is not part of the B200 engine (weights are bf16)")
    torch.cuda.set_device(int(local_rank) if local_rank is not None else 0)
    # the CLIP / InternVideo2 towers are resolved from config.image_mm_vision_tower (local HF directory or a hub id in the local HF cache) and
    # config.mm_vision_tower inside from_pretrained; an unresolvable path raises there / in initialize_vision_modules, naming the path
    model = VideoGLaMMForCausalLM.from_pretrained(model_base, low_cpu_mem_usage=False, use_sam2_video_branch=use_sam2_video_branch,
                                                  torch_dtype={"bf16": torch.bfloat16, "fp16": torch.float16}.get(precision, torch.float32))
    tokenizer = AutoTokenizer.from_pretrained(model_base, use_fast=False)
    tokenizer.pad_token = tokenizer.unk_token
    tokenizer.add_tokens("[SEG]")
    seg_token_idx = tokenizer.convert_tokens_to_ids("[SEG]")
    model.resize_token_embeddings(len(tokenizer))                      # chat.py:300 (a no-op for exported checkpoints, which carry the resized matrices)
    model.config.seg_token_idx = seg_token_idx                         # chat.py:303
    model.config.eos_token_id = tokenizer.eos_token_id          # chat.py:305; the ids of generation_config.json keep stopping generation (_eos_ids)
    model.config.bos_token_id = tokenizer.bos_token_id
    model.config.pad_token_id = tokenizer.pad_token_id
    if tokenizer.pad_token_id is None:                                 # chat.py:310-313 (llama-3.1 tokenizers)
        tokenizer.pad_token = tokenizer.eos_token
        tokenizer.pad_token_id = tokenizer.eos_token_id
    model.get_model().initialize_vision_modules(model.get_model().config)   # chat.py:315 (strict: the tower weights must be resolvable)
    model = model.bfloat16().cuda()
    model.eval()
    num_frames = int(os.environ.get("NUM_FRAMES", 16))                 # model/videogpt_plus/constants.py
    enc_preprocessor = EncPreprocessor_VideoGPTPlus(num_frames=num_frames)
    conv_generator = ConvGenerator_VideoGPTPlus(num_frames=num_frames)
    sam_preprocessor = SAM_v2_Preprocess()
    return model, tokenizer, enc_preprocessor, conv_generator, sam_preprocessor
