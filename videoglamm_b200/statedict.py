"""The state_dict contract of the reference checkpoints: key -> shape tables for the SAM2 side
(`model.visual_model.*` in a VideoGLaMM checkpoint) and the LMM side, used by the loader to validate checkpoints and by
bench.py / tests to build synthetic weights of the named architecture (there are no checkpoints on the box)."""
from __future__ import annotations

from collections import OrderedDict

import torch


def _attn(shapes, p, dim, internal, kv_in=None):
    kv_in = kv_in or dim
    for n, (o, i) in dict(q_proj=(internal, dim), k_proj=(internal, kv_in), v_proj=(internal, kv_in), out_proj=(dim, internal)).items():
        shapes[f"{p}.{n}.weight"] = (o, i)
        shapes[f"{p}.{n}.bias"] = (o,)


def _lin(shapes, p, o, i, bias=True):
    shapes[p + ".weight"] = (o, i)
    if bias:
        shapes[p + ".bias"] = (o,)


def _norm(shapes, p, c):
    shapes[p + ".weight"] = (c,)
    shapes[p + ".bias"] = (c,)


def _mlp(shapes, p, dims):
    for j, (i, o) in enumerate(zip(dims[:-1], dims[1:])):
        _lin(shapes, f"{p}.layers.{j}", o, i)


def sam2_param_shapes(cfg) -> "OrderedDict[str, tuple]":
    """Key -> shape for SAM2VideoPredictor(sam2_hiera_l.yaml) — modeling/sam2_base.py:__init__, _build_sam_heads."""
    s: OrderedDict[str, tuple] = OrderedDict()
    C, M = cfg.d_model, cfg.mem_dim
    s["maskmem_tpos_enc"] = (cfg.num_maskmem, 1, 1, M)
    s["no_mem_embed"] = (1, 1, C)
    s["no_mem_pos_enc"] = (1, 1, C)
    s["no_obj_ptr"] = (1, C)
    t = "image_encoder.trunk"
    s[t + ".pos_embed"] = (1, cfg.embed_dim, *cfg.window_pos_embed_bkg_spatial_size)
    s[t + ".pos_embed_window"] = (1, cfg.embed_dim, cfg.window_spec[0], cfg.window_spec[0])
    s[t + ".patch_embed.proj.weight"] = (cfg.embed_dim, 3, 7, 7)
    s[t + ".patch_embed.proj.bias"] = (cfg.embed_dim,)
    for i, (dim, dim_out, heads, window, qs) in enumerate(cfg.block_specs()):
        b = f"{t}.blocks.{i}"
        _norm(s, b + ".norm1", dim)
        _lin(s, b + ".attn.qkv", 3 * dim_out, dim)
        _lin(s, b + ".attn.proj", dim_out, dim_out)
        _norm(s, b + ".norm2", dim_out)
        _mlp(s, b + ".mlp", [dim_out, 4 * dim_out, dim_out])
        if dim != dim_out:
            _lin(s, b + ".proj", dim_out, dim)
    for j, ch in enumerate(cfg.channel_list):
        s[f"image_encoder.neck.convs.{j}.conv.weight"] = (C, ch, 1, 1)
        s[f"image_encoder.neck.convs.{j}.conv.bias"] = (C,)
    s["mask_downsample.weight"] = (1, 1, 4, 4)
    s["mask_downsample.bias"] = (1,)
    for i in range(cfg.mem_attn_layers):
        l = f"memory_attention.layers.{i}"
        _attn(s, l + ".self_attn", C, C)
        _attn(s, l + ".cross_attn_image", C, C, kv_in=M)
        _lin(s, l + ".linear1", cfg.mem_attn_ffn, C)
        _lin(s, l + ".linear2", C, cfg.mem_attn_ffn)
        for n in ("norm1", "norm2", "norm3"):
            _norm(s, f"{l}.{n}", C)
    _norm(s, "memory_attention.norm", C)
    me = "memory_encoder"
    ch = 1
    for j in range(4):
        s[f"{me}.mask_downsampler.encoder.{3 * j}.weight"] = (ch * 4, ch, 3, 3)
        s[f"{me}.mask_downsampler.encoder.{3 * j}.bias"] = (ch * 4,)
        _norm(s, f"{me}.mask_downsampler.encoder.{3 * j + 1}", ch * 4)
        ch *= 4
    s[f"{me}.mask_downsampler.encoder.12.weight"] = (C, ch, 1, 1)
    s[f"{me}.mask_downsampler.encoder.12.bias"] = (C,)
    s[f"{me}.pix_feat_proj.weight"] = (C, C, 1, 1)
    s[f"{me}.pix_feat_proj.bias"] = (C,)
    for j in range(2):
        l = f"{me}.fuser.layers.{j}"
        s[l + ".weight"] = (C,)
        s[l + ".dwconv.weight"] = (C, 1, 7, 7)
        s[l + ".dwconv.bias"] = (C,)
        _norm(s, l + ".norm", C)
        _lin(s, l + ".pwconv1", 4 * C, C)
        _lin(s, l + ".pwconv2", C, 4 * C)
    s[f"{me}.out_proj.weight"] = (M, C, 1, 1)
    s[f"{me}.out_proj.bias"] = (M,)
    pe = "sam_prompt_encoder"
    s[pe + ".pe_layer.positional_encoding_gaussian_matrix"] = (2, C // 2)
    for j in range(4):
        s[f"{pe}.point_embeddings.{j}.weight"] = (1, C)
    s[pe + ".not_a_point_embed.weight"] = (1, C)
    s[pe + ".mask_downscaling.0.weight"] = (4, 1, 2, 2)
    s[pe + ".mask_downscaling.0.bias"] = (4,)
    _norm(s, pe + ".mask_downscaling.1", 4)
    s[pe + ".mask_downscaling.3.weight"] = (16, 4, 2, 2)
    s[pe + ".mask_downscaling.3.bias"] = (16,)
    _norm(s, pe + ".mask_downscaling.4", 16)
    s[pe + ".mask_downscaling.6.weight"] = (C, 16, 1, 1)
    s[pe + ".mask_downscaling.6.bias"] = (C,)
    s[pe + ".no_mask_embed.weight"] = (1, C)
    md = "sam_mask_decoder"
    for i in range(cfg.dec_depth):
        l = f"{md}.transformer.layers.{i}"
        _attn(s, l + ".self_attn", C, C)
        _norm(s, l + ".norm1", C)
        _attn(s, l + ".cross_attn_token_to_image", C, C // 2)
        _norm(s, l + ".norm2", C)
        _mlp(s, l + ".mlp", [C, cfg.dec_mlp, C])
        _norm(s, l + ".norm3", C)
        _norm(s, l + ".norm4", C)
        _attn(s, l + ".cross_attn_image_to_token", C, C // 2)
    _attn(s, md + ".transformer.final_attn_token_to_image", C, C // 2)
    _norm(s, md + ".transformer.norm_final_attn", C)
    s[md + ".iou_token.weight"] = (1, C)
    s[md + ".mask_tokens.weight"] = (4, C)
    s[md + ".obj_score_token.weight"] = (1, C)
    s[md + ".output_upscaling.0.weight"] = (C, C // 4, 2, 2)
    s[md + ".output_upscaling.0.bias"] = (C // 4,)
    _norm(s, md + ".output_upscaling.1", C // 4)
    s[md + ".output_upscaling.3.weight"] = (C // 4, C // 8, 2, 2)
    s[md + ".output_upscaling.3.bias"] = (C // 8,)
    s[md + ".conv_s0.weight"] = (C // 8, C, 1, 1)
    s[md + ".conv_s0.bias"] = (C // 8,)
    s[md + ".conv_s1.weight"] = (C // 4, C, 1, 1)
    s[md + ".conv_s1.bias"] = (C // 4,)
    for j in range(4):
        _mlp(s, f"{md}.output_hypernetworks_mlps.{j}", [C, C, C, C // 8])
    _mlp(s, md + ".iou_prediction_head", [C, 256, 256, 4])
    _mlp(s, md + ".pred_obj_score_head", [C, C, C, 1])
    _mlp(s, "obj_ptr_proj", [C, C, C, C])
    return s



def lmm_param_shapes(cfg) -> "OrderedDict[str, tuple]":
    """Key -> shape for the LMM side of VideoGLaMMForCausalLM: Phi-3 (HF names), projectors,
    text_hidden_fcs, InternVideo2 vision encoder (keys used at inference) and the HF CLIP vision tower."""
    s: OrderedDict[str, tuple] = OrderedDict()
    ph, cl, iv = cfg.phi, cfg.clip, cfg.iv2
    H = ph.hidden
    s["model.embed_tokens.weight"] = (ph.vocab, H)
    for l in range(ph.layers):
        p = f"model.layers.{l}"
        s[p + ".self_attn.qkv_proj.weight"] = (3 * H, H)
        s[p + ".self_attn.o_proj.weight"] = (H, H)
        s[p + ".mlp.gate_up_proj.weight"] = (2 * ph.inter, H)
        s[p + ".mlp.down_proj.weight"] = (H, ph.inter)
        s[p + ".input_layernorm.weight"] = (H,)
        s[p + ".post_attention_layernorm.weight"] = (H,)
    s["model.norm.weight"] = (H,)
    s["lm_head.weight"] = (ph.vocab, H)
    for name, din in (("model.mm_projector", iv.width), ("model.image_mm_projector", cl.width)):
        _lin(s, name + ".0", H, din)
        for j in range(1, cfg.proj_depth):
            _lin(s, f"{name}.{2 * j}", H, H)
    _lin(s, "model.text_hidden_fcs.0.0", H, H)
    _lin(s, "model.text_hidden_fcs.0.2", cfg.out_dim, H)
    e = "model.vision_tower.vision_encoder"
    g2 = iv.grid ** 2
    s[e + ".cls_token"] = (1, 1, iv.width)
    s[e + ".pos_embed"] = (1, 1 + iv.frames * g2, iv.width)
    s[e + ".img_pos_embed"] = (1, 1 + g2, iv.width)
    s[e + ".patch_embed.proj.weight"] = (iv.width, 3, 1, iv.patch, iv.patch)
    s[e + ".patch_embed.proj.bias"] = (iv.width,)
    for i in range(iv.depth):
        b = f"{e}.blocks.{i}"
        s[b + ".norm1.weight"] = (iv.width,)
        s[b + ".attn.qkv.weight"] = (3 * iv.width, iv.width)
        _lin(s, b + ".attn.proj", iv.width, iv.width)
        s[b + ".attn.q_norm.weight"] = (iv.width,)
        s[b + ".attn.k_norm.weight"] = (iv.width,)
        s[b + ".ls1.gamma"] = (iv.width,)
        s[b + ".norm2.weight"] = (iv.width,)
        _lin(s, b + ".mlp.fc1", iv.mlp, iv.width)
        _lin(s, b + ".mlp.fc2", iv.width, iv.mlp)
        s[b + ".ls2.gamma"] = (iv.width,)
    v = "model.image_vision_tower.vision_tower.vision_model"
    s[v + ".embeddings.class_embedding"] = (cl.width,)
    s[v + ".embeddings.patch_embedding.weight"] = (cl.width, 3, cl.patch, cl.patch)
    s[v + ".embeddings.position_embedding.weight"] = (1 + cl.grid ** 2, cl.width)
    _norm(s, v + ".pre_layrnorm", cl.width)
    for i in range(cl.layers):
        l = f"{v}.encoder.layers.{i}"
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            _lin(s, f"{l}.self_attn.{n}", cl.width, cl.width)
        _norm(s, l + ".layer_norm1", cl.width)
        _lin(s, l + ".mlp.fc1", cl.mlp, cl.width)
        _lin(s, l + ".mlp.fc2", cl.width, cl.mlp)
        _norm(s, l + ".layer_norm2", cl.width)
    _norm(s, v + ".post_layernorm", cl.width)
    return s




def videoglamm_param_shapes(lmm_cfg, sam_cfg) -> "OrderedDict[str, tuple]":
    s = lmm_param_shapes(lmm_cfg)
    for k, v in sam2_param_shapes(sam_cfg).items():
        s["model.visual_model." + k] = v
    return s


def validate_state_dict(sd: dict, shapes: "OrderedDict[str, tuple]") -> None:
    """Raise with a precise message if a checkpoint lacks a key the engine needs or a shape differs."""
    missing = [k for k in shapes if k not in sd]
    if missing:
        raise KeyError(f"state_dict is missing {len(missing)} keys, e.g. {missing[:5]}")
    bad = [(k, tuple(sd[k].shape), tuple(shapes[k])) for k in shapes if tuple(sd[k].shape) != tuple(shapes[k])]
    if bad:
        raise ValueError(f"state_dict shape mismatches, e.g. {bad[:3]}")


def random_state_dict(shapes: "OrderedDict[str, tuple]", device: str = "cuda", seed: int = 1234, dtype=torch.bfloat16) -> dict:
    """Fast on-device random weights of the given architecture for throughput runs (values are irrelevant to timing):
    matrices ~ U(+-sqrt(3/fan_in)), vectors: norm scales ~ 1, everything else small."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if len(shp) >= 2 and k.endswith("weight"):
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = (torch.rand(shp, device=device, generator=g, dtype=torch.float32) * 2 - 1) * (3.0 / fan_in) ** 0.5
            out[k] = t.to(dtype)
        elif len(shp) >= 2:
            out[k] = (0.02 * torch.randn(shp, device=device, generator=g, dtype=torch.float32)).to(dtype)
        elif "norm" in k and k.endswith("weight"):
            out[k] = torch.ones(shp, device=device, dtype=torch.float32)
        elif k.endswith("gamma") or (k.endswith(".weight") and len(shp) == 1):
            out[k] = torch.full(shp, 0.1, device=device, dtype=torch.float32)
        else:
            out[k] = 0.02 * torch.randn(shp, device=device, generator=g, dtype=torch.float32)
    return out
