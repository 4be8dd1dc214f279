"""TEST INFRASTRUCTURE (oracle) — CPU fp32 restatement of the reference's SAM2 path, functional over a flat state_dict.

Not product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may import it.
Parity pinning: the reference has no tests or golden vectors for this path, so this restatement is
pinned against outputs of the reference's own modules run here (tests/test_oracle_vs_reference.py, fixtures made by
tests/golden/make_golden.py). Every function cites the reference file:line it follows
(paths relative to /root/reference/VideoGLaMM/model/segment_anything_2/sam2/).

Tensors follow the reference layouts (NCHW feature maps, [B,N,C] token tensors).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

NO_OBJ_SCORE = -1024.0  # modeling/sam2_base.py:21


@dataclass
class Sam2Cfg:
    """sam2_configs/sam2_hiera_l.yaml (defaults) — shrinkable for fast tests."""
    image_size: int = 1024
    embed_dim: int = 144
    num_heads: int = 2
    stages: tuple = (2, 6, 36, 4)
    global_att_blocks: tuple = (23, 33, 43)
    window_pos_embed_bkg_spatial_size: tuple = (7, 7)
    window_spec: tuple = (8, 4, 16, 8)
    d_model: int = 256
    mem_dim: int = 64
    num_maskmem: int = 7
    max_obj_ptrs_in_encoder: int = 16
    mem_attn_layers: int = 4
    mem_attn_ffn: int = 2048
    dec_depth: int = 2
    dec_heads: int = 8
    dec_mlp: int = 2048
    sigmoid_scale_for_mem_enc: float = 20.0
    sigmoid_bias_for_mem_enc: float = -10.0
    q_pool: int = 3

    @property
    def stage_ends(self):
        return [sum(self.stages[:i]) - 1 for i in range(1, len(self.stages) + 1)]

    @property
    def q_pool_blocks(self):
        return [x + 1 for x in self.stage_ends[:-1]][: self.q_pool]

    def block_specs(self):
        """(dim, dim_out, heads, window, q_stride?) per block — backbones/hieradet.py:217-245."""
        specs, dim, heads, cur = [], self.embed_dim, self.num_heads, 1
        for i in range(sum(self.stages)):
            dim_out = dim
            window = self.window_spec[cur - 1]
            if i in self.global_att_blocks:
                window = 0
            if i - 1 in self.stage_ends:
                dim_out, heads, cur = dim * 2, heads * 2, cur + 1
            specs.append((dim, dim_out, heads, window, i in self.q_pool_blocks))
            dim = dim_out
        return specs

    @property
    def channel_list(self):
        specs = self.block_specs()
        return [specs[i][1] for i in self.stage_ends[::-1]]

    @property
    def feat_size(self):
        return self.image_size // 16


TINY = dict(image_size=256, embed_dim=32, num_heads=2, stages=(1, 2, 3, 1), global_att_blocks=(4,), mem_attn_layers=2)


def linear(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def layer_norm(sd, p, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def layer_norm2d(sd, p, x, eps=1e-6):
    """modeling/sam2_utils.py:137-149"""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    x = (x - u) / torch.sqrt(s + eps)
    return sd[p + ".weight"][:, None, None] * x + sd[p + ".bias"][:, None, None]


def mlp(sd, p, x, n, act=F.relu, sigmoid_output=False):
    """modeling/sam2_utils.py:108-132"""
    for i in range(n):
        x = linear(sd, f"{p}.layers.{i}", x)
        if i < n - 1:
            x = act(x)
    return torch.sigmoid(x) if sigmoid_output else x


# ------------------------------------------------------------------------------------------------ Hiera + FPN
def window_partition(x, ws):
    """backbones/utils.py:16-40"""
    B, H, W, C = x.shape
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    if ph or pw:
        x = F.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.view(B, Hp // ws, ws, Wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)
    return x, (Hp, Wp)


def window_unpartition(w, ws, pad_hw, hw):
    """backbones/utils.py:43-62"""
    Hp, Wp = pad_hw
    H, W = hw
    B = w.shape[0] // (Hp * Wp // ws // ws)
    x = w.view(B, Hp // ws, Wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :]


def _pool(x):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)


def hiera_block(sd, p, x, spec):
    """MultiScaleBlock.forward backbones/hieradet.py:136-168 + MultiScaleAttention.forward :58-83"""
    dim, dim_out, heads, window, q_stride = spec
    shortcut = x
    x = layer_norm(sd, p + ".norm1", x, 1e-6)
    if dim != dim_out:
        shortcut = linear(sd, p + ".proj", x)
        if q_stride:
            shortcut = _pool(shortcut)
    ws = window
    if ws > 0:
        H, W = x.shape[1:3]
        x, pad_hw = window_partition(x, ws)
    B, Hh, Ww, _ = x.shape
    qkv = linear(sd, p + ".attn.qkv", x).reshape(B, Hh * Ww, 3, heads, -1)
    q, k, v = qkv.unbind(2)
    if q_stride:
        q = _pool(q.reshape(B, Hh, Ww, -1))
        Hh, Ww = q.shape[1:3]
        q = q.reshape(B, Hh * Ww, heads, -1)
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    x = linear(sd, p + ".attn.proj", o.transpose(1, 2).reshape(B, Hh, Ww, -1))
    if q_stride:
        ws = window // 2
        H, W = shortcut.shape[1:3]
        pad_hw = (H + (ws - H % ws) % ws, W + (ws - W % ws) % ws)
    if window > 0:
        x = window_unpartition(x, ws, pad_hw, (H, W))
    x = shortcut + x
    return x + mlp(sd, p + ".mlp", layer_norm(sd, p + ".norm2", x, 1e-6), 2, act=F.gelu)


def hiera_pos_embed(sd, p, hw):
    """Hiera._get_pos_embed backbones/hieradet.py:269-277"""
    pe = F.interpolate(sd[p + ".pos_embed"], size=hw, mode="bicubic")
    win = sd[p + ".pos_embed_window"]
    pe = pe + win.tile([x // y for x, y in zip(pe.shape, win.shape)])
    return pe.permute(0, 2, 3, 1)


def hiera(sd, p, cfg: Sam2Cfg, img):
    """Hiera.forward backbones/hieradet.py:279-295; PatchEmbed backbones/utils.py:65-95"""
    x = F.conv2d(img, sd[p + ".patch_embed.proj.weight"], sd[p + ".patch_embed.proj.bias"], stride=4, padding=3).permute(0, 2, 3, 1)
    x = x + hiera_pos_embed(sd, p, x.shape[1:3])
    outs = []
    for i, spec in enumerate(cfg.block_specs()):
        x = hiera_block(sd, f"{p}.blocks.{i}", x, spec)
        if i in cfg.stage_ends:
            outs.append(x.permute(0, 3, 1, 2))
    return outs


def pos_embed_sine(shape_hw, num_pos_feats, temperature=10000.0):
    """PositionEmbeddingSine.forward modeling/position_encoding.py:79-112 (normalize=True, scale=2*pi) -> [C,H,W]"""
    H, W = shape_hw
    npf = num_pos_feats // 2
    y = torch.arange(1, H + 1, dtype=torch.float32).view(H, 1).repeat(1, W)
    x = torch.arange(1, W + 1, dtype=torch.float32).view(1, W).repeat(H, 1)
    eps, scale = 1e-6, 2 * math.pi
    y = y / (y[-1:, :] + eps) * scale
    x = x / (x[:, -1:] + eps) * scale
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / npf)
    px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).permute(2, 0, 1)


def forward_image(sd, cfg: Sam2Cfg, img):
    """SAM2Base.forward_image modeling/sam2_base.py:465-477; ImageEncoder.forward backbones/image_encoder.py:29-42;
    FpnNeck.forward :101-133 (nearest 2x top-down on levels 2,3; scalp=1). Returns (fpn [s0,s1,top], pos [..])."""
    xs = hiera(sd, "image_encoder.trunk", cfg, img)
    n = len(xs) - 1
    out, prev = [None] * len(xs), None
    for i in range(n, -1, -1):
        lat = F.conv2d(xs[i], sd[f"image_encoder.neck.convs.{n - i}.conv.weight"], sd[f"image_encoder.neck.convs.{n - i}.conv.bias"])
        if i in (2, 3) and prev is not None:
            prev = lat + F.interpolate(prev.float(), scale_factor=2.0, mode="nearest")
        else:
            prev = lat
        out[i] = prev
    feats = out[:-1]  # scalp = 1
    pos = [pos_embed_sine(f.shape[-2:], cfg.d_model)[None].expand(f.shape[0], -1, -1, -1) for f in feats]
    feats[0] = F.conv2d(feats[0], sd["sam_mask_decoder.conv_s0.weight"], sd["sam_mask_decoder.conv_s0.bias"])
    feats[1] = F.conv2d(feats[1], sd["sam_mask_decoder.conv_s1.weight"], sd["sam_mask_decoder.conv_s1.bias"])
    return feats, pos


# ------------------------------------------------------------------------------------------------ prompt encoder
def dense_pe(sd, cfg: Sam2Cfg):
    """PromptEncoder.get_dense_pe / PositionEmbeddingRandom.forward modeling/sam/prompt_encoder.py:68-77,192-229 -> [1,C,H,W]"""
    h = w = cfg.feat_size
    g = sd["sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    grid = torch.ones((h, w), dtype=torch.float32)
    y = (grid.cumsum(0) - 0.5) / h
    x = (grid.cumsum(1) - 0.5) / w
    c = (2 * torch.stack([x, y], dim=-1) - 1) @ g
    c = 2 * math.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1).permute(2, 0, 1)[None]


def prompt_encode(sd, cfg: Sam2Cfg, B, text_embeds=None):
    """PromptEncoder.forward modeling/sam/prompt_encoder.py:143-189 for the only prompt kinds on the VideoGLaMM path:
    one padding point (label -1, padded again -> two not-a-point tokens, :79-101) + optional text token (:179-180);
    dense = no_mask_embed (:185-187)."""
    nap = sd["sam_prompt_encoder.not_a_point_embed.weight"]  # [1,C]
    sparse = nap[None].expand(B, 2, -1)
    if text_embeds is not None:
        sparse = torch.cat([sparse, text_embeds], dim=1)
    dense = sd["sam_prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(B, -1, cfg.feat_size, cfg.feat_size)
    return sparse, dense


# ------------------------------------------------------------------------------------------------ mask decoder
def attention(sd, p, q, k, v, heads):
    """Attention.forward modeling/sam/transformer.py:233-260"""
    q, k, v = linear(sd, p + ".q_proj", q), linear(sd, p + ".k_proj", k), linear(sd, p + ".v_proj", v)
    sep = lambda t: t.reshape(t.shape[0], t.shape[1], heads, -1).transpose(1, 2)
    o = F.scaled_dot_product_attention(sep(q), sep(k), sep(v))
    o = o.transpose(1, 2).reshape(q.shape[0], q.shape[1], -1)
    return linear(sd, p + ".out_proj", o)


def two_way_transformer(sd, p, cfg: Sam2Cfg, src, pos, tokens):
    """TwoWayTransformer.forward modeling/sam/transformer.py:63-115; TwoWayAttentionBlock.forward :160-193"""
    keys = src.flatten(2).permute(0, 2, 1)
    key_pe = pos.flatten(2).permute(0, 2, 1)
    queries, query_pe, H = tokens, tokens, cfg.dec_heads
    for i in range(cfg.dec_depth):
        lp = f"{p}.layers.{i}"
        if i == 0:
            queries = attention(sd, lp + ".self_attn", queries, queries, queries, H)
        else:
            q = queries + query_pe
            queries = queries + attention(sd, lp + ".self_attn", q, q, queries, H)
        queries = layer_norm(sd, lp + ".norm1", queries)
        q, k = queries + query_pe, keys + key_pe
        queries = layer_norm(sd, lp + ".norm2", queries + attention(sd, lp + ".cross_attn_token_to_image", q, k, keys, H))
        queries = layer_norm(sd, lp + ".norm3", queries + mlp(sd, lp + ".mlp", queries, 2))
        q, k = queries + query_pe, keys + key_pe
        keys = layer_norm(sd, lp + ".norm4", keys + attention(sd, lp + ".cross_attn_image_to_token", k, q, queries, H))
    q, k = queries + query_pe, keys + key_pe
    queries = queries + attention(sd, p + ".final_attn_token_to_image", q, k, keys, H)
    return layer_norm(sd, p + ".norm_final_attn", queries), keys


def mask_decoder(sd, cfg: Sam2Cfg, image_embeddings, image_pe, sparse, dense, high_res_features, multimask_output, repeat_image=False):
    """MaskDecoder.forward/predict_masks modeling/sam/mask_decoder.py:110-245 with pred_obj_scores(+mlp),
    use_high_res_features, iou sigmoid, use_multimask_token_for_obj_ptr (sam2_hiera_l.yaml:98-116).
    Returns (masks [B,M,4h,4w], iou [B,M], sam_tokens [B,M|1,C], object_score_logits [B,1], all 4 masks, all 4 ious)."""
    p = "sam_mask_decoder"
    B = sparse.shape[0]
    out_tokens = torch.cat([sd[p + ".obj_score_token.weight"], sd[p + ".iou_token.weight"], sd[p + ".mask_tokens.weight"]], 0)
    tokens = torch.cat([out_tokens[None].expand(B, -1, -1), sparse], dim=1)
    src = torch.repeat_interleave(image_embeddings, B, dim=0) if repeat_image else image_embeddings
    src = src + dense
    pos_src = torch.repeat_interleave(image_pe, B, dim=0)
    b, c, h, w = src.shape
    hs, src = two_way_transformer(sd, p + ".transformer", cfg, src, pos_src, tokens)
    iou_token_out = hs[:, 1, :]
    mask_tokens_out = hs[:, 2:6, :]
    src = src.transpose(1, 2).reshape(b, c, h, w)
    feat_s0, feat_s1 = high_res_features
    up = F.conv_transpose2d(src, sd[p + ".output_upscaling.0.weight"], sd[p + ".output_upscaling.0.bias"], stride=2) + feat_s1
    up = F.gelu(layer_norm2d(sd, p + ".output_upscaling.1", up))
    up = F.gelu(F.conv_transpose2d(up, sd[p + ".output_upscaling.3.weight"], sd[p + ".output_upscaling.3.bias"], stride=2) + feat_s0)
    hyper_in = torch.stack([mlp(sd, f"{p}.output_hypernetworks_mlps.{i}", mask_tokens_out[:, i, :], 3) for i in range(4)], dim=1)
    b, c, h, w = up.shape
    masks = (hyper_in @ up.view(b, c, h * w)).view(b, -1, h, w)
    iou_pred = mlp(sd, p + ".iou_prediction_head", iou_token_out, 3, sigmoid_output=True)
    obj_logits = mlp(sd, p + ".pred_obj_score_head", hs[:, 0, :], 3)
    all_masks, all_iou = masks, iou_pred
    if multimask_output:
        masks, iou_pred, sam_tokens = masks[:, 1:], iou_pred[:, 1:], mask_tokens_out[:, 1:]
    else:
        masks, iou_pred = dynamic_multimask_via_stability(masks, iou_pred)
        sam_tokens = mask_tokens_out[:, 0:1]
    return masks, iou_pred, sam_tokens, obj_logits, all_masks, all_iou


def dynamic_multimask_via_stability(all_mask_logits, all_iou, delta=0.05, thresh=0.98):
    """MaskDecoder._dynamic_multimask_via_stability modeling/sam/mask_decoder.py:247-295 (build_sam.py:56-60 enables it)."""
    mm, mi = all_mask_logits[:, 1:], all_iou[:, 1:]
    best = torch.argmax(mi, dim=-1)
    bi = torch.arange(mi.size(0))
    best_m, best_i = mm[bi, best].unsqueeze(1), mi[bi, best].unsqueeze(1)
    sm, si = all_mask_logits[:, 0:1], all_iou[:, 0:1]
    flat = sm.flatten(-2)
    area_i = torch.sum(flat > delta, dim=-1).float()
    area_u = torch.sum(flat > -delta, dim=-1).float()
    stab = torch.where(area_u > 0, area_i / area_u, 1.0)
    ok = stab >= thresh
    return torch.where(ok[..., None, None].expand_as(sm), sm, best_m), torch.where(ok.expand_as(si), si, best_i)


def forward_sam_heads(sd, cfg: Sam2Cfg, backbone_features, high_res_features, text_inputs=None, multimask_output=True):
    """SAM2Base._forward_sam_heads modeling/sam2_base.py:251-411 (no point / mask prompts on this path)."""
    B = backbone_features.size(0)
    sparse, dense = prompt_encode(sd, cfg, B, text_inputs)
    low_multi, ious, sam_tokens, obj_logits, _, _ = mask_decoder(sd, cfg, backbone_features, dense_pe(sd, cfg), sparse, dense,
                                                                 high_res_features, multimask_output)
    is_obj = obj_logits > 0
    low_multi = torch.where(is_obj[:, None, None], low_multi, torch.tensor(NO_OBJ_SCORE)).float()
    high_multi = F.interpolate(low_multi, size=(cfg.image_size, cfg.image_size), mode="bilinear", align_corners=False)
    tok = sam_tokens[:, 0]
    if multimask_output:
        best = torch.argmax(ious, dim=-1)
        bi = torch.arange(B)
        low, high = low_multi[bi, best].unsqueeze(1), high_multi[bi, best].unsqueeze(1)
        if sam_tokens.size(1) > 1:
            tok = sam_tokens[bi, best]
    else:
        low, high = low_multi, high_multi
    obj_ptr = mlp(sd, "obj_ptr_proj", tok, 3)
    lam = is_obj.float()
    obj_ptr = lam * obj_ptr + (1 - lam) * sd["no_obj_ptr"]
    return dict(low_res_multimasks=low_multi, ious=ious, low_res_masks=low, high_res_masks=high, obj_ptr=obj_ptr,
                object_score_logits=obj_logits)


# ------------------------------------------------------------------------------------------------ memory encoder
def memory_encoder(sd, cfg: Sam2Cfg, pix_feat, mask_for_mem):
    """MemoryEncoder.forward modeling/memory_encoder.py:159-182 (skip_mask_sigmoid=True); MaskDownSampler :17-58
    (4x [3x3 s2 p1 conv, LayerNorm2d, GELU] then 1x1); CXBlock :62-118 x2; out_proj 1x1 -> mem_dim."""
    p = "memory_encoder"
    m = mask_for_mem
    for j in range(4):
        m = F.conv2d(m, sd[f"{p}.mask_downsampler.encoder.{3 * j}.weight"], sd[f"{p}.mask_downsampler.encoder.{3 * j}.bias"], stride=2, padding=1)
        m = F.gelu(layer_norm2d(sd, f"{p}.mask_downsampler.encoder.{3 * j + 1}", m))
    m = F.conv2d(m, sd[f"{p}.mask_downsampler.encoder.12.weight"], sd[f"{p}.mask_downsampler.encoder.12.bias"])
    x = F.conv2d(pix_feat, sd[p + ".pix_feat_proj.weight"], sd[p + ".pix_feat_proj.bias"]) + m
    for j in range(2):
        lp = f"{p}.fuser.layers.{j}"
        y = F.conv2d(x, sd[lp + ".dwconv.weight"], sd[lp + ".dwconv.bias"], padding=3, groups=x.shape[1])
        y = layer_norm2d(sd, lp + ".norm", y).permute(0, 2, 3, 1)
        y = linear(sd, lp + ".pwconv2", F.gelu(linear(sd, lp + ".pwconv1", y)))
        x = x + (sd[lp + ".weight"] * y).permute(0, 3, 1, 2)
    x = F.conv2d(x, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])
    pos = pos_embed_sine(x.shape[-2:], cfg.mem_dim)[None].expand(x.shape[0], -1, -1, -1)
    return x, pos


def encode_new_memory(sd, cfg: Sam2Cfg, pix_feat, pred_masks_high_res, is_mask_from_pts):
    """SAM2Base._encode_new_memory modeling/sam2_base.py:666-704 (binarize_mask_from_pts_for_mem_enc=true, build_sam.py:62)."""
    m = (pred_masks_high_res > 0).float() if is_mask_from_pts else torch.sigmoid(pred_masks_high_res)
    m = m * cfg.sigmoid_scale_for_mem_enc + cfg.sigmoid_bias_for_mem_enc
    return memory_encoder(sd, cfg, pix_feat, m)


# ------------------------------------------------------------------------------------------------ memory attention
def axial_cis(dim, end_x, end_y, theta=10000.0):
    """compute_axial_cis modeling/position_encoding.py:174-192"""
    fr = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    tx, ty = (t % end_x).float(), torch.div(t, end_x, rounding_mode="floor").float()
    fx, fy = torch.outer(tx, fr), torch.outer(ty, fr)
    return torch.cat([torch.polar(torch.ones_like(fx), fx), torch.polar(torch.ones_like(fy), fy)], dim=-1)


def apply_rope(x, cis, repeat=False):
    """apply_rotary_enc modeling/position_encoding.py:194-216 (one tensor at a time)."""
    x_ = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    if repeat:
        cis = cis.repeat(x_.shape[-2] // cis.shape[-2], 1)
    return torch.view_as_real(x_ * cis).flatten(-2)


def rope_attention(sd, p, q, k, v, num_k_exclude_rope=0, rope_k_repeat=False):
    """RoPEAttention.forward modeling/sam/transformer.py:285-327 (1 head; freqs recomputed for the actual grid :299-302)."""
    q, k, v = linear(sd, p + ".q_proj", q), linear(sd, p + ".k_proj", k), linear(sd, p + ".v_proj", v)
    q, k, v = q[:, None], k[:, None], v[:, None]
    side = int(math.sqrt(q.shape[-2]))
    cis = axial_cis(q.shape[-1], side, side)
    nk = k.shape[-2] - num_k_exclude_rope
    q = apply_rope(q, cis)
    k = torch.cat([apply_rope(k[:, :, :nk], cis, repeat=rope_k_repeat), k[:, :, nk:]], dim=2) if nk > 0 else k
    o = F.scaled_dot_product_attention(q, k, v)[:, 0]
    return linear(sd, p + ".out_proj", o)


def memory_attention(sd, cfg: Sam2Cfg, curr, curr_pos, memory, memory_pos, num_obj_ptr_tokens):
    """MemoryAttention.forward modeling/memory_attention.py:119-169; MemoryAttentionLayer :58-99.
    curr [HW,B,C], memory [Nk,B,mem_dim] (sequence-first as in the reference) -> [HW,B,C]."""
    x = (curr + 0.1 * curr_pos).transpose(0, 1)
    mem, mpos = memory.transpose(0, 1), memory_pos.transpose(0, 1)
    for i in range(cfg.mem_attn_layers):
        lp = f"memory_attention.layers.{i}"
        t = layer_norm(sd, lp + ".norm1", x)
        x = x + rope_attention(sd, lp + ".self_attn", t, t, t)
        t = layer_norm(sd, lp + ".norm2", x)
        x = x + rope_attention(sd, lp + ".cross_attn_image", t, mem + mpos, mem, num_obj_ptr_tokens, rope_k_repeat=True)
        t = layer_norm(sd, lp + ".norm3", x)
        x = x + linear(sd, lp + ".linear2", F.relu(linear(sd, lp + ".linear1", t)))
    return layer_norm(sd, "memory_attention.norm", x).transpose(0, 1)


# ------------------------------------------------------------------------------------------------ video session
@dataclass
class FrameOut:
    pred_masks: torch.Tensor            # [B,1,256,256] low-res logits
    obj_ptr: torch.Tensor               # [B,C]
    maskmem_features: torch.Tensor | None = None  # [B,mem_dim,h,w] (rounded to bf16 like sam2_video_predictor.py:967)
    maskmem_pos_enc: torch.Tensor | None = None


@dataclass
class VideoState:
    feats: dict = field(default_factory=dict)
    cond: dict = field(default_factory=dict)       # frame_idx -> FrameOut (conditioning frames)
    non_cond: dict = field(default_factory=dict)   # frame_idx -> FrameOut


def _flat(x):
    return x.flatten(2).permute(2, 0, 1)


def prepare_memory_conditioned_features(sd, cfg: Sam2Cfg, frame_idx, feats_top, pos_top, st: VideoState, num_frames, B):
    """SAM2Base._prepare_memory_conditioned_features modeling/sam2_base.py:495-664 for forward tracking of a
    non-initial frame (memory_temporal_stride 1, max_cond_frames -1, add_tpos_enc_to_obj_ptrs false)."""
    C, hw = cfg.d_model, cfg.feat_size
    mems, mem_pos = [], []
    t_pos_and_prevs = [(0, out) for out in st.cond.values()]
    for t_pos in range(1, cfg.num_maskmem):
        t_rel = cfg.num_maskmem - t_pos
        prev_idx = frame_idx - t_rel
        t_pos_and_prevs.append((t_pos, st.non_cond.get(prev_idx)))
    for t_pos, prev in t_pos_and_prevs:
        if prev is None:
            continue
        mems.append(_flat(prev.maskmem_features.float()))
        mem_pos.append(_flat(prev.maskmem_pos_enc) + sd["maskmem_tpos_enc"][cfg.num_maskmem - t_pos - 1])
    max_ptrs = min(num_frames, cfg.max_obj_ptrs_in_encoder)
    ptrs = [out.obj_ptr for t, out in st.cond.items() if t <= frame_idx]
    for t_diff in range(1, max_ptrs):
        t = frame_idx - t_diff
        if t < 0:
            break
        out = st.non_cond.get(t)
        if out is not None:
            ptrs.append(out.obj_ptr)
    n_ptr_tokens = 0
    if ptrs:
        op = torch.stack(ptrs, dim=0)  # [P,B,C]
        op = op.reshape(-1, B, C // cfg.mem_dim, cfg.mem_dim).permute(0, 2, 1, 3).flatten(0, 1)
        mems.append(op)
        mem_pos.append(op.new_zeros(op.shape))
        n_ptr_tokens = op.shape[0]
    memory, memory_pos = torch.cat(mems, 0), torch.cat(mem_pos, 0)
    out = memory_attention(sd, cfg, _flat(feats_top), _flat(pos_top), memory, memory_pos, n_ptr_tokens)
    return out.permute(1, 2, 0).reshape(B, C, hw, hw)


def run_video(sd, cfg: Sam2Cfg, images, text_embeds, video_hw, collect=None):
    """The VideoGLaMM video-branch call order (VideoGLaMM.py:846-875): init_state_from_tensor -> add_new_text(frame 0)
    per object -> propagate_in_video. sam2_video_predictor.py:109-180,416-495,675-729,745-827,921-1017.
    images [T,3,S,S]; text_embeds [n_obj,1,C]. Returns list over frames of video-res logits [n_obj,1,H,W]."""
    T, B = images.shape[0], text_embeds.shape[0]
    st = VideoState()

    def get_feats(t):
        if t not in st.feats:
            st.feats = {t: forward_image(sd, cfg, images[t:t + 1].float())}  # cache of ONE frame (:901)
        return st.feats[t]

    # add_new_text on frame 0, one object at a time (batch 1), init cond frame: + no_mem_embed (sam2_base.py:639-645)
    fpn, pos = get_feats(0)
    per_obj = []
    for o in range(B):
        pix = fpn[2] + sd["no_mem_embed"].reshape(1, -1, 1, 1)
        r = forward_sam_heads(sd, cfg, pix, [fpn[0], fpn[1]], text_inputs=text_embeds[o:o + 1], multimask_output=True)
        per_obj.append(r)
    # preflight: consolidate, binarised memory encoding with batch B (sam2_video_predictor.py:675-729, 520-636)
    low = torch.cat([r["low_res_masks"] for r in per_obj], 0)
    ptr = torch.cat([r["obj_ptr"] for r in per_obj], 0)
    high = F.interpolate(low, size=(cfg.image_size, cfg.image_size), mode="bilinear", align_corners=False)
    mf, mp = encode_new_memory(sd, cfg, fpn[2].expand(B, -1, -1, -1), high, is_mask_from_pts=True)
    st.cond[0] = FrameOut(low, ptr, mf.to(torch.bfloat16), mp)
    outs = []
    H, W = video_hw
    for t in range(T):
        if t in st.cond:
            pred = st.cond[t].pred_masks
        else:
            fpn, pos = get_feats(t)
            ex = lambda x: x.expand(B, -1, -1, -1)
            pix = prepare_memory_conditioned_features(sd, cfg, t, ex(fpn[2]), ex(pos[2]), st, T, B)
            r = forward_sam_heads(sd, cfg, pix, [ex(fpn[0]), ex(fpn[1])], text_inputs=None, multimask_output=True)
            mf, mp = encode_new_memory(sd, cfg, ex(fpn[2]), r["high_res_masks"], is_mask_from_pts=False)
            st.non_cond[t] = FrameOut(r["low_res_masks"], r["obj_ptr"], mf.to(torch.bfloat16), mp)
            pred = r["low_res_masks"]
            if collect is not None:
                collect.setdefault("pix_feat_with_mem", {})[t] = pix
        if collect is not None:
            collect.setdefault("low_res", {})[t] = pred
        outs.append(F.interpolate(pred, size=(H, W), mode="bilinear", align_corners=False))
    return outs


def run_framewise(sd, cfg: Sam2Cfg, images, text_embeds, video_hw, n_zero_dense: int = 0):
    """SAM2 part of VideoGLaMM_SAM2.inference_framewise (VideoGLaMM.py:598-768) with get_visual_embs_sam2 (:205-241): per frame,
    image_embed = top feature + no_mem_embed; sam_prompt_encoder(points=None, boxes=None, masks=None, text_embeds) -> the text token only;
    sam_mask_decoder(multimask_output=False, repeat_image=True) (stability fallback on); postprocess_masks (:147-153)."""
    B = text_embeds.shape[0]
    outs = []
    for t in range(images.shape[0]):
        fpn, _ = forward_image(sd, cfg, images[t:t + 1].float())
        image_embed = fpn[2] + sd["no_mem_embed"].reshape(1, -1, 1, 1)
        dense = sd["sam_prompt_encoder.no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(B, -1, cfg.feat_size, cfg.feat_size)
        if n_zero_dense:   # model_forward pads the objects to 4 with ZERO sparse and ZERO dense embeddings (VideoGLaMM.py:416-433): the last n rows
            dense = torch.cat([dense[: B - n_zero_dense], torch.zeros_like(dense[B - n_zero_dense:])], 0)
        hi = [fpn[0].expand(B, -1, -1, -1), fpn[1].expand(B, -1, -1, -1)]
        low, _, _, _, _, _ = mask_decoder(sd, cfg, image_embed, dense_pe(sd, cfg), text_embeds, dense, hi, multimask_output=False, repeat_image=True)
        outs.append(F.interpolate(low.float(), video_hw, mode="bilinear", align_corners=False))
    return outs
