"""TEST INFRASTRUCTURE (oracle) — the SAM2 mask decoder (oracle/sam2.py:mask_decoder, i.e. sam/mask_decoder.py:110-245 + sam/transformer.py:63-193)
re-scheduled the way the planned fused B200 kernel will run it, to prove on CPU that the schedule is exact:

  * the 64x64 image tokens are cut into `n_blocks` row blocks that never exchange data except through THREE reductions
    (the token->image attentions of layer 0, layer 1 and the final one), each carrying per (object, head, token) a running max, a
    running sum and a [head_dim] partial output — the flash-decoding merge;
  * everything else on the image side (k/v projections, image->token attention, norm4, both transposed convolutions with their
    high-resolution skip inputs, LayerNorm2d, GELU, the hyper-network product) is local to a block of rows;
  * the 8-9-token chain (self-attention, MLP, the token side of the cross attentions, hyper-network / IoU / object-score MLPs) is tiny and
    global.

`mask_decoder_rowblock` returns the same tuple as `oracle.sam2.mask_decoder` (multimask_output=True); tests/test_mask_decoder_schedule_cpu.py
checks equality to fp32 reassociation noise for several block counts."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .sam2 import Sam2Cfg, layer_norm, linear, mlp


def _heads(t, H):
    return t.reshape(t.shape[0], t.shape[1], H, -1).transpose(1, 2)          # [B, H, N, d]


def _t2i_partial(q, k_blk, v_blk):
    """One row block's contribution to a token->image attention: q [B,H,Nt,d], k/v [B,H,rows,d] -> (m [B,H,Nt], l [B,H,Nt], o [B,H,Nt,d])."""
    s = (q @ k_blk.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    m = s.amax(-1)
    p = torch.exp(s - m[..., None])
    return m, p.sum(-1), p @ v_blk


def _merge(parts):
    """Grid-wide reduction #k: combine the per-block (m, l, o) partials."""
    m_all = torch.stack([p[0] for p in parts]).amax(0)
    l_all = sum(p[1] * torch.exp(p[0] - m_all) for p in parts)
    o_all = sum(p[2] * torch.exp(p[0] - m_all)[..., None] for p in parts)
    return o_all / l_all[..., None]


def mask_decoder_rowblock(sd, cfg: Sam2Cfg, image_embeddings, image_pe, sparse, dense, high_res_features, n_blocks: int = 16):
    p, tp = "sam_mask_decoder", "sam_mask_decoder.transformer"
    B, H = sparse.shape[0], cfg.dec_heads
    out_tokens = torch.cat([sd[p + ".obj_score_token.weight"], sd[p + ".iou_token.weight"], sd[p + ".mask_tokens.weight"]], 0)
    tokens = torch.cat([out_tokens[None].expand(B, -1, -1), sparse], dim=1)
    b, c, h, w = image_embeddings.shape
    assert h % n_blocks == 0 or n_blocks % h == 0
    rows_per = max(h // n_blocks, 1)                                            # image rows (of w tokens) per block
    blocks = [slice(r * w, (r + rows_per) * w) for r in range(0, h, rows_per)]   # token ranges
    keys = [(image_embeddings + dense).flatten(2).permute(0, 2, 1)[:, s] for s in blocks]      # per block [B, rows*w, C], resident for the whole decoder
    pe = [image_pe.flatten(2).permute(0, 2, 1).expand(B, -1, -1)[:, s] for s in blocks]
    queries, query_pe = tokens, tokens

    def t2i(prefix, queries_in):
        """token->image attention = block-local partials + ONE reduction."""
        q = _heads(linear(sd, prefix + ".q_proj", queries_in + query_pe), H)
        parts = []
        for kb, pb in zip(keys, pe):
            kk = _heads(linear(sd, prefix + ".k_proj", kb + pb), H)
            vv = _heads(linear(sd, prefix + ".v_proj", kb), H)
            parts.append(_t2i_partial(q, kk, vv))
        o = _merge(parts).transpose(1, 2).reshape(B, queries_in.shape[1], -1)
        return linear(sd, prefix + ".out_proj", o)

    for i in range(cfg.dec_depth):
        lp = f"{tp}.layers.{i}"
        # ---- token chain (global, tiny)
        sa = lambda q_, k_, v_: linear(sd, lp + ".self_attn.out_proj", F.scaled_dot_product_attention(
            _heads(linear(sd, lp + ".self_attn.q_proj", q_), H), _heads(linear(sd, lp + ".self_attn.k_proj", k_), H),
            _heads(linear(sd, lp + ".self_attn.v_proj", v_), H)).transpose(1, 2).reshape(B, queries.shape[1], -1))
        queries = sa(queries, queries, queries) if i == 0 else queries + sa(queries + query_pe, queries + query_pe, queries)
        queries = layer_norm(sd, lp + ".norm1", queries)
        queries = layer_norm(sd, lp + ".norm2", queries + t2i(lp + ".cross_attn_token_to_image", queries))        # reduction i + 1
        queries = layer_norm(sd, lp + ".norm3", queries + mlp(sd, lp + ".mlp", queries, 2))
        # ---- image->token attention + norm4: local to each row block (keys/values are the 8-9 tokens)
        ip = lp + ".cross_attn_image_to_token"
        kt = _heads(linear(sd, ip + ".k_proj", queries + query_pe), H)
        vt = _heads(linear(sd, ip + ".v_proj", queries), H)
        new_keys = []
        for kb, pb in zip(keys, pe):
            qi = _heads(linear(sd, ip + ".q_proj", kb + pb), H)
            o = F.scaled_dot_product_attention(qi, kt, vt).transpose(1, 2).reshape(B, kb.shape[1], -1)
            new_keys.append(layer_norm(sd, lp + ".norm4", kb + linear(sd, ip + ".out_proj", o)))
        keys = new_keys
    queries = layer_norm(sd, tp + ".norm_final_attn", queries + t2i(tp + ".final_attn_token_to_image", queries))   # reduction 3
    hs = queries
    mask_tokens_out = hs[:, 2:6]
    hyper_in = torch.stack([mlp(sd, f"{p}.output_hypernetworks_mlps.{i}", mask_tokens_out[:, i, :], 3) for i in range(4)], dim=1)   # [B,4,C/8]
    # ---- upscaling + mask product: local to each row block (ConvTranspose k=2,s=2 maps one input pixel to its own 2x2 output patch)
    feat_s0, feat_s1 = high_res_features
    mask_rows = []
    for bi, kb in enumerate(keys):
        r0 = bi * rows_per
        x = kb.transpose(1, 2).reshape(B, c, rows_per, w)
        up = F.conv_transpose2d(x, sd[p + ".output_upscaling.0.weight"], sd[p + ".output_upscaling.0.bias"], stride=2) + feat_s1[:, :, 2 * r0:2 * (r0 + rows_per)]
        u = up.mean(1, keepdim=True)
        v = (up - u).pow(2).mean(1, keepdim=True)
        up = F.gelu(sd[p + ".output_upscaling.1.weight"][:, None, None] * ((up - u) / torch.sqrt(v + 1e-6)) + sd[p + ".output_upscaling.1.bias"][:, None, None])
        up = F.gelu(F.conv_transpose2d(up, sd[p + ".output_upscaling.3.weight"], sd[p + ".output_upscaling.3.bias"], stride=2) + feat_s0[:, :, 4 * r0:4 * (r0 + rows_per)])
        bb, cc, hh, ww = up.shape
        mask_rows.append((hyper_in @ up.reshape(bb, cc, hh * ww)).view(bb, -1, hh, ww))
    masks = torch.cat(mask_rows, dim=2)
    iou_pred = mlp(sd, p + ".iou_prediction_head", hs[:, 1, :], 3, sigmoid_output=True)
    obj_logits = mlp(sd, p + ".pred_obj_score_head", hs[:, 0, :], 3)
    return masks[:, 1:], iou_pred[:, 1:], mask_tokens_out[:, 1:], obj_logits, masks, iou_pred
