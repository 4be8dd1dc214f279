"""TEST INFRASTRUCTURE (oracle) — CPU fp32 restatement of VideoGLaMMForCausalLM.model_forward (/root/reference/VideoGLaMM/model/VideoGLaMM.py:325-558)
over the flat state_dict, composed from oracle.lmm / oracle.sam2. Pinned against the reference's own model_forward run here
(tests/test_oracle_vs_reference.py::test_model_forward_live_reference_equals_oracle). Only tests may import it."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import lmm as M, sam2 as O

MASK_IGNORE_INDEX, MAX_SEG, IGNORE_INDEX, IMAGE_TOKEN_INDEX = -1, 4, -100, -200


def dice_loss(inputs, targets, num_masks, ignore_index=None, scale=1000, eps=1e-6):
    """VideoGLaMM.py:20-52"""
    inputs, targets = inputs.sigmoid().flatten(1, 2), targets.flatten(1, 2)
    if ignore_index is not None:
        mask = targets != ignore_index
        inputs, targets = inputs * mask, targets * mask
    num = 2 * (inputs / scale * targets).sum(-1)
    den = (inputs / scale).sum(-1) + (targets / scale).sum(-1)
    return (1 - (num + eps) / (den + eps)).sum() / (num_masks + 1e-8)


def sigmoid_ce_loss(inputs, targets, num_masks, ignore_index=None):
    """VideoGLaMM.py:55-81"""
    loss = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none").flatten(1, 2)
    if ignore_index is not None:
        loss = loss * (targets.flatten(1, 2) != ignore_index)
    return loss.mean(1).sum() / (num_masks + 1e-8)


def model_forward(sd, lcfg, scfg, images_for_sam, images, context_images, input_ids, labels, attention_masks, offset, masks_list, label_list,
                  inference=False, weights=(1.0, 1.0, 1.0)):
    """Returns the reference's dict: {"pred_masks", "gt_masks"} (inference) or the five losses."""
    ssd = {k[len("model.visual_model."):]: v for k, v in sd.items() if k.startswith("model.visual_model.")}
    batch = len(images_for_sam)
    pred_emb, ce_num, ce_den = [], 0.0, 0
    for b in range(batch):
        vis = M.visual_tokens(sd, lcfg, images[b], None if context_images is None else context_images[b])
        embs = []
        for c in range(int(offset[b]), int(offset[b + 1])):
            keep = attention_masks[c].bool() if attention_masks is not None else torch.ones_like(input_ids[c], dtype=torch.bool)
            ids = input_ids[c][keep]
            hidden = M.phi3_forward(sd, lcfg.phi, M.splice(sd, lcfg, ids, vis))
            n_new = hidden.shape[0] - ids.shape[0]
            img = torch.where(ids == IMAGE_TOKEN_INDEX)[0]
            start, end = int(img[0]), int(img[-1])
            seg_txt = torch.cat([ids[1:] == lcfg.seg_token_idx, torch.zeros(1, dtype=torch.bool)])          # arch.py:215-216
            seg = torch.cat([seg_txt[:start], torch.zeros(vis.shape[0], dtype=torch.bool), seg_txt[end + 1:]])   # new_seg_token_masks
            assert seg.shape[0] == hidden.shape[0] and n_new == vis.shape[0] - (end - start + 1)
            feats = M.lin(sd, "model.text_hidden_fcs.0.2", F.relu(M.lin(sd, "model.text_hidden_fcs.0.0", hidden)))
            embs.append(feats[seg])
            if not inference:
                lab = labels[c][keep]
                lab = torch.cat([lab[:start], torch.full((vis.shape[0],), IGNORE_INDEX), lab[end + 1:]])
                logits = F.linear(hidden, sd["lm_head.weight"])
                tgt = lab[1:]
                valid = tgt != IGNORE_INDEX
                ce_num = ce_num + F.cross_entropy(logits[:-1][valid], tgt[valid], reduction="sum")
                ce_den += int(valid.sum())
        pred_emb.append(torch.cat(embs, 0))
    pred_masks = []
    for b in range(batch):
        emb = pred_emb[b]
        n_real = emb.shape[0]
        n_pad = MAX_SEG if n_real == 0 else max(MAX_SEG - n_real, 0)
        text = torch.cat([emb, torch.zeros(n_pad, emb.shape[1])], 0)[:, None, :]
        H, W = label_list[b].shape[-2:]
        frames = O.run_framewise(ssd, scfg, images_for_sam[b], text, (int(H), int(W)), n_zero_dense=n_pad)
        pred_masks.append([f[:, 0] for f in frames])
    if inference:
        return {"pred_masks": pred_masks, "gt_masks": masks_list}
    ce = ce_num / max(ce_den, 1) * weights[0]
    bce, dice, n_masks = 0.0, 0.0, 0
    for b in range(batch):
        gt = masks_list[b].float()
        gt = gt[:, None] if gt.dim() == 3 else gt
        for t, pm in enumerate(pred_masks[b]):
            g = gt[:, t]
            n = min(g.shape[0], pm.shape[0])
            bce = bce + sigmoid_ce_loss(pm[:n], g[:n], n, MASK_IGNORE_INDEX) * n
            dice = dice + dice_loss(pm[:n], g[:n], n, MASK_IGNORE_INDEX) * n
            n_masks += n
    bce, dice = weights[2] * bce / (n_masks + 1e-8), weights[1] * dice / (n_masks + 1e-8)
    return {"loss": ce + bce + dice, "ce_loss": ce, "mask_bce_loss": bce, "mask_dice_loss": dice, "mask_loss": bce + dice}
