"""Training / validation forward of the reference (`VideoGLaMMForCausalLM.model_forward`, /root/reference/VideoGLaMM/model/VideoGLaMM.py:325-558) on
the B200 engine, taking the dict `collate_fn` builds (utils/dataset.py:28-110: images_for_sam, images, context_images, input_ids, labels,
attention_masks, offset, masks_list, label_list, resize_list, inference).

FORWARD ONLY: the engine has no backward pass, so the training path returns the loss VALUES (ce / bce / dice, the reference's weights and
normalisations) of one teacher-forced forward — what a validation loop or a loss-parity check needs — not gradients.
Steps (same numbering as the reference):
  1. SAM2 image embeddings of every frame (get_visual_embs_sam2 :205-241: top-level feature + no_mem_embed, high-res maps);
  3. one teacher-forced LMM forward per conversation (no generation): last-layer hidden states [L, hidden] and, on the training path, the logits;
  4. text_hidden_fcs on the rows that PREDICT a [SEG] token (seg_token_mask = ids[1:] == seg, arch.py:215-216) -> [n_seg, 256];
  5. per sample: pad to MAX_NUM_SEG_TOKENS_PER_SAMPLE = 4 objects with zero sparse AND zero dense embeddings (:416-433), decode every frame
     independently with multimask_output=False, repeat_image=True (:480-494), bilinear to label_list[b].shape (postprocess_masks :147-153);
  6. (training) ce_loss * w_ce + (bce * w_bce + dice * w_dice) over min(len(gt), len(pred)) objects per frame (:505-545).
The loss reductions (log-softmax over the vocabulary, BCE / DICE sums over pixels) are scalar reductions of kernel outputs, done with torch."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops

MASK_IGNORE_INDEX = -1                      # VideoGLaMM.py:14
MAX_NUM_SEG_TOKENS_PER_SAMPLE = 4           # VideoGLaMM.py:15
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
BF, F32 = torch.bfloat16, torch.float32


def dice_loss(inputs, targets, num_masks, ignore_index=None, scale=1000, eps=1e-6):
    """VideoGLaMM.py:20-52"""
    inputs = inputs.sigmoid().flatten(1, 2)
    targets = targets.flatten(1, 2)
    if ignore_index is not None:
        mask = targets != ignore_index
        inputs, targets = inputs * mask, targets * mask
    numerator = 2 * (inputs / scale * targets).sum(-1)
    denominator = (inputs / scale).sum(-1) + (targets / scale).sum(-1)
    return (1 - (numerator + eps) / (denominator + eps)).sum() / (num_masks + 1e-8)


def sigmoid_ce_loss(inputs, targets, num_masks, ignore_index=None):
    """VideoGLaMM.py:55-81"""
    loss = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none").flatten(1, 2)
    if ignore_index is not None:
        loss = loss * (targets.flatten(1, 2) != ignore_index)
    return loss.mean(1).sum() / (num_masks + 1e-8)


def _seg_rows(ids_row: torch.Tensor, seg_idx: int, n_visual: int) -> list[int]:
    """Rows of the EMBEDDED sequence whose hidden state predicts a [SEG]: text position p with ids[p+1] == seg (arch.py:215-216), shifted by the
    expansion of the run of -200 sentinels into n_visual rows (new_seg_token_masks, arch.py:283-379)."""
    ids = ids_row.tolist()
    img = [i for i, t in enumerate(ids) if t == IMAGE_TOKEN_INDEX]
    start, end = (img[0], img[-1]) if img else (len(ids), len(ids) - 1)
    rows = []
    for p in range(len(ids) - 1):
        if ids[p + 1] == seg_idx and ids[p] != IMAGE_TOKEN_INDEX:
            rows.append(p if p < start else p - (end - start + 1) + n_visual)
    return rows


@torch.no_grad()
def model_forward(model, images_for_sam, images, context_images, input_ids, labels=None, attention_masks=None, offset=None, masks_list=None,
                  label_list=None, resize_list=None, inference: bool = False, **_):
    model._ensure_engines()
    lmm, sam, dev = model.lmm, model.sam, model.device
    seg_idx = int(model.config.seg_token_idx)
    batch = len(images_for_sam)
    assert offset is not None and batch == len(offset) - 1                                     # VideoGLaMM.py:357
    if inference:
        assert len(images) == 1 and input_ids.shape[0] == 1                                    # __inference_path :246-247
    ph = lmm.cfg.phi
    P = lmm.P
    # ---- 3 / 4: teacher-forced LMM forward per conversation, [SEG] rows -> text_hidden_fcs
    pred_embeddings, ce_num, ce_den = [], torch.zeros((), device=dev), 0
    for b in range(batch):
        ctx = None if (context_images is None or context_images[b] is None) else context_images[b].to(dev, F32)
        vis = lmm.visual_tokens_any(images[b].to(dev, F32), ctx)                               # once per sample (the reference recomputes it per conversation)
        embs = []
        for c in range(int(offset[b]), int(offset[b + 1])):
            ids = input_ids[c]
            if attention_masks is not None:
                ids = ids[attention_masks[c].bool().to(ids.device)]                            # right padding carries no information under causal attention
            x = lmm.embed_prompt(ids.to(dev), vis)
            lmm._alloc_cache(x.shape[0] + 1)
            xr = lmm._llm(x, 0, raw=True)
            rows = _seg_rows(ids.cpu(), seg_idx, vis.shape[0])
            if rows:
                h = ops.norm(xr[rows].contiguous(), lmm.final_norm, eps=ph.eps, rms=True, out_dtype=F32)   # output_hidden_states[-1] rows
                h = h if P else ops.cast(h, BF)
                e = ops.gemm(ops.gemm(h, lmm.fc[0].W(P), bias=lmm.fc[0].b, act="relu", out_dtype=F32 if P else BF), lmm.fc[1].W(P), bias=lmm.fc[1].b, out_dtype=F32)
                embs.append(e)
            if not inference and labels is not None:                                           # HF causal-LM loss: shifted cross entropy, mean over valid targets
                lab = labels[c]
                if attention_masks is not None:
                    lab = lab[attention_masks[c].bool().to(lab.device)]
                lab = lab.tolist()
                img = [i for i, t in enumerate(ids.tolist()) if t == IMAGE_TOKEN_INDEX]
                if img:                                                                        # labels of the visual block are IGNORE_INDEX (arch.py:338-346)
                    lab = lab[:img[0]] + [IGNORE_INDEX] * vis.shape[0] + lab[img[-1] + 1:]
                tgt = torch.tensor(lab[1:], device=dev)
                valid = tgt != IGNORE_INDEX
                if bool(valid.any()):
                    logits = lmm.logits_all(xr[:-1][valid].contiguous())
                    ce_num = ce_num + F.cross_entropy(logits, tgt[valid], reduction="sum")
                    ce_den += int(valid.sum())
        pred_embeddings.append(torch.cat(embs, 0) if embs else torch.zeros((0, lmm.cfg.out_dim), device=dev))
    # ---- 1 / 5: per sample, every frame decoded independently, objects padded to MAX_NUM_SEG_TOKENS_PER_SAMPLE
    pred_masks = []
    for b in range(batch):
        emb = pred_embeddings[b]
        n_real = emb.shape[0]
        n_pad = max(MAX_NUM_SEG_TOKENS_PER_SAMPLE - n_real, 0) if n_real else MAX_NUM_SEG_TOKENS_PER_SAMPLE
        H, W = label_list[b].shape[-2:]
        frames = images_for_sam[b].to(dev, F32)
        real = sam.run_framewise(frames, emb[:, None, :], (int(H), int(W))) if n_real else None
        pad = sam.run_framewise(frames, torch.zeros((n_pad, 1, emb.shape[1]), device=dev), (int(H), int(W)), zero_dense=True) if n_pad else None
        pred_masks.append([torch.cat([x[t] for x in (real, pad) if x is not None], 0) for t in range(frames.shape[0])])
    if inference:
        return {"pred_masks": pred_masks, "gt_masks": masks_list}
    # ---- 6: losses (forward values)
    ce_loss = ce_num / max(ce_den, 1) * float(getattr(model, "ce_loss_weight", 1.0) or 1.0)
    bce, dice, num_masks = torch.zeros((), device=dev), torch.zeros((), device=dev), 0
    for b in range(batch):
        gt = masks_list[b].to(dev, F32)
        gt = gt[:, None] if gt.dim() == 3 else gt                                              # image samples: [n, H, W] -> [n, 1, H, W]
        assert gt.shape[1] == len(pred_masks[b]), f"len(gt_mask): {gt.shape[1]}, len(pred_mask): {len(pred_masks[b])}"
        for t, pm in enumerate(pred_masks[b]):
            g = gt[:, t]
            n = min(g.shape[0], pm.shape[0])
            if n == 0:
                continue
            bce = bce + sigmoid_ce_loss(pm[:n], g[:n], num_masks=n, ignore_index=MASK_IGNORE_INDEX) * n
            dice = dice + dice_loss(pm[:n], g[:n], num_masks=n, ignore_index=MASK_IGNORE_INDEX) * n
            num_masks += n
    mask_bce_loss = float(getattr(model, "bce_loss_weight", 1.0) or 1.0) * bce / (num_masks + 1e-8)
    mask_dice_loss = float(getattr(model, "dice_loss_weight", 1.0) or 1.0) * dice / (num_masks + 1e-8)
    mask_loss = mask_bce_loss + mask_dice_loss
    return {"loss": ce_loss + mask_loss, "ce_loss": ce_loss, "mask_bce_loss": mask_bce_loss, "mask_dice_loss": mask_dice_loss, "mask_loss": mask_loss}
