"""Parity at BASELINE.json's FULL widths (Hiera-L 144..1152 / head_dim 72, Phi-3 3072 x 32 heads x 96, CLIP-L 1024 x 16 x 64,
InternVideo2 1408 x 16 x 88, SAM2 decoder / memory 256) against the CPU fp32 oracle. Depth is reduced where the CPU oracle would take
minutes (the layers are identical); the SAM2 image encoder runs at its true depth (48 blocks) on one 1024x1024 frame."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import lmm as M, sam2 as O, weights as Wt  # noqa: E402
from videoglamm_b200 import ops  # noqa: E402
from videoglamm_b200.config import Sam2Config, make_lmm_config  # noqa: E402
from videoglamm_b200.lmm_engine import LmmEngine  # noqa: E402
from videoglamm_b200.sam2_engine import Sam2Engine  # noqa: E402


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def tok(x):
    return x[0].permute(1, 2, 0).reshape(-1, x.shape[1])


def test_hiera_l_full_depth_one_frame_and_heads():
    ocfg = O.Sam2Cfg()
    sd = Wt.make_sam2_state_dict(ocfg, 4321)
    sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
    eng = Sam2Engine(sd, Sam2Config())
    torch.manual_seed(1)
    img = torch.randn(1, 3, 1024, 1024)
    with torch.no_grad():
        fpn, pos = O.forward_image(sd, ocfg, img)
    f = eng.forward_image(img.cuda())
    assert f["top"].shape == (4096, 256) and f["s0"].shape == (65536, 32) and f["s1"].shape == (16384, 64)
    assert rel(f["top"], tok(fpn[2])) < 2e-2 and rel(f["s1"], tok(fpn[1])) < 2e-2 and rel(f["s0"], tok(fpn[0])) < 2e-2
    # heads + memory encoder + one memory-attention call on the oracle's features (identical inputs), 2 objects
    text = torch.randn(2, 1, 256)
    pix = (fpn[2] + sd["no_mem_embed"].reshape(1, -1, 1, 1)).expand(2, -1, -1, -1)
    hi = [fpn[0].expand(2, -1, -1, -1), fpn[1].expand(2, -1, -1, -1)]
    with torch.no_grad():
        oh = O.forward_sam_heads(sd, ocfg, pix, hi, text_inputs=text)
        mf, _ = O.encode_new_memory(sd, ocfg, fpn[2].expand(2, -1, -1, -1), oh["high_res_masks"], is_mask_from_pts=False)
    s0, s1 = tok(fpn[0]).cuda().bfloat16().contiguous(), tok(fpn[1]).cuda().bfloat16().contiguous()
    eh = eng.forward_sam_heads(pix.permute(0, 2, 3, 1).reshape(-1, 256).cuda().contiguous(), s0, s1, text.cuda(), 2)
    assert rel(eh["low_res_masks"], oh["low_res_masks"][:, 0]) < 2e-2
    # the fused mask-decoder kernel at full size (HW = 4096 -> 128 row blocks per object, 2 objects) on fp32 high-res maps: north_star's 1e-3 bar
    ef = eng.forward_sam_heads(pix.permute(0, 2, 3, 1).reshape(-1, 256).cuda().contiguous(), tok(fpn[0]).cuda().contiguous(), tok(fpn[1]).cuda().contiguous(),
                               text.cuda(), 2)
    e_masks, e_ptr = rel(ef["low_res_masks"], oh["low_res_masks"][:, 0]), rel(ef["obj_ptr"], oh["obj_ptr"])
    print(f"PARITY full-size fused mask decoder (HW=4096, B=2): low_res_masks {e_masks:.2e}, obj_ptr {e_ptr:.2e}")
    assert e_masks < 1e-3 and e_ptr < 1e-3
    emf = eng.encode_memory(tok(fpn[2]).cuda().contiguous(), oh["low_res_masks"][:, 0].cuda().contiguous(), binarize=False)
    assert rel(emf, mf.permute(0, 2, 3, 1).reshape(2, -1, 64)) < 1.5e-2
    Nk = 3 * 4096 + 12
    mem = torch.randn(Nk, 2, 64).bfloat16().float()
    mpos = torch.randn(Nk, 1, 64).expand(-1, 2, -1)
    with torch.no_grad():
        oma = O.memory_attention(sd, ocfg, O._flat(fpn[2]).expand(-1, 2, -1), O._flat(pos[2]).expand(-1, 2, -1), mem, mpos, 12)
    ema = eng.memory_attention(tok(fpn[2]).cuda().contiguous(), 2, mem.permute(1, 0, 2).contiguous().cuda().bfloat16(), mpos[:, 0].contiguous().cuda(), 12)
    assert rel(ema.view(2, 4096, 256), oma.permute(1, 0, 2)) < 1e-2


def test_full_width_towers_and_phi3_reduced_depth():
    kw = dict(clip=dict(layers=3), iv2=dict(depth=3), phi=dict(layers=2, sliding_window=2047))
    ocfg, ecfg = M.make_cfg(**kw, seg_token_idx=32011), make_lmm_config(**kw, seg_token_idx=32011)
    sd = Wt.make_lmm_state_dict(ocfg, 99)
    sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
    eng = LmmEngine(sd, ecfg)
    torch.manual_seed(2)
    T = 4
    images, ctx = torch.randn(T, 3, 224, 224), torch.randn(T, 3, 336, 336)
    with torch.no_grad():
        vis = M.visual_tokens(sd, ocfg, images, ctx)
    evis = eng.visual_tokens(images.cuda(), ctx.cuda())
    assert evis.shape == vis.shape == (T * 208, 3072) and rel(evis, vis) < 1.5e-2
    ids = torch.cat([torch.randint(3, 32000, (9,)), torch.full((T,), -200), torch.randint(3, 32000, (14,))])
    with torch.no_grad():
        ids_ref, emb_ref, _ = M.generate(sd, ocfg, ids, images, ctx, 4, forced={1: 32011})
    out_ids, emb = eng.generate(ids[None].cuda(), images.cuda(), ctx.cuda(), 4, forced={1: 32011})
    assert torch.equal(out_ids[0].cpu(), ids_ref)          # greedy ids bit-exact at full width (855-token sequence, KV-cached vs no-cache)
    assert rel(emb, emb_ref) < 2e-2
    # batched decode at full width: S = 3 sequences (different prompt lengths) through the multi-row weight-streaming kernel
    # (csrc/gemv_ring.cu gemv_ring_mma_kernel: qkv / gate_up / lm_head with fused RMSNorm on fp32 rows, o_proj / down_proj on bf16 rows),
    # teacher-forced to the oracle's tokens: ids reproduced, [SEG] hidden state within the same bar as the single-sequence path
    forced_all = {st: int(ids_ref[ids.numel() + st]) for st in range(4)}
    reqs = [(ids[None].cuda(), images.cuda(), ctx.cuda()),
            (torch.cat([ids[:9], ids[9:9 + T], ids[9 + T:][:5]])[None].cuda(), images.cuda(), ctx.cuda()),
            (ids[None].cuda(), images.cuda(), ctx.cuda())]
    n0 = ops.launch_count()
    outs = eng.generate_batch(reqs, 4, forced=forced_all)
    assert ops.launch_count() > n0
    for k in (0, 2):
        assert torch.equal(outs[k][0][0].cpu(), ids_ref)
        assert rel(outs[k][1], emb_ref) < 2e-2
    single = eng.generate(*reqs[1], 4, forced=forced_all)
    assert torch.equal(outs[1][0], single[0]) and rel(outs[1][1], single[1]) < 5e-3


def test_phi3_prefill_and_decode_beyond_the_sliding_window():
    """BASELINE.json configs[3]/[4] (C4 / C5): T_enc = 16 -> 3,328 visual tokens, L = 3,352 > sliding_window + 1: the band cuts in the prefill
    (tcgen05 attention with band-edge masking) and in every decode step (single-query kernel over the last 2,048 keys). Full Phi-3 width,
    one layer (the CPU oracle re-forwards the whole sequence per token); greedy ids bit-exact, [SEG] embedding close."""
    kw = dict(clip=dict(layers=2), iv2=dict(depth=2), phi=dict(layers=1, sliding_window=2047))
    ocfg, ecfg = M.make_cfg(**kw, seg_token_idx=32011), make_lmm_config(**kw, seg_token_idx=32011)
    sd = Wt.make_lmm_state_dict(ocfg, 77)
    sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
    eng = LmmEngine(sd, ecfg)
    torch.manual_seed(3)
    T = 16
    images, ctx = torch.randn(T, 3, 224, 224), torch.randn(T, 3, 336, 336)
    ids = torch.cat([torch.randint(3, 32000, (9,)), torch.full((T,), -200), torch.randint(3, 32000, (15,))])
    with torch.no_grad():
        ids_ref, emb_ref, hidden = M.generate(sd, ocfg, ids, images, ctx, 3, forced={1: 32011})
    assert hidden.shape[0] > 2048 + 1300
    out_ids, emb = eng.generate(ids[None].cuda(), images.cuda(), ctx.cuda(), 3, forced={1: 32011})
    assert torch.equal(out_ids[0].cpu(), ids_ref)
    assert rel(emb, emb_ref) < 2e-2
