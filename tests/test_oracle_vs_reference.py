"""Oracle vs the LIVE reference modules (needs /root/reference; skipped on the GPU box). Also checks that the synthetic
state dicts carry exactly the reference's keys/shapes (strict load)."""
import pytest
import torch

from oracle import ref_harness as H, sam2 as O, weights as Wt

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not H.available(), reason="reference tree not present")]


def test_sam2_state_dict_contract_and_video_parity():
    import sys
    sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__file__), "golden"))
    from make_golden import ref_sam2
    cfg = O.Sam2Cfg(**O.TINY)
    ref = ref_sam2(tiny=True)
    sd = Wt.make_sam2_state_dict(cfg, 77)
    rsd = ref.state_dict()
    assert set(sd) == set(rsd)
    assert all(tuple(sd[k].shape) == tuple(rsd[k].shape) for k in sd)
    ref.load_state_dict(sd, strict=True)
    torch.manual_seed(3)
    images, text = torch.randn(3, 3, cfg.image_size, cfg.image_size), torch.randn(2, 1, 256)
    with torch.inference_mode():
        st = ref.init_state_from_tensor(images, 48, 64)
        for o in range(2):
            ref.add_new_text(st, 0, o, text[o:o + 1])
        want = torch.stack([m for _, _, m in ref.propagate_in_video(st)])
        got = torch.stack(O.run_video(sd, cfg, images, text, (48, 64)))
    assert (got - want).abs().max() < 5e-3 * want.abs().max().clamp_min(1.0)


def test_full_size_sam2_key_contract():
    """Full Hiera-L shapes (no forward): the synthetic table must equal the reference's state_dict layout."""
    from make_golden import ref_sam2
    ref = ref_sam2(tiny=False)
    shapes = Wt.sam2_param_shapes(O.Sam2Cfg())
    rsd = ref.state_dict()
    assert set(shapes) == set(rsd)
    assert all(tuple(shapes[k]) == tuple(rsd[k].shape) for k in shapes)


def test_e2e_inference_live_reference_equals_golden_and_oracle():
    """Runs the reference's own VideoGLaMMForCausalLM.inference here and checks (a) the committed fixture is what it produces and
    (b) the oracle composition reproduces it (ids bit-exact, masks identical)."""
    import numpy as np
    import make_golden as MG
    from test_oracle_golden import _e2e_oracle, load_e2e_golden
    ref, *_ = MG.ref_videoglamm()
    out_ids, masks = MG.ref_e2e_run(ref)
    want_ids, want_masks = load_e2e_golden()
    assert np.array_equal(out_ids, want_ids.numpy()) and np.array_equal(masks, want_masks.numpy())
    ids, logits = _e2e_oracle()
    assert torch.equal(ids, want_ids) and torch.equal(logits > 0, want_masks)


def test_still_image_branch_live_reference_equals_oracle():
    """context_images=None (chat.py:458-487): one CLIP-preprocessed 336 px image, 576 un-pooled tokens at the single <image> sentinel, one SAM2
    frame. Token ids bit-exact and masks identical between the reference's inference() and the oracle composition."""
    import os
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as MG
    from oracle import lmm as M
    ref, sd, ocfg, scfg = MG.ref_videoglamm()
    torch.manual_seed(5)
    image, sam = torch.randn(1, 3, 336, 336), torch.randn(1, 3, 256, 256)
    ids = torch.cat([torch.tensor([1]), torch.randint(3, 290, (5,)), torch.tensor([-200]), torch.randint(3, 290, (6,))])
    with torch.no_grad():
        out_ids, segs = ref.inference([image], None, [sam], ids[None], None, [MG.E2E_HW], max_new_tokens=MG.E2E_NEW, use_sam2_video_branch=True)
        o_ids, emb, _ = M.generate(sd, ocfg, ids, image, None, MG.E2E_NEW, forced=MG.E2E_FORCED)
        ssd = {k[len("model.visual_model."):]: v for k, v in sd.items() if k.startswith("model.visual_model.")}
        logits = O.run_video(ssd, scfg, sam, emb[:, None, :], MG.E2E_HW)
    assert torch.equal(out_ids[0], o_ids)
    for o in range(emb.shape[0]):
        assert np.array_equal(segs[0][0][o], (logits[0][o, 0] > 0).numpy())


def test_model_forward_live_reference_equals_oracle():
    """The reference's own model_forward (validation path inference=True AND the training path's loss values) on the collate_fn dict vs
    oracle/model_forward.py: padded-to-4 per-frame masks and ce / bce / dice losses."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as MG
    from oracle import model_forward as OMF
    # get_visual_embs_sam2 hard-codes the 1024 px feature sizes (VideoGLaMM.py:225): tiny Hiera WIDTH / DEPTH, full 1024 px image
    kw = {k: v for k, v in O.TINY.items() if k != "image_size"}
    ref, sd, ocfg, scfg = MG.ref_videoglamm(sam2_overrides=[o for o in MG.SAM2_TINY_OVERRIDES if "image_size" not in o], sam2_cfg_kw=kw)
    ref.ce_loss_weight, ref.dice_loss_weight, ref.bce_loss_weight = 1.0, 0.5, 2.0
    torch.manual_seed(9)
    images, ctx, _, _ = MG.e2e_inputs()
    sam = torch.randn(2, 3, 1024, 1024)
    seg = MG.E2E_SEG
    ids = torch.cat([torch.tensor([1]), torch.randint(3, 290, (5,)), torch.full((4,), -200), torch.randint(3, 290, (4,)), torch.tensor([seg]),
                     torch.randint(3, 290, (3,)), torch.tensor([seg, 2])])[None]
    labels = ids.clone()
    labels[:, :12] = -100
    am = torch.ones_like(ids, dtype=torch.bool)
    gt = (torch.rand(2, 2, 48, 64) > 0.5).float()
    batch = dict(images_for_sam=[sam], images=[images], context_images=[ctx], input_ids=ids, labels=labels, attention_masks=am,
                 offset=torch.LongTensor([0, 1]), masks_list=[gt], label_list=[torch.zeros(48, 64)], resize_list=[(1024, 1024)])
    with torch.no_grad():
        want_v = ref.model_forward(**batch, inference=True)
        ref.training = True      # top module only: phi3.py:104-107 returns the per-layer hidden-state tuple only in training mode (submodules stay in eval)
        want_t = ref.model_forward(**batch, inference=False)
        ref.training = False
        got_v = OMF.model_forward(sd, ocfg, scfg, [sam], [images], [ctx], ids, labels, am, batch["offset"], [gt], batch["label_list"], inference=True)
        got_t = OMF.model_forward(sd, ocfg, scfg, [sam], [images], [ctx], ids, labels, am, batch["offset"], [gt], batch["label_list"], inference=False,
                                  weights=(1.0, 0.5, 2.0))
    assert len(want_v["pred_masks"][0]) == 2 and want_v["pred_masks"][0][0].shape == (4, 48, 64)          # padded to 4 objects
    for t in range(2):
        a, b = got_v["pred_masks"][0][t], want_v["pred_masks"][0][t].float()
        assert (a - b).abs().max() < 1e-3 * b.abs().max().clamp_min(1.0), (t, (a - b).abs().max())
    for k in ("ce_loss", "mask_bce_loss", "mask_dice_loss", "loss"):
        assert abs(float(got_t[k]) - float(want_t[k])) < 1e-4 * max(1.0, abs(float(want_t[k]))), (k, float(got_t[k]), float(want_t[k]))
