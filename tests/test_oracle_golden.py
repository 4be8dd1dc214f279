"""Pins the oracle (oracle/*.py) against golden vectors produced by the UNMODIFIED reference modules
(tests/golden/make_golden.py). Runs on CPU anywhere — the fixtures, not the reference tree, are needed."""
import os

import numpy as np
import torch

from oracle import lmm as M, sam2 as O, weights as Wt

G = os.path.join(os.path.dirname(__file__), "golden")


def test_sam2_video_session_matches_reference_golden():
    g = np.load(os.path.join(G, "sam2_tiny_video.npz"))
    cfg = O.Sam2Cfg(**O.TINY)
    for name, bias, T, nobj in (("present", 4.0, 9, 2), ("absent", -100.0, 2, 1)):
        sd = Wt.make_sam2_state_dict(cfg, 1234, obj_score_bias=bias)
        torch.manual_seed(5678)
        images, text = torch.randn(T, 3, cfg.image_size, cfg.image_size), torch.randn(nobj, 1, 256)
        with torch.inference_mode():
            out = torch.stack(O.run_video(sd, cfg, images, text, (60, 100)))
        ref = torch.from_numpy(g[name])
        assert out.shape == ref.shape
        # fp32 reassociation noise only (maskmem is rounded to bf16 on both sides, which amplifies 1e-6 to ~1e-3 late in the clip)
        assert (out - ref).abs().max() < 5e-3 * ref.abs().max().clamp_min(1.0)
        assert ((out > 0) == (ref > 0)).float().mean() > 0.9999
        if name == "absent":
            assert torch.all(out == O.NO_OBJ_SCORE)


def test_sam2_framewise_matches_reference_golden():
    g = np.load(os.path.join(G, "sam2_tiny_framewise.npz"))
    cfg = O.Sam2Cfg(**O.TINY)
    sd = Wt.make_sam2_state_dict(cfg, 1234)
    torch.manual_seed(9)
    images, text = torch.randn(2, 3, 256, 256), torch.randn(3, 1, 256)
    with torch.no_grad():
        out = torch.stack(O.run_framewise(sd, cfg, images, text, (50, 70)))
    ref = torch.from_numpy(g["masks"])
    assert out.shape == ref.shape and (out - ref).abs().max() < 1e-4 * ref.abs().max()


def test_towers_and_llm_match_reference_golden():
    g = np.load(os.path.join(G, "towers_tiny.npz"))
    cfg = M.make_cfg(**M.TINY_LMM)
    sd = Wt.make_lmm_state_dict(cfg, 1)
    torch.manual_seed(0)
    xv = torch.randn(2, 4, 3, cfg.iv2.image_size, cfg.iv2.image_size)
    xc = torch.randn(4, 3, cfg.clip.image_size, cfg.clip.image_size)
    with torch.no_grad():
        iv2 = M.iv2_tower(sd, "model.vision_tower", cfg.iv2, xv)
        clip = M.clip_tower(sd, "model.image_vision_tower.vision_tower", cfg.clip, xc)
        ids = torch.from_numpy(g["ids"])[0]
        lg = torch.nn.functional.linear(M.phi3_forward(sd, cfg.phi, sd["model.embed_tokens.weight"][ids]), sd["lm_head.weight"])
    for got, key in ((iv2, "iv2"), (clip, "clip"), (lg, "phi3_logits")):
        ref = torch.from_numpy(g[key])
        assert (got - ref).abs().max() < 2e-5 * ref.abs().max(), key
    assert torch.equal(lg.argmax(-1), torch.from_numpy(g["phi3_logits"]).argmax(-1))  # token ids bit-exact


def test_phi3_sliding_window_rule_is_transformers_4_41():
    """keys j >= i - sliding_window stay visible (sw+1 keys incl. self), SURVEY.md §7."""
    cfg = M.make_cfg(**M.TINY_LMM)
    sd = Wt.make_lmm_state_dict(cfg, 1)
    x = sd["model.embed_tokens.weight"][torch.arange(12)]
    cfg.phi.sliding_window = 100
    full = M.phi3_forward(sd, cfg.phi, x)
    cfg.phi.sliding_window = 11
    assert torch.allclose(M.phi3_forward(sd, cfg.phi, x), full)           # window covers everything: j >= i - 11 for i <= 11
    cfg.phi.sliding_window = 10
    assert not torch.allclose(M.phi3_forward(sd, cfg.phi, x)[-1], full[-1])  # row 11 loses key 0


def _e2e_oracle():
    import sys
    sys.path.insert(0, G)
    import make_golden as MG
    ocfg, scfg = M.make_cfg(**MG.E2E_LMM, seg_token_idx=MG.E2E_SEG), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=11)
    images, ctx, sam, ids = MG.e2e_inputs()
    with torch.no_grad():
        out_ids, emb, _ = M.generate(sd, ocfg, ids, images, ctx, MG.E2E_NEW, forced=MG.E2E_FORCED)
        ssd = {k[len("model.visual_model."):]: v for k, v in sd.items() if k.startswith("model.visual_model.")}
        logits = O.run_video(ssd, scfg, sam, emb[:, None, :], MG.E2E_HW)
    return out_ids, torch.stack([l[:, 0] for l in logits])          # ids [L], logits [T, n_obj, H, W]


def load_e2e_golden():
    g = np.load(os.path.join(G, "e2e_tiny.npz"))
    shape = tuple(int(x) for x in g["mask_shape"])
    return torch.from_numpy(g["output_ids"]), torch.from_numpy(np.unpackbits(g["masks"], axis=-1, count=shape[-1]).astype(bool).reshape(shape))


def test_e2e_composition_matches_reference_golden():
    """The whole path as ONE call of the unmodified reference (VideoGLaMMForCausalLM.inference, video branch, 2 forced [SEG] -> 2 objects,
    3 frames): output ids bit-exact, every bool mask pixel identical."""
    want_ids, want_masks = load_e2e_golden()
    ids, logits = _e2e_oracle()
    assert torch.equal(ids, want_ids)
    assert tuple(logits.shape) == tuple(want_masks.shape)
    assert torch.equal(logits > 0, want_masks)
    want_logits = torch.from_numpy(np.load(os.path.join(G, "e2e_tiny.npz"))["logits"])   # the reference's own video-res logits (propagate_in_video)
    err = ((logits - want_logits).norm() / want_logits.norm()).item()
    assert err < 1e-4, err
