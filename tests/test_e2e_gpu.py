"""End to end through the public API (`VideoGLaMMForCausalLM.inference`) vs the oracle: token ids identical, masks agree."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_no_seg_returns_empty_dict_and_errors():
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    ocfg, scfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=300), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=300), Sam2Config(**O.TINY))
    T = 4
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.full((T,), -200), torch.randint(3, 290, (7,))])[None].cuda()
    args = ([torch.randn(T, 3, 56, 56).cuda()], [torch.randn(T, 3, 56, 56).cuda()], [torch.randn(2, 3, 256, 256).cuda()], ids, None, [(48, 64)])
    out_ids, segs = model.inference(*args, max_new_tokens=3, use_sam2_video_branch=True, forced_tokens={0: 5, 1: 6, 2: 7})
    assert segs == [{}] and out_ids.shape[1] == ids.shape[1] + 3            # VideoGLaMM.py:840-842
    model.use_sam2 = False
    with pytest.raises(ValueError):                                          # VideoGLaMM.py:585
        model.inference(*args, use_sam2_video_branch=True)
    bad = dict(sd)
    bad.pop("model.norm.weight")
    with pytest.raises(KeyError):
        VideoGLaMMForCausalLM(bad, make_lmm_config(**M.TINY_LMM, seg_token_idx=300), Sam2Config(**O.TINY))


def test_framewise_dispatch_through_public_api():
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    ocfg, scfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=300), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=300), Sam2Config(**O.TINY))
    torch.manual_seed(3)
    T = 4
    images, ctx, sam = torch.randn(T, 3, 56, 56), torch.randn(T, 3, 56, 56), torch.randn(2, 3, 256, 256)
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.full((T,), -200), torch.randint(3, 290, (7,))])
    forced = {1: 300}
    out_ids, segs = model.inference([images.cuda()], [ctx.cuda()], [sam.cuda()], ids[None].cuda(), None, [(48, 64)], max_new_tokens=4,
                                    use_sam2_video_branch=False, forced_tokens=forced)
    with torch.no_grad():
        ids_ref, emb_ref, _ = M.generate(sd, ocfg, ids, images, ctx, 4, forced=forced)
        ssd = {k[len("model.visual_model."):]: v for k, v in sd.items() if k.startswith("model.visual_model.")}
        ref = O.run_framewise(ssd, scfg, sam, emb_ref[:, None, :], (48, 64))
    assert torch.equal(out_ids[0].cpu(), ids_ref)
    for t in range(2):
        assert (torch.from_numpy(segs[0][t][0]) == (ref[t][0, 0] > 0)).float().mean() > 0.98


def test_e2e_against_reference_golden_verification_mode():
    """precise="all" vs the fixture frozen from ONE call of the UNMODIFIED reference's VideoGLaMMForCausalLM.inference (fp32): token ids
    bit-exact, EVERY bool mask pixel identical, video-resolution mask logits within 1e-3 relative (north_star's bar; measured ~1e-4) —
    through both towers, the projectors, Phi-3 prefill + KV-cached decode, the [SEG] projection, Hiera, and the 3-frame 2-object session."""
    import os, sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as MG
    from test_oracle_golden import load_e2e_golden
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    want_ids, want_masks = load_e2e_golden()
    want_logits = torch.from_numpy(np.load(os.path.join(os.path.dirname(__file__), "golden", "e2e_tiny.npz"))["logits"])
    ocfg, scfg = M.make_cfg(**MG.E2E_LMM, seg_token_idx=MG.E2E_SEG), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=11)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**MG.E2E_LMM, seg_token_idx=MG.E2E_SEG), Sam2Config(**O.TINY), precise="all")
    images, ctx, sam, ids = MG.e2e_inputs()
    out_ids, segs, logits = model.inference([images.cuda()], [ctx.cuda()], [sam.cuda()], ids[None].cuda(), None, [MG.E2E_HW], max_new_tokens=MG.E2E_NEW,
                                            use_sam2_video_branch=True, forced_tokens=MG.E2E_FORCED, return_logits=True)
    assert torch.equal(out_ids[0].cpu(), want_ids)
    T, n_obj = want_masks.shape[:2]
    got = torch.stack([torch.stack([torch.from_numpy(segs[0][t][o]) for o in range(n_obj)]) for t in range(T)])
    err = ((logits - want_logits).norm() / want_logits.norm()).item()
    print(f"PARITY e2e verification mode vs reference golden: logits rel {err:.2e}, mask pixels differing {(got != want_masks).sum().item()}/{got.numel()}")
    assert err < 1e-3, err
    assert torch.equal(got, want_masks), (got != want_masks).sum().item()


def test_e2e_against_reference_golden():
    """PRODUCT-DEFAULT arithmetic (bf16 operands in the towers, Phi-3, Hiera and the memory path; fp32-grade mask decoder) vs the same fixture:
    token ids bit-exact; bool masks agree except for pixels whose fp32 logit is within bf16 noise of 0; the logit error is the bf16 error of the
    LMM + Hiera stages (the reference's own bf16 autocast run differs from its fp32 run by the same order), reported here."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as MG
    from test_oracle_golden import _e2e_oracle, load_e2e_golden
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    want_ids, want_masks = load_e2e_golden()
    ocfg, scfg = M.make_cfg(**MG.E2E_LMM, seg_token_idx=MG.E2E_SEG), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=11)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**MG.E2E_LMM, seg_token_idx=MG.E2E_SEG), Sam2Config(**O.TINY))
    images, ctx, sam, ids = MG.e2e_inputs()
    out_ids, segs, got_logits = model.inference([images.cuda()], [ctx.cuda()], [sam.cuda()], ids[None].cuda(), None, [MG.E2E_HW], max_new_tokens=MG.E2E_NEW,
                                                use_sam2_video_branch=True, forced_tokens=MG.E2E_FORCED, return_logits=True)
    assert torch.equal(out_ids[0].cpu(), want_ids)
    _, logits = _e2e_oracle()
    bf16_err = ((got_logits - logits).norm() / logits.norm()).item()
    print(f"PARITY e2e product default (bf16 LMM + Hiera) vs fp32 reference: logits rel {bf16_err:.2e}")
    assert bf16_err < 6e-2, bf16_err
    T, n_obj = want_masks.shape[:2]
    assert sorted(segs[0]) == list(range(T)) and all(sorted(segs[0][t]) == list(range(n_obj)) for t in range(T))
    got = torch.stack([torch.stack([torch.from_numpy(segs[0][t][o]) for o in range(n_obj)]) for t in range(T)])
    agree = (got == want_masks).float().mean().item()
    decided = logits.abs() > 0.05 * logits.abs().max()          # pixels whose sign bf16 noise cannot flip
    assert agree > 0.98 and torch.equal(got[decided], want_masks[decided]), agree


def test_still_image_input_matches_oracle():
    """context_images=None: the still-image branch (arch.py:110-119, 193-203; chat.py:458-487) through the public API vs the oracle."""
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    arch = dict(M.TINY_LMM)
    ocfg, scfg = M.make_cfg(**arch, seg_token_idx=300), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**arch, seg_token_idx=300), Sam2Config(**O.TINY))
    torch.manual_seed(8)
    side = ocfg.clip.image_size
    image, sam = torch.randn(1, 3, side, side), torch.randn(1, 3, 256, 256)
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.tensor([-200]), torch.randint(3, 290, (7,))])
    forced = {1: 300}
    for ctx in (None, [None]):
        out_ids, segs = model.inference([image.cuda()], ctx, [sam.cuda()], ids[None].cuda(), None, [(48, 64)], max_new_tokens=4, use_sam2_video_branch=True,
                                        forced_tokens=forced)
        with torch.no_grad():
            ids_ref, emb_ref, _ = M.generate(sd, ocfg, ids, image, None, 4, forced=forced)
            ssd = {k[len("model.visual_model."):]: v for k, v in sd.items() if k.startswith("model.visual_model.")}
            ref = O.run_video(ssd, scfg, sam, emb_ref[:, None, :], (48, 64))
        assert torch.equal(out_ids[0].cpu(), ids_ref)
        assert out_ids.shape[1] == ids.shape[0] + 4 and sorted(segs[0]) == [0]
        assert (torch.from_numpy(segs[0][0][0]) == (ref[0][0, 0] > 0)).float().mean() > 0.98


def test_from_pretrained_round_trip(tmp_path):
    """An exported HF directory (config.json + pytorch_model.bin + generation_config.json, CLIP in its own HF directory, InternVideo2 .pt) loaded with
    from_pretrained, taken through the calls chat.py makes (resize_token_embeddings, initialize_vision_modules, tower handles, .bfloat16().cuda()),
    gives exactly the outputs of the engine built directly from the same state dict; the multi-id EOS list of generation_config.json stops generation."""
    import json
    from safetensors.torch import save_file
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    arch = dict(M.TINY_LMM)
    seg = 300
    ocfg, scfg = M.make_cfg(**arch, seg_token_idx=seg), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    cp, ip = "model.image_vision_tower.vision_tower.", "model.vision_tower."
    ckpt, clip_dir = tmp_path / "ckpt", tmp_path / "clip"
    ckpt.mkdir(); clip_dir.mkdir()
    save_file({k[len(cp):]: v.contiguous() for k, v in sd.items() if k.startswith(cp)}, str(clip_dir / "model.safetensors"))
    torch.save({k[len(ip):]: v for k, v in sd.items() if k.startswith(ip)}, str(tmp_path / "iv2.pt"))
    torch.save({k: v for k, v in sd.items() if not k.startswith(cp) and not k.startswith(ip)}, str(ckpt / "pytorch_model.bin"))
    p = ocfg.phi
    json.dump(dict(vocab_size=p.vocab, hidden_size=p.hidden, num_hidden_layers=p.layers, num_attention_heads=p.heads, intermediate_size=p.inter,
                   rope_theta=p.rope_theta, rms_norm_eps=p.eps, sliding_window=p.sliding_window, seg_token_idx=seg, out_dim=256, use_sam2=True,
                   mm_vision_tower=str(tmp_path / "iv2.pt"), image_mm_vision_tower=str(clip_dir), eos_token_id=2,
                   vg_arch_overrides={k: v for k, v in arch.items() if k != "phi"},
                   vg_sam2_overrides={k: (list(v) if isinstance(v, tuple) else v) for k, v in O.TINY.items()}), open(ckpt / "config.json", "w"))
    torch.manual_seed(3)
    T = 4
    images, ctx, sam = torch.randn(T, 3, 56, 56), torch.randn(T, 3, 56, 56), torch.randn(2, 3, 256, 256)
    ids = torch.cat([torch.randint(3, 290, (5,)), torch.full((T,), -200), torch.randint(3, 290, (7,))])[None].cuda()
    direct = VideoGLaMMForCausalLM(sd, make_lmm_config(**arch, seg_token_idx=seg), Sam2Config(**O.TINY))
    forced = {1: seg}
    want_ids, want_segs = direct.inference([images.cuda()], [ctx.cuda()], [sam.cuda()], ids, None, [(48, 64)], max_new_tokens=6, use_sam2_video_branch=True,
                                           forced_tokens=forced)
    natural = [int(t) for t in want_ids[0, ids.shape[1]:]]
    json.dump(dict(eos_token_id=[2, natural[3]]), open(ckpt / "generation_config.json", "w"))       # the 4th generated token is an EOS of the generation config
    m = VideoGLaMMForCausalLM.from_pretrained(str(ckpt), low_cpu_mem_usage=False, use_sam2_video_branch=True, torch_dtype=torch.bfloat16)
    assert m.lmm is None                                                           # weights on the host until .cuda()
    m.resize_token_embeddings(p.vocab)
    m.config.eos_token_id = 2                                                      # chat.py:305 overwrites it with the tokenizer's single id
    m.get_model().initialize_vision_modules(m.get_model().config)
    m.get_model().get_vision_tower().to(dtype=torch.bfloat16, device=0)
    m.get_model().get_image_vision_tower().to(dtype=torch.bfloat16, device=0)
    m = m.bfloat16().cuda()
    m.eval()
    got_ids, got_segs = m.inference([images.cuda()], [ctx.cuda()], [sam.cuda()], ids, None, [(48, 64)], max_new_tokens=6, use_sam2_video_branch=True,
                                    forced_tokens=forced)
    n = got_ids.shape[1] - ids.shape[1]
    assert n == 4 and torch.equal(got_ids[0], want_ids[0, : ids.shape[1] + 4])     # stopped at the generation_config EOS id
    import numpy as np
    for t in want_segs[0]:
        assert np.array_equal(got_segs[0][t][0], want_segs[0][t][0])
    with pytest.raises(FileNotFoundError):                                          # unresolvable tower path is named, not a generic missing-keys error
        cfgj = json.load(open(ckpt / "config.json")); cfgj["image_mm_vision_tower"] = "openai/clip-vit-large-patch14-336"
        json.dump(cfgj, open(ckpt / "config.json", "w"))
        VideoGLaMMForCausalLM.from_pretrained(str(ckpt)).initialize_vision_modules(None)


@pytest.mark.parametrize("precise,tol", [("all", 2e-3), ("heads", 6e-2)])
def test_model_forward_validation_and_loss_values(precise, tol):
    """model_forward on the collate_fn dict (VideoGLaMM.py:325-558) through `model(**batch)` (forward dispatch :897-903): validation path
    (inference=True: per-frame masks padded to 4 objects, zero sparse + zero dense padding rows) and the training path's loss values vs
    oracle/model_forward.py (pinned against the reference's own model_forward in tests/test_oracle_vs_reference.py)."""
    from oracle import lmm as M, model_forward as OMF, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    seg = 300
    ocfg, scfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg), Sam2Config(**O.TINY), precise=precise)
    model.ce_loss_weight, model.dice_loss_weight, model.bce_loss_weight = 1.0, 0.5, 2.0
    torch.manual_seed(9)
    T = 4
    images, ctx, sam = torch.randn(T, 3, 56, 56), torch.randn(T, 3, 56, 56), torch.randn(2, 3, 256, 256)
    ids = torch.cat([torch.tensor([1]), torch.randint(3, 290, (5,)), torch.full((T,), -200), torch.randint(3, 290, (4,)), torch.tensor([seg]),
                     torch.randint(3, 290, (3,)), torch.tensor([seg, 2, 0, 0])])[None]
    labels = ids.clone()
    labels[:, :12] = -100
    am = torch.ones_like(ids, dtype=torch.bool)
    am[:, -2:] = False                                                  # right padding
    gt = (torch.rand(2, 2, 48, 64) > 0.5).float()
    batch = dict(images_for_sam=[sam], images=[images], context_images=[ctx], input_ids=ids, labels=labels, attention_masks=am,
                 offset=torch.LongTensor([0, 1]), masks_list=[gt], label_list=[torch.zeros(48, 64)], resize_list=[(256, 256)])
    got_v = model(**batch, inference=True)
    got_t = model(**batch, inference=False)
    with torch.no_grad():
        want_v = OMF.model_forward(sd, ocfg, scfg, [sam], [images], [ctx], ids, labels, am, batch["offset"], [gt], batch["label_list"], inference=True)
        want_t = OMF.model_forward(sd, ocfg, scfg, [sam], [images], [ctx], ids, labels, am, batch["offset"], [gt], batch["label_list"], inference=False,
                                   weights=(1.0, 0.5, 2.0))
    assert len(got_v["pred_masks"][0]) == 2 and tuple(got_v["pred_masks"][0][0].shape) == (4, 48, 64)
    for t in range(2):
        a, b = got_v["pred_masks"][0][t].cpu(), want_v["pred_masks"][0][t]
        e = ((a - b).norm() / b.norm()).item()
        assert e < tol, (t, e)
    for k in ("ce_loss", "mask_bce_loss", "mask_dice_loss", "loss"):
        assert abs(float(got_t[k]) - float(want_t[k])) < tol * max(1.0, abs(float(want_t[k]))), (k, float(got_t[k]), float(want_t[k]))
