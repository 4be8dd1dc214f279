"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules (imported from /root/reference through
oracle/ref_harness.py; HF transformers classes for CLIP / Phi-3) on seeded synthetic weights and inputs.
Run here (CPU, reference present):  python tests/golden/make_golden.py
The fixtures travel to the GPU box; tests/test_oracle_golden.py checks oracle/*.py against them without the reference."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lmm as M, ref_harness as H, sam2 as O, weights as Wt  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SAM2_TINY_OVERRIDES = ["++model.image_size=256", "++model.image_encoder.trunk.embed_dim=32", "++model.image_encoder.trunk.stages=[1,2,3,1]",
                       "++model.image_encoder.trunk.global_att_blocks=[4]", "++model.image_encoder.neck.backbone_channel_list=[256,128,64,32]",
                       "++model.memory_attention.num_layers=2"]


def ref_sam2(tiny=True):
    H.install_shims()
    from model.segment_anything_2.sam2.build_sam import build_sam2_video_predictor
    ref = build_sam2_video_predictor("sam2_hiera_l.yaml", ckpt_path=None, device=None, hydra_overrides_extra=SAM2_TINY_OVERRIDES if tiny else []).float().eval()
    return H.patch_predictor_cpu(ref)


def sam2_video_golden():
    cfg = O.Sam2Cfg(**O.TINY)
    ref = ref_sam2()
    out = {}
    for name, bias, T, nobj in (("present", 4.0, 9, 2), ("absent", -100.0, 2, 1)):
        sd = Wt.make_sam2_state_dict(cfg, 1234, obj_score_bias=bias)
        ref.load_state_dict(sd, strict=True)
        torch.manual_seed(5678)
        images, text = torch.randn(T, 3, cfg.image_size, cfg.image_size), torch.randn(nobj, 1, 256)
        with torch.inference_mode():
            st = ref.init_state_from_tensor(images, 60, 100)
            for o in range(nobj):
                ref.add_new_text(st, 0, o, text[o:o + 1])
            masks = torch.stack([m for _, _, m in ref.propagate_in_video(st)])
        out[name] = masks.numpy()
    np.savez_compressed(os.path.join(OUT, "sam2_tiny_video.npz"), **out)
    print("sam2_tiny_video", {k: v.shape for k, v in out.items()})


def sam2_framewise_golden():
    """The reference's modules called in the order VideoGLaMM.py:205-241 (get_visual_embs_sam2) and :690-745 (framewise decode) call them."""
    cfg = O.Sam2Cfg(**O.TINY)
    ref = ref_sam2()
    ref.load_state_dict(Wt.make_sam2_state_dict(cfg, 1234), strict=True)
    torch.manual_seed(9)
    images, text = torch.randn(2, 3, 256, 256), torch.randn(3, 1, 256)
    outs = []
    with torch.no_grad():
        for t in range(2):
            bo = ref.forward_image(images[t:t + 1])
            _, emb, _, _ = ref._prepare_backbone_features(bo)
            emb[-1] = emb[-1] + ref.no_mem_embed
            sizes = [(64, 64), (32, 32), (16, 16)]
            feats = [f.permute(1, 2, 0).view(1, -1, *s) for f, s in zip(emb[::-1], sizes[::-1])][::-1]
            sp, de = ref.sam_prompt_encoder(points=None, boxes=None, masks=None, text_embeds=text)
            low, _, _, _ = ref.sam_mask_decoder(image_embeddings=feats[-1], image_pe=ref.sam_prompt_encoder.get_dense_pe(), sparse_prompt_embeddings=sp,
                                                dense_prompt_embeddings=de, multimask_output=False, repeat_image=True, high_res_features=feats[:-1])
            outs.append(torch.nn.functional.interpolate(low.float(), (50, 70), mode="bilinear", align_corners=False))
    np.savez_compressed(os.path.join(OUT, "sam2_tiny_framewise.npz"), masks=torch.stack(outs).numpy())
    print("sam2_tiny_framewise", torch.stack(outs).shape)


def towers_golden():
    H.install_shims()
    from model.videogpt_plus.model.internvideo.internvideo2 import PretrainInternVideo2
    from transformers import CLIPVisionConfig, CLIPVisionModel, Phi3Config, Phi3ForCausalLM
    cfg = M.make_cfg(**M.TINY_LMM)
    sd = Wt.make_lmm_state_dict(cfg, 1)
    iv, c, p = cfg.iv2, cfg.clip, cfg.phi
    ref = PretrainInternVideo2(in_chans=3, img_size=iv.image_size, patch_size=iv.patch, embed_dim=iv.width, depth=iv.depth, num_heads=iv.heads,
                               mlp_ratio=iv.mlp / iv.width, clip_embed_dim=32, attn_pool_num_heads=4, qkv_bias=False, drop_path_rate=0.25, init_values=1e-5,
                               qk_normalization=True, use_flash_attn=False, use_fused_rmsnorm=False, use_fused_mlp=False, layerscale_no_force_fp32=False,
                               num_frames=4, tubelet_size=1, sep_pos_embed=False, sep_image_video_pos_embed=True, clip_teacher_embed_dim=32,
                               clip_teacher_final_dim=16, clip_return_layer=1, clip_student_return_interval=1).eval()
    pre = "model.vision_tower.vision_encoder."
    r = ref.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=False)
    assert not r.unexpected_keys and all(k.startswith(("clip_", "final_clip")) for k in r.missing_keys)
    torch.manual_seed(0)
    xv = torch.randn(2, 4, 3, iv.image_size, iv.image_size)
    xc = torch.randn(4, 3, c.image_size, c.image_size)
    ids = torch.randint(0, p.vocab, (1, 37))
    hf = CLIPVisionModel(CLIPVisionConfig(hidden_size=c.width, intermediate_size=c.mlp, num_hidden_layers=c.layers, num_attention_heads=c.heads,
                                          image_size=c.image_size, patch_size=c.patch, hidden_act="quick_gelu")).eval()
    pre = "model.image_vision_tower.vision_tower."
    hf.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)
    hfp = Phi3ForCausalLM(Phi3Config(vocab_size=p.vocab, hidden_size=p.hidden, intermediate_size=p.inter, num_hidden_layers=p.layers,
                                     num_attention_heads=p.heads, num_key_value_heads=p.heads, rms_norm_eps=p.eps, rope_theta=p.rope_theta, sliding_window=None,
                                     max_position_embeddings=4096, original_max_position_embeddings=4096, pad_token_id=0, attn_implementation="eager")).eval()
    hfp.load_state_dict({k: v for k, v in sd.items() if k.startswith("model.layers") or k in ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight")},
                        strict=True)
    with torch.no_grad():
        a = ref(xv.permute(0, 2, 1, 3, 4), None, False, x_vis_return_idx=-2, x_vis_only=True)
        b = hf(xc, output_hidden_states=True).hidden_states[-2][:, 1:]
        lg = hfp(input_ids=ids, use_cache=False).logits[0]
    np.savez_compressed(os.path.join(OUT, "towers_tiny.npz"), iv2=a.numpy(), clip=b.numpy(), phi3_logits=lg.numpy(), ids=ids.numpy())
    print("towers_tiny", a.shape, b.shape, lg.shape)


E2E_SEG = 300
E2E_LMM = dict(clip=dict(layers=2), iv2=dict(depth=3), phi=dict(vocab=320, hidden=128, layers=2, heads=4, inter=256, sliding_window=None))
E2E_FORCED = {1: E2E_SEG, 3: E2E_SEG}
E2E_NEW, E2E_HW = 5, (48, 64)


def e2e_inputs():
    """Seeded inputs of the end-to-end pin (shared by the generator and the tests)."""
    torch.manual_seed(21)
    T, Ts = 4, 3
    images, ctx, sam = torch.randn(T, 3, 224, 224), torch.randn(T, 3, 336, 336), torch.randn(Ts, 3, 256, 256)
    ids = torch.cat([torch.tensor([1]), torch.randint(3, 290, (5,)), torch.full((T,), -200), torch.randint(3, 290, (7,))])
    return images, ctx, sam, ids


def ref_videoglamm(workdir=None, sam2_overrides=None, sam2_cfg_kw=None):
    """The reference's full VideoGLaMMForCausalLM (video branch) on CPU: true IV2 width/grid and CLIP width/336 px (hard-coded in the
    reference's composition), reduced depths, tiny Phi-3 and tiny SAM2. Weights = oracle.weights seed 11."""
    import tempfile
    ocfg, scfg = M.make_cfg(**E2E_LMM, seg_token_idx=E2E_SEG), O.Sam2Cfg(**(sam2_cfg_kw or O.TINY))
    c = ocfg.clip
    os.environ["NUM_FRAMES"] = "4"
    ref = H.build_videoglamm(workdir or tempfile.mkdtemp(), phi=E2E_LMM["phi"], iv2_depth=E2E_LMM["iv2"]["depth"],
                             sam2_overrides=SAM2_TINY_OVERRIDES if sam2_overrides is None else sam2_overrides,
                             clip=dict(width=c.width, mlp=c.mlp, layers=c.layers, heads=c.heads, image_size=c.image_size, patch=c.patch), seg_token_idx=E2E_SEG)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=11)
    r = ref.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys and all(("clip_" in k or "final_clip" in k or "vision_proj" in k) for k in r.missing_keys), r
    H.fix_clip_hidden_states(ref)
    H.force_tokens(ref, E2E_FORCED)
    return ref, sd, ocfg, scfg


def ref_e2e_run(ref, with_logits=False):
    """One call of the reference's inference(). with_logits: also records the fp32 video-resolution mask logits that propagate_in_video yields
    (the tensors the reference thresholds at VideoGLaMM.py:870-875) by wrapping the predictor's generator — the reference code itself is untouched."""
    images, ctx, sam, ids = e2e_inputs()
    seen = []
    vm = ref.model.visual_model
    orig = vm.propagate_in_video

    def recording(*a, **k):
        for fi, obj_ids, m in orig(*a, **k):
            seen.append(m.detach().float().clone())
            yield fi, obj_ids, m

    vm.propagate_in_video = recording
    try:
        with torch.no_grad():
            out_ids, segs = ref.inference([images], [ctx], [sam], ids[None], None, [E2E_HW], max_new_tokens=E2E_NEW, use_sam2_video_branch=True)
    finally:
        vm.propagate_in_video = orig
    masks = np.stack([np.stack([segs[0][t][o] for o in sorted(segs[0][t])]) for t in sorted(segs[0])])  # [T, n_obj, H, W] bool
    if with_logits:
        return out_ids[0].numpy(), masks, torch.stack(seen)[:, :, 0].numpy()                            # logits [T, n_obj, H, W] fp32
    return out_ids[0].numpy(), masks


def e2e_golden():
    """VideoGLaMMForCausalLM.inference(use_sam2_video_branch=True) VideoGLaMM.py:560-597, 770-875 of the UNMODIFIED reference: pins the
    composition (splice, greedy generate with use_cache=False, [SEG] mask shift, text_hidden_fcs gather, per-object add_new_text, propagate,
    threshold). Two forced [SEG] tokens -> two objects."""
    ref, *_ = ref_videoglamm()
    out_ids, masks, logits = ref_e2e_run(ref, with_logits=True)
    assert np.array_equal(logits > 0, masks)
    np.savez_compressed(os.path.join(OUT, "e2e_tiny.npz"), output_ids=out_ids, masks=np.packbits(masks, axis=-1), mask_shape=np.array(masks.shape),
                        logits=logits.astype(np.float32))
    print("e2e_tiny", out_ids.tolist(), masks.shape, masks.mean(), "logits rms", float(np.sqrt((logits ** 2).mean())))


if __name__ == "__main__":
    if "--e2e-only" in sys.argv:
        e2e_golden()
        sys.exit(0)
    sam2_video_golden()
    sam2_framewise_golden()
    towers_golden()
    e2e_golden()
