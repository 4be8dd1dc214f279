"""The two-stream clip pipeline returns exactly what clip-by-clip `inference` returns."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipeline_matches_sequential_inference():
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    from videoglamm_b200.pipeline import ClipPipeline
    seg = 300
    ocfg, scfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg), Sam2Config(**O.TINY))
    clips = []
    for i in range(5):
        g = torch.Generator().manual_seed(100 + i)
        T = 4
        clips.append(dict(images=torch.randn(T, 3, 56, 56, generator=g), context_images=torch.randn(T, 3, 56, 56, generator=g),
                          images_for_sam=torch.randn(3, 3, 256, 256, generator=g),
                          input_ids=torch.cat([torch.randint(3, 290, (5,), generator=g), torch.full((T,), -200), torch.randint(3, 290, (7,), generator=g)])[None],
                          original_size=(48, 64)))
    forced = {1: seg, 3: seg}
    seq = [model.inference([c["images"]], [c["context_images"]], [c["images_for_sam"]], c["input_ids"].cuda(), None, [c["original_size"]],
                           max_new_tokens=5, use_sam2_video_branch=True, forced_tokens=forced) for c in clips]
    seq = [(ids.cpu().clone(), {t: {o: m.copy() for o, m in fr.items()} for t, fr in segs[0].items()}) for ids, segs in seq]
    out = ClipPipeline(model).run(clips, 5, forced_tokens=forced)
    for (ids_a, seg_a), (ids_b, seg_b, packed) in zip(seq, out):
        assert torch.equal(ids_a, ids_b.cpu())
        back = packed.to_segments()                                   # the device-packed bits (multi-GPU gather payload) carry the same masks
        assert all(np.array_equal(back[t][o], seg_b[t][o]) for t in seg_b for o in seg_b[t])
        assert sorted(seg_a) == sorted(seg_b)
        for t in seg_a:
            for o in seg_a[t]:
                assert np.array_equal(seg_a[t][o], seg_b[t][o])


def test_pipeline_cold_start_with_heterogeneous_shapes():
    """Clips whose prompt length, frame count, object count and video size all differ arrive at a COLD pipeline (no prior warm-up): every stage
    meets uncaptured shape keys while the other stage thread is running (first sight eager, second sight captured under the process-wide capture
    lock in thread-local capture mode), keys recur so that captures and replays interleave, and the LRU bound evicts. Results equal clip-by-clip
    inference on a second, independently built engine."""
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    from videoglamm_b200.pipeline import ClipPipeline
    seg = 300
    ocfg, scfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    mk = lambda: VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg), Sam2Config(**O.TINY))
    model, ref_model = mk(), mk()
    model.sam.g_video.max_entries = 2                                  # force evictions
    variants = [(4, 2, 7, (48, 64), {1: seg}), (8, 3, 11, (40, 56), {1: seg, 3: seg}), (4, 1, 5, (32, 32), {0: seg}), (4, 2, 7, (48, 64), {1: seg})]
    clips, forced_of = [], []
    for rep in range(3):
        for i, (T, Ts, nq, hw, forced) in enumerate(variants):
            g = torch.Generator().manual_seed(1000 + i)
            clips.append(dict(images=torch.randn(T, 3, 56, 56, generator=g), context_images=torch.randn(T, 3, 56, 56, generator=g),
                              images_for_sam=torch.randn(Ts, 3, 256, 256, generator=g),
                              input_ids=torch.cat([torch.randint(3, 290, (5,), generator=g), torch.full((T,), -200), torch.randint(3, 290, (nq,), generator=g)])[None],
                              original_size=hw))
            forced_of.append(forced)
    # the pipeline takes ONE forced-token map per call: group consecutive clips by their map (the shapes inside a group still differ)
    out = []
    pipe = ClipPipeline(model)
    i = 0
    while i < len(clips):
        j = i
        while j < len(clips) and forced_of[j] == forced_of[i]:
            j += 1
        out += pipe.run(clips[i:j], 5, forced_tokens=forced_of[i])
        i = j
    for c, forced, (ids_b, seg_b, _) in zip(clips, forced_of, out):
        ids_a, segs_a = ref_model.inference([c["images"]], [c["context_images"]], [c["images_for_sam"]], c["input_ids"].cuda(), None, [c["original_size"]],
                                            max_new_tokens=5, use_sam2_video_branch=True, forced_tokens=forced)
        assert torch.equal(ids_a.cpu(), ids_b.cpu())
        assert sorted(segs_a[0]) == sorted(seg_b)
        for t in seg_b:
            for o in seg_b[t]:
                assert np.array_equal(segs_a[0][t][o], seg_b[t][o])
    st = model.sam.g_video.stats
    assert st["captures"] >= 2 and st["evictions"] >= 1 and st["replays"] >= 1, st
