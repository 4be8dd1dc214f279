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


def test_batched_decode_equals_per_clip_generation():
    """LmmEngine.generate_batch (S sequences of different prompt lengths decode together: skinny GEMMs with M = S, per-sequence KV-cache slot, key count
    and RoPE position in the single-query attention) returns exactly what S calls of generate return; EOS stops one sequence while the others go on.
    Then the same through ClipPipeline(decode_batch=3)."""
    from oracle import lmm as M, sam2 as O, weights as Wt
    from videoglamm_b200.config import Sam2Config, make_lmm_config
    from videoglamm_b200.model import VideoGLaMMForCausalLM
    from videoglamm_b200.pipeline import ClipPipeline
    seg = 300
    ocfg, scfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg), O.Sam2Cfg(**O.TINY)
    sd = Wt.make_videoglamm_state_dict(ocfg, scfg, seed=7)
    model = VideoGLaMMForCausalLM(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg), Sam2Config(**O.TINY))
    reqs, clips = [], []
    for i, (T, nq) in enumerate([(4, 7), (8, 3), (4, 12), (4, 5), (8, 9)]):
        g = torch.Generator().manual_seed(50 + i)
        images, ctx = torch.randn(T, 3, 56, 56, generator=g), torch.randn(T, 3, 56, 56, generator=g)
        ids = torch.cat([torch.randint(3, 290, (5,), generator=g), torch.full((T,), -200), torch.randint(3, 290, (nq,), generator=g)])[None]
        reqs.append((ids.cuda(), images.cuda(), ctx.cuda()))
        clips.append(dict(images=images, context_images=ctx, images_for_sam=torch.randn(2, 3, 256, 256, generator=g), input_ids=ids, original_size=(32, 48)))
    forced = {2: seg, 9: seg}
    single = [model.lmm.generate(*r, 14, forced=forced) for r in reqs]
    eos = (int(single[1][0][0, -6]),)                                     # a token that sequence 1 emits mid-way becomes an EOS id
    single = [model.lmm.generate(*r, 14, eos_ids=eos, forced=forced) for r in reqs]
    assert len({o.shape[1] - r[0].shape[1] for (o, _), r in zip(single, reqs)}) > 1   # the sequences really stop at different lengths
    # (a) fused-RMS skinny GEMVs (the single-sequence arithmetic, M = S rows): free-running greedy decoding is bit-identical to per-clip generation
    model.lmm.batch_tc = False
    for rep in range(3):                                                  # eager first sight, capture, replay
        batch = model.lmm.generate_batch(reqs, 14, eos_ids=eos, forced=forced)
        for (ia, ea), (ib, eb) in zip(single, batch):
            assert torch.equal(ia, ib) and torch.equal(ea, eb)
    # (b) the alternative path (VG_BATCH_TC_MIN): the weight passes go through the TMA-fed tcgen05 GEMM on bf16-normalised operands (the prefill's
    # arithmetic). Near-tied logits may legitimately resolve differently from the fp32-operand GEMV, so every step is teacher-forced to the
    # single-sequence tokens and the hidden states that predicted the [SEG] tokens are compared
    model.lmm.batch_tc, model.lmm.batch_tc_min = True, 2
    model.lmm.g_decode_batch.clear()
    n_new = min(o.shape[1] - r[0].shape[1] for (o, _), r in zip(single, reqs))
    common = {st: int(single[0][0][0, reqs[0][0].shape[1] + st]) for st in range(n_new)}
    common.update(forced)
    ref_forced = [model.lmm.generate(*r, n_new, forced=common) for r in reqs]
    for rep in range(3):
        batch = model.lmm.generate_batch(reqs, n_new, forced=common)
        for (ia, ea), (ib, eb) in zip(ref_forced, batch):
            assert torch.equal(ia, ib)
            assert ea.shape == eb.shape and ((ea - eb).norm() / ea.norm().clamp_min(1e-9)).item() < 2e-2
    model.lmm.batch_tc = False
    model.lmm.g_decode_batch.clear()
    seq = [model.inference([c["images"]], [c["context_images"]], [c["images_for_sam"]], c["input_ids"].cuda(), None, [c["original_size"]],
                           max_new_tokens=14, use_sam2_video_branch=True, forced_tokens=forced) for c in clips]
    seq = [(ids.cpu().clone(), {t: {o: m.copy() for o, m in fr.items()} for t, fr in segs[0].items()}) for ids, segs in seq]
    out = ClipPipeline(model, decode_batch=3).run(clips, 14, forced_tokens=forced)
    for (ids_a, seg_a), (ids_b, seg_b, _) in zip(seq, out):
        assert torch.equal(ids_a, ids_b.cpu())
        for t in seg_a:
            for o in seg_a[t]:
                assert np.array_equal(seg_a[t][o], seg_b[t][o])


def test_batched_towers_and_prefill_for_equally_shaped_clips():
    """Equally shaped clips with equally long prompts: generate_batch runs ONE pass of the vision towers over all frames and ONE batched prefill into
    cache slots 0..S-1 (`_prefill_last_multi`); tokens and [SEG] embeddings equal per-clip generation (eager first sight, capture, replay)."""
    from oracle import lmm as M, weights as Wt
    from videoglamm_b200.config import make_lmm_config
    from videoglamm_b200.lmm_engine import LmmEngine
    seg = 300
    ocfg = M.make_cfg(**M.TINY_LMM, seg_token_idx=seg)
    sd = Wt.make_lmm_state_dict(ocfg, seed=11)
    eng = LmmEngine(sd, make_lmm_config(**M.TINY_LMM, seg_token_idx=seg))
    reqs = []
    for i in range(3):
        g = torch.Generator().manual_seed(80 + i)
        images, ctx = torch.randn(4, 3, 56, 56, generator=g), torch.randn(4, 3, 56, 56, generator=g)
        ids = torch.cat([torch.randint(3, 290, (5,), generator=g), torch.full((4,), -200), torch.randint(3, 290, (9,), generator=g)])[None]
        reqs.append((ids.cuda(), images.cuda(), ctx.cuda()))
    forced = {2: seg, 6: seg}
    single = [eng.generate(*r, 10, forced=forced) for r in reqs]
    for rep in range(3):
        batch = eng.generate_batch(reqs, 10, forced=forced)
        for (ia, ea), (ib, eb) in zip(single, batch):
            assert torch.equal(ia, ib)
            assert ea.shape == eb.shape and ((ea - eb).norm() / ea.norm().clamp_min(1e-9)).item() < 1e-5
    assert any(k[1] == "multi" for k in list(eng.g_prefill.entries) + list(eng.g_prefill.seen) if len(k) == 3)
