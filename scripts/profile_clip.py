"""One full-size config-2 clip between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys, json
import torch
sys.path.insert(0, ".")
import bench
from videoglamm_b200 import ops
from videoglamm_b200.config import Sam2Config, make_lmm_config
from videoglamm_b200.model import VideoGLaMMForCausalLM
from videoglamm_b200.statedict import random_state_dict, videoglamm_param_shapes
lmm_cfg, sam_cfg = make_lmm_config(), Sam2Config()
sd = random_state_dict(videoglamm_param_shapes(lmm_cfg, sam_cfg), "cuda", seed=1234)
model = VideoGLaMMForCausalLM(sd, lmm_cfg, sam_cfg, validate=False); del sd
inp = [t.cuda() for t in bench.synth_inputs(0, pinned=False)]
forced = {st: lmm_cfg.seg_token_idx for st in bench.CFG["seg_steps"]}
def step():
    return model.inference([inp[0]], [inp[1]], [inp[2]], inp[3], None, [bench.VIDEO_HW], max_new_tokens=bench.N_NEW, use_sam2_video_branch=True, forced_tokens=forced)
n_warm = int(os.environ.get("WARM", 2))
for _ in range(n_warm): step()
torch.cuda.synchronize()
if os.environ.get("SAM_STAGES"):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    sam = model.sam
    emb = torch.randn(1, 1, 256, device="cuda")
    T = 8
    tm = {}
    def timeit(name, fn):
        a, b = ev(), ev(); a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        tm[name] = tm.get(name, 0.0) + a.elapsed_time(b); return r
    for rep in range(2):
        tm.clear()
        fb = timeit("hiera+fpn x8", lambda: sam.forward_image(inp[2].float()))
        feats = [{k: v.view(8, -1, v.shape[-1])[i] for k, v in fb.items()} for i in range(8)]
        f0 = feats[0]
        pix0 = ops.add(f0["top"], sam.no_mem_embed)
        r0 = timeit("sam_heads(frame0)", lambda: sam.forward_sam_heads(pix0, f0["s0"], f0["s1"], emb, 1))
        from videoglamm_b200.sam2_engine import FrameMem
        cond = {0: FrameMem(r0["low_res_masks"], r0["obj_ptr"], timeit("encode_memory", lambda: sam.encode_memory(f0["top"], r0["low_res_masks"], True)))}
        non = {}
        for t in range(1, T):
            f = feats[t]
            memory, mem_pos, n_ptr = sam._memory_bank(t, cond, non, T, 1)
            pix = timeit("memory_attention", lambda: sam.memory_attention(f["top"], 1, memory, mem_pos, n_ptr))
            r = timeit("sam_heads", lambda: sam.forward_sam_heads(pix, f["s0"], f["s1"], None, 1))
            non[t] = FrameMem(r["low_res_masks"], r["obj_ptr"], timeit("encode_memory", lambda: sam.encode_memory(f["top"], r["low_res_masks"], False)))
    print({k: round(v, 2) for k, v in tm.items()})
elif os.environ.get("E2E_STAGES"):
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        e = [ev() for _ in range(3)]
        e[0].record()
        out_ids, emb = model.lmm.generate(inp[3], inp[0], inp[1], bench.N_NEW, forced=forced)
        e[1].record()
        frames = model.sam.run_video(inp[2], emb[:, None, :], bench.VIDEO_HW, binary=True)
        e[2].record(); torch.cuda.synchronize()
        print(f"lmm.generate {e[0].elapsed_time(e[1]):.2f} ms   sam.run_video {e[1].elapsed_time(e[2]):.2f} ms")
    import time
    from videoglamm_b200 import ops as _o
    vis = model.lmm.visual_tokens(inp[0], inp[1]); torch.cuda.synchronize()
    for name, fn in (("g_vis", lambda: model.lmm.g_vis((tuple(inp[0].shape), tuple(inp[1].shape)), model.lmm.visual_tokens, images=inp[0], context_images=inp[1])),
                     ("g_prefill", lambda: model.lmm.g_prefill((1696,), model.lmm._prefill_last, x=torch.zeros(1696, 3072, device="cuda"))),
                     ("g_decode", lambda: model.lmm.g_decode((2048,), lambda: model.lmm._decode_step(2048), mutable_state=(model.lmm.dec_tok, model.lmm.dec_pos, model.lmm.dec_nk)))):
        fn(); torch.cuda.synchronize()
        a, b = ev(), ev(); a.record()
        for _ in range(5): fn()
        b.record(); torch.cuda.synchronize()
        print(f"{name}: {a.elapsed_time(b)/5:.3f} ms")
elif os.environ.get("STAGES"):
    # coarse per-stage CUDA-event timing
    ev = lambda: torch.cuda.Event(enable_timing=True)
    dev = "cuda"
    marks = []
    def mark(name):
        e = ev(); e.record(); marks.append((name, e))
    mark("start")
    vis_iv2 = model.lmm.iv2(inp[0].view(2, 4, 3, 224, 224)); mark("iv2")
    vis_clip = model.lmm.clip(inp[1]); mark("clip")
    vis = model.lmm.visual_tokens(inp[0], inp[1]); mark("visual_tokens(all again)")
    x = model.lmm.embed_prompt(inp[3], vis); model.lmm._alloc_cache(x.shape[0] + 20)
    h = model.lmm._llm(x, 0); mark("phi3_prefill")
    e = ops.embedding(model.lmm.embed, torch.tensor([5], device=dev))
    for i in range(12):
        hh = model.lmm._llm(e.clone(), x.shape[0] + i)
        lg = ops.gemm(ops.cast(hh, torch.bfloat16), model.lmm.lm_head.w, out_dtype=torch.float32); ops.argmax(lg)
    mark("phi3_decode_x12")
    feats = [model.sam.forward_image(inp[2][t:t+1]) for t in range(8)]; mark("hiera_x8")
    emb = torch.randn(1, 1, 256, device=dev)
    out = model.sam.run_video(inp[2], emb, bench.VIDEO_HW, binary=True); mark("sam2_run_video(incl hiera)")
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        print(f"{n1:32s} {e0.elapsed_time(e1):9.3f} ms")
else:
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
