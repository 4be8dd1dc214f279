"""Timeline of the two pipeline stages (CUDA events) for a stream of C2 clips at decode_batch 1 and 4."""
import sys, torch
sys.path.insert(0, ".")
import bench
from videoglamm_b200 import build
from videoglamm_b200.config import Sam2Config, make_lmm_config
from videoglamm_b200.model import VideoGLaMMForCausalLM
from videoglamm_b200.pipeline import ClipPipeline
from videoglamm_b200.statedict import random_state_dict, videoglamm_param_shapes
build.build()
lc, sc = make_lmm_config(), Sam2Config()
model = VideoGLaMMForCausalLM(random_state_dict(videoglamm_param_shapes(lc, sc), "cuda", seed=1), lc, sc, device="cuda", validate=False)
forced = {5: lc.seg_token_idx}
im, ctx, sam, ids = [t.cuda() for t in bench.synth_inputs(0, pinned=False)]
clips = [dict(images=im, context_images=ctx, images_for_sam=sam, input_ids=ids, original_size=(336, 336)) for _ in range(16)]
for db in (1, 4):
    pipe = ClipPipeline(model, decode_batch=db)
    for _ in range(2):
        pipe.run(clips[:8], 12, forced_tokens=forced)
    torch.cuda.synchronize()
    pipe.trace = []
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    pipe.run(clips, 12, forced_tokens=forced)
    t1 = torch.cuda.Event(enable_timing=True); t1.record(); torch.cuda.synchronize()
    print(f"decode_batch={db}: total {t0.elapsed_time(t1):.1f} ms for {len(clips)} clips")
    marks = {(tag, i): t0.elapsed_time(ev) for tag, i, ev in pipe.trace}
    for (tag, i), t in sorted(marks.items(), key=lambda kv: kv[1]):
        if tag.endswith("0"):
            t_end = marks[(tag[:-1] + "1", i)]
            print(f"  {tag[:-1]:4s} {i:3d}: {t:8.1f} -> {t_end:8.1f}  ({t_end - t:6.1f} ms)")
