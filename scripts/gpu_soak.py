"""Soak: 3 x 64 config-2 clips with varying prompt lengths and object counts through ClipPipeline(decode_batch=8); device memory must plateau."""
import sys, time, torch
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
im, ctx, sam, ids = bench.synth_inputs(0, pinned=True)
raw = bench.synth_raw_frames(0)
clips = []
g = torch.Generator().manual_seed(0)
for i in range(64):
    extra = torch.randint(3, 30000, (int(torch.randint(0, 6, (1,), generator=g)),), generator=g)
    clips.append(dict(images=im, context_images=ctx, images_for_sam=raw, input_ids=torch.cat([ids, extra[None]], 1), original_size=(336, 336)))
pipe = ClipPipeline(model, decode_batch=8)
for rep in range(3):
    forced = {5: lc.seg_token_idx} if rep != 1 else {3: lc.seg_token_idx, 8: lc.seg_token_idx}
    t0 = time.time()
    out = pipe.run(clips, 12, forced_tokens=forced)
    torch.cuda.synchronize()
    n_masks = sum(len(fr) for _, segs, _ in out for fr in segs.values())
    print(f"pass {rep}: {len(out) / (time.time() - t0):.2f} clips/s, masks {n_masks}, allocated {torch.cuda.memory_allocated() / 2**30:.2f} GiB, "
          f"reserved {torch.cuda.memory_reserved() / 2**30:.2f} GiB, prefill graphs {len(model.lmm.g_prefill.entries)} "
          f"(evictions {model.lmm.g_prefill.stats['evictions']}), video graphs {len(model.sam.g_video.entries)}", flush=True)
