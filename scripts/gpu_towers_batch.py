"""Vision towers + projectors: 8 frames (one clip) against 64 frames (8 clips in one call): ms per clip."""
import sys, torch
sys.path.insert(0, ".")
from videoglamm_b200 import build
from videoglamm_b200.config import make_lmm_config
from videoglamm_b200.lmm_engine import LmmEngine
from videoglamm_b200.statedict import lmm_param_shapes, random_state_dict
build.build()
cfg = make_lmm_config(phi=dict(layers=1))
eng = LmmEngine(random_state_dict(lmm_param_shapes(cfg), "cuda", seed=1), cfg)
def t(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for S in (1, 2, 4, 8):
    im, ctx = torch.randn(8 * S, 3, 224, 224, device="cuda"), torch.randn(8 * S, 3, 336, 336, device="cuda")
    ms = t(lambda: eng.g_vis((tuple(im.shape), tuple(ctx.shape)), eng.visual_tokens, images=im, context_images=ctx))
    ms_iv2 = t(lambda: eng.iv2(im.view(2 * S, 4, 3, 224, 224)))
    ms_clip = t(lambda: eng.clip(ctx))
    print(f"S={S}: visual_tokens {ms:.2f} ms = {ms / S:.2f} ms per clip (eager: iv2 {ms_iv2 / S:.2f}, clip {ms_clip / S:.2f} per clip)")
