"""LMM stage alone at C2 shapes: 4 x generate (one clip at a time) against generate_batch of the same 4 clips; per-phase split of the batch."""
import sys, time, torch
sys.path.insert(0, ".")
import bench
from videoglamm_b200 import build, ops
from videoglamm_b200.config import make_lmm_config
from videoglamm_b200.lmm_engine import LmmEngine
from videoglamm_b200.statedict import lmm_param_shapes, random_state_dict
build.build()
cfg = make_lmm_config()
eng = LmmEngine(random_state_dict(lmm_param_shapes(cfg), "cuda", seed=1), cfg)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reqs = []
for j in range(4):
    im, ctx, sam, ids = bench.synth_inputs(j, pinned=False)
    reqs.append((ids.cuda(), im.cuda(), ctx.cuda()))
forced = {5: cfg.seg_token_idx}
def t(fn, n=3):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
a = t(lambda: [eng.generate(*r, N, forced=forced) for r in reqs])
b = t(lambda: eng.generate_batch(reqs, N, forced=forced))
c = t(lambda: eng.generate_batch(reqs, 1, forced=forced))
d = t(lambda: [eng.generate(*r, 1, forced=forced) for r in reqs])
print(f"N_new={N}: 4 x generate {a:.1f} ms; generate_batch(4) {b:.1f} ms; towers+prefill only: batch {c:.1f} ms, single x4 {d:.1f} ms")
print(eng.g_decode_batch.stats, eng.g_prefill.stats, eng.g_vis.stats)
