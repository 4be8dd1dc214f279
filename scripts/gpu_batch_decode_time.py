"""Times one batched decode step (S sequences, full-size Phi-3) against the single-sequence step: ms per step, per token, streamed GB/s."""
import sys, time, torch
sys.path.insert(0, ".")
from videoglamm_b200 import build, ops
from videoglamm_b200.config import make_lmm_config
from videoglamm_b200.lmm_engine import LmmEngine
from videoglamm_b200.statedict import lmm_param_shapes, random_state_dict
build.build()
cfg = make_lmm_config(clip=dict(layers=1), iv2=dict(depth=2))
sd = random_state_dict(lmm_param_shapes(cfg), "cuda", seed=1)
eng = LmmEngine(sd, cfg)
ph = cfg.phi
L = 1704
wbytes = ph.layers * 2 * (4 * ph.hidden * ph.hidden + 3 * ph.inter * ph.hidden) + 2 * ph.vocab * ph.hidden
for S in (1, 2, 4, 8):
    eng._alloc_cache(L + 64, slots=max(S, 1))
    if S == 1:
        eng.dec_pos.fill_(L); eng.dec_nk.fill_(L + 1); eng.dec_tok.fill_(5)
        step = lambda: eng.g_decode((2048,), lambda: eng._decode_step(2048), mutable_state=(eng.dec_tok, eng.dec_pos, eng.dec_nk))
    else:
        eng._batch_state[S] = dict(tok=torch.full((S,), 5, device="cuda", dtype=torch.int64), pos=torch.full((S,), L, device="cuda", dtype=torch.int32),
                                   nk=torch.full((S,), L + 1, device="cuda", dtype=torch.int32), qkv=torch.zeros((S, 3 * ph.hidden), device="cuda", dtype=torch.bfloat16),
                                   ws=ops.decode_workspace(S, ph.heads, ph.head_dim, "cuda"))
        st = eng._batch_state[S]
        step = lambda: eng.g_decode_batch((S, 2048), lambda: eng._decode_step_batch(S, 2048), mutable_state=(st["tok"], st["pos"], st["nk"]))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    kv = S * ph.layers * 2 * L * ph.hidden * 2
    print(f"S={S}: {ms:.3f} ms/step, {ms / S:.3f} ms/token, {(wbytes + kv) / ms / 1e6:.0f} GB/s streamed")
