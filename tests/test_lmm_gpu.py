"""LMM engine (CUDA) vs the CPU fp32 oracle: towers, visual tokens, Phi-3 prefill, KV-cached greedy decode vs the oracle's
no-cache decode (token ids must be identical), [SEG] embeddings. bf16-operand tolerances as in test_sam2_gpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import lmm as M, weights as Wt  # noqa: E402  (checker only)
from videoglamm_b200.config import make_lmm_config  # noqa: E402
from videoglamm_b200.lmm_engine import LmmEngine  # noqa: E402

SEG = 300
MID = dict(clip=dict(image_size=56, width=128, layers=3, heads=2, mlp=256), iv2=dict(image_size=56, width=176, depth=3, heads=2, mlp=352),
           phi=dict(vocab=320, hidden=192, layers=2, heads=2, inter=256, sliding_window=None))  # true head dims 64 / 88 / 96


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm()).item()


def build(kw):
    ocfg = M.make_cfg(**kw, seg_token_idx=SEG)
    ecfg = make_lmm_config(**kw, seg_token_idx=SEG)
    sd = Wt.make_lmm_state_dict(ocfg, 1)
    sd = {k: v.bfloat16().float() if v.dim() >= 2 else v for k, v in sd.items()}
    return ocfg, sd, LmmEngine(sd, ecfg)


@pytest.fixture(scope="module", params=["tiny", "mid"])
def ctx(request):
    kw = M.TINY_LMM if request.param == "tiny" else MID
    ocfg, sd, eng = build(kw)
    torch.manual_seed(11)
    T = 8
    images = torch.randn(T, 3, ocfg.iv2.image_size, ocfg.iv2.image_size)
    context = torch.randn(T, 3, ocfg.clip.image_size, ocfg.clip.image_size)
    ids = torch.cat([torch.randint(3, 290, (6,)), torch.full((T,), M.IMAGE_TOKEN_INDEX), torch.randint(3, 290, (9,))])
    return dict(ocfg=ocfg, sd=sd, eng=eng, images=images, context=context, ids=ids)


def test_towers(ctx):
    o, sd, eng = ctx["ocfg"], ctx["sd"], ctx["eng"]
    with torch.no_grad():
        ref_c = M.clip_tower(sd, "model.image_vision_tower.vision_tower", o.clip, ctx["context"])
        ref_v = M.iv2_tower(sd, "model.vision_tower", o.iv2, ctx["images"].reshape(2, 4, *ctx["images"].shape[1:]))
    assert rel(eng.clip(ctx["context"].cuda())[:, 1:], ref_c) < 1e-2
    assert rel(eng.iv2(ctx["images"].cuda().view(2, 4, *ctx["images"].shape[1:])), ref_v) < 1e-2


def test_visual_tokens_and_prefill(ctx):
    o, sd, eng = ctx["ocfg"], ctx["sd"], ctx["eng"]
    with torch.no_grad():
        vis = M.visual_tokens(sd, o, ctx["images"], ctx["context"])
        x = M.splice(sd, o, ctx["ids"], vis)
        hid = M.phi3_forward(sd, o.phi, x)
    evis = eng.visual_tokens(ctx["images"].cuda(), ctx["context"].cuda())
    assert evis.shape == vis.shape and rel(evis, vis) < 1e-2
    ex = eng.embed_prompt(ctx["ids"].cuda(), vis.cuda())       # identical visual tokens -> isolates the LLM
    assert rel(ex, x) < 1e-6
    eng._alloc_cache(ex.shape[0] + 4)
    assert rel(eng._llm(ex.clone(), 0), hid) < 1e-2


def test_sliding_window_prefill(ctx):
    o, sd, eng = ctx["ocfg"], ctx["sd"], ctx["eng"]
    x = sd["model.embed_tokens.weight"][torch.randint(0, 300, (150,))]
    o.phi.sliding_window, eng.cfg.phi.sliding_window = 40, 40
    try:
        with torch.no_grad():
            hid = M.phi3_forward(sd, o.phi, x)
        eng._alloc_cache(200)
        assert rel(eng._llm(x.cuda().clone(), 0), hid) < 1e-2
    finally:
        o.phi.sliding_window, eng.cfg.phi.sliding_window = None, None


def test_generate_token_ids_and_seg_embeddings(ctx):
    """KV-cached decode on the GPU vs the oracle's use_cache=False decode: identical greedy ids; [SEG] forced at steps 2 and 5."""
    o, sd, eng = ctx["ocfg"], ctx["sd"], ctx["eng"]
    forced = {2: SEG, 5: SEG}
    with torch.no_grad():
        ids_ref, emb_ref, _ = M.generate(sd, o, ctx["ids"], ctx["images"], ctx["context"], 8, forced=forced)
    ids, emb = eng.generate(ctx["ids"].cuda()[None], ctx["images"].cuda(), ctx["context"].cuda(), 8, forced=forced)
    assert torch.equal(ids[0].cpu(), ids_ref), (ids[0].cpu()[-8:], ids_ref[-8:])
    assert emb.shape == emb_ref.shape == (2, o.out_dim)
    assert rel(emb, emb_ref) < 2e-2


def test_generate_stops_at_eos(ctx):
    o, sd, eng = ctx["ocfg"], ctx["sd"], ctx["eng"]
    with torch.no_grad():
        ids_ref, _, _ = M.generate(sd, o, ctx["ids"], ctx["images"], ctx["context"], 8, eos_ids=(7,), forced={3: 7})
    ids, emb = eng.generate(ctx["ids"].cuda()[None], ctx["images"].cuda(), ctx["context"].cuda(), 8, eos_ids=(7,), forced={3: 7})
    assert torch.equal(ids[0].cpu(), ids_ref) and ids.shape[1] == ctx["ids"].numel() + 4 and emb.shape[0] == 0
