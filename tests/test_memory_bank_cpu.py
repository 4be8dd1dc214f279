"""Host logic of the SAM2 session on CPU: the engine's memory-bank assembly (`Sam2Engine._memory_bank`: which frames, which temporal slots, which
object pointers, in which order) against the oracle's `_prepare_memory_conditioned_features` (sam2_base.py:495-664), whose bank is captured by
intercepting its memory-attention call. No GPU: the method is pure tensor bookkeeping, so it is exercised on CPU tensors."""
import torch

from oracle import sam2 as O
from videoglamm_b200.config import Sam2Config
from videoglamm_b200.sam2_engine import FrameMem, Sam2Engine


def test_bank_selection_matches_oracle(monkeypatch):
    ocfg, ecfg = O.Sam2Cfg(**O.TINY), Sam2Config(**O.TINY)
    C, md, fs, B, T = ocfg.d_model, ocfg.mem_dim, ocfg.feat_size, 2, 20
    HW = fs * fs
    g = torch.Generator().manual_seed(0)
    tpos = torch.randn(ocfg.num_maskmem, 1, 1, md, generator=g)
    pos_mem = torch.randn(HW, md, generator=g)                         # maskmem_pos_enc of one frame, tokens x channels
    eng = object.__new__(Sam2Engine)                                   # no CUDA: only the attributes _memory_bank reads
    eng.cfg, eng.dev, eng.p_all = ecfg, torch.device("cpu"), False
    eng.mem_pos_slot = (pos_mem[None] + tpos.reshape(ocfg.num_maskmem, 1, md)).contiguous()
    # per-frame outputs shared by both sides
    mem = {t: torch.randn(B, HW, md, generator=g).bfloat16() for t in range(T)}
    ptr = {t: torch.randn(B, C, generator=g) for t in range(T)}
    captured = {}

    def fake_memory_attention(sd, cfg, curr, curr_pos, memory, memory_pos, n_ptr):
        captured.update(memory=memory, pos=memory_pos, n_ptr=n_ptr)
        return torch.zeros(HW, B, C)

    monkeypatch.setattr(O, "memory_attention", fake_memory_attention)
    sd = {"maskmem_tpos_enc": tpos}
    to_map = lambda x: x.float().permute(0, 2, 1).reshape(B, md, fs, fs)         # [B,HW,md] -> [B,md,h,w]
    pos_map = pos_mem.t().reshape(1, md, fs, fs).expand(B, -1, -1, -1)
    for t in (1, 3, 6, 7, 8, 12, 16, 17, 19):
        st = O.VideoState()
        st.cond[0] = O.FrameOut(None, ptr[0], to_map(mem[0]), pos_map)
        cond = {0: FrameMem(None, ptr[0], mem[0])}
        non = {}
        for u in range(1, t):
            st.non_cond[u] = O.FrameOut(None, ptr[u], to_map(mem[u]), pos_map)
            non[u] = FrameMem(None, ptr[u], mem[u])
        O.prepare_memory_conditioned_features(sd, ocfg, t, torch.zeros(B, C, fs, fs), torch.zeros(B, C, fs, fs), st, T, B)
        memory, mem_pos, n_ptr = eng._memory_bank(t, cond, non, T, B)
        assert n_ptr == captured["n_ptr"]
        assert tuple(memory.shape) == (B, captured["memory"].shape[0], md)
        # oracle layout [Nk, B, md]; the pointer tokens are fp32 there and rounded to bf16 in the engine's bank
        assert torch.equal(memory.float().permute(1, 0, 2), captured["memory"].bfloat16().float())
        assert torch.allclose(mem_pos, captured["pos"][:, 0], atol=0, rtol=0)
        assert torch.equal(captured["pos"][:, 0], captured["pos"][:, 1])        # positions are shared by the objects
