"""CPU study (no GPU): which bf16 roundings of the SAM2 decoder / memory path keep the mask logits from meeting 1e-3 vs the fp32 oracle.
Monkeypatches the oracle's functional ops with operand-rounding emulations of the engine's kernels (fp32 accumulate everywhere).

  python scripts/precision_study.py
"""
import os
import sys
import types

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import sam2 as O, weights as Wt  # noqa: E402

bf = lambda t: t.to(torch.bfloat16).float()


def split(t, terms):
    """value seen by a `terms`-way bf16 split (hi [+ lo [+ lo2]]); terms = "h": one fp16 term, "h2": two fp16 terms"""
    if terms == "h":
        return t.half().float()
    if terms == "h2":
        h = t.half().float()
        return h + (t - h).half().float()
    out, r = torch.zeros_like(t), t
    for _ in range(terms):
        h = bf(r)
        out, r = out + h, r - h
    return out


class Emu:
    """F-like namespace: lin = #bf16 terms for GEMM A operands (0 = fp32), w = terms for weights, qk/p/v for attention operands."""

    def __init__(self, lin=1, w=1, qk=1, p=1, v=1, conv=None, small_exact=False):
        self.lin, self.w, self.qk, self.p, self.v = lin, w, qk, p, v
        self.small_exact = small_exact   # decoder attention (<= 16 queries or keys) in exact fp32
        self.conv = lin if conv is None else conv

    def r(self, t, terms):
        return t if terms == 0 else split(t, terms)

    def __getattr__(self, name):
        return getattr(F, name)

    def linear(self, x, w, b=None):
        return F.linear(self.r(x, self.lin), self.r(w, self.w), b)

    def conv2d(self, x, w, b=None, **kw):
        if kw.get("groups", 1) != 1:
            return F.conv2d(x, w, b, **kw)  # depthwise: fp32 SIMT in the engine
        return F.conv2d(self.r(x, self.conv), self.r(w, self.w), b, **kw)

    def conv_transpose2d(self, x, w, b=None, **kw):
        return F.conv_transpose2d(self.r(x, self.conv), self.r(w, self.w), b, **kw)

    def scaled_dot_product_attention(self, q, k, v):
        if self.small_exact and (q.shape[-2] <= 16 or k.shape[-2] <= 16):
            return F.scaled_dot_product_attention(q, k, v)
        q, k, v = self.r(q, self.qk), self.r(k, self.qk), self.r(v, self.v)
        s = (q @ k.transpose(-1, -2)) * q.shape[-1] ** -0.5
        m = s.amax(-1, keepdim=True)
        p = (s - m).exp()
        l = p.sum(-1, keepdim=True)          # row sum from the fp32 p (as the kernels do)
        return (self.r(p, self.p) @ v) / l


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def main():
    torch.manual_seed(0)
    cfg = O.Sam2Cfg(**O.TINY)
    sd = Wt.make_sam2_state_dict(cfg, 1234)
    sd = {k: bf(v) if v.dim() >= 2 else v for k, v in sd.items()}
    torch.manual_seed(5678)
    T = int(os.environ.get("T", 5))
    images, text = torch.randn(T, 3, cfg.image_size, cfg.image_size), torch.randn(2, 1, 256)
    real_F = O.F
    with torch.no_grad():
        feats = [O.forward_image(sd, cfg, images[t:t + 1]) for t in range(T)]   # identical fp32 features for every variant
        orig_fi = O.forward_image
        idx = {"t": 0}

        def fi(sd_, cfg_, img):
            for t in range(T):
                if torch.equal(img, images[t:t + 1]):
                    return feats[t]
            raise KeyError

        O.forward_image = fi
        ref = O.run_video(sd, cfg, images, text, (60, 100))
        variants = {
            "heads x2 + exact dec attn; memattn all bf16; memenc x2": (dict(lin=2, w=1, small_exact=True), dict(lin=1, w=1), None),
            "heads x2 + exact dec attn; memattn all bf16; memenc bf16": (dict(lin=2, w=1, small_exact=True), dict(lin=1, w=1), dict(lin=1, w=1)),
            "heads x2 + exact dec attn; memattn gemm x2 attn bf16; memenc bf16": (dict(lin=2, w=1, small_exact=True), dict(lin=2, w=1), dict(lin=1, w=1)),
            "heads bf16 gemm + exact dec attn; rest x2": (dict(lin=1, w=1, small_exact=True), dict(lin=2, w=1), dict(lin=2, w=1)),
            "heads x2 + bf16 dec attn; rest x2": (dict(lin=2, w=1), dict(lin=2, w=1), dict(lin=2, w=1)),
        }
        orig_ma, orig_me = O.memory_attention, O.memory_encoder
        for name, (kw, kw_ma, kw_me) in variants.items():
            base = Emu(**kw)

            def wrap(fn, kws):
                def f(*a, **k):
                    O.F = Emu(**kws) if kws is not None else base
                    try:
                        return fn(*a, **k)
                    finally:
                        O.F = base
                return f

            O.memory_attention, O.memory_encoder = wrap(orig_ma, kw_ma), wrap(orig_me, kw_me)
            O.F = base
            got = O.run_video(sd, cfg, images, text, (60, 100))
            O.F = real_F
            O.memory_attention, O.memory_encoder = orig_ma, orig_me
            errs = [rel(got[t], ref[t]) for t in range(T)]
            flips = sum(((got[t] > 0) != (ref[t] > 0)).sum().item() for t in range(T))
            print(f"{name:48s} " + " ".join(f"{e:.1e}" for e in errs) + f"  flips {flips}")
        O.forward_image = orig_fi


if __name__ == "__main__":
    main()
