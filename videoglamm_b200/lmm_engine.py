"""B200 engine for the LMM side of VideoGLaMM's grounded generation: CLIP ViT-L/14-336 + InternVideo2-1B towers, V-L projectors
and pooling, multimodal splice, Phi-3 prefill + KV-cached greedy decode, [SEG] hidden-state extraction and the L-V projection.

Host orchestration only; activations are produced by libvglamm kernels (videoglamm_b200.ops). Weights come from a state_dict
with the reference's keys. Unlike the reference (use_cache=False: every new token re-runs both vision
towers and the full-sequence LLM, VideoGLaMM.py:790-799) the towers run ONCE and decode is KV-cached; results are defined
against the no-cache forward, which is mathematically identical under causal attention.

Reference call sites cited per method (paths relative to VideoGLaMM/).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch

from . import ops
from .config import ClipConfig, Iv2Config, LmmConfig, PhiConfig
from .graphs import GraphedStage, MAX_ENTRIES as _GRAPH_MAX

BF, F32 = torch.bfloat16, torch.float32
IMAGE_TOKEN_INDEX = -200


@dataclass
class _Lin:
    w: torch.Tensor
    b: torch.Tensor | None
    w3: "ops.W3 | None" = None      # split-bf16 copy of the fp32 weight (verification mode, ops.gemm3)

    def W(self, precise: bool):
        if precise:
            assert self.w3 is not None, "weight was packed without its split-bf16 copy"
            return self.w3
        return self.w


class _Packer:
    def __init__(self, sd: dict, dev: torch.device, precise: bool = False):
        self.sd, self.dev, self.precise = sd, dev, precise

    def mk(self, w: torch.Tensor, b) -> _Lin:
        w = w.contiguous()
        return _Lin(w.to(BF).contiguous(), b, ops.W3(w) if self.precise else None)

    def t(self, k):
        return self.sd[k].detach().to(self.dev, F32)

    def lin(self, p, kpad: int | None = None) -> _Lin:
        w = self.t(p + ".weight")
        w = w.reshape(w.shape[0], -1)
        if kpad is not None and kpad != w.shape[1]:
            wp = torch.zeros(w.shape[0], kpad, device=self.dev)
            wp[:, : w.shape[1]] = w
            w = wp
        b = self.sd.get(p + ".bias")
        return self.mk(w, None if b is None else b.detach().to(self.dev, F32).contiguous())

    def cat_lin(self, ps) -> _Lin:
        ws = [self.t(p + ".weight") for p in ps]
        bs = [self.sd.get(p + ".bias") for p in ps]
        return self.mk(torch.cat([w.reshape(w.shape[0], -1) for w in ws]),
                       torch.cat([b.detach().to(self.dev, F32) for b in bs]).contiguous() if bs[0] is not None else None)

    def vec(self, k):
        return self.t(k).reshape(-1).contiguous()


def _patchify_gemm(images: torch.Tensor, patch: int, w: _Lin, kpad: int, out_rows: torch.Tensor, resid: torch.Tensor, P: bool = False):
    """Non-overlapping patch conv as im2col + GEMM, one image group at a time; out_rows/resid are [rows, width] views."""
    cols, _, _ = ops.im2col(images, patch, patch, 0, kpad, out_dtype=F32 if P else BF)
    ops.gemm(cols, w.W(P), bias=w.b, resid=resid, out=out_rows)


class ClipTower:
    """HF CLIPVisionModel as used by CLIPVisionTower.forward/feature_select model/videogpt_plus/model/multimodal_encoder/clip_encoder.py:34-72."""

    def __init__(self, pk: _Packer, p: str, cfg: ClipConfig):
        self.cfg, v = cfg, p + ".vision_model"
        self.P = pk.precise
        self.kpad = (3 * cfg.patch ** 2 + 7) // 8 * 8
        self.patch = pk.lin(v + ".embeddings.patch_embedding", self.kpad)
        pos = pk.t(v + ".embeddings.position_embedding.weight")
        self.cls_row = (pk.t(v + ".embeddings.class_embedding").reshape(1, -1) + pos[:1]).contiguous()
        self.pos_patches = pos[1:].contiguous()
        self.pre = (pk.vec(v + ".pre_layrnorm.weight"), pk.vec(v + ".pre_layrnorm.bias"))
        self.n_run = cfg.layers + 1 + cfg.select_layer
        self.layers = []
        for i in range(self.n_run):
            l = f"{v}.encoder.layers.{i}"
            self.layers.append(dict(ln1=(pk.vec(l + ".layer_norm1.weight"), pk.vec(l + ".layer_norm1.bias")),
                                    qkv=pk.cat_lin([f"{l}.self_attn.{n}_proj" for n in "qkv"]), out=pk.lin(l + ".self_attn.out_proj"),
                                    ln2=(pk.vec(l + ".layer_norm2.weight"), pk.vec(l + ".layer_norm2.bias")),
                                    fc1=pk.lin(l + ".mlp.fc1"), fc2=pk.lin(l + ".mlp.fc2")))

    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        """images [T,3,S,S] fp32 -> fp32 [T, 1+g^2, width] (hidden_states[select_layer]; the caller drops the CLS row)."""
        cfg = self.cfg
        T, W, N = images.shape[0], cfg.width, 1 + cfg.grid ** 2
        x = torch.empty((T, N, W), device=images.device, dtype=F32)
        x[:, 0] = self.cls_row
        P = self.P
        adt = F32 if P else BF
        for t in range(T):
            _patchify_gemm(images[t:t + 1].contiguous(), cfg.patch, self.patch, self.kpad, x[t, 1:], self.pos_patches, P)
        x = x.view(T * N, W)
        x = ops.norm(x, *self.pre, eps=1e-5, out_dtype=F32)
        hd = W // cfg.heads
        for L in self.layers:
            h = ops.norm(x, *L["ln1"], eps=1e-5, out_dtype=adt)
            qkv = ops.gemm(h, L["qkv"].W(P), bias=L["qkv"].b, out_dtype=adt).view(T, N, 3, cfg.heads, hd)
            o = ops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], scale=hd ** -0.5)
            x = ops.gemm(o.view(T * N, W), L["out"].W(P), bias=L["out"].b, resid=x, out=x)
            h = ops.norm(x, *L["ln2"], eps=1e-5, out_dtype=adt)
            h = ops.gemm(h, L["fc1"].W(P), bias=L["fc1"].b, act="quick_gelu", out_dtype=adt)
            x = ops.gemm(h, L["fc2"].W(P), bias=L["fc2"].b, resid=x, out=x)
        return x.view(T, N, W)


class Iv2Tower:
    """PretrainInternVideo2.forward(x_vis_only=True, x_vis_return_idx=-2) model/videogpt_plus/model/internvideo/internvideo2.py:585-651
    with Block :306-310, Attention._naive_attn :190-209 (q/k RMSNorm over the full width), LayerScale :148-163."""

    def __init__(self, pk: _Packer, p: str, cfg: Iv2Config):
        self.cfg, e = cfg, p + ".vision_encoder"
        self.P = pk.precise
        self.kpad = (3 * cfg.patch ** 2 + 7) // 8 * 8
        self.patch = pk.lin(e + ".patch_embed.proj", self.kpad)
        pos = pk.t(e + ".pos_embed")[0]
        self.cls_row = (pk.t(e + ".cls_token").reshape(1, -1) + pos[:1]).contiguous()
        self.pos_patches = pos[1:].contiguous()
        self.blocks = []
        for i in range(cfg.depth - 1):  # the reference breaks after block depth-2
            b = f"{e}.blocks.{i}"
            self.blocks.append(dict(n1=pk.vec(b + ".norm1.weight"), qkv=pk.lin(b + ".attn.qkv"), qn=pk.vec(b + ".attn.q_norm.weight"),
                                    kn=pk.vec(b + ".attn.k_norm.weight"), proj=pk.lin(b + ".attn.proj"), ls1=pk.vec(b + ".ls1.gamma"),
                                    n2=pk.vec(b + ".norm2.weight"), fc1=pk.lin(b + ".mlp.fc1"), fc2=pk.lin(b + ".mlp.fc2"), ls2=pk.vec(b + ".ls2.gamma")))

    def __call__(self, chunks: torch.Tensor) -> torch.Tensor:
        """chunks [n,4,3,S,S] fp32 -> fp32 [n, 1+4*g^2, width]"""
        cfg = self.cfg
        n, W = chunks.shape[0], cfg.width
        N = 1 + cfg.frames * cfg.grid ** 2
        x = torch.empty((n, N, W), device=chunks.device, dtype=F32)
        x[:, 0] = self.cls_row
        P = self.P
        adt = F32 if P else BF
        for c in range(n):  # Conv3d(1 x p x p) == per-frame 2D patch conv; the 4 frames of a chunk are consecutive token blocks
            _patchify_gemm(chunks[c].contiguous(), cfg.patch, self.patch, self.kpad, x[c, 1:], self.pos_patches, P)
        x = x.view(n * N, W)
        hd = W // cfg.heads
        for B in self.blocks:
            h = ops.norm(x, B["n1"], eps=1e-6, rms=True, out_dtype=adt)
            qkv = ops.gemm(h, B["qkv"].W(P), out_dtype=adt)
            ops.norm(qkv[:, :W], B["qn"], eps=1e-6, rms=True, out=qkv[:, :W])
            ops.norm(qkv[:, W:2 * W], B["kn"], eps=1e-6, rms=True, out=qkv[:, W:2 * W])
            q5 = qkv.view(n, N, 3, cfg.heads, hd)
            o = ops.attention(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], scale=hd ** -0.5)
            x = ops.gemm(o.view(n * N, W), B["proj"].W(P), bias=B["proj"].b, col_scale=B["ls1"], resid=x, out=x)
            h = ops.norm(x, B["n2"], eps=1e-6, rms=True, out_dtype=adt)
            h = ops.gemm(h, B["fc1"].W(P), bias=B["fc1"].b, act="gelu", out_dtype=adt)
            x = ops.gemm(h, B["fc2"].W(P), bias=B["fc2"].b, col_scale=B["ls2"], resid=x, out=x)
        return x.view(n, N, W)


class LmmEngine:
    def __init__(self, sd: dict, cfg: LmmConfig, device: str = "cuda", precise: bool = False):
        """precise=True: verification mode — every contraction is a split-bf16 (fp32-grade) GEMM, every attention exact fp32, the KV cache fp32, no
        decode fusions or graphs (parity with the fp32 reference at ~1e-5; ~3x the GEMM work). Default: bf16 operands, fp32 accumulation."""
        self.cfg, self.dev = cfg, torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("LmmEngine needs a CUDA device (there is no CPU path)")
        self.P = precise
        pk = _Packer(sd, self.dev, precise)
        self.clip = ClipTower(pk, "model.image_vision_tower.vision_tower", cfg.clip)
        self.iv2 = Iv2Tower(pk, "model.vision_tower", cfg.iv2)
        self.vid_proj = [pk.lin(f"model.mm_projector.{2 * j}") for j in range(cfg.proj_depth)]
        self.ctx_proj = [pk.lin(f"model.image_mm_projector.{2 * j}") for j in range(cfg.proj_depth)]
        self.fc = [pk.lin("model.text_hidden_fcs.0.0"), pk.lin("model.text_hidden_fcs.0.2")]
        ph = cfg.phi
        self.embed = pk.t("model.embed_tokens.weight").to(BF).contiguous()
        self.embed32 = pk.t("model.embed_tokens.weight").contiguous() if precise else None   # verification mode: un-rounded rows (a gather, no arithmetic)
        self.final_norm = pk.vec("model.norm.weight")
        # RMSNorm weights are folded into the following projection (W' = W diag(g), in fp32 before the bf16 rounding): the prefill norm
        # kernel then runs weight-less and the decode GEMV fuses the whole normalisation (x * rsqrt(mean x^2 + eps)) into its A loads.
        self.lm_head = pk.mk(pk.t("lm_head.weight") * self.final_norm[None], None)
        self.ones = torch.ones(ph.hidden, device=self.dev)
        self.layers = []
        I = ph.inter
        for l in range(ph.layers):
            p = f"model.layers.{l}"
            gu = pk.t(p + ".mlp.gate_up_proj.weight")
            # interleave [16 gate | 16 up] row blocks so the GEMM epilogue computes silu(gate)*up in registers (include/vglamm.h, glu)
            gu = torch.stack([gu[:I].view(I // 16, 16, -1), gu[I:].view(I // 16, 16, -1)], dim=1).reshape(2 * I, -1)
            n1, n2 = pk.vec(p + ".input_layernorm.weight"), pk.vec(p + ".post_attention_layernorm.weight")
            qkv = pk.mk(pk.t(p + ".self_attn.qkv_proj.weight") * n1[None], None)
            self.layers.append(dict(qkv=qkv, o=pk.lin(p + ".self_attn.o_proj"), gu=pk.mk(gu * n2[None], None), down=pk.lin(p + ".mlp.down_proj")))
        self.kv = None  # per-layer packed qkv cache [max_len, 3*hidden] bf16 (q rows are scratch)
        # graph-captured stages + the device-side decode state they read (token, position, key count, qkv scratch row)
        # prefill graphs are keyed by (prompt length, KV-cache slot): a decode batch of 8 alone needs 8 live entries
        self.g_vis, self.g_prefill, self.g_decode = GraphedStage("visual_tokens"), GraphedStage("prefill", max_entries=2 * _GRAPH_MAX + 8), GraphedStage("decode")
        self.g_decode_batch = GraphedStage("decode_batch")
        # batched decode: S <= 8 sequences stream the weights through the multi-row TMA-ring kernel (csrc/gemv_ring.cu: mma.sync on fp32 activations
        # split hi + lo, fused RMSNorm: the single-sequence step's arithmetic up to summation order). S >= batch_tc_min (none by default) goes through
        # the tcgen05 GEMM with one-wave tile counts instead (measured 4.6 / 5.3 ms per step at S = 4 / 8 against 3.1 / 4.1 ms for the ring)
        self.batch_tc = os.environ.get("VG_BATCH_DECODE_TC", "1") != "0" and ph.hidden % 64 == 0
        self.batch_tc_min = int(os.environ.get("VG_BATCH_TC_MIN", "9"))
        self.kv_slots = None
        self._batch_state = {}
        self.dec_tok = torch.zeros(1, device=self.dev, dtype=torch.int64)
        self.dec_pos = torch.zeros(1, device=self.dev, dtype=torch.int32)
        self.dec_nk = torch.zeros(1, device=self.dev, dtype=torch.int32)
        self.dec_qkv = torch.zeros((1, 3 * ph.hidden), device=self.dev, dtype=BF)
        self.dec_ws = ops.decode_workspace(1, ph.heads, ph.head_dim, self.dev)

    # ------------------------------------------------------------------------------------------------ vision side
    def _project(self, x3: torch.Tensor, layers) -> torch.Tensor:
        """mlpNx_gelu projector (multimodal_projector/builder.py:38-46) on a [B, n, width] fp32 view (CLS already sliced off)."""
        B, n, W = x3.shape
        P = self.P
        a = torch.empty((B, n, W), device=self.dev, dtype=F32 if P else BF)
        for b in range(B):
            ops.add(x3[b], None, out=a[b])
        h = a.view(B * n, W)
        for j, L in enumerate(layers):
            last = j == len(layers) - 1
            h = ops.gemm(h, L.W(P), bias=L.b, act=None if last else "gelu", out_dtype=F32 if (last or P) else BF)
        return h

    def visual_tokens(self, images: torch.Tensor, context_images: torch.Tensor, clips: int = 1) -> torch.Tensor:
        """encode_videos model/videogpt_plus/model/arch.py:121-151 + project(input_type="video") :164-191:
        -> fp32 [T*(gc/2)^2 + T*(gv/2)^2, hidden], context block first. clips > 1: the frames of that many equally shaped clips stacked along
        dim 0 (one pass of the towers with 8x the rows fills the GEMM waves better: 11.4 -> 9.3 ms per clip at 8 clips) -> [clips, tokens, hidden]."""
        cfg = self.cfg
        T = images.shape[0]
        assert T % cfg.iv2.frames == 0, "encoder frames must be a multiple of CHUNK_SIZE=4 (arch.py:133)"
        vid = self.iv2(images.view(T // cfg.iv2.frames, cfg.iv2.frames, *images.shape[1:]))[:, 1:]
        ctx = self.clip(context_images)[:, 1:]
        gv, gc = cfg.iv2.grid, cfg.clip.grid
        v = self._project(vid, self.vid_proj)                       # [T*gv^2, hidden] (chunk-major == frame-major)
        v = ops.pool2x2(v, T, gv, gv, is_max=False)
        c = self._project(ctx, self.ctx_proj)
        c = ops.pool2x2(c, T, gc, gc, is_max=False)
        if clips > 1:                                               # frames of `clips` clips stacked: [clips, tokens per clip, hidden]
            return torch.cat([c.view(clips, -1, c.shape[1]), v.view(clips, -1, v.shape[1])], 1)
        return torch.cat([c, v], 0)

    def image_tokens(self, image: torch.Tensor) -> torch.Tensor:
        """Still-image input (context_images=None): encode_images arch.py:110-119 (the CLIP tower on the 336 px image, patch features) +
        project(input_type="image") :193-203 (image_mm_projector, NO pooling) -> fp32 [n*576, hidden]."""
        ctx = self.clip(image)[:, 1:]
        return self._project(ctx, self.ctx_proj)

    def visual_tokens_any(self, images: torch.Tensor, context_images: torch.Tensor | None) -> torch.Tensor:
        return self.image_tokens(images) if context_images is None else self.visual_tokens(images, context_images)

    # ------------------------------------------------------------------------------------------------ Phi-3
    def _alloc_cache(self, max_len: int, slots: int = 1):
        """Per layer one [slots, max_len, 3*hidden] buffer (slot = one sequence of a decode batch); self.kv[l] is slot 0 (single-sequence paths)."""
        ph = self.cfg.phi
        if self.kv is None or self.kv[0].shape[0] < max_len or self.kv_slots[0].shape[0] < slots:
            n = max(max_len, ph.max_len)                 # allocate once at the model's context length: graphs keep these pointers
            ns = max(slots, 1 if self.kv is None else self.kv_slots[0].shape[0])
            self.kv_slots = [torch.zeros((ns, n, 3 * ph.hidden), device=self.dev, dtype=F32 if self.P else BF) for _ in range(ph.layers)]
            self.kv = [t[0] for t in self.kv_slots]
            for g in (self.g_prefill, self.g_decode, self.g_decode_batch):
                g.clear()

    PREFILL_BUCKET = 64   # prompts are zero-padded to a multiple of this many rows: graph keys (and batched-prefill groups) per bucket, not per length

    def _bucket(self, n: int) -> int:
        return -(-n // self.PREFILL_BUCKET) * self.PREFILL_BUCKET

    @staticmethod
    def _pad_rows(x: torch.Tensor, n: int) -> torch.Tensor:
        return x if x.shape[0] == n else torch.cat([x, x.new_zeros((n - x.shape[0], x.shape[1]))], 0)

    def _prefill_last(self, x: torch.Tensor, slot: int = 0, last: torch.Tensor | None = None) -> torch.Tensor:
        """Prefill rows [0, n); returns the un-normalised residual of the LAST real row only ([1, hidden] fp32) — all that decoding needs. `last`
        (int64 [1], device): index of that row when x carries zero rows behind the prompt (bucketed length: under causal attention the pad rows
        change nothing in front of them, and the decode steps overwrite their cache rows)."""
        xr = self._llm(x, 0, raw=True, slot=slot)
        return xr[-1:].contiguous() if last is None else xr.index_select(0, last)

    @staticmethod
    def _ksplit(K: int, want: int) -> int:
        """K slices for the skinny o_proj / down_proj of the batched decode: `want` if its slices are multiples of 64 and >= 512, else fewer."""
        for p in range(want, 1, -1):
            if K % p == 0 and (K // p) % 64 == 0 and K // p >= 512:
                return p
        return 1

    def _decode_step_batch(self, S: int, kv_bound: int):
        """One KV-cached decode step for S sequences at once (C3: the clips queued on a GPU decode together): every weight matrix is streamed ONCE for
        the S tokens (skinny GEMMs with M = S), each sequence attends over its own cache slot with its own key count (dec_nk[s]), RoPE and the cache
        append fused into the attention kernel. Device-side state per sequence: st["tok"], st["pos"], st["nk"]."""
        ph = self.cfg.phi
        H, hd, D = ph.heads, ph.head_dim, ph.hidden
        st = self._batch_state[S]
        x = ops.embedding(self.embed, st["tok"])                                             # [S, hidden] fp32
        # M = S rows: the weight-streaming work goes through the TMA-fed tcgen05 GEMM (force_block_n=64 -> N/64 tiles, every SM streams a ring of
        # weight tiles; the A tile is 4 real rows + TMA zero fill), o_proj / down_proj as K-split batched GEMMs so that they too fill the machine
        for l, L in enumerate(self.layers):
            cache = self.kv_slots[l]
            if self.batch_tc and S >= self.batch_tc_min:
                h = ops.norm(x, self.ones, eps=ph.eps, rms=True)
                ops.gemm(h, L["qkv"].w, out=st["qkv"], force_block_n=64)
                q = st["qkv"][:, :D].unflatten(1, (H, hd))[:, None]
                k = cache[:S, :kv_bound, D:2 * D].unflatten(2, (H, hd))
                v = cache[:S, :kv_bound, 2 * D:].unflatten(2, (H, hd))
                o = ops.attention(q, k, v, scale=hd ** -0.5, causal=True, window=ph.sliding_window or 0, nk_dev=st["nk"], decode_ws=st["ws"],
                                  fused_rope_theta=ph.rope_theta)
                # tile counts are chosen to fill ONE wave of the 148 SMs (each tile streams its slice of the weights): Phi-3-mini: qkv 9216/64 = 144,
                # o 3 K-slices x 3072/64 = 144, gate_up 16384/128 = 128, down 4 K-slices x 3072/96 = 128
                x = ops.gemm_ksplit(o.view(S, D), L["o"].w, self._ksplit(D, 3), resid=x, out=x)
                h = ops.norm(x, self.ones, eps=ph.eps, rms=True)
                h = ops.gemm(h, L["gu"].w, glu=True, act="silu", force_block_n=128 if 2 * ph.inter >= 128 * 96 else 64)
                x = ops.gemm_ksplit(h, L["down"].w, self._ksplit(ph.inter, 4), resid=x, out=x, block_n=96 if D % 96 == 0 else 64)
                continue
            ops.gemm(x, L["qkv"].w, out=st["qkv"], a_rms_eps=ph.eps)
            q = st["qkv"][:, :D].unflatten(1, (H, hd))[:, None]                              # [S, 1, H, hd], batch stride 3*hidden
            k = cache[:S, :kv_bound, D:2 * D].unflatten(2, (H, hd))
            v = cache[:S, :kv_bound, 2 * D:].unflatten(2, (H, hd))
            o = ops.attention(q, k, v, scale=hd ** -0.5, causal=True, window=ph.sliding_window or 0, nk_dev=st["nk"], decode_ws=st["ws"],
                              fused_rope_theta=ph.rope_theta)
            x = ops.gemm(o.view(S, D), L["o"].w, resid=x, out=x)
            h = ops.gemm(x, L["gu"].w, glu=True, act="silu", a_rms_eps=ph.eps)
            x = ops.gemm(h, L["down"].w, resid=x, out=x)
        nxt = self._logits_argmax(x)
        st["tok"].copy_(nxt)
        st["pos"].add_(1)
        st["nk"].add_(1)
        return x, nxt

    def _decode_step(self, kv_bound: int):
        """One KV-cached decode step whose launch sequence does not depend on the position: token id, position and key count live in
        device memory (dec_tok / dec_pos / dec_nk), the new K/V row is scattered into the cache by a device-indexed copy, and the
        single-query attention reads its key count from dec_nk. Captured once and replayed for every generated token."""
        ph = self.cfg.phi
        H, hd, D = ph.heads, ph.head_dim, ph.hidden
        x = ops.embedding(self.embed, self.dec_tok)
        for l, L in enumerate(self.layers):
            cache = self.kv[l]
            ops.gemm(x, L["qkv"].w, out=self.dec_qkv, a_rms_eps=ph.eps)                       # input RMSNorm fused
            q = self.dec_qkv[:, :D].unflatten(1, (H, hd))[None]
            k = cache[:kv_bound, D:2 * D].unflatten(1, (H, hd))[None]
            v = cache[:kv_bound, 2 * D:].unflatten(1, (H, hd))[None]
            # RoPE(q, k_new) and the KV-cache append happen inside the attention kernel (position = dec_nk - 1 = dec_pos)
            o = ops.attention(q, k, v, scale=hd ** -0.5, causal=True, window=ph.sliding_window or 0, nk_dev=self.dec_nk, decode_ws=self.dec_ws,
                              fused_rope_theta=ph.rope_theta)
            x = ops.gemm(o.view(1, D), L["o"].w, resid=x, out=x)
            h = ops.gemm(x, L["gu"].w, glu=True, act="silu", a_rms_eps=ph.eps)                 # post-attention RMSNorm fused
            x = ops.gemm(h, L["down"].w, resid=x, out=x)
        nxt = self._logits_argmax(x)
        self.dec_tok.copy_(nxt)                            # greedy continuation; the host overwrites dec_tok when a token is forced
        self.dec_pos.add_(1)
        self.dec_nk.add_(1)
        return x, nxt

    def logits_all(self, x_resid: torch.Tensor) -> torch.Tensor:
        """Full-sequence logits from the un-normalised residual rows [n, hidden]: RMSNorm without weight, then the lm_head that carries the folded
        final-norm weight (the same arithmetic as the decode path) -> fp32 [n, vocab]."""
        P = self.P
        h = ops.norm(x_resid, self.ones, eps=self.cfg.phi.eps, rms=True, out_dtype=F32 if P else BF)
        return ops.gemm(h, self.lm_head.W(P), out_dtype=F32)

    def _llm(self, x: torch.Tensor, pos0: int, last_only: bool = False, raw: bool = False, slot: int = 0) -> torch.Tensor:
        """HF Phi3Model.forward (called at language_model/phi3.py:72-81) on rows [pos0, pos0+n) given a cache holding rows [0,pos0).
        x fp32 [n, hidden] (modified in place as the residual stream) -> final-norm hidden fp32 [n, hidden]."""
        ph = self.cfg.phi
        n, H, hd, D = x.shape[0], ph.heads, ph.head_dim, ph.hidden
        pos = torch.arange(pos0, pos0 + n, device=self.dev, dtype=torch.int32)
        end = pos0 + n
        P = self.P
        adt = F32 if P else BF
        for l, L in enumerate(self.layers):
            cache = self.kv_slots[l][slot]
            h = ops.norm(x, self.ones, eps=ph.eps, rms=True, out_dtype=adt)
            rows = cache[pos0:end]
            ops.gemm(h, L["qkv"].W(P), out=rows)
            ops.rope_neox_(rows, pos, 2 * H, hd, ph.rope_theta)
            q = rows[:, :D].unflatten(1, (H, hd))[None]
            k = cache[:end, D:2 * D].unflatten(1, (H, hd))[None]
            v = cache[:end, 2 * D:].unflatten(1, (H, hd))[None]
            o = ops.attention(q, k, v, scale=hd ** -0.5, causal=True, window=ph.sliding_window or 0)
            x = ops.gemm(o.view(n, D), L["o"].W(P), resid=x, out=x)
            h = ops.norm(x, self.ones, eps=ph.eps, rms=True, out_dtype=adt)
            h = ops.gemm(h, L["gu"].W(P), glu=True, act="silu", out_dtype=adt)
            x = ops.gemm(h, L["down"].W(P), resid=x, out=x)
        if raw:
            return x                                        # un-normalised residual rows (logits_all / final norm by the caller)
        if last_only:
            return x[-1:].contiguous()                      # un-normalised residual row: the consumers fuse the final norm
        return ops.norm(x, self.final_norm, eps=ph.eps, rms=True, out_dtype=F32)

    def _prefill_last_multi(self, x: torch.Tensor, last: torch.Tensor) -> torch.Tensor:
        """Prefill of S prompts of one length bucket at once into cache slots 0..S-1: x fp32 [S, n, hidden] (zero rows behind the shorter prompts),
        last int64 [S] = index of the last real row of every sequence -> its un-normalised residual row [S, hidden]. Everything runs on the flattened [S*n, hidden] rows (the N = 3072 projections of ONE 1696-token prompt are 224
        tiles = 1.5 waves of 148 SMs; eight prompts are 12 waves) except the qkv projection, a batched GEMM that writes the rows of every
        sequence straight into its cache slot, RoPE (per slot) and attention (batch dimension = slot). Same arithmetic per element as `_llm`."""
        ph = self.cfg.phi
        S, n, _ = x.shape
        H, hd, D = ph.heads, ph.head_dim, ph.hidden
        pos = torch.arange(0, n, device=self.dev, dtype=torch.int32)
        x = x.reshape(S * n, D)
        for l, L in enumerate(self.layers):
            cache = self.kv_slots[l]
            h = ops.norm(x, self.ones, eps=ph.eps, rms=True, out_dtype=BF)
            rows = cache[:S, :n]                                                     # [S, n, 3*hidden], batch stride = one cache slot
            ops.gemm(h.view(S, n, D), L["qkv"].w, out=rows)
            for s in range(S):
                ops.rope_neox_(rows[s], pos, 2 * H, hd, ph.rope_theta)
            q = rows[:, :, :D].unflatten(2, (H, hd))
            k = rows[:, :, D:2 * D].unflatten(2, (H, hd))
            v = rows[:, :, 2 * D:].unflatten(2, (H, hd))
            o = ops.attention(q, k, v, scale=hd ** -0.5, causal=True, window=ph.sliding_window or 0)
            x = ops.gemm(o.view(S * n, D), L["o"].w, resid=x, out=x)
            h = ops.norm(x, self.ones, eps=ph.eps, rms=True, out_dtype=BF)
            h = ops.gemm(h, L["gu"].w, glu=True, act="silu", out_dtype=BF)
            x = ops.gemm(h, L["down"].w, resid=x, out=x)
        return x.index_select(0, last + torch.arange(S, device=self.dev, dtype=torch.int64) * n)   # row last[s] of sequence s (pad rows follow it)

    def _logits_argmax(self, x_last: torch.Tensor) -> torch.Tensor:
        """final RMSNorm (fused, weight folded into lm_head) -> lm_head -> argmax on one residual row [1, hidden] fp32."""
        if self.P:
            h = ops.norm(x_last, self.ones, eps=self.cfg.phi.eps, rms=True, out_dtype=F32)
            return ops.argmax(ops.gemm(h, self.lm_head.w3, out_dtype=F32))
        if x_last.shape[0] >= self.batch_tc_min and self.batch_tc:      # batched decode: stream lm_head through the TMA-fed tensor-core GEMM (see _decode_step_batch)
            h = ops.norm(x_last, self.ones, eps=self.cfg.phi.eps, rms=True)
            return ops.argmax(ops.gemm(h, self.lm_head.w, out_dtype=F32, force_block_n=64))
        return ops.argmax(ops.gemm(x_last, self.lm_head.w, out_dtype=F32, a_rms_eps=self.cfg.phi.eps))

    def _embed(self, ids: torch.Tensor) -> torch.Tensor:
        return self.embed32.index_select(0, ids) if self.P else ops.embedding(self.embed, ids)

    def embed_prompt(self, input_ids: torch.Tensor, vis: torch.Tensor) -> torch.Tensor:
        """prepare_inputs_labels_for_multimodal arch.py:283-379 (video case): embed(ids before the -200 run) + visual + embed(rest)."""
        ids_host = input_ids.reshape(-1).cpu()            # one small D2H per clip; the prompt structure is host control flow
        idx = torch.where(ids_host == IMAGE_TOKEN_INDEX)[0]
        start, end = int(idx[0]), int(idx[-1])
        text = torch.cat([ids_host[:start], ids_host[end + 1:]]).to(self.dev)
        emb = self._embed(text) if text.numel() else torch.zeros((0, self.embed.shape[1]), device=self.dev)
        return torch.cat([emb[:start], vis, emb[start:]], 0).contiguous()

    @torch.no_grad()
    def generate_batch(self, requests: list[tuple], max_new_tokens: int, eos_ids=(), forced: dict | None = None) -> list[tuple]:
        """`generate` for S = len(requests) <= 8 independent clips at once: requests = [(input_ids, images, context_images)]. Towers and prefill run
        per clip (tensor-bound, each fills the GPU alone); the HBM-bound decode loop runs ONCE for all S sequences (`_decode_step_batch`), with one
        host synchronisation per step for the whole batch. Same arithmetic as S calls of `generate` up to the fp32 summation order of the
        weight-streaming kernels (bit-identical where both take the generic skinny kernel): [(output_ids, pred_embeddings)]."""
        cfg, S = self.cfg, len(requests)
        assert 1 <= S <= 8 and not self.P
        if S == 1:
            return [self.generate(*requests[0], max_new_tokens, eos_ids=eos_ids, forced=forced)]
        ph = cfg.phi
        xs = []
        same_vis = all(r[2] is not None and r[1].shape == requests[0][1].shape and r[2].shape == requests[0][2].shape for r in requests)
        if same_vis:                                            # equally shaped clips: ONE pass of the towers over all their frames
            im_all, ctx_all = torch.cat([r[1] for r in requests], 0), torch.cat([r[2] for r in requests], 0)
            vis_all = self.g_vis((tuple(im_all.shape), tuple(ctx_all.shape), S), lambda images, context_images: self.visual_tokens(images, context_images, clips=S),
                                 images=im_all, context_images=ctx_all)
            xs = [self.embed_prompt(ids, vis_all[s]) for s, (ids, _, _) in enumerate(requests)]
        else:
            for ids, images, ctx in requests:
                if ctx is None:
                    vis = self.g_vis((tuple(images.shape), None), self.image_tokens, image=images)
                else:
                    vis = self.g_vis((tuple(images.shape), tuple(ctx.shape)), self.visual_tokens, images=images, context_images=ctx)
                xs.append(self.embed_prompt(ids, vis))
        L0 = [x.shape[0] for x in xs]
        Lb = [self._bucket(n) for n in L0]
        self._alloc_cache(max(max(Lb), max(L0) + max_new_tokens + 1), slots=S)
        last = torch.tensor([n - 1 for n in L0], device=self.dev, dtype=torch.int64)
        if len(set(Lb)) == 1:                                   # prompts of one length bucket: one batched prefill into slots 0..S-1
            x_last = self.g_prefill((Lb[0], "multi", S), self._prefill_last_multi, x=torch.stack([self._pad_rows(x, Lb[0]) for x in xs], 0), last=last).clone()
        else:
            x_last = torch.cat([self.g_prefill((Lb[s], s), lambda x, last, s=s: self._prefill_last(x, slot=s, last=last), x=self._pad_rows(x, Lb[s]),
                                               last=last[s:s + 1]).clone() for s, x in enumerate(xs)], 0)   # [S, hidden]
        if S not in self._batch_state:
            self._batch_state[S] = dict(tok=torch.zeros(S, device=self.dev, dtype=torch.int64), pos=torch.zeros(S, device=self.dev, dtype=torch.int32),
                                        nk=torch.zeros(S, device=self.dev, dtype=torch.int32), qkv=torch.zeros((S, 3 * ph.hidden), device=self.dev, dtype=BF),
                                        ws=ops.decode_workspace(S, ph.heads, ph.head_dim, self.dev))
        st = self._batch_state[S]
        kv_bound = min(self.kv[0].shape[0], (max(L0) + max_new_tokens + 511) // 512 * 512)
        new_ids = [[] for _ in range(S)]
        seg_hidden = [[] for _ in range(S)]
        done = [False] * S
        nxt_dev = self._logits_argmax(x_last)
        for step in range(max_new_tokens):
            nxt = nxt_dev.tolist()                              # ONE host sync per step for the whole batch
            if forced and step in forced:
                nxt = [forced[step]] * S
            for s in range(S):
                if done[s]:
                    continue
                new_ids[s].append(nxt[s])
                if nxt[s] == cfg.seg_token_idx:
                    seg_hidden[s].append(ops.norm(x_last[s:s + 1], self.final_norm, eps=ph.eps, rms=True, out_dtype=F32))
                if nxt[s] in eos_ids:
                    done[s] = True
            if all(done) or step == max_new_tokens - 1:
                break
            if step == 0:
                st["pos"].copy_(torch.tensor(L0, dtype=torch.int32), non_blocking=True)
                st["nk"].copy_(torch.tensor([l + 1 for l in L0], dtype=torch.int32), non_blocking=True)
            if step == 0 or (forced and step in forced):
                st["tok"].copy_(torch.tensor(nxt, dtype=torch.int64), non_blocking=True)
            x_last, nxt_dev = self.g_decode_batch((S, kv_bound), lambda: self._decode_step_batch(S, kv_bound), mutable_state=(st["tok"], st["pos"], st["nk"]))
        out = []
        for s, (ids, _, _) in enumerate(requests):
            out_ids = torch.cat([ids.reshape(1, -1), torch.tensor([new_ids[s]], device=ids.device, dtype=torch.int64)], dim=1)
            if seg_hidden[s]:
                hs = ops.cast(torch.cat(seg_hidden[s], 0), BF)
                emb = ops.gemm(ops.gemm(hs, self.fc[0].w, bias=self.fc[0].b, act="relu"), self.fc[1].w, bias=self.fc[1].b, out_dtype=F32)
            else:
                emb = torch.zeros((0, cfg.out_dim), device=self.dev)
            out.append((out_ids, emb))
        return out

    @torch.no_grad()
    def generate(self, input_ids: torch.Tensor, images: torch.Tensor, context_images: torch.Tensor, max_new_tokens: int, eos_ids=(),
                 forced: dict | None = None, return_debug: bool = False):
        """Steps (A)-(D) of VideoGLaMM_SAM2.inference_video_branch VideoGLaMM.py:782-831 with greedy decoding:
        returns (output_ids int64 [1, L_txt+N] incl. -200 sentinels, pred_embeddings fp32 [n_seg, out_dim]).
        The hidden state that PREDICTED each [SEG] (mask shifted by one, :803-806) is the final-norm hidden of the step that emitted it."""
        cfg = self.cfg
        if context_images is None:      # still image (chat.py:458-487): CLIP-preprocessed image -> 576 un-pooled tokens at the single <image> sentinel
            vis = self.image_tokens(images) if self.P else self.g_vis((tuple(images.shape), None), self.image_tokens, image=images)
        elif self.P:
            vis = self.visual_tokens(images, context_images)
        else:
            vis = self.g_vis((tuple(images.shape), tuple(context_images.shape)), self.visual_tokens, images=images, context_images=context_images)
        x = self.embed_prompt(input_ids, vis)
        L0 = x.shape[0]
        self._alloc_cache(L0 + max_new_tokens + 1)
        dbg = None
        if return_debug:
            raise NotImplementedError("return_debug was removed with the fused decode path; use _llm() directly")
        elif self.P:
            x_last = self._prefill_last(x)
        else:
            Lb = self._bucket(L0)
            self._alloc_cache(max(Lb, L0 + max_new_tokens + 1))
            x_last = self.g_prefill((Lb,), lambda x, last: self._prefill_last(x, last=last), x=self._pad_rows(x, Lb),
                                    last=torch.tensor([L0 - 1], device=self.dev, dtype=torch.int64))
        new_ids, seg_hidden = [], []
        kv_bound = min(self.kv[0].shape[0], (L0 + max_new_tokens + 511) // 512 * 512)
        nxt_dev = self._logits_argmax(x_last)
        for step in range(max_new_tokens):
            nxt = int(nxt_dev[0])                           # host sync per token: EOS / [SEG] bookkeeping
            if forced and step in forced:
                nxt = forced[step]
            new_ids.append(nxt)
            if nxt == cfg.seg_token_idx:  # the hidden state that predicted [SEG]: final-norm of the residual row that produced it
                seg_hidden.append(ops.norm(x_last, self.final_norm, eps=cfg.phi.eps, rms=True, out_dtype=F32))
            if nxt in eos_ids or step == max_new_tokens - 1:
                break
            if self.P:   # verification mode: one un-fused, un-captured KV-cached forward per token
                tok = torch.tensor([nxt], device=self.dev, dtype=torch.int64)
                x_last = self._llm(self._embed(tok), L0 + step, last_only=True)
                nxt_dev = self._logits_argmax(x_last)
                continue
            if step == 0:
                self.dec_pos.fill_(L0)
                self.dec_nk.fill_(L0 + 1)
            if step == 0 or (forced and step in forced):
                self.dec_tok.fill_(nxt)
            x_last, nxt_dev = self.g_decode((kv_bound,), lambda: self._decode_step(kv_bound), mutable_state=(self.dec_tok, self.dec_pos, self.dec_nk))
        out_ids = torch.cat([input_ids.reshape(1, -1), torch.tensor([new_ids], device=input_ids.device, dtype=torch.int64)], dim=1)
        if seg_hidden:
            P = self.P
            hs = torch.cat(seg_hidden, 0) if P else ops.cast(torch.cat(seg_hidden, 0), BF)
            emb = ops.gemm(ops.gemm(hs, self.fc[0].W(P), bias=self.fc[0].b, act="relu", out_dtype=F32 if P else BF), self.fc[1].W(P), bias=self.fc[1].b,
                           out_dtype=F32)
        else:
            emb = torch.zeros((0, cfg.out_dim), device=self.dev)
        return (out_ids, emb, dbg) if return_debug else (out_ids, emb)
