"""B200 engine for the SAM2 side of VideoGLaMM's grounded-generation path: Hiera-L + FPN image encoder, prompt encoder,
TwoWay-transformer mask decoder, memory encoder, RoPE memory attention and the per-video tracking session.

Host orchestration only — every tensor operation on activations is a libvglamm kernel (videoglamm_b200.ops). Weights are
taken from a state_dict with the reference's keys (`model.visual_model.` prefix stripped by the caller).
Layout: all activations are channels-last token matrices [tokens, C]; residual streams fp32, GEMM operands bf16.

Reference call sites are cited per method (paths relative to VideoGLaMM/model/segment_anything_2/sam2/).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

from . import ops
from .graphs import GraphedStage
from .config import Sam2Config

NO_OBJ_SCORE = -1024.0
_OUT_MODE = {False: 0, True: 3, "packed": 4}   # vg_bilinear output map of the final upsampling: fp32 logits / uint8 (x>0) / bit-packed (x>0)
BF, F32 = torch.bfloat16, torch.float32


@dataclass
class _Lin:
    w: torch.Tensor  # bf16 [N, K]
    b: torch.Tensor | None  # fp32 [N]
    w3: "ops.W3 | None" = None  # split-bf16 copy of the fp32 weight (fp32-grade GEMMs, ops.gemm3)

    def W(self, precise: bool):
        if precise:
            assert self.w3 is not None, "weight was packed without its split-bf16 copy"
            return self.w3
        return self.w


@dataclass
class FrameMem:
    pred_masks: torch.Tensor          # [B, 256, 256] fp32 low-res logits
    obj_ptr: torch.Tensor             # [B, C] fp32
    maskmem: torch.Tensor | None      # [B, HW, mem_dim] bf16


class Sam2Engine:
    def __init__(self, sd: dict, cfg: Sam2Config | None = None, device: str = "cuda", prefix: str = "", precise: str = "heads"):
        """precise: which stages run in fp32-grade arithmetic (split-bf16 tensor-core GEMMs, exact fp32 attention; include/vglamm.h vg_split3):
        "heads" (default) = the mask decoder + IoU / object-score / object-pointer MLPs, whose rounding dominates the mask-logit error
        (scripts/precision_study.py: bf16 there -> 5e-3, fp32-grade there -> <= 4e-4 with everything else bf16); "none" = bf16 operands everywhere;
        "all" = every stage incl. Hiera and the memory path (verification mode: parity to the fp32 reference at ~1e-5, ~3x the GEMM work)."""
        self.cfg = cfg or Sam2Config()
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("Sam2Engine needs a CUDA device (there is no CPU path)")
        assert precise in ("none", "heads", "all"), precise
        self.precise = precise
        self.p_heads, self.p_all = precise != "none", precise == "all"
        self.fused_decoder = self.p_heads and (self.cfg.d_model, self.cfg.dec_heads, self.cfg.dec_mlp, self.cfg.dec_depth) == (256, 8, 2048, 2)
        self.frames_per_encode = 8
        self.g_video = GraphedStage("sam2_video")
        self._sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        self._pack()
        del self._sd

    # ------------------------------------------------------------------------------------------------ weight packing
    def _t(self, k):
        return self._sd[k].detach().to(self.dev, F32)

    def _mk(self, w: torch.Tensor, b, precise: bool) -> _Lin:
        w = w.contiguous()
        return _Lin(w.to(BF).contiguous(), b, ops.W3(w) if precise else None)

    def _lin(self, p, precise: bool | None = None) -> _Lin:
        w = self._t(p + ".weight")
        w = w.reshape(w.shape[0], -1)
        b = self._sd.get(p + ".bias")
        return self._mk(w, None if b is None else b.detach().to(self.dev, F32).contiguous(), self.p_all if precise is None else precise)

    def _ln(self, p):
        return self._t(p + ".weight").contiguous(), self._t(p + ".bias").contiguous()

    def _attn_w(self, p, precise: bool | None = None):
        return {n: self._lin(f"{p}.{n}_proj", precise) for n in ("q", "k", "v", "out")}

    def _pack(self):
        cfg, t = self.cfg, "image_encoder.trunk"
        C = cfg.d_model
        # Hiera patch embedding (backbones/utils.py:65-95): [C,3,7,7] -> [C,147] padded to K=152
        w = self._t(t + ".patch_embed.proj.weight").reshape(cfg.embed_dim, -1)
        self.pe_kpad = (w.shape[1] + 7) // 8 * 8
        wp = torch.zeros(cfg.embed_dim, self.pe_kpad, device=self.dev)
        wp[:, : w.shape[1]] = w
        self.patch = self._mk(wp, self._t(t + ".patch_embed.proj.bias"), self.p_all)
        # Hiera._get_pos_embed (backbones/hieradet.py:269-277): constant of the weights, folded at load time
        g = cfg.image_size // 4
        pe = F.interpolate(self._t(t + ".pos_embed"), size=(g, g), mode="bicubic")
        win = self._t(t + ".pos_embed_window")
        pe = pe + win.tile([x // y for x, y in zip(pe.shape, win.shape)])
        self.hiera_pos = pe.permute(0, 2, 3, 1).reshape(g * g, cfg.embed_dim).contiguous()
        self.blocks = []
        for i, (dim, dim_out, heads, window, qs) in enumerate(cfg.block_specs()):
            b = f"{t}.blocks.{i}"
            blk = dict(n1=self._ln(b + ".norm1"), qkv=self._lin(b + ".attn.qkv"), proj=self._lin(b + ".attn.proj"), n2=self._ln(b + ".norm2"),
                       fc1=self._lin(b + ".mlp.layers.0"), fc2=self._lin(b + ".mlp.layers.1"), spec=(dim, dim_out, heads, window, qs))
            if dim != dim_out:
                blk["dproj"] = self._lin(b + ".proj")
            self.blocks.append(blk)
        # FPN neck (backbones/image_encoder.py:101-133), scalp=1, top-down only into level 2. Levels 0/1 feed only conv_s0/conv_s1
        # (sam2_base.py:465-477): two chained 1x1 convs with no nonlinearity between them are folded into one at load time (fp32).
        neck = [self._lin(f"image_encoder.neck.convs.{j}.conv") for j in range(4)]  # j -> stage 3-j
        self.neck3, self.neck2 = neck[0], neck[1]

        def fold(outer: str, inner: _Lin):
            wo, bo = self._t(outer + ".weight").reshape(-1, C), self._t(outer + ".bias")
            return self._mk(wo @ inner.w.float(), (wo @ inner.b + bo).contiguous(), self.p_all)

        w2 = self._t("image_encoder.neck.convs.2.conv.weight").reshape(C, -1)
        w3 = self._t("image_encoder.neck.convs.3.conv.weight").reshape(C, -1)
        self.s1 = fold("sam_mask_decoder.conv_s1", _Lin(w2, self._t("image_encoder.neck.convs.2.conv.bias")))
        self.s0 = fold("sam_mask_decoder.conv_s0", _Lin(w3, self._t("image_encoder.neck.convs.3.conv.bias")))
        fs = cfg.feat_size
        self.HW = fs * fs
        self.pos_top = self._sine_pos(fs, fs, C)              # [HW, C] vision_pos_enc of the top level
        self.rope_table = ops.rope2d_table(C, fs, fs, self.dev)   # (cos, sin) of the memory attention's axial RoPE, filled once
        self.pos_mem = self._sine_pos(fs, fs, cfg.mem_dim)    # [HW, mem_dim] maskmem_pos_enc
        tpos = self._t("maskmem_tpos_enc").reshape(cfg.num_maskmem, cfg.mem_dim)
        self.mem_pos_slot = (self.pos_mem[None] + tpos[:, None]).contiguous()  # [7, HW, mem_dim] (sam2_base.py:577-581)
        self.no_mem_embed = self._t("no_mem_embed").reshape(1, C).contiguous()
        self.no_obj_ptr = self._t("no_obj_ptr").reshape(1, C)
        # prompt encoder constants (sam/prompt_encoder.py:68-77,192-229)
        gm = self._t("sam_prompt_encoder.pe_layer.positional_encoding_gaussian_matrix")
        grid = torch.ones((fs, fs), device=self.dev)
        yy, xx = (grid.cumsum(0) - 0.5) / fs, (grid.cumsum(1) - 0.5) / fs
        cc = 2 * math.pi * ((2 * torch.stack([xx, yy], -1) - 1) @ gm)
        self.dense_pe = torch.cat([cc.sin(), cc.cos()], -1).reshape(self.HW, C).contiguous()
        self.not_a_point = self._t("sam_prompt_encoder.not_a_point_embed.weight")
        self.no_mask_embed = self._t("sam_prompt_encoder.no_mask_embed.weight").reshape(1, C).contiguous()
        # mask decoder
        md = "sam_mask_decoder"
        self.out_tokens = torch.cat([self._t(md + ".obj_score_token.weight"), self._t(md + ".iou_token.weight"), self._t(md + ".mask_tokens.weight")], 0)
        self.dec_layers = []
        for i in range(cfg.dec_depth):
            l = f"{md}.transformer.layers.{i}"
            PH = self.p_heads
            self.dec_layers.append(dict(sa=self._attn_w(l + ".self_attn", PH), t2i=self._attn_w(l + ".cross_attn_token_to_image", PH),
                                        i2t=self._attn_w(l + ".cross_attn_image_to_token", PH), n1=self._ln(l + ".norm1"), n2=self._ln(l + ".norm2"),
                                        n3=self._ln(l + ".norm3"), n4=self._ln(l + ".norm4"), fc1=self._lin(l + ".mlp.layers.0", PH),
                                        fc2=self._lin(l + ".mlp.layers.1", PH)))
        PH = self.p_heads
        self.dec_final = self._attn_w(md + ".transformer.final_attn_token_to_image", PH)
        self.dec_nf = self._ln(md + ".transformer.norm_final_attn")
        # ConvTranspose2d weight [Cin, Cout, 2, 2] -> GEMM weight rows (dy,dx,co), cols ci (mask_decoder.py:66-77)
        for name, key in (("up1", ".output_upscaling.0"), ("up2", ".output_upscaling.3")):
            wt = self._t(md + key + ".weight")
            setattr(self, name, self._mk(wt.permute(2, 3, 1, 0).reshape(-1, wt.shape[0]), self._t(md + key + ".bias"), PH))
        self.up_ln = self._ln(md + ".output_upscaling.1")
        # 4 hypernetwork MLPs as batched GEMMs: [4, out, in] weights, [4, out] biases
        self.hyper = []
        for j in range(3):
            ws = torch.stack([self._t(f"{md}.output_hypernetworks_mlps.{i}.layers.{j}.weight") for i in range(4)])
            bs = torch.stack([self._t(f"{md}.output_hypernetworks_mlps.{i}.layers.{j}.bias") for i in range(4)])
            self.hyper.append(self._mk(ws, bs.contiguous(), PH))
        self.iou_head = [self._lin(f"{md}.iou_prediction_head.layers.{j}", PH) for j in range(3)]
        self.obj_head = [self._lin(f"{md}.pred_obj_score_head.layers.{j}", PH) for j in range(3)]
        self.obj_ptr_proj = [self._lin(f"obj_ptr_proj.layers.{j}", PH) for j in range(3)]
        if self.fused_decoder:
            self._pack_fused_decoder()
        # memory encoder (memory_encoder.py)
        me = "memory_encoder"
        self.ds = []
        for j in range(4):
            wt = self._t(f"{me}.mask_downsampler.encoder.{3 * j}.weight")  # [cout, cin, 3, 3]
            self.ds.append(dict(w=wt.permute(2, 3, 1, 0).reshape(-1, wt.shape[0]).contiguous(), b=self._t(f"{me}.mask_downsampler.encoder.{3 * j}.bias"),
                                wg=wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1).to(BF).contiguous(),   # [cout, (ky,kx,ci)] for the GEMM path
                                ln=self._ln(f"{me}.mask_downsampler.encoder.{3 * j + 1}"), cin=wt.shape[1], cout=wt.shape[0]))
        self.ds_out = self._lin(f"{me}.mask_downsampler.encoder.12")
        self.pix_proj = self._lin(me + ".pix_feat_proj")
        self.fuser = []
        for j in range(2):
            l = f"{me}.fuser.layers.{j}"
            self.fuser.append(dict(dw=self._t(l + ".dwconv.weight").reshape(C, 7, 7).contiguous(), dwb=self._t(l + ".dwconv.bias"), ln=self._ln(l + ".norm"),
                                   pw1=self._lin(l + ".pwconv1"), pw2=self._lin(l + ".pwconv2"), gamma=self._t(l + ".weight").contiguous()))
        self.mem_out = self._lin(me + ".out_proj")
        # memory attention (memory_attention.py); self-attention q/k/v fused into one [3C, C] GEMM
        self.ma_layers = []
        for i in range(cfg.mem_attn_layers):
            l = f"memory_attention.layers.{i}"
            sa, ca = self._attn_w(l + ".self_attn"), self._attn_w(l + ".cross_attn_image")
            qkv = self._mk(torch.cat([self._t(f"{l}.self_attn.{n}_proj.weight") for n in "qkv"]), torch.cat([sa["q"].b, sa["k"].b, sa["v"].b]).contiguous(),
                           self.p_all)
            self.ma_layers.append(dict(qkv=qkv, sa_out=sa["out"], ca=ca, n1=self._ln(l + ".norm1"), n2=self._ln(l + ".norm2"), n3=self._ln(l + ".norm3"),
                                       fc1=self._lin(l + ".linear1"), fc2=self._lin(l + ".linear2")))
        self.ma_norm = self._ln("memory_attention.norm")

    def _pack_fused_decoder(self):
        """Table of fp32 device pointers for vg_mask_decoder (include/vglamm.h VgMaskDecoderWeights): nn.Linear layout for the token-side GEMVs,
        transposed copies for the row-local image-side contractions, and the load-time constants dense_pe @ Wk^T / dense_pe @ Wq^T."""
        from . import _lib as L
        md = "sam_mask_decoder"
        keep = self._fused_keep = []

        def lin(key, transpose=False, w=None):
            w = self._t(key + ".weight") if w is None else w
            w = w.reshape(w.shape[0], -1)
            w = (w.t() if transpose else w).contiguous()
            b = self._t(key + ".bias").contiguous()
            keep.extend([w, b])
            return L.VgLinear(w.data_ptr(), b.data_ptr())

        def attn(p, tq=False, tk=False, tv=False, tout=False):
            return L.VgMdAttn(lin(p + ".q_proj", tq), lin(p + ".k_proj", tk), lin(p + ".v_proj", tv), lin(p + ".out_proj", tout))

        def const(x):
            x = x.contiguous()
            keep.append(x)
            return x.data_ptr()

        W = L.VgMaskDecoderWeights()
        for i in range(2):
            l = f"{md}.transformer.layers.{i}"
            W.layer[i] = L.VgMdLayer(attn(l + ".self_attn"), attn(l + ".cross_attn_token_to_image", tk=True, tv=True),
                                     attn(l + ".cross_attn_image_to_token", tq=True, tout=True), lin(l + ".norm1"), lin(l + ".norm2"), lin(l + ".norm3"),
                                     lin(l + ".norm4"), lin(l + ".mlp.layers.0"), lin(l + ".mlp.layers.1"))
            W.pek[i] = const(self.dense_pe @ self._t(l + ".cross_attn_token_to_image.k_proj.weight").t())
            W.peq[i] = const(self.dense_pe @ self._t(l + ".cross_attn_image_to_token.q_proj.weight").t())
        f = md + ".transformer.final_attn_token_to_image"
        W.final_t2i = attn(f, tk=True, tv=True)
        W.pek[2] = const(self.dense_pe @ self._t(f + ".k_proj.weight").t())
        W.norm_final = lin(md + ".transformer.norm_final_attn")
        for name, key in (("up1", ".output_upscaling.0"), ("up2", ".output_upscaling.3")):
            wt = self._t(md + key + ".weight")                                   # ConvTranspose2d [Cin, Cout, 2, 2] -> T[Cin][(dy,dx,co)]
            setattr(W, name, lin(md + key, w=wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1)))
        W.up_ln = lin(md + ".output_upscaling.1")
        for i in range(4):
            for j in range(3):
                W.hyper[i][j] = lin(f"{md}.output_hypernetworks_mlps.{i}.layers.{j}")
        for j in range(3):
            W.iou[j] = lin(f"{md}.iou_prediction_head.layers.{j}")
            W.obj[j] = lin(f"{md}.pred_obj_score_head.layers.{j}")
        # tensor-core path: the 12 image-side matrices in nn.Linear layout, split into bf16 hi / lo (include/vglamm.h img_hi / img_lo)
        def split(i, w):
            w = w.reshape(w.shape[0], -1).contiguous()
            hi = w.to(BF)
            lo = (w - hi.float()).to(BF)
            keep.extend([hi, lo])
            W.img_hi[i], W.img_lo[i] = hi.data_ptr(), lo.data_ptr()

        for i in range(2):
            l = f"{md}.transformer.layers.{i}"
            split(4 * i + 0, self._t(l + ".cross_attn_token_to_image.k_proj.weight"))
            split(4 * i + 1, self._t(l + ".cross_attn_token_to_image.v_proj.weight"))
            split(4 * i + 2, self._t(l + ".cross_attn_image_to_token.q_proj.weight"))
            split(4 * i + 3, self._t(l + ".cross_attn_image_to_token.out_proj.weight"))
        split(8, self._t(f + ".k_proj.weight"))
        split(9, self._t(f + ".v_proj.weight"))
        for i, key in ((10, ".output_upscaling.0"), (11, ".output_upscaling.3")):
            wt = self._t(md + key + ".weight")                                   # [Cin, Cout, 2, 2] -> [(dy,dx,co)][Cin]
            split(i, wt.permute(2, 3, 1, 0).reshape(-1, wt.shape[0]))
        self.fused_w = W
        self.fused_exact = False            # True: fp32-FMA row-local contractions (vg_mask_decoder force_exact)
        self.no_mask_vec = self.no_mask_embed.reshape(-1).contiguous()

    def _sine_pos(self, H, W, num_pos_feats):
        """PositionEmbeddingSine (position_encoding.py:79-112): a constant table per feature size -> [H*W, C] fp32."""
        npf = num_pos_feats // 2
        y = torch.arange(1, H + 1, dtype=F32, device=self.dev).view(H, 1).repeat(1, W)
        x = torch.arange(1, W + 1, dtype=F32, device=self.dev).view(1, W).repeat(H, 1)
        y = y / (y[-1:, :] + 1e-6) * (2 * math.pi)
        x = x / (x[:, -1:] + 1e-6) * (2 * math.pi)
        dim_t = torch.arange(npf, dtype=F32, device=self.dev)
        dim_t = 10000.0 ** (2 * (dim_t // 2) / npf)
        px, py = x[:, :, None] / dim_t, y[:, :, None] / dim_t
        px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
        py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
        return torch.cat((py, px), dim=2).reshape(H * W, num_pos_feats).contiguous()

    # ------------------------------------------------------------------------------------------------ image encoder
    def _hiera_block(self, x, blk, n, H, W):
        """MultiScaleBlock.forward backbones/hieradet.py:136-168. x fp32 [n*H*W, dim] -> (fp32 [n*H'*W', dim_out], H', W')."""
        dim, dim_out, heads, window, qs = blk["spec"]
        P = self.p_all
        adt = F32 if P else BF
        hd = dim_out // heads
        xn = ops.norm(x, *blk["n1"], eps=1e-6, out_dtype=adt)
        if dim != dim_out:
            sc = ops.gemm(xn, blk["dproj"].W(P), bias=blk["dproj"].b, out_dtype=F32)
            if qs:
                sc = ops.pool2x2(sc, n, H, W, is_max=True)
        else:
            sc = x
        qkv = ops.gemm(xn, blk["qkv"].W(P), bias=blk["qkv"].b, out_dtype=adt)  # [rows, 3*dim_out] = (3, heads, hd)
        q, k, v = (qkv[:, j * dim_out:(j + 1) * dim_out].unflatten(1, (heads, hd)) for j in range(3))
        Hq, Wq = H, W
        if qs:
            q = ops.pool2x2(qkv[:, :dim_out], n, H, W, is_max=True).unflatten(1, (heads, hd))
            Hq, Wq = H // 2, W // 2
        if window > 0:
            wq = window // 2 if qs else window
            nwx, nwy = W // window, H // window
            o = ops.attention(q[None], k[None], v[None], scale=hd ** -0.5, q_win=(wq, Wq, nwx, nwy), k_win=(window, W, nwx, nwy))
        else:
            o = ops.attention(q.unflatten(0, (n, Hq * Wq)), k.unflatten(0, (n, H * W)), v.unflatten(0, (n, H * W)), scale=hd ** -0.5)
        o = o.reshape(n * Hq * Wq, dim_out)
        x = ops.gemm(o, blk["proj"].W(P), bias=blk["proj"].b, resid=sc, out=sc, out_dtype=F32)  # in place: aliasing residual -> TMA reduce-add epilogue
        xn = ops.norm(x, *blk["n2"], eps=1e-6, out_dtype=adt)
        h = ops.gemm(xn, blk["fc1"].W(P), bias=blk["fc1"].b, act="gelu", out_dtype=adt)
        x = ops.gemm(h, blk["fc2"].W(P), bias=blk["fc2"].b, resid=x, out=x, out_dtype=F32)
        return x, Hq, Wq

    def forward_image(self, img: torch.Tensor) -> dict:
        """SAM2Base.forward_image sam2_base.py:465-477 (Hiera hieradet.py:279-295 + FpnNeck image_encoder.py:101-133 + conv_s0/s1).
        img [n,3,S,S] fp32 -> {"s0": [n*(S/4)^2, 32], "s1": [n*(S/8)^2, 64] (bf16; fp32 when the heads are precise), "top": fp32 [n*(S/16)^2, 256]}"""
        cfg = self.cfg
        P = self.p_all
        adt = F32 if P else BF
        n = img.shape[0]
        assert img.shape[1:] == (3, cfg.image_size, cfg.image_size) and img.is_cuda
        img = img.contiguous()
        g = cfg.image_size // 4
        x = torch.empty((n * g * g, cfg.embed_dim), device=self.dev, dtype=F32)
        for i in range(n):  # pos-embed residual is per image
            cols, _, _ = ops.im2col(img[i:i + 1], 7, 4, 3, self.pe_kpad, out_dtype=adt)
            ops.gemm(cols, self.patch.W(P), bias=self.patch.b, resid=self.hiera_pos, out=x[i * g * g:(i + 1) * g * g])
        H = W = g
        stage_out = []
        ends = cfg.stage_ends
        for i, blk in enumerate(self.blocks):
            x, H, W = self._hiera_block(x, blk, n, H, W)
            if i in ends:
                stage_out.append((x.clone() if P else ops.cast(x, BF), H, W))
        (x0, H0, W0), (x1, H1, W1), (x2, H2, W2), (x3, H3, W3) = stage_out
        lat3 = ops.gemm(x3, self.neck3.W(P), bias=self.neck3.b, out_dtype=F32)
        lat2 = ops.gemm(x2, self.neck2.W(P), bias=self.neck2.b, out_dtype=F32)
        top = ops.upsample2x_add(lat2, lat3, n, H3, W3)
        hdt = F32 if self.p_heads else BF     # the high-res maps are summands of the upscaled mask features: keep them unrounded for the precise heads
        s1 = ops.gemm(x1, self.s1.W(P), bias=self.s1.b, out_dtype=hdt)
        s0 = ops.gemm(x0, self.s0.W(P), bias=self.s0.b, out_dtype=hdt)
        return dict(s0=s0, s1=s1, top=top)

    # ------------------------------------------------------------------------------------------------ mask decoder
    def _attn(self, w, q_in, k_in, v_in, B, Nq, Nk, resid=None, P=False):
        """Attention.forward sam/transformer.py:233-260 on [B*N, C] inputs (bf16, or fp32 when P) -> fp32 [B*Nq, C] (+ residual)."""
        H = self.cfg.dec_heads
        adt = F32 if P else BF
        q = ops.gemm(q_in, w["q"].W(P), bias=w["q"].b, out_dtype=adt)
        k = ops.gemm(k_in, w["k"].W(P), bias=w["k"].b, out_dtype=adt)
        v = ops.gemm(v_in, w["v"].W(P), bias=w["v"].b, out_dtype=adt)
        d = q.shape[1] // H
        o = ops.attention(q.view(B, Nq, H, d), k.view(B, Nk, H, d), v.view(B, Nk, H, d), scale=d ** -0.5)
        return ops.gemm(o.view(B * Nq, H * d), w["out"].W(P), bias=w["out"].b, resid=resid, out=resid, out_dtype=F32)  # in place when residual

    def _mlp3(self, layers, x, final_act=None, P=False):
        adt = F32 if P else BF
        h = ops.gemm(x, layers[0].W(P), bias=layers[0].b, act="relu", out_dtype=adt)
        h = ops.gemm(h, layers[1].W(P), bias=layers[1].b, act="relu", out_dtype=adt)
        return ops.gemm(h, layers[2].W(P), bias=layers[2].b, act=final_act, out_dtype=F32)

    def mask_decoder(self, pix: torch.Tensor, s0: torch.Tensor, s1: torch.Tensor, sparse: torch.Tensor, debug: dict | None = None, zero_dense: bool = False):
        """PromptEncoder dense part + MaskDecoder.predict_masks (sam/mask_decoder.py:167-245), TwoWayTransformer
        (sam/transformer.py:63-115, 160-193). pix fp32 [B*HW, C] image embedding (one row block per object), s0/s1 the
        high-res maps of the ONE image shared by the B objects, sparse fp32 [B, Ns, C].
        Returns all-4 masks fp32 [B,4,(4fs)^2], iou [B,4] (sigmoid), mask tokens fp32 [B,4,C], object score logits [B,1].
        With precise heads (default) every contraction is a split-bf16 GEMM and every attention exact fp32 (activations stay fp32)."""
        cfg, HW, C = self.cfg, self.HW, self.cfg.d_model
        P = self.p_heads
        adt = F32 if P else BF
        act = (lambda t: t) if P else (lambda t: ops.cast(t, BF))     # GEMM-operand view of an fp32 activation
        B = sparse.shape[0]
        if self.fused_decoder and debug is None and s0.dtype == F32 and pix.shape[0] == B * HW:
            tokens = torch.cat([self.out_tokens[None].expand(B, -1, -1), sparse], dim=1).contiguous()
            no_mask = torch.zeros_like(self.no_mask_vec) if zero_dense else self.no_mask_vec   # zero dense embedding: the padded objects of model_forward
            return ops.mask_decoder_fused(self.fused_w, pix, self.dense_pe, no_mask, s0, s1, tokens, cfg.feat_size, exact=self.fused_exact)
        tokens = torch.cat([self.out_tokens[None].expand(B, -1, -1), sparse], dim=1).reshape(-1, C).contiguous()
        Nt = tokens.shape[0] // B
        keys = pix.clone() if zero_dense else ops.add(pix, self.no_mask_embed)    # src = image_embeddings + dense (no_mask_embed)
        queries = tokens
        k_pe = v_k = None
        for i, L in enumerate(self.dec_layers):
            if i == 0:
                qb = act(queries)
                queries = self._attn(L["sa"], qb, qb, qb, B, Nt, Nt, P=P)
            else:
                q = ops.add(queries, tokens, out_dtype=adt)
                queries = self._attn(L["sa"], q, q, act(queries), B, Nt, Nt, resid=queries, P=P)
            queries = ops.norm(queries, *L["n1"], eps=1e-5, out_dtype=F32)
            q = ops.add(queries, tokens, out_dtype=adt)
            k_pe = ops.add(keys, self.dense_pe, out_dtype=adt)
            v_k = act(keys)
            queries = ops.norm(self._attn(L["t2i"], q, k_pe, v_k, B, Nt, HW, resid=queries, P=P), *L["n2"], eps=1e-5, out_dtype=F32)
            h = ops.gemm(act(queries), L["fc1"].W(P), bias=L["fc1"].b, act="relu", out_dtype=adt)
            queries = ops.norm(ops.gemm(h, L["fc2"].W(P), bias=L["fc2"].b, resid=queries, out=queries, out_dtype=F32), *L["n3"], eps=1e-5, out_dtype=F32)
            q = ops.add(queries, tokens, out_dtype=adt)
            keys = ops.norm(self._attn(L["i2t"], k_pe, q, act(queries), B, HW, Nt, resid=keys, P=P), *L["n4"], eps=1e-5, out_dtype=F32)
        q = ops.add(queries, tokens, out_dtype=adt)
        k_pe = ops.add(keys, self.dense_pe, out_dtype=adt)
        v_k = act(keys)
        queries = ops.norm(self._attn(self.dec_final, q, k_pe, v_k, B, Nt, HW, resid=queries, P=P), *self.dec_nf, eps=1e-5, out_dtype=F32)
        hs = queries.view(B, Nt, C)
        # upscaling: ConvT(k2,s2) as GEMM + pixel-shuffle epilogue (+feat_s1, LayerNorm2d, GELU), then again (+feat_s0, GELU)
        fs = cfg.feat_size
        g1 = ops.gemm(v_k, self.up1.W(P), out_dtype=F32)
        up1 = ops.convt2x2_epilogue(g1, self.up1.b, s1, self.up_ln[0], self.up_ln[1], 1e-6, B, fs, fs, C // 4, feat_bcast=True, out_dtype=adt)
        g2 = ops.gemm(up1, self.up2.W(P), out_dtype=F32)
        up2 = ops.convt2x2_epilogue(g2, self.up2.b, s0, None, None, 0.0, B, 2 * fs, 2 * fs, C // 8, feat_bcast=True, out_dtype=adt)
        # hypernetwork MLPs, batched over the 4 mask tokens
        mt = hs[:, 2:6]                                                # [B,4,C]
        h = act(mt.transpose(0, 1).contiguous())                       # [4,B,C]
        h = ops.gemm(h, self.hyper[0].W(P), bias=self.hyper[0].b, act="relu", out_dtype=adt)
        h = ops.gemm(h, self.hyper[1].W(P), bias=self.hyper[1].b, act="relu", out_dtype=adt)
        hyper_in = ops.gemm(h, self.hyper[2].W(P), bias=self.hyper[2].b, out_dtype=adt).transpose(0, 1).contiguous()   # [B,4,C/8]
        if P:
            masks = ops.gemm3(hyper_in, up2.view(B, -1, C // 8), out_dtype=F32)                        # both operands are activations: split both
        else:
            masks = ops.gemm(hyper_in, up2.view(B, -1, C // 8), out_dtype=F32)                         # [B,4,(4fs)^2]
        if debug is not None:
            debug.update(hs=hs, keys=keys, up1=up1, up2=up2, hyper_in=hyper_in)
        iou = self._mlp3(self.iou_head, act(hs[:, 1].contiguous()), final_act="sigmoid", P=P)
        obj = self._mlp3(self.obj_head, act(hs[:, 0].contiguous()), P=P)
        return masks, iou, mt, obj

    def forward_sam_heads(self, pix, s0, s1, text_embeds: torch.Tensor | None, B: int):
        """SAM2Base._forward_sam_heads sam2_base.py:251-411 for the prompt kinds on the VideoGLaMM path (padding points + optional
        text token), multimask output with best-IoU selection."""
        sparse = self.not_a_point[None].expand(B, 2, -1)
        if text_embeds is not None:
            sparse = torch.cat([sparse, text_embeds.to(F32)], dim=1)
        masks, iou, mt, obj = self.mask_decoder(pix, s0, s1, sparse)
        side = 4 * self.cfg.feat_size
        masks, iou, toks = masks[:, 1:], iou[:, 1:], mt[:, 1:]
        is_obj = obj > 0                                                                    # [B,1]
        low_multi = torch.where(is_obj[:, :, None], masks, torch.full_like(masks, NO_OBJ_SCORE))
        best = torch.argmax(iou, dim=-1)
        bi = torch.arange(B, device=self.dev)
        low = low_multi[bi, best].reshape(B, side, side).contiguous()
        tok = toks[bi, best].contiguous()
        ptr = self._mlp3(self.obj_ptr_proj, tok if self.p_heads else ops.cast(tok, BF), P=self.p_heads)
        lam = is_obj.float()
        ptr = lam * ptr + (1 - lam) * self.no_obj_ptr
        return dict(low_res_masks=low, obj_ptr=ptr, ious=iou, object_score_logits=obj, low_res_multimasks=low_multi.view(B, 3, side, side))

    # ------------------------------------------------------------------------------------------------ memory encoder
    def encode_memory(self, top: torch.Tensor, low_res_masks: torch.Tensor, binarize: bool) -> torch.Tensor:
        """_encode_new_memory sam2_base.py:666-704 + MemoryEncoder.forward memory_encoder.py:159-182.
        top fp32 [HW, C] (raw top-level feature of the frame), low_res_masks fp32 [B, 4fs, 4fs] -> maskmem bf16 [B, HW, mem_dim]."""
        cfg, C = self.cfg, self.cfg.d_model
        P = self.p_all
        adt = F32 if P else BF
        B, S = low_res_masks.shape[0], cfg.image_size
        m = ops.bilinear(low_res_masks, S, S, mode=2 if binarize else 1, scale=cfg.sigmoid_scale_for_mem_enc, bias=cfg.sigmoid_bias_for_mem_enc)
        m = m.view(B * S * S, 1)
        side = S
        for j, d in enumerate(self.ds):
            gemm_path = d["cin"] >= 16 and not P                                                   # wide stages: im2col + tcgen05 GEMM (bf16 operands)
            conv = ops.conv3x3s2_ln_gelu_gemm if gemm_path else ops.conv3x3s2_ln_gelu
            m = conv(m, d["wg"] if gemm_path else d["w"], d["b"], d["ln"][0], d["ln"][1], 1e-6, B, side, side, d["cin"], d["cout"],
                     out_dtype=adt if j == 3 else F32)
            side //= 2
        m = ops.gemm(m, self.ds_out.W(P), bias=self.ds_out.b, out_dtype=F32)                                # [B*HW, C]
        px = ops.gemm(top if P else ops.cast(top, BF), self.pix_proj.W(P), bias=self.pix_proj.b, out_dtype=F32)   # [HW, C] shared by objects
        x = ops.add(m, px)
        fs = cfg.feat_size
        for f in self.fuser:
            y = ops.dwconv(x, f["dw"], f["dwb"], B, fs, fs, 7)
            yn = ops.norm(y, *f["ln"], eps=1e-6, out_dtype=adt)
            h = ops.gemm(yn, f["pw1"].W(P), bias=f["pw1"].b, act="gelu", out_dtype=adt)
            x = ops.gemm(h, f["pw2"].W(P), bias=f["pw2"].b, col_scale=f["gamma"], resid=x, out=x, out_dtype=F32)
        out = ops.gemm(x if P else ops.cast(x, BF), self.mem_out.W(P), bias=self.mem_out.b)                  # bf16 (sam2_video_predictor.py:967)
        return out.view(B, self.HW, cfg.mem_dim)

    # ------------------------------------------------------------------------------------------------ memory attention
    def memory_attention(self, top: torch.Tensor, B: int, memory: torch.Tensor, mem_pos: torch.Tensor, n_ptr_tokens: int) -> torch.Tensor:
        """MemoryAttention.forward memory_attention.py:119-169 / MemoryAttentionLayer :58-99 / RoPEAttention sam/transformer.py:285-327.
        top fp32 [HW, C]; memory bf16 (fp32 when precise="all") [B, Nk, mem_dim]; mem_pos fp32 [Nk, mem_dim] (shared by objects) -> fp32 [B*HW, C]."""
        cfg, C, HW = self.cfg, self.cfg.d_model, self.HW
        P = self.p_all
        adt = F32 if P else BF
        fs = cfg.feat_size
        Nk = memory.shape[1]
        n_rope = Nk - n_ptr_tokens
        x = ops.add(top, self.pos_top, alpha=0.1, rows=B * HW)                                           # curr + 0.1*curr_pos, tiled over objects
        mem2 = memory.reshape(B * Nk, cfg.mem_dim)                                                       # bf16, or fp32 in verification mode (_memory_bank)
        k_in = ops.add(mem2, mem_pos, out_dtype=adt)
        for L in self.ma_layers:
            t = ops.norm(x, *L["n1"], eps=1e-5, out_dtype=adt)
            qkv = ops.gemm(t, L["qkv"].W(P), bias=L["qkv"].b, out_dtype=adt)                               # [B*HW, 3C]
            ops.rope2d_(qkv[:, :2 * C], C, fs, fs, table=self.rope_table, nblk=2)                         # q and k in one launch
            q, k, v = (qkv[:, j * C:(j + 1) * C].view(B, HW, 1, C) for j in range(3))
            o = ops.attention(q, k, v, scale=C ** -0.5)
            x = ops.gemm(o.view(B * HW, C), L["sa_out"].W(P), bias=L["sa_out"].b, resid=x, out=x, out_dtype=F32)
            t = ops.norm(x, *L["n2"], eps=1e-5, out_dtype=adt)
            ca = L["ca"]
            q = ops.gemm(t, ca["q"].W(P), bias=ca["q"].b, out_dtype=adt)
            ops.rope2d_(q, C, fs, fs, table=self.rope_table)
            kk = ops.gemm(k_in, ca["k"].W(P), bias=ca["k"].b, out_dtype=adt).view(B, Nk, C)
            for b in range(B):
                ops.rope2d_(kk[b, :n_rope], C, fs, fs, table=self.rope_table)                             # rope_k_repeat: position = row % HW
            vv = ops.gemm(mem2, ca["v"].W(P), bias=ca["v"].b, out_dtype=adt)
            o = ops.attention(q.view(B, HW, 1, C), kk.view(B, Nk, 1, C), vv.view(B, Nk, 1, C), scale=C ** -0.5)
            x = ops.gemm(o.view(B * HW, C), ca["out"].W(P), bias=ca["out"].b, resid=x, out=x, out_dtype=F32)
            t = ops.norm(x, *L["n3"], eps=1e-5, out_dtype=adt)
            h = ops.gemm(t, L["fc1"].W(P), bias=L["fc1"].b, act="relu", out_dtype=adt)
            x = ops.gemm(h, L["fc2"].W(P), bias=L["fc2"].b, resid=x, out=x, out_dtype=F32)
        return ops.norm(x, *self.ma_norm, eps=1e-5, out_dtype=F32)

    def _memory_bank(self, t: int, cond: dict, non_cond: dict, num_frames: int, B: int):
        """_prepare_memory_conditioned_features sam2_base.py:495-664 (forward tracking, stride 1): conditioning frame(s) at t_pos 0,
        the 6 previous frames, then up to 16 object pointers split into C/mem_dim tokens each."""
        cfg = self.cfg
        mems, poss = [], []
        sel = [(0, o) for o in cond.values()] + [(tp, non_cond.get(t - (cfg.num_maskmem - tp))) for tp in range(1, cfg.num_maskmem)]
        for tp, prev in sel:
            if prev is None:
                continue
            mems.append(ops.cast(prev.maskmem, F32) if self.p_all else prev.maskmem)   # bf16 storage (the reference rounds it, sam2_video_predictor.py:967)
            poss.append(self.mem_pos_slot[cfg.num_maskmem - tp - 1])
        ptrs = [o.obj_ptr for ti, o in cond.items() if ti <= t]
        for d in range(1, min(num_frames, cfg.max_obj_ptrs_in_encoder)):
            if t - d < 0:
                break
            o = non_cond.get(t - d)
            if o is not None:
                ptrs.append(o.obj_ptr)
        n_ptr = 0
        if ptrs:
            split = cfg.d_model // cfg.mem_dim
            op = torch.stack(ptrs, dim=1).reshape(B, len(ptrs) * split, cfg.mem_dim)   # [B, P*4, mem_dim]
            mems.append(op if self.p_all else op.to(BF))                        # pointers stay fp32 in the reference; bf16 here is the GEMM operand rounding
            poss.append(torch.zeros((op.shape[1], cfg.mem_dim), device=self.dev))
            n_ptr = op.shape[1]
        return torch.cat(mems, dim=1).contiguous(), torch.cat(poss, dim=0).contiguous(), n_ptr

    # ------------------------------------------------------------------------------------------------ video session
    @torch.no_grad()
    def run_video(self, images: torch.Tensor, text_embeds: torch.Tensor, video_hw: tuple, binary: bool | str = False, collect: dict | None = None):
        """Graph-captured entry: the whole session's launch sequence is a function of (T, n_obj, video size) only. Returns a list of
        per-frame [n_obj, H, W] tensors (views of graph-owned memory, valid until the next call with the same shapes)."""
        if collect is not None:
            return self._run_video(images, text_embeds, video_hw, binary, collect)
        # the graph ends at the low-res mask logits [T, n_obj, 4fs, 4fs]; the upsampling to the video's own resolution (one launch for all planes)
        # stays outside, so that neither the resolution nor the output format is part of the key: eval sets change resolution per video
        key = (tuple(images.shape), tuple(text_embeds.shape), images.dtype)
        low = self.g_video(key, lambda images, text_embeds: torch.stack(self._run_video(images, text_embeds, None, False, None)),
                           images=images, text_embeds=text_embeds.to(F32))
        T, B, side = low.shape[0], low.shape[1], low.shape[-1]
        out = ops.bilinear(low.view(T * B, side, side), int(video_hw[0]), int(video_hw[1]), mode=_OUT_MODE[binary])
        return list(out.view(T, B, *out.shape[1:]).unbind(0))

    def _run_video(self, images: torch.Tensor, text_embeds: torch.Tensor, video_hw: tuple, binary: bool | str = False, collect: dict | None = None):
        """The VideoGLaMM video-branch session (VideoGLaMM.py:846-875 -> sam2_video_predictor.py init_state_from_tensor :109-180,
        add_new_text :416-495 on frame 0 per object, propagate_in_video :745-827 incl. preflight :675-729).
        images [T,3,S,S] fp32 cuda; text_embeds [n_obj,1,C]. Returns per frame [n_obj,H,W]: fp32 logits (binary=False) or uint8 (x>0); binary="packed": per frame uint8 [n_obj, ceil(H*W/8)],
        the (x>0) bits packed 8 per byte by the upsampling kernel (videoglamm_b200.parallel wire format)."""
        cfg = self.cfg
        T, B = images.shape[0], text_embeds.shape[0]
        HW, C = self.HW, cfg.d_model
        # video_hw None: return the low-res logits [n_obj, 4fs, 4fs] per frame (run_video upsamples all planes at once outside the graph)
        # no temporal dependence in the image encoder: encode all frames up front, `frames_per_encode` at a time (fills GEMM waves)
        feats = []
        for t0 in range(0, T, self.frames_per_encode):
            fb = self.forward_image(images[t0:t0 + self.frames_per_encode].float())
            nb = fb["top"].shape[0] // self.HW
            for i in range(nb):
                feats.append({k: v.view(nb, -1, v.shape[-1])[i] for k, v in fb.items()})
        f0 = feats[0]
        pix0 = ops.add(f0["top"], self.no_mem_embed)                               # init cond frame: + no_mem_embed (sam2_base.py:641-645)
        per_obj = [self.forward_sam_heads(pix0, f0["s0"], f0["s1"], text_embeds[o:o + 1], 1) for o in range(B)]
        low = torch.cat([r["low_res_masks"] for r in per_obj], 0)
        ptr = torch.cat([r["obj_ptr"] for r in per_obj], 0)
        cond = {0: FrameMem(low, ptr, self.encode_memory(f0["top"], low, binarize=True))}
        non_cond: dict[int, FrameMem] = {}
        outs = []
        for t in range(T):
            if t in cond:
                pred = cond[t].pred_masks
            else:
                f = feats[t]
                memory, mem_pos, n_ptr = self._memory_bank(t, cond, non_cond, T, B)
                pix = self.memory_attention(f["top"], B, memory, mem_pos, n_ptr)
                r = self.forward_sam_heads(pix, f["s0"], f["s1"], None, B)
                pred = r["low_res_masks"]
                non_cond[t] = FrameMem(pred, r["obj_ptr"], self.encode_memory(f["top"], pred, binarize=False))
                if collect is not None:
                    collect.setdefault("pix_feat_with_mem", {})[t] = pix
                for old in [k for k in non_cond if k < t - max(cfg.num_maskmem, cfg.max_obj_ptrs_in_encoder)]:
                    non_cond[old].maskmem = None                                   # ring buffer: nothing older is ever read again
            if collect is not None:
                collect.setdefault("low_res", {})[t] = pred
            outs.append(pred if video_hw is None else ops.bilinear(pred, int(video_hw[0]), int(video_hw[1]), mode=_OUT_MODE[binary]))
        return outs

    # ------------------------------------------------------------------------------------------------ framewise branch
    @torch.no_grad()
    def run_framewise(self, images: torch.Tensor, text_embeds: torch.Tensor, video_hw: tuple, binary: bool | str = False, zero_dense: bool = False):
        """SAM2 part of VideoGLaMM_SAM2.inference_framewise (VideoGLaMM.py:598-768 with get_visual_embs_sam2 :205-241): every frame is
        decoded independently — image_embed = top-level feature + no_mem_embed, prompt = the text token only (no padding points),
        multimask_output=False with the stability fallback (mask_decoder.py:247-295), repeat_image over the objects,
        postprocess_masks bilinear to the original size (:147-153). Returns per frame [n_obj, H, W] logits (or uint8 masks)."""
        cfg = self.cfg
        T, B = images.shape[0], text_embeds.shape[0]
        Hv, Wv = video_hw
        side = 4 * cfg.feat_size
        outs = []
        sparse = text_embeds.to(F32).reshape(B, 1, cfg.d_model)
        for t0 in range(0, T, self.frames_per_encode):
            fb = self.forward_image(images[t0:t0 + self.frames_per_encode].float())
            nb = fb["top"].shape[0] // self.HW
            for i in range(nb):
                f = {k: v.view(nb, -1, v.shape[-1])[i] for k, v in fb.items()}
                pix = ops.add(f["top"], self.no_mem_embed, rows=B * self.HW)
                masks, iou, _, _ = self.mask_decoder(pix, f["s0"], f["s1"], sparse, zero_dense=zero_dense)
                sel = ops.multimask_select(masks, iou.contiguous())
                low = masks[torch.arange(B, device=self.dev), sel].reshape(B, side, side).contiguous()
                outs.append(ops.bilinear(low, Hv, Wv, mode=_OUT_MODE[binary]))
        return outs

    # ------------------------------------------------------------------------------------------------ predictor surface
    @torch.no_grad()
    def init_state_from_tensor(self, images: torch.Tensor, video_height: int, video_width: int) -> dict:
        """SAM2VideoPredictor.init_state_from_tensor sam2_video_predictor.py:109-180: the frames are encoded up front (no one-frame cache)."""
        feats = []
        for t0 in range(0, images.shape[0], self.frames_per_encode):
            fb = self.forward_image(images[t0:t0 + self.frames_per_encode].to(self.dev).float())
            nb = fb["top"].shape[0] // self.HW
            feats += [{k: v.view(nb, -1, v.shape[-1])[i] for k, v in fb.items()} for i in range(nb)]
        return dict(feats=feats, num_frames=images.shape[0], video_height=video_height, video_width=video_width, obj_ids=[], per_obj={},
                    tracking_has_started=False)

    def reset_state(self, state: dict) -> None:
        state.update(obj_ids=[], per_obj={}, tracking_has_started=False)

    @torch.no_grad()
    def add_new_text(self, inference_state: dict, frame_idx: int, obj_id, text: torch.Tensor):
        """SAM2VideoPredictor.add_new_text :416-495 (the VideoGLaMM fork's text prompt) on the conditioning frame 0.
        Returns (frame_idx, obj_ids, video_res_masks fp32 [n_obj,1,H,W]) like the reference."""
        st = inference_state
        if frame_idx != 0:
            raise NotImplementedError("text prompts are supported on frame 0 only (the only call pattern of VideoGLaMM.py:846-863)")
        if st["tracking_has_started"] and obj_id not in st["per_obj"]:
            raise RuntimeError(f"Cannot add new object id {obj_id} after tracking starts.")
        f0 = st["feats"][0]
        pix0 = ops.add(f0["top"], self.no_mem_embed)
        r = self.forward_sam_heads(pix0, f0["s0"], f0["s1"], text.to(self.dev, F32).reshape(1, 1, -1), 1)
        if obj_id not in st["per_obj"]:
            st["obj_ids"].append(obj_id)
        st["per_obj"][obj_id] = r
        low = torch.cat([st["per_obj"][o]["low_res_masks"] for o in st["obj_ids"]], 0)
        masks = ops.bilinear(low, st["video_height"], st["video_width"])
        return frame_idx, list(st["obj_ids"]), masks[:, None]

    @torch.no_grad()
    def propagate_in_video(self, inference_state: dict):
        """SAM2VideoPredictor.propagate_in_video :745-827 (forward, from frame 0): generator of (frame_idx, obj_ids, masks [n_obj,1,H,W])."""
        st = inference_state
        if not st["obj_ids"]:
            raise RuntimeError("No points are provided; please add points first")
        st["tracking_has_started"] = True
        B, T = len(st["obj_ids"]), st["num_frames"]
        Hv, Wv = st["video_height"], st["video_width"]
        f0 = st["feats"][0]
        low = torch.cat([st["per_obj"][o]["low_res_masks"] for o in st["obj_ids"]], 0)
        ptr = torch.cat([st["per_obj"][o]["obj_ptr"] for o in st["obj_ids"]], 0)
        cond = {0: FrameMem(low, ptr, self.encode_memory(f0["top"], low, binarize=True))}
        non_cond: dict[int, FrameMem] = {}
        for t in range(T):
            if t in cond:
                pred = cond[t].pred_masks
            else:
                f = st["feats"][t]
                memory, mem_pos, n_ptr = self._memory_bank(t, cond, non_cond, T, B)
                pix = self.memory_attention(f["top"], B, memory, mem_pos, n_ptr)
                r = self.forward_sam_heads(pix, f["s0"], f["s1"], None, B)
                pred = r["low_res_masks"]
                non_cond[t] = FrameMem(pred, r["obj_ptr"], self.encode_memory(f["top"], pred, binarize=False))
            yield t, list(st["obj_ids"]), ops.bilinear(pred, Hv, Wv)[:, None]
