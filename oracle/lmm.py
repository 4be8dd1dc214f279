"""TEST INFRASTRUCTURE (oracle) — CPU fp32 restatement of the LMM side of the path: CLIP ViT-L/14-336 (HF `CLIPVisionModel`,
hidden_states[-2]), InternVideo2-1B video tower, V-L projectors + pooling, multimodal splice, Phi-3 decoder, greedy decode
with use_cache=False semantics, [SEG] hidden-state extraction and the L-V projection.

Not product code (see oracle/sam2.py header). Third-party arithmetic that is NOT under /root/reference: HuggingFace
`transformers` (pinned ==4.41.0 in /root/reference/README.md:76; 5.5.0 installed here) provides CLIPVisionModel and
Phi3Model; their published algorithms are restated below and pinned against the installed classes in
tests/test_oracle_vs_reference.py. Paths cited are relative to /root/reference/VideoGLaMM/.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = -200  # model/videogpt_plus/constants.py
CHUNK = 4                 # CHUNK_SIZE, model/videogpt_plus/constants.py:6


@dataclass
class ClipCfg:
    """openai/clip-vit-large-patch14-336 vision tower."""
    image_size: int = 336
    patch: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    select_layer: int = -2  # mm_vision_select_layer

    @property
    def grid(self):
        return self.image_size // self.patch


@dataclass
class Iv2Cfg:
    """InternVideo2-Stage2_1B-224p-f4 vision encoder (model/videogpt_plus/model/internvideo/internvideo2.py:698-723)."""
    image_size: int = 224
    patch: int = 14
    width: int = 1408
    depth: int = 40
    heads: int = 16
    mlp: int = 6144
    frames: int = 4

    @property
    def grid(self):
        return self.image_size // self.patch


@dataclass
class PhiCfg:
    """Phi-3-mini-4k-instruct."""
    vocab: int = 32064
    hidden: int = 3072
    layers: int = 32
    heads: int = 32
    inter: int = 8192
    rope_theta: float = 10000.0
    eps: float = 1e-5
    sliding_window: int | None = 2047

    @property
    def head_dim(self):
        return self.hidden // self.heads


@dataclass
class LmmCfg:
    clip: ClipCfg
    iv2: Iv2Cfg
    phi: PhiCfg
    out_dim: int = 256
    seg_token_idx: int = 32011
    proj_depth: int = 2  # mlp2x_gelu


TINY_LMM = dict(clip=dict(image_size=56, width=64, layers=3, heads=4, mlp=128), iv2=dict(image_size=56, width=64, depth=4, heads=4, mlp=128),
                phi=dict(vocab=320, hidden=128, layers=2, heads=4, inter=256, sliding_window=None))


def make_cfg(clip=None, iv2=None, phi=None, **kw) -> LmmCfg:
    return LmmCfg(ClipCfg(**(clip or {})), Iv2Cfg(**(iv2 or {})), PhiCfg(**(phi or {})), **kw)


def lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


# ------------------------------------------------------------------------------------------------ CLIP (HF CLIPVisionModel)
def clip_tower(sd, p, cfg: ClipCfg, images):
    """CLIPVisionTower.forward + feature_select('patch') model/videogpt_plus/model/multimodal_encoder/clip_encoder.py:34-72:
    HF CLIPVisionTransformer embeddings -> pre_layrnorm -> encoder layers; hidden_states[select_layer][:, 1:]. images [T,3,S,S]."""
    v = p + ".vision_model"
    x = F.conv2d(images, sd[v + ".embeddings.patch_embedding.weight"], None, stride=cfg.patch).flatten(2).transpose(1, 2)
    cls = sd[v + ".embeddings.class_embedding"].expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], 1) + sd[v + ".embeddings.position_embedding.weight"][None]
    x = F.layer_norm(x, (cfg.width,), sd[v + ".pre_layrnorm.weight"], sd[v + ".pre_layrnorm.bias"], 1e-5)
    n_run = cfg.layers + 1 + cfg.select_layer  # hidden_states has layers+1 entries; [-2] = after layers-1 blocks
    hd = cfg.width // cfg.heads
    for i in range(n_run):
        l = f"{v}.encoder.layers.{i}"
        h = F.layer_norm(x, (cfg.width,), sd[l + ".layer_norm1.weight"], sd[l + ".layer_norm1.bias"], 1e-5)
        B, N, _ = h.shape
        q, k, vv = (lin(sd, f"{l}.self_attn.{n}_proj", h).view(B, N, cfg.heads, hd).transpose(1, 2) for n in "qkv")
        a = torch.softmax((q * hd ** -0.5) @ k.transpose(-1, -2), -1) @ vv
        x = x + lin(sd, l + ".self_attn.out_proj", a.transpose(1, 2).reshape(B, N, -1))
        h = F.layer_norm(x, (cfg.width,), sd[l + ".layer_norm2.weight"], sd[l + ".layer_norm2.bias"], 1e-5)
        h = lin(sd, l + ".mlp.fc1", h)
        x = x + lin(sd, l + ".mlp.fc2", h * torch.sigmoid(1.702 * h))  # quick_gelu
    return x[:, 1:]


# ------------------------------------------------------------------------------------------------ InternVideo2
def rms_norm(x, w, eps):
    """RMSNorm model/videogpt_plus/model/internvideo/internvideo2.py:134-145 (== HF Phi3RMSNorm)."""
    xf = x.float()
    return w * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)


def iv2_tower(sd, p, cfg: Iv2Cfg, chunks):
    """InternVideo2_Stage2V.forward internvideo/utils.py:229-238 -> PretrainInternVideo2.forward(x_vis_only=True, x_vis_return_idx=-2)
    internvideo2.py:585-651; Block (non-fused) :306-310; Attention._naive_attn :190-209 (q/k RMSNorm over the full width).
    chunks [n_chunks, 4, 3, S, S] -> [n_chunks, 1 + 4*grid^2, width] (cls kept; caller drops it, arch.py:145)."""
    e = p + ".vision_encoder"
    x = chunks.permute(0, 2, 1, 3, 4)  # B,C,T,H,W
    x = F.conv3d(x, sd[e + ".patch_embed.proj.weight"], sd[e + ".patch_embed.proj.bias"], stride=(1, cfg.patch, cfg.patch))
    x = x.flatten(3).permute(0, 2, 3, 1)  # B,T,HW,C
    B = x.shape[0]
    x = x.reshape(B, -1, cfg.width)
    x = torch.cat([sd[e + ".cls_token"].expand(B, -1, -1), x], 1) + sd[e + ".pos_embed"]  # T=4 > 1 -> video pos_embed
    hd = cfg.width // cfg.heads
    for i in range(cfg.depth - 1):  # break after idx == depth - 2
        b = f"{e}.blocks.{i}"
        h = rms_norm(x, sd[b + ".norm1.weight"], 1e-6)
        N = h.shape[1]
        qkv = F.linear(h, sd[b + ".attn.qkv.weight"]).reshape(B, N, 3, cfg.heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        q = rms_norm(q.transpose(1, 2).flatten(-2, -1), sd[b + ".attn.q_norm.weight"], 1e-6).view(B, N, cfg.heads, hd).transpose(1, 2)
        k = rms_norm(k.transpose(1, 2).flatten(-2, -1), sd[b + ".attn.k_norm.weight"], 1e-6).view(B, N, cfg.heads, hd).transpose(1, 2)
        a = torch.softmax((q * hd ** -0.5) @ k.transpose(-2, -1), -1) @ v
        a = lin(sd, b + ".attn.proj", a.transpose(1, 2).reshape(B, N, cfg.width))
        x = x + a * sd[b + ".ls1.gamma"]
        h = rms_norm(x, sd[b + ".norm2.weight"], 1e-6)
        h = lin(sd, b + ".mlp.fc2", F.gelu(lin(sd, b + ".mlp.fc1", h)))
        x = x + h * sd[b + ".ls2.gamma"]
    return x


# ------------------------------------------------------------------------------------------------ projectors + splice
def projector(sd, p, x, depth):
    """mlpNx_gelu model/videogpt_plus/model/multimodal_projector/builder.py:38-46"""
    x = lin(sd, p + ".0", x)
    for j in range(1, depth):
        x = lin(sd, f"{p}.{2 * j}", F.gelu(x))
    return x


def avg_pool_tokens(x, out_side):
    """apply_adaptive_avg_pooling model/videogpt_plus/model/arch.py:88-96"""
    b, n, c = x.shape
    h = int(math.sqrt(n))
    x = F.adaptive_avg_pool2d(x.permute(0, 2, 1).reshape(b, c, h, h), (out_side, out_side))
    return x.flatten(2).transpose(1, 2)


def visual_tokens(sd, cfg: LmmCfg, images, context_images):
    """encode_videos arch.py:121-151 + project(input_type='video') :164-191. images [T,3,224,224] (IV2), context_images
    [T,3,336,336] (CLIP) -> [T*(gc/2)^2 + T*(gv/2)^2, hidden] with the CONTEXT block first."""
    if context_images is None:   # still image: encode_images arch.py:110-119 (CLIP tower, patch features) + project(input_type="image") :193-203 (no pooling)
        ctx = clip_tower(sd, "model.image_vision_tower.vision_tower", cfg.clip, images)
        return projector(sd, "model.image_mm_projector", ctx, cfg.proj_depth).reshape(-1, cfg.phi.hidden)
    T = images.shape[0]
    vid = iv2_tower(sd, "model.vision_tower", cfg.iv2, images.reshape(T // CHUNK, CHUNK, *images.shape[1:]))[:, 1:]
    ctx = clip_tower(sd, "model.image_vision_tower.vision_tower", cfg.clip, context_images)
    vid = projector(sd, "model.mm_projector", vid, cfg.proj_depth)                      # [chunks, 4*gv^2, D]
    vid = vid.reshape(T, cfg.iv2.grid ** 2, -1)
    vid = avg_pool_tokens(vid, cfg.iv2.grid // 2).reshape(-1, vid.shape[-1])
    ctx = projector(sd, "model.image_mm_projector", ctx, cfg.proj_depth)
    ctx = avg_pool_tokens(ctx, cfg.clip.grid // 2).reshape(-1, ctx.shape[-1])
    return torch.cat([ctx, vid], 0)


def splice(sd, cfg: LmmCfg, input_ids, vis):
    """prepare_inputs_labels_for_multimodal arch.py:283-379 (video case, no im_start/end): embed(ids before the run of -200)
    + visual block + embed(ids after). input_ids [L_txt] int64 -> [L, hidden]."""
    idx = torch.where(input_ids == IMAGE_TOKEN_INDEX)[0]
    start, end = int(idx[0]), int(idx[-1])
    emb = sd["model.embed_tokens.weight"]
    return torch.cat([emb[input_ids[:start]], vis, emb[input_ids[end + 1:]]], 0)


# ------------------------------------------------------------------------------------------------ Phi-3
def phi3_forward(sd, cfg: PhiCfg, x):
    """HF Phi3Model.forward on inputs_embeds x [L, hidden] (called at model/videogpt_plus/model/language_model/phi3.py:72-81):
    causal (optionally sliding-window, transformers 4.41 rule: keys j >= i - sliding_window) MHA with rotate-half RoPE, SwiGLU MLP,
    RMSNorm. Returns the final-norm hidden states [L, hidden]."""
    L = x.shape[0]
    hd, H = cfg.head_dim, cfg.heads
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    ang = torch.arange(L, dtype=torch.float32)[:, None] * inv[None]
    cos, sin = torch.cat([ang, ang], -1).cos()[None], torch.cat([ang, ang], -1).sin()[None]
    i, j = torch.arange(L)[:, None], torch.arange(L)[None]
    mask = j <= i
    if cfg.sliding_window is not None:
        mask = mask & (j >= i - cfg.sliding_window)
    rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., : hd // 2]], -1)
    for l in range(cfg.layers):
        p = f"model.layers.{l}"
        h = rms_norm(x, sd[p + ".input_layernorm.weight"], cfg.eps)
        qkv = F.linear(h, sd[p + ".self_attn.qkv_proj.weight"])
        q, k, v = (qkv[:, n * cfg.hidden:(n + 1) * cfg.hidden].view(L, H, hd).transpose(0, 1) for n in range(3))
        q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
        s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        s = s.masked_fill(~mask, float("-inf"))
        a = torch.softmax(s.float(), -1) @ v
        x = x + F.linear(a.transpose(0, 1).reshape(L, -1), sd[p + ".self_attn.o_proj.weight"])
        h = rms_norm(x, sd[p + ".post_attention_layernorm.weight"], cfg.eps)
        gu = F.linear(h, sd[p + ".mlp.gate_up_proj.weight"])
        gate, up = gu.chunk(2, dim=-1)
        x = x + F.linear(up * F.silu(gate), sd[p + ".mlp.down_proj.weight"])
    return rms_norm(x, sd["model.norm.weight"], cfg.eps)


def generate(sd, cfg: LmmCfg, input_ids, images, context_images, max_new_tokens, eos_ids=(), forced: dict | None = None):
    """VideoGLaMM_SAM2.inference_video_branch VideoGLaMM.py:782-831 (steps A-D): greedy HF generate with use_cache=False
    (every step re-runs the full multimodal forward; here the visual tokens are computed once since they are identical each step),
    [SEG] mask shifted by one (:803-806), text_hidden_fcs on the LAST step's last-layer hidden states (:813), gather.
    `forced` {step: token_id} overwrites the sampled token at that step. Returns (output_ids, pred_embeddings, hidden)."""
    vis = visual_tokens(sd, cfg, images, context_images)
    ids = input_ids.clone()
    hidden = None
    for step in range(max_new_tokens):
        hidden = phi3_forward(sd, cfg.phi, splice(sd, cfg, ids, vis))
        logits = F.linear(hidden[-1], sd["lm_head.weight"])
        nxt = int(torch.argmax(logits))
        if forced and step in forced:
            nxt = forced[step]
        ids = torch.cat([ids, torch.tensor([nxt])])
        if nxt in eos_ids:
            break
    n_new = hidden.shape[0] - (ids.shape[0] - 1)          # num_newly_added_tokens (VideoGLaMM.py:786)
    seg = torch.cat([torch.zeros(n_new, dtype=torch.bool), ids[1:] == cfg.seg_token_idx])
    feats = lin(sd, "model.text_hidden_fcs.0.2", F.relu(lin(sd, "model.text_hidden_fcs.0.0", hidden)))
    return ids, feats[seg], hidden
