"""Typed configuration of the engine (the constants the reference spreads over Hydra YAML, HF config.json and python configs)."""
from __future__ import annotations

from dataclasses import dataclass


@dataclass
class Sam2Config:
    """sam2_hiera_l.yaml of the reference (segment_anything_2/sam2_configs/sam2_hiera_l.yaml) as plain fields."""
    image_size: int = 1024
    embed_dim: int = 144
    num_heads: int = 2
    stages: tuple = (2, 6, 36, 4)
    global_att_blocks: tuple = (23, 33, 43)
    window_pos_embed_bkg_spatial_size: tuple = (7, 7)
    window_spec: tuple = (8, 4, 16, 8)
    d_model: int = 256
    mem_dim: int = 64
    num_maskmem: int = 7
    max_obj_ptrs_in_encoder: int = 16
    mem_attn_layers: int = 4
    dec_depth: int = 2
    dec_heads: int = 8
    sigmoid_scale_for_mem_enc: float = 20.0
    sigmoid_bias_for_mem_enc: float = -10.0
    q_pool: int = 3

    @property
    def stage_ends(self) -> list[int]:
        return [sum(self.stages[:i]) - 1 for i in range(1, len(self.stages) + 1)]

    def block_specs(self) -> list[tuple]:
        """(dim, dim_out, heads, window, has_q_stride) per Hiera block."""
        q_pool_blocks = [x + 1 for x in self.stage_ends[:-1]][: self.q_pool]
        specs, dim, heads, stage = [], self.embed_dim, self.num_heads, 0
        for i in range(sum(self.stages)):
            window = 0 if i in self.global_att_blocks else self.window_spec[stage]
            dim_out = dim
            if i - 1 in self.stage_ends:
                dim_out, heads, stage = dim * 2, heads * 2, stage + 1
            specs.append((dim, dim_out, heads, window, i in q_pool_blocks))
            dim = dim_out
        return specs

    @property
    def feat_size(self) -> int:
        return self.image_size // 16

    @property
    def channel_list(self) -> list[int]:
        specs = self.block_specs()
        return [specs[i][1] for i in self.stage_ends[::-1]]

    mem_attn_ffn: int = 2048
    dec_mlp: int = 2048


@dataclass
class ClipConfig:
    """HF CLIPVisionConfig fields used (openai/clip-vit-large-patch14-336)."""
    image_size: int = 336
    patch: int = 14
    width: int = 1024
    layers: int = 24
    heads: int = 16
    mlp: int = 4096
    select_layer: int = -2

    @property
    def grid(self) -> int:
        return self.image_size // self.patch


@dataclass
class Iv2Config:
    """InternVideo2-Stage2_1B-224p-f4 vision encoder."""
    image_size: int = 224
    patch: int = 14
    width: int = 1408
    depth: int = 40
    heads: int = 16
    mlp: int = 6144
    frames: int = 4

    @property
    def grid(self) -> int:
        return self.image_size // self.patch


@dataclass
class PhiConfig:
    """HF Phi3Config fields used (Phi-3-mini-4k-instruct)."""
    vocab: int = 32064
    hidden: int = 3072
    layers: int = 32
    heads: int = 32
    inter: int = 8192
    rope_theta: float = 10000.0
    eps: float = 1e-5
    sliding_window: int | None = 2047
    max_len: int = 4096

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads


@dataclass
class LmmConfig:
    clip: ClipConfig
    iv2: Iv2Config
    phi: PhiConfig
    out_dim: int = 256
    seg_token_idx: int = 32011
    proj_depth: int = 2


def make_lmm_config(clip: dict | None = None, iv2: dict | None = None, phi: dict | None = None, **kw) -> LmmConfig:
    return LmmConfig(ClipConfig(**(clip or {})), Iv2Config(**(iv2 or {})), PhiConfig(**(phi or {})), **kw)
