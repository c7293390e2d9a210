"""Architecture configuration of the denoising step (plain data).

``sdxl()`` = SDXL-base UNet + InstantIR adapter geometry (BASELINE.json configs 2-5);
``tiny()`` = BASELINE.json config 1 (same topology, scaled-down widths, head_dim 64).
Field names follow diffusers' UNet2DConditionModel config where one exists.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Tuple


@dataclass
class ModelConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    down_block_types: Tuple[str, ...] = ("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20)
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    pooled_dim: int = 1280
    time_embed_dim: int = 1280
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    num_ip_tokens: int = 64
    image_embed_dim: int = 1024
    image_seq_len: int = 257
    resampler_dim: int = 1280
    resampler_depth: int = 4
    resampler_heads: int = 20
    resampler_dim_head: int = 64
    resampler_ff_mult: int = 4
    ip_scale: float = 1.0
    sft_hidden: int = 128
    lora_rank: int = 64
    text_seq_len: int = 77

    def __post_init__(self):
        self.block_out_channels = tuple(self.block_out_channels)
        self.down_block_types = tuple(self.down_block_types)
        self.transformer_layers_per_block = tuple(self.transformer_layers_per_block)
        self.num_attention_heads = tuple(self.num_attention_heads)
        for c, h in zip(self.block_out_channels, self.num_attention_heads):
            if c != 64 * h:
                raise ValueError(f"head_dim must be 64 (the sm_100a attention kernel): C={c}, heads={h}")
        if self.resampler_dim_head != 64:
            raise ValueError("resampler_dim_head must be 64")

    @property
    def projection_class_embeddings_input_dim(self) -> int:
        return self.pooled_dim + 6 * self.addition_time_embed_dim

    def to_dict(self):
        return asdict(self)


def sdxl() -> ModelConfig:
    return ModelConfig()


def tiny() -> ModelConfig:
    return ModelConfig(
        block_out_channels=(64, 128, 256),
        transformer_layers_per_block=(1, 1, 2),
        num_attention_heads=(1, 2, 4),
        cross_attention_dim=256,
        addition_time_embed_dim=32,
        pooled_dim=64,
        time_embed_dim=256,
        num_ip_tokens=16,
        image_embed_dim=64,
        image_seq_len=33,
        resampler_dim=128,
        resampler_depth=2,
        resampler_heads=2,
        resampler_dim_head=64,
        lora_rank=8,
    )
