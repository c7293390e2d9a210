"""Architecture configuration shared by the oracle modules (plain data, no behaviour).

``sdxl()`` is the SDXL-base UNet + InstantIR adapter geometry;
``tiny()`` is BASELINE.json config 1: same topology, scaled-down widths, head_dim kept at 64.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import Tuple


@dataclass
class StepConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    down_block_types: Tuple[str, ...] = ("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D")
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    num_attention_heads: Tuple[int, ...] = (5, 10, 20)  # diffusers' "attention_head_dim" for SDXL
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    pooled_dim: int = 1280
    time_embed_dim: int = 1280  # 4 * block_out_channels[0] in diffusers
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    # IP adapter / Resampler (module/ip_adapter/utils.py:138-157)
    num_ip_tokens: int = 64
    image_embed_dim: int = 1024      # DINOv2-L hidden size
    image_seq_len: int = 257
    resampler_dim: int = 1280
    resampler_depth: int = 4
    resampler_heads: int = 20
    resampler_dim_head: int = 64
    resampler_ff_mult: int = 4
    ip_scale: float = 1.0
    # Aggregator SFT hidden width (module/aggregator.py:61)
    sft_hidden: int = 128
    # previewer LoRA (pipelines/sdxl_instantir.py:376-381)
    lora_rank: int = 64
    text_seq_len: int = 77

    @property
    def projection_class_embeddings_input_dim(self) -> int:
        return self.pooled_dim + 6 * self.addition_time_embed_dim

    def to_dict(self):
        return asdict(self)


def sdxl() -> StepConfig:
    return StepConfig()


def tiny() -> StepConfig:
    """BASELINE.json configs[0]: scaled-down SDXL-architecture step that the CPU finishes in seconds."""
    return StepConfig(
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
