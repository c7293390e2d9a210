"""Shared test helpers: build oracle models with seeded weights and hand the same weights to the
CUDA product through its StateDictSource."""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from seeding import seeded_init  # noqa: E402

from oracle import config as ocfg  # noqa: E402
from oracle import lora as olora  # noqa: E402
from oracle import model as om  # noqa: E402


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def export_state(module):
    """(state dict with peft's '.base_layer' stripped, lora dict '<module>.lora_{A,B}.weight')."""
    sd, lora = {}, {}
    for k, v in module.state_dict().items():
        if ".lora_A." in k or ".lora_B." in k:
            lora[k] = v
        else:
            sd[k.replace(".base_layer.", ".")] = v
    return sd, lora


def build_oracle(cfg, seed=0, lora_alpha=None):
    """oracle UNet (+adapter, +previewer LoRA) and Aggregator with every tensor seeded non-zero."""
    unet = om.load_adapter(om.UNet2DConditionModel(cfg))
    if lora_alpha is not None:
        olora.add_previewer_lora(unet, cfg.lora_rank, lora_alpha)
    seeded_init(unet, seed)
    agg = om.Aggregator(cfg)
    om.remove_attn2(agg)
    seeded_init(agg, seed + 1)
    # keep the injected residuals O(1) relative to the skips they are added to
    with torch.no_grad():
        for head in list(agg.controlnet_down_blocks) + [agg.controlnet_mid_block]:
            head[1].weight.mul_(0.5)
    return unet.eval(), agg.eval()


def make_inputs(cfg, B=1, h=32, w=32, seed=1234):
    """Seeded synthetic conditioning.  Positive and negative prompt embeddings share a large common
    component (real CLIP embeddings of two prompts are strongly correlated); with independent random
    prompts the guided eps = e_u + g (e_c - e_u) would be ~5x the latent norm, which no trained model
    produces and which multiplies every rounding error by the same factor."""
    g = torch.Generator().manual_seed(seed)

    def r(*s):
        return torch.randn(*s, generator=g)

    base_p, base_q = r(B, cfg.text_seq_len, cfg.cross_attention_dim), r(B, cfg.pooled_dim)
    return dict(
        image=r(B, 4, h, w) * 0.8,
        prompt_embeds=base_p + 0.15 * r(B, cfg.text_seq_len, cfg.cross_attention_dim),
        negative_prompt_embeds=base_p + 0.15 * r(B, cfg.text_seq_len, cfg.cross_attention_dim),
        pooled_prompt_embeds=base_q + 0.15 * r(B, cfg.pooled_dim),
        negative_pooled_prompt_embeds=base_q + 0.15 * r(B, cfg.pooled_dim),
        ip=torch.stack([0.3 * r(B, cfg.image_seq_len, cfg.image_embed_dim), r(B, cfg.image_seq_len, cfg.image_embed_dim)]),
        time_ids=torch.tensor([[h * 8.0, w * 8.0, 0.0, 0.0, h * 8.0, w * 8.0]]).repeat(B, 1),
    )
