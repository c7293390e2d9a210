"""Parameter inventory and weight sources.

The state-dict key layout is the reference's: diffusers SDXL UNet keys,
``...attn2.processor.{to_k_ip,to_v_ip,ln_k_ip.linear,ln_v_ip.linear}`` for the IP-adapter
processors (module/ip_adapter/attention_processor.py:1089-1092), the Resampler under
``encoder_hid_proj.image_projection_layers.0`` (module/ip_adapter/utils.py:138-157), Aggregator keys
of module/aggregator.py:414-471, and peft-style ``<module>.lora_A.weight / lora_B.weight`` for the
previewer LoRA (pipelines/sdxl_instantir.py:141-162,376-385).

A *source* hands out fp32 tensors by key: ``StateDictSource`` wraps a loaded checkpoint,
``RandomSource`` synthesises random-init weights of the right shapes on the device (there is no
network for checkpoints; BASELINE.json asks for random-init weights of the architecture).
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from .config import ModelConfig

PREVIEWER_LORA_MODULES = [
    "to_q", "to_kv", "0.to_out", "attn1.to_k", "attn1.to_v", "to_k_ip", "to_v_ip", "ln_k_ip.linear",
    "ln_v_ip.linear", "to_out.0", "proj_in", "proj_out", "ff.net.0.proj", "ff.net.2", "conv1", "conv2",
    "conv_shortcut", "downsamplers.0.conv", "upsamplers.0.conv", "time_emb_proj",
]

Shapes = "OrderedDict[str, Tuple[int, ...]]"


def _lin(d, name, n_in, n_out, bias=True):
    d[name + ".weight"] = (n_out, n_in)
    if bias:
        d[name + ".bias"] = (n_out,)


def _conv(d, name, c_in, c_out, k):
    d[name + ".weight"] = (c_out, c_in, k, k)
    d[name + ".bias"] = (c_out,)


def _norm(d, name, c):
    d[name + ".weight"] = (c,)
    d[name + ".bias"] = (c,)


def _resnet(d, p, cfg, c_in, c_out):
    _norm(d, p + ".norm1", c_in)
    _conv(d, p + ".conv1", c_in, c_out, 3)
    _lin(d, p + ".time_emb_proj", cfg.time_embed_dim, c_out)
    _norm(d, p + ".norm2", c_out)
    _conv(d, p + ".conv2", c_out, c_out, 3)
    if c_in != c_out:
        _conv(d, p + ".conv_shortcut", c_in, c_out, 1)


def _attention(d, p, c, kv_dim):
    _lin(d, p + ".to_q", c, c, bias=False)
    _lin(d, p + ".to_k", kv_dim, c, bias=False)
    _lin(d, p + ".to_v", kv_dim, c, bias=False)
    _lin(d, p + ".to_out.0", c, c)


def _t2d(d, p, cfg, c, n_layers, cross, adapter):
    _norm(d, p + ".norm", c)
    _lin(d, p + ".proj_in", c, c)
    for k in range(n_layers):
        b = f"{p}.transformer_blocks.{k}"
        _norm(d, b + ".norm1", c)
        _attention(d, b + ".attn1", c, c)
        if cross:
            _norm(d, b + ".norm2", c)
            _attention(d, b + ".attn2", c, cfg.cross_attention_dim)
            if adapter:
                q = b + ".attn2.processor"
                _lin(d, q + ".to_k_ip", cfg.cross_attention_dim, c, bias=False)
                _lin(d, q + ".to_v_ip", cfg.cross_attention_dim, c, bias=False)
                _lin(d, q + ".ln_k_ip.linear", cfg.time_embed_dim, 2 * c)
                _lin(d, q + ".ln_v_ip.linear", cfg.time_embed_dim, 2 * c)
        _norm(d, b + ".norm3", c)
        _lin(d, b + ".ff.net.0.proj", c, 8 * c)
        _lin(d, b + ".ff.net.2", 4 * c, c)
    _lin(d, p + ".proj_out", c, c)


def _embeddings(d, cfg):
    ch0 = cfg.block_out_channels[0]
    _lin(d, "time_embedding.linear_1", ch0, cfg.time_embed_dim)
    _lin(d, "time_embedding.linear_2", cfg.time_embed_dim, cfg.time_embed_dim)
    _lin(d, "add_embedding.linear_1", cfg.projection_class_embeddings_input_dim, cfg.time_embed_dim)
    _lin(d, "add_embedding.linear_2", cfg.time_embed_dim, cfg.time_embed_dim)


def _down_and_mid(d, cfg, cross, adapter):
    ch = cfg.block_out_channels
    out = ch[0]
    for i, t in enumerate(cfg.down_block_types):
        inp, out = out, ch[i]
        for j in range(cfg.layers_per_block):
            _resnet(d, f"down_blocks.{i}.resnets.{j}", cfg, inp if j == 0 else out, out)
            if t == "CrossAttnDownBlock2D":
                _t2d(d, f"down_blocks.{i}.attentions.{j}", cfg, out, cfg.transformer_layers_per_block[i], cross, adapter)
        if i != len(ch) - 1:
            _conv(d, f"down_blocks.{i}.downsamplers.0.conv", out, out, 3)
    _resnet(d, "mid_block.resnets.0", cfg, ch[-1], ch[-1])
    _t2d(d, "mid_block.attentions.0", cfg, ch[-1], cfg.transformer_layers_per_block[-1], cross, adapter)
    _resnet(d, "mid_block.resnets.1", cfg, ch[-1], ch[-1])


def resampler_param_shapes(cfg: ModelConfig, prefix="encoder_hid_proj.image_projection_layers.0"):
    d = OrderedDict()
    dim, inner = cfg.resampler_dim, cfg.resampler_dim_head * cfg.resampler_heads
    d[prefix + ".latents"] = (1, cfg.num_ip_tokens, dim)
    _lin(d, prefix + ".proj_in", cfg.image_embed_dim, dim)
    _lin(d, prefix + ".proj_out", dim, cfg.cross_attention_dim)
    _norm(d, prefix + ".norm_out", cfg.cross_attention_dim)
    for i in range(cfg.resampler_depth):
        a = f"{prefix}.layers.{i}.0"
        _norm(d, a + ".norm1", dim)
        _norm(d, a + ".norm2", dim)
        _lin(d, a + ".to_q", dim, inner, bias=False)
        _lin(d, a + ".to_kv", dim, 2 * inner, bias=False)
        _lin(d, a + ".to_out", inner, dim, bias=False)
        f = f"{prefix}.layers.{i}.1"
        _norm(d, f + ".0", dim)
        _lin(d, f + ".1", dim, dim * cfg.resampler_ff_mult, bias=False)
        _lin(d, f + ".3", dim * cfg.resampler_ff_mult, dim, bias=False)
    return d


def unet_param_shapes(cfg: ModelConfig, adapter: bool = True):
    d = OrderedDict()
    ch = cfg.block_out_channels
    _conv(d, "conv_in", cfg.in_channels, ch[0], 3)
    _embeddings(d, cfg)
    _down_and_mid(d, cfg, cross=True, adapter=adapter)
    rch, rtx = list(reversed(ch)), list(reversed(cfg.transformer_layers_per_block))
    rtypes = list(reversed(cfg.down_block_types))
    out = rch[0]
    n = cfg.layers_per_block + 1
    for i in range(len(ch)):
        prev, out = out, rch[i]
        inp = rch[min(i + 1, len(ch) - 1)]
        for j in range(n):
            skip = inp if j == n - 1 else out
            _resnet(d, f"up_blocks.{i}.resnets.{j}", cfg, (prev if j == 0 else out) + skip, out)
            if rtypes[i] == "CrossAttnDownBlock2D":
                _t2d(d, f"up_blocks.{i}.attentions.{j}", cfg, out, rtx[i], True, adapter)
        if i != len(ch) - 1:
            _conv(d, f"up_blocks.{i}.upsamplers.0.conv", out, out, 3)
    _norm(d, "conv_norm_out", ch[0])
    _conv(d, "conv_out", ch[0], cfg.out_channels, 3)
    if adapter:
        d.update(resampler_param_shapes(cfg))
    return d


def aggregator_param_shapes(cfg: ModelConfig):
    """Aggregator after remove_attn2 (no attn2/norm2 keys; gradio_demo/app.py:81 loads strict)."""
    d = OrderedDict()
    ch = cfg.block_out_channels
    _conv(d, "conv_in", cfg.in_channels, ch[0], 3)
    _conv(d, "ref_conv_in", cfg.in_channels, ch[0], 3)
    _embeddings(d, cfg)
    _down_and_mid(d, cfg, cross=False, adapter=False)

    def head(p, c):
        _conv(d, p + ".0.mlp_shared.0", c, cfg.sft_hidden, 3)
        _conv(d, p + ".0.mul", cfg.sft_hidden, c, 3)
        _conv(d, p + ".0.add", cfg.sft_hidden, c, 3)
        _conv(d, p + ".1", c, c, 1)

    idx = 0
    head(f"controlnet_down_blocks.{idx}", ch[0])
    for i in range(len(ch)):
        for _ in range(cfg.layers_per_block + (0 if i == len(ch) - 1 else 1)):
            idx += 1
            head(f"controlnet_down_blocks.{idx}", ch[i])
    head("controlnet_mid_block", ch[-1])
    return d


def lora_targets(shapes) -> "list[str]":
    """module names (without .weight) the previewer LoRA wraps — peft's suffix rule."""
    mods = sorted({k[: -len(".weight")] for k in shapes if k.endswith(".weight") and len(shapes[k]) in (2, 4)})
    return [m for m in mods if any(m == t or m.endswith("." + t) for t in PREVIEWER_LORA_MODULES)]


def lora_param_shapes(cfg: ModelConfig, shapes):
    d = OrderedDict()
    r = cfg.lora_rank
    for m in lora_targets(shapes):
        s = shapes[m + ".weight"]
        if len(s) == 2:
            d[m + ".lora_A.weight"] = (r, s[1])
            d[m + ".lora_B.weight"] = (s[0], r)
        else:
            d[m + ".lora_A.weight"] = (r, s[1], s[2], s[3])
            d[m + ".lora_B.weight"] = (s[0], r, 1, 1)
    return d


# ----------------------------------------------------------------------------------- sources
class StateDictSource:
    """Weights from a loaded state dict (any device/dtype; handed out as fp32 on `device`)."""

    def __init__(self, sd: Dict[str, torch.Tensor], device, lora: Optional[Dict[str, torch.Tensor]] = None,
                 lora_scale: float = 1.0):
        self.sd, self.device, self.lora, self.lora_scale = sd, device, lora, lora_scale

    def has(self, key):
        return key in self.sd

    def get(self, key) -> torch.Tensor:
        if key not in self.sd:
            raise KeyError(f"missing weight '{key}'")
        return self.sd[key].detach().to(device=self.device, dtype=torch.float32)

    def get_lora(self, module):
        """(A, B, alpha/r) for a wrapped module, or None."""
        if self.lora is None or module + ".lora_A.weight" not in self.lora:
            return None
        a = self.lora[module + ".lora_A.weight"].detach().to(device=self.device, dtype=torch.float32)
        b = self.lora[module + ".lora_B.weight"].detach().to(device=self.device, dtype=torch.float32)
        return a, b, self.lora_scale


class RandomSource:
    """Random-init weights generated on the device, one tensor at a time, deterministically per key.
    Matrices ~ N(0, 1/fan_in) (keeps activations O(1) through 70 blocks), norm gains ~ 1 + N(0,0.05),
    other vectors ~ N(0, 0.05); zero-init tensors of the reference (zero convs, adaLN linears,
    LoRA B) are randomised too so every sub-path carries signal."""

    def __init__(self, shapes, device, seed: int = 0, lora_shapes=None, lora_scale: float = 1.0):
        self.shapes, self.device, self.seed = shapes, device, seed
        self.lora_shapes, self.lora_scale = lora_shapes, lora_scale

    def has(self, key):
        return key in self.shapes

    def _draw(self, key, shape):
        h = int.from_bytes(hashlib.sha256(f"{self.seed}:{key}".encode()).digest()[:6], "little")
        g = torch.Generator(device=self.device).manual_seed(h)
        t = torch.randn(shape, generator=g, device=self.device, dtype=torch.float32)
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            if key.endswith("lora_B.weight"):
                return t * (0.3 * fan_in ** -0.5)
            return t * fan_in ** -0.5
        if "norm" in key and key.endswith("weight"):
            return 1.0 + 0.05 * t
        return 0.05 * t

    def get(self, key):
        if key not in self.shapes:
            raise KeyError(f"missing weight '{key}'")
        return self._draw(key, self.shapes[key])

    def get_lora(self, module):
        if self.lora_shapes is None or module + ".lora_A.weight" not in self.lora_shapes:
            return None
        a = self._draw(module + ".lora_A.weight", self.lora_shapes[module + ".lora_A.weight"])
        b = self._draw(module + ".lora_B.weight", self.lora_shapes[module + ".lora_B.weight"])
        return a, b, self.lora_scale


def merge_lora(w: torch.Tensor, lora) -> torch.Tensor:
    """W + (alpha/r) * B·A (peft 0.10.0 semantics); conv: B is 1x1."""
    a, b, s = lora
    if w.ndim == 2:
        return w + s * (b @ a)
    return w + s * torch.einsum("or,rikl->oikl", b[:, :, 0, 0], a)
