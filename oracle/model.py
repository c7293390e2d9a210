"""Oracle modules: SDXL UNet, attention processors, Resampler, Aggregator (PyTorch fp32, CPU).

Module/parameter names equal the reference checkpoint keys so that a state
dict moves unchanged between this oracle, the reference and the CUDA product.
Each class cites the reference lines it restates.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .config import StepConfig


# ----------------------------------------------------------------------------- embeddings
class Timesteps(nn.Module):
    """[cos | sin] sinusoid, fp32 — module/min_sdxl.py:205-224 (flip_sin_to_cos=True, shift 0)."""

    def __init__(self, num_channels: int):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - 0.0)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    """Linear-SiLU-Linear — module/min_sdxl.py:227-239."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, out_features)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(out_features, out_features)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------------- resnet
class ResnetBlock2D(nn.Module):
    """module/min_sdxl.py:242-283 (GN eps 1e-5, shortcut 1x1 conv iff channels differ)."""

    def __init__(self, in_channels, out_channels, temb_channels, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


# ------------------------------------------------------------------------------ attention
class AdaLayerNorm(nn.Module):
    """module/ip_adapter/attention_processor.py:6-26 (zero-init linear; chunk -> shift, scale)."""

    def __init__(self, embedding_dim, time_embedding_dim=None):
        super().__init__()
        time_embedding_dim = time_embedding_dim or embedding_dim
        self.silu = nn.SiLU()
        self.linear = nn.Linear(time_embedding_dim, 2 * embedding_dim)
        nn.init.zeros_(self.linear.weight)
        nn.init.zeros_(self.linear.bias)
        self.norm = nn.LayerNorm(embedding_dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, timestep_embedding):
        emb = self.linear(self.silu(timestep_embedding))
        shift, scale = emb.view(len(x), 1, -1).chunk(2, dim=-1)
        return self.norm(x) * (1 + scale) + shift


def _sdpa(q, k, v, heads):
    b = q.shape[0]
    d = q.shape[-1] // heads
    q = q.view(b, -1, heads, d).transpose(1, 2)
    k = k.view(b, -1, heads, d).transpose(1, 2)
    v = v.view(b, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    return o.transpose(1, 2).reshape(b, -1, heads * d)


class AttnProcessor2_0(nn.Module):
    """Self-attention path — module/ip_adapter/attention_processor.py:337-414 (no mask, no norms)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None,
                 external_kv=None, temb=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = attn.to_q(hidden_states)
        k = attn.to_k(ctx)
        v = attn.to_v(ctx)
        out = _sdpa(q, k, v, attn.heads)
        return attn.to_out[1](attn.to_out[0](out))


class TA_IPAttnProcessor2_0(nn.Module):
    """Decoupled text + image cross-attention with time-aware adaLN on the image K/V —
    module/ip_adapter/attention_processor.py:1063-1207."""

    def __init__(self, hidden_size, cross_attention_dim=None, time_embedding_dim=None, scale=1.0, num_tokens=4):
        super().__init__()
        self.hidden_size, self.cross_attention_dim = hidden_size, cross_attention_dim
        self.scale, self.num_tokens = scale, num_tokens
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.ln_k_ip = AdaLayerNorm(hidden_size, time_embedding_dim)
        self.ln_v_ip = AdaLayerNorm(hidden_size, time_embedding_dim)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None,
                 external_kv=None, temb=None):
        assert temb is not None, "Timestep embedding is needed for a time-aware attention processor."
        if not isinstance(encoder_hidden_states, tuple):
            end_pos = encoder_hidden_states.shape[1] - self.num_tokens
            encoder_hidden_states, ip_hidden_states = (encoder_hidden_states[:, :end_pos, :],
                                                       encoder_hidden_states[:, end_pos:, :])
        else:
            ip_hidden_states = encoder_hidden_states[1][0]
            encoder_hidden_states = encoder_hidden_states[0]
        q = attn.to_q(hidden_states)
        out = _sdpa(q, attn.to_k(encoder_hidden_states), attn.to_v(encoder_hidden_states), attn.heads)
        ip_key = self.ln_k_ip(self.to_k_ip(ip_hidden_states), temb)
        ip_value = self.ln_v_ip(self.to_v_ip(ip_hidden_states), temb)
        out = out + self.scale * _sdpa(q, ip_key, ip_value, attn.heads)
        return attn.to_out[1](attn.to_out[0](out))


class Attention(nn.Module):
    """diffusers Attention as configured for SDXL: q/k/v no bias, out bias,
    head_dim = C/heads, processor called with the cross_attention_kwargs it accepts."""

    def __init__(self, query_dim, heads, cross_attention_dim=None):
        super().__init__()
        self.heads = heads
        kv_dim = cross_attention_dim or query_dim
        self.to_q = nn.Linear(query_dim, query_dim, bias=False)
        self.to_k = nn.Linear(kv_dim, query_dim, bias=False)
        self.to_v = nn.Linear(kv_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()
        # fields the reference processors read (attention_processor.py:348-375,409-412)
        self.spatial_norm = self.group_norm = None
        self.norm_cross = False
        self.residual_connection = False
        self.rescale_output_factor = 1.0

    def set_processor(self, processor):
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, **cross_attention_kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              **cross_attention_kwargs)


class GEGLU(nn.Module):
    """module/min_sdxl.py:502-510 (exact-erf GELU)."""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x1, x2 = self.proj(x).chunk(2, dim=-1)
        return x1 * F.gelu(x2)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for layer in self.net:
            x = layer(x)
        return x


class BasicTransformerBlock(nn.Module):
    """module/min_sdxl.py:531-562; attn2/norm2 may be deleted (Aggregator,
    pipelines/sdxl_instantir.py:165-177) in which case cross-attention is skipped."""

    def __init__(self, dim, heads, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, encoder_hidden_states=None, cross_attention_kwargs=None):
        kw = cross_attention_kwargs or {}
        x = self.attn1(self.norm1(x), encoder_hidden_states=None, **kw) + x
        if self.attn2 is not None:
            x = self.attn2(self.norm2(x), encoder_hidden_states=encoder_hidden_states, **kw) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    """module/min_sdxl.py:565-595 (GN eps 1e-6, linear projections)."""

    def __init__(self, channels, heads, n_layers, cross_attention_dim, groups=32):
        super().__init__()
        self.norm = nn.GroupNorm(groups, channels, eps=1e-6)
        self.proj_in = nn.Linear(channels, channels)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(channels, heads, cross_attention_dim) for _ in range(n_layers)])
        self.proj_out = nn.Linear(channels, channels)

    def forward(self, x, encoder_hidden_states=None, cross_attention_kwargs=None):
        b, c, h, w = x.shape
        res = x
        y = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        y = self.proj_in(y)
        for blk in self.transformer_blocks:
            y = blk(y, encoder_hidden_states, cross_attention_kwargs)
        y = self.proj_out(y)
        return y.reshape(b, h, w, c).permute(0, 3, 1, 2).contiguous() + res


class Downsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    """DownBlock2D / CrossAttnDownBlock2D — module/min_sdxl.py:621-682."""

    def __init__(self, cfg: StepConfig, in_ch, out_ch, heads, n_tx, has_attn, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_ch if j == 0 else out_ch, out_ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
            for j in range(cfg.layers_per_block)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(out_ch, heads, n_tx, cfg.cross_attention_dim, cfg.norm_num_groups)
            for _ in range(cfg.layers_per_block)]) if has_attn else None
        self.downsamplers = nn.ModuleList([Downsample2D(out_ch)]) if add_downsample else None

    def forward(self, x, temb, encoder_hidden_states=None, cross_attention_kwargs=None):
        outs = []
        for j, resnet in enumerate(self.resnets):
            x = resnet(x, temb)
            if self.attentions is not None:
                x = self.attentions[j](x, encoder_hidden_states, cross_attention_kwargs)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    """UNetMidBlock2DCrossAttn — module/min_sdxl.py:764-786."""

    def __init__(self, cfg: StepConfig, ch, heads, n_tx):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
                                      for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, n_tx, cfg.cross_attention_dim, cfg.norm_num_groups)])

    def forward(self, x, temb, encoder_hidden_states=None, cross_attention_kwargs=None):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, encoder_hidden_states, cross_attention_kwargs)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    """UpBlock2D / CrossAttnUpBlock2D — module/min_sdxl.py:685-760: pops skips from the end,
    cat([hidden, skip], dim=1)."""

    def __init__(self, cfg: StepConfig, in_ch, out_ch, prev_ch, heads, n_tx, has_attn, add_upsample):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList()
        for j in range(n):
            skip_ch = in_ch if j == n - 1 else out_ch
            res_in = prev_ch if j == 0 else out_ch
            self.resnets.append(ResnetBlock2D(res_in + skip_ch, out_ch, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps))
        self.attentions = nn.ModuleList([
            Transformer2DModel(out_ch, heads, n_tx, cfg.cross_attention_dim, cfg.norm_num_groups)
            for _ in range(n)]) if has_attn else None
        self.upsamplers = nn.ModuleList([Upsample2D(out_ch)]) if add_upsample else None

    def forward(self, x, skips: List[torch.Tensor], temb, encoder_hidden_states=None, cross_attention_kwargs=None):
        for j, resnet in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = resnet(x, temb)
            if self.attentions is not None:
                x = self.attentions[j](x, encoder_hidden_states, cross_attention_kwargs)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


# ------------------------------------------------------------------------------ resampler
class PerceiverAttention(nn.Module):
    """module/ip_adapter/resampler.py:34-78 (q and k each scaled by dim_head^-1/4, fp32 softmax)."""

    def __init__(self, dim, dim_head, heads):
        super().__init__()
        self.dim_head, self.heads = dim_head, heads
        inner = dim_head * heads
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_kv = nn.Linear(dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)

    def forward(self, x, latents):
        x = self.norm1(x)
        latents = self.norm2(latents)
        b, l, _ = latents.shape
        q = self.to_q(latents)
        k, v = self.to_kv(torch.cat((x, latents), dim=-2)).chunk(2, dim=-1)

        def split(t):
            return t.view(b, t.shape[1], self.heads, -1).transpose(1, 2)

        q, k, v = split(q), split(k), split(v)
        scale = 1 / math.sqrt(math.sqrt(self.dim_head))
        weight = (q * scale) @ (k * scale).transpose(-2, -1)
        weight = torch.softmax(weight.float(), dim=-1).type(weight.dtype)
        out = (weight @ v).permute(0, 2, 1, 3).reshape(b, l, -1)
        return self.to_out(out)


def _resampler_ff(dim, mult):
    inner = int(dim * mult)
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, inner, bias=False), nn.GELU(),
                         nn.Linear(inner, dim, bias=False))


class Resampler(nn.Module):
    """module/ip_adapter/resampler.py:81-147 (no pos-emb, no mean-pooled latents: the live config)."""

    def __init__(self, dim, depth, dim_head, heads, num_queries, embedding_dim, output_dim, ff_mult=4):
        super().__init__()
        self.latents = nn.Parameter(torch.randn(1, num_queries, dim) / dim ** 0.5)
        self.proj_in = nn.Linear(embedding_dim, dim)
        self.proj_out = nn.Linear(dim, output_dim)
        self.norm_out = nn.LayerNorm(output_dim)
        self.layers = nn.ModuleList([
            nn.ModuleList([PerceiverAttention(dim, dim_head, heads), _resampler_ff(dim, ff_mult)])
            for _ in range(depth)])

    def forward(self, x):
        latents = self.latents.repeat(x.size(0), 1, 1)
        x = self.proj_in(x)
        for attn, ff in self.layers:
            latents = attn(x, latents) + latents
            latents = ff(latents) + latents
        return self.norm_out(self.proj_out(latents))


class MultiIPAdapterImageProjection(nn.Module):
    """module/ip_adapter/ip_adapter.py:63-90."""

    def __init__(self, layers):
        super().__init__()
        self.image_projection_layers = nn.ModuleList(layers)

    def forward(self, image_embeds):
        if not isinstance(image_embeds, list):
            image_embeds = [image_embeds.unsqueeze(1)]
        out = []
        for e, layer in zip(image_embeds, self.image_projection_layers):
            b, n = e.shape[0], e.shape[1]
            out.append(layer(e.reshape((b * n,) + e.shape[2:])))
        return out


def make_resampler(cfg: StepConfig) -> Resampler:
    return Resampler(cfg.resampler_dim, cfg.resampler_depth, cfg.resampler_dim_head, cfg.resampler_heads,
                     cfg.num_ip_tokens, cfg.image_embed_dim, cfg.cross_attention_dim, cfg.resampler_ff_mult)


# ----------------------------------------------------------------------------------- UNet
class _Cfg:
    """attribute view of the config fields the reference pipeline reads from unet.config."""

    def __init__(self, cfg: StepConfig):
        self.in_channels = cfg.in_channels
        self.time_cond_proj_dim = None
        self.addition_time_embed_dim = cfg.addition_time_embed_dim
        self.cross_attention_dim = cfg.cross_attention_dim
        self.block_out_channels = cfg.block_out_channels
        self.encoder_hid_dim_type = None


class UNet2DConditionModel(nn.Module):
    """diffusers UNet2DConditionModel (0.28.1) as configured for SDXL, restated from
    module/min_sdxl.py:789-914 (structure) and module/unet/unet_2d_ZeroSFT.py:998-1122,1184-1397
    (forward plumbing), with the stock ControlNet residual adds."""

    def __init__(self, cfg: StepConfig):
        super().__init__()
        self.cfg = cfg
        self.config = _Cfg(cfg)
        ch = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_proj = Timesteps(ch[0])
        self.time_embedding = TimestepEmbedding(ch[0], cfg.time_embed_dim)
        self.add_time_proj = Timesteps(cfg.addition_time_embed_dim)
        self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, cfg.time_embed_dim)
        self.time_embed_act = None
        self.encoder_hid_proj = None
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, t in enumerate(cfg.down_block_types):
            inp, out = out, ch[i]
            self.down_blocks.append(DownBlock(cfg, inp, out, cfg.num_attention_heads[i],
                                              cfg.transformer_layers_per_block[i], t == "CrossAttnDownBlock2D",
                                              i != len(ch) - 1))
        self.mid_block = MidBlock(cfg, ch[-1], cfg.num_attention_heads[-1], cfg.transformer_layers_per_block[-1])
        self.up_blocks = nn.ModuleList()
        rch = list(reversed(ch))
        rheads = list(reversed(cfg.num_attention_heads))
        rtx = list(reversed(cfg.transformer_layers_per_block))
        rtypes = list(reversed(cfg.down_block_types))
        out = rch[0]
        for i in range(len(ch)):
            prev, out = out, rch[i]
            inp = rch[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(UpBlock(cfg, inp, out, prev, rheads[i], rtx[i],
                                          rtypes[i] == "CrossAttnDownBlock2D", i != len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    # --- pieces the pipeline calls directly (pipelines/sdxl_instantir.py:1516-1529)
    def get_time_embed(self, sample, timestep):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.int64, device=sample.device)
        elif t.ndim == 0:
            t = t[None].to(sample.device)
        t = t.expand(sample.shape[0])
        return self.time_proj(t).to(dtype=sample.dtype)

    def get_aug_embed(self, emb, encoder_hidden_states, added_cond_kwargs):
        text_embeds = added_cond_kwargs["text_embeds"]
        time_ids = added_cond_kwargs["time_ids"]
        time_embeds = self.add_time_proj(time_ids.flatten()).reshape((text_embeds.shape[0], -1))
        add_embeds = torch.cat([text_embeds, time_embeds], dim=-1).to(emb.dtype)
        return self.add_embedding(add_embeds)

    def process_encoder_hidden_states(self, encoder_hidden_states, added_cond_kwargs):
        if self.encoder_hid_proj is not None and self.config.encoder_hid_dim_type == "ip_image_proj":
            image_embeds = self.encoder_hid_proj(added_cond_kwargs["image_embeds"])
            encoder_hidden_states = (encoder_hidden_states, image_embeds)
        return encoder_hidden_states

    @property
    def attn_processors(self):
        procs = {}
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                procs[f"{name}.processor"] = m.processor
        return procs

    def set_attn_processor(self, procs):
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                m.set_processor(procs[f"{name}.processor"])

    def forward(self, sample, timestep, encoder_hidden_states, timestep_cond=None, cross_attention_kwargs=None,
                added_cond_kwargs=None, down_block_additional_residuals=None, mid_block_additional_residual=None,
                return_dict=False):
        t_emb = self.get_time_embed(sample, timestep)
        emb = self.time_embedding(t_emb, timestep_cond)
        emb = emb + self.get_aug_embed(emb, encoder_hidden_states, added_cond_kwargs)
        encoder_hidden_states = self.process_encoder_hidden_states(encoder_hidden_states, added_cond_kwargs)
        sample = self.conv_in(sample)
        skips = [sample]
        for blk in self.down_blocks:
            sample, outs = blk(sample, emb, encoder_hidden_states, cross_attention_kwargs)
            skips += outs
        is_controlnet = mid_block_additional_residual is not None and down_block_additional_residuals is not None
        if is_controlnet:
            skips = [s + r for s, r in zip(skips, down_block_additional_residuals)]
        sample = self.mid_block(sample, emb, encoder_hidden_states, cross_attention_kwargs)
        if is_controlnet:
            sample = sample + mid_block_additional_residual
        for blk in self.up_blocks:
            sample = blk(sample, skips, emb, encoder_hidden_states, cross_attention_kwargs)
        sample = self.conv_out(self.conv_act(self.conv_norm_out(sample)))
        return (sample,)


def init_attn_proc(unet: UNet2DConditionModel, ip_adapter_tokens, time_embedding_dim, scale=1.0):
    """module/ip_adapter/attention_processor.py:1364-1415 with use_adaln=True: attn1 keeps the plain
    processor, attn2 gets TA_IPAttnProcessor2_0 whose to_k_ip/to_v_ip start as copies of to_k/to_v.
    (The reference hard-codes time_embedding_dim=1280; here it follows the config.)"""
    procs = {}
    for name, m in unet.named_modules():
        if not isinstance(m, Attention):
            continue
        if name.endswith("attn1"):
            procs[f"{name}.processor"] = AttnProcessor2_0()
        else:
            p = TA_IPAttnProcessor2_0(m.to_q.in_features, unet.cfg.cross_attention_dim,
                                      time_embedding_dim=time_embedding_dim, scale=scale,
                                      num_tokens=ip_adapter_tokens)
            p.to_k_ip.weight.data.copy_(m.to_k.weight.data)
            p.to_v_ip.weight.data.copy_(m.to_v.weight.data)
            procs[f"{name}.processor"] = p
    return procs


def load_adapter(unet: UNet2DConditionModel):
    """What module/ip_adapter/utils.py:73-161 does to the UNet: install processors + Resampler and
    flip encoder_hid_dim_type to "ip_image_proj" (:160)."""
    cfg = unet.cfg
    unet.set_attn_processor(init_attn_proc(unet, cfg.num_ip_tokens, cfg.time_embed_dim, cfg.ip_scale))
    unet.encoder_hid_proj = MultiIPAdapterImageProjection([make_resampler(cfg)])
    unet.config.encoder_hid_dim_type = "ip_image_proj"
    return unet


# ------------------------------------------------------------------------------ aggregator
def zero_module(module):
    for p in module.parameters():
        nn.init.zeros_(p)
    return module


class SFT(nn.Module):
    """module/aggregator.py:51-90."""

    def __init__(self, label_nc, norm_nc, nhidden=128):
        super().__init__()
        self.mlp_shared = nn.Sequential(nn.Conv2d(label_nc, nhidden, 3, padding=1), nn.SiLU())
        self.mul = nn.Conv2d(nhidden, norm_nc, 3, padding=1)
        self.add = nn.Conv2d(nhidden, norm_nc, 3, padding=1)

    def forward(self, hidden_states):
        c, h = hidden_states
        actv = self.mlp_shared(c)
        return h * (self.mul(actv) + 1) + self.add(actv)


def remove_attn2(model):
    """pipelines/sdxl_instantir.py:165-177."""
    for m in model.modules():
        if isinstance(m, BasicTransformerBlock):
            m.attn2 = None
            m.norm2 = None


class Aggregator(nn.Module):
    """module/aggregator.py:158-977: SDXL down+mid blocks on a 2h x w canvas (LQ latent features on
    top, preview latent features below, cat_dim=-2), then SFT + zero 1x1 conv per skip."""

    def __init__(self, cfg: StepConfig):
        super().__init__()
        self.cfg = cfg
        ch = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.ref_conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_proj = Timesteps(ch[0])
        self.time_embedding = TimestepEmbedding(ch[0], cfg.time_embed_dim)
        self.add_time_proj = Timesteps(cfg.addition_time_embed_dim)
        self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, cfg.time_embed_dim)

        def head(c):
            return nn.Sequential(SFT(c, c, cfg.sft_hidden), zero_module(nn.Conv2d(c, c, 1)))

        self.down_blocks = nn.ModuleList()
        self.controlnet_down_blocks = nn.ModuleList([head(ch[0])])
        out = ch[0]
        for i, t in enumerate(cfg.down_block_types):
            inp, out = out, ch[i]
            last = i == len(ch) - 1
            self.down_blocks.append(DownBlock(cfg, inp, out, cfg.num_attention_heads[i],
                                              cfg.transformer_layers_per_block[i], t == "CrossAttnDownBlock2D", not last))
            for _ in range(cfg.layers_per_block):
                self.controlnet_down_blocks.append(head(out))
            if not last:
                self.controlnet_down_blocks.append(head(out))
        self.controlnet_mid_block = head(ch[-1])
        self.mid_block = MidBlock(cfg, ch[-1], cfg.num_attention_heads[-1], cfg.transformer_layers_per_block[-1])

    @classmethod
    def from_unet(cls, unet: UNet2DConditionModel):
        """module/aggregator.py:503-578 (load_weights_from_unet=True)."""
        agg = cls(unet.cfg)
        agg.conv_in.load_state_dict(unet.conv_in.state_dict())
        agg.ref_conv_in.load_state_dict(unet.conv_in.state_dict())
        agg.time_embedding.load_state_dict(unet.time_embedding.state_dict())
        agg.add_embedding.load_state_dict(unet.add_embedding.state_dict())
        sd = {k: v for k, v in unet.down_blocks.state_dict().items() if ".processor." not in k}
        agg.down_blocks.load_state_dict(sd)
        sd = {k: v for k, v in unet.mid_block.state_dict().items() if ".processor." not in k}
        agg.mid_block.load_state_dict(sd)
        return agg

    def forward(self, sample, timestep, encoder_hidden_states, controlnet_cond, cat_dim=-2,
                conditioning_scale=1.0, added_cond_kwargs=None, cross_attention_kwargs=None, return_dict=False):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.int64, device=sample.device)
        elif t.ndim == 0:
            t = t[None].to(sample.device)
        t = t.expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(t).to(sample.dtype))
        text_embeds, time_ids = added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"]
        time_embeds = self.add_time_proj(time_ids.flatten()).reshape((text_embeds.shape[0], -1))
        emb = emb + self.add_embedding(torch.cat([text_embeds, time_embeds], dim=-1).to(emb.dtype))
        assert cat_dim in (-2, 2), "the live path concatenates along H (module/aggregator.py:902)"
        sample = torch.cat([self.conv_in(sample), self.ref_conv_in(controlnet_cond)], dim=-2)
        skips = [sample]
        for blk in self.down_blocks:
            sample, outs = blk(sample, emb, None, cross_attention_kwargs)
            skips += outs
        sample = self.mid_block(sample, emb, None, cross_attention_kwargs)

        def split(x):
            h = x.shape[2]
            return x[:, :, :h // 2, :], x[:, :, -(h // 2):, :]

        down = [head(split(s)) * conditioning_scale for s, head in zip(skips, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(split(sample)) * conditioning_scale
        return down, mid
