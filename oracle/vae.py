"""CPU fp32 restatement of the VAE encode / decode the pipeline runs once per image — TEST
INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/, tests/golden/ and bench.py's CPU legs.

Follows the reference's vendored decoder plumbing ``module/diffusers_vae/vae.py:185-350`` (Decoder:
conv_in -> UNetMidBlock2D -> UpDecoderBlock2D x4 -> GroupNorm -> SiLU -> conv_out) and
``module/diffusers_vae/autoencoder_kl.py:270-281`` (``_decode``: post_quant_conv then the decoder), as called by
``pipelines/sdxl_instantir.py:1670-1704`` (latents / scaling_factor -> decode -> VaeImageProcessor.postprocess).

The blocks themselves live in ``diffusers==0.28.1`` (``models/unets/unet_2d_blocks.py``, ``models/resnet.py``,
``models/upsampling.py``, ``models/attention_processor.py``), which is absent from /root/reference: they are
restated here from their published definitions.  Pinning (tests/golden/make_golden_vae.py):
  * Decoder.forward and Encoder.forward (module/diffusers_vae/vae.py:46-182) run VERBATIM from the reference file
    over these blocks, DiagonalGaussianDistribution verbatim -> plumbing and sampling pinned;
  * the mid-block attention arithmetic is the reference's own ``AttnProcessor2_0``
    (module/ip_adapter/attention_processor.py:337-414, 4-D input + group_norm + residual branch) -> pinned;
  * ResnetBlock2D (temb=None) / Upsample2D / Downsample2D(padding=0) arithmetic: parity unpinned (no reference
    source or vectors).
"""
from __future__ import annotations

from dataclasses import asdict, dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class VaeConfig:
    """AutoencoderKL config fields the decode path reads (module/diffusers_vae/autoencoder_kl.py:71-90)."""

    in_channels: int = 3
    out_channels: int = 3
    latent_channels: int = 4
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32
    scaling_factor: float = 0.13025
    force_upcast: bool = True

    def to_dict(self):
        return asdict(self)


def sdxl_vae() -> VaeConfig:
    return VaeConfig()


def tiny_vae() -> VaeConfig:
    """scaled-down decoder with the same topology (3 up blocks, two of them upsampling: 32² latent -> 128² image)"""
    return VaeConfig(block_out_channels=(64, 128, 128), layers_per_block=1)


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D with temb_channels=None, eps=1e-6, output_scale_factor=1 (as built by
    UNetMidBlock2D / UpDecoderBlock2D for the VAE)."""

    def __init__(self, c_in, c_out, groups, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, c_in, eps=eps)
        self.conv1 = nn.Conv2d(c_in, c_out, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, c_out, eps=eps)
        self.conv2 = nn.Conv2d(c_out, c_out, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(c_in, c_out, 1) if c_in != c_out else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class VaeAttention(nn.Module):
    """diffusers Attention as UNetMidBlock2D builds it: one head of dim C, GroupNorm(eps 1e-6) on the input,
    biased q/k/v/out projections, residual connection, rescale_output_factor 1.  Exposes the duck-typed fields
    the reference's AttnProcessor2_0 reads, so that processor can drive it verbatim."""

    def __init__(self, C, groups, eps=1e-6):
        super().__init__()
        self.heads = 1
        self.group_norm = nn.GroupNorm(groups, C, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(C, C), nn.Linear(C, C), nn.Linear(C, C)
        self.to_out = nn.ModuleList([nn.Linear(C, C), nn.Dropout(0.0)])
        self.spatial_norm = None
        self.norm_cross = False
        self.residual_connection = True
        self.rescale_output_factor = 1.0

    def forward(self, x, temb=None):
        b, c, h, w = x.shape
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)
        q, k, v = self.to_q(t), self.to_k(t), self.to_v(t)
        p = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1)
        o = self.to_out[0](p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class UNetMidBlock2D(nn.Module):
    def __init__(self, C, groups, eps=1e-6, attention_cls=VaeAttention):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(C, C, groups, eps), ResnetBlock2D(C, C, groups, eps)])
        self.attentions = nn.ModuleList([attention_cls(C, groups, eps)])

    def forward(self, x, temb=None):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, temb=temb)
        return self.resnets[1](x, temb)


class Upsample2D(nn.Module):
    """nearest 2x then 3x3 conv (diffusers Upsample2D, use_conv=True)."""

    def __init__(self, C):
        super().__init__()
        self.conv = nn.Conv2d(C, C, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class UpDecoderBlock2D(nn.Module):
    def __init__(self, num_layers, c_in, c_out, add_upsample, groups, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c_in if i == 0 else c_out, c_out, groups, eps) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(c_out)]) if add_upsample else None

    def forward(self, x, temb=None):
        for r in self.resnets:
            x = r(x, temb)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class Downsample2D(nn.Module):
    """diffusers Downsample2D(use_conv=True, padding=0) as DownEncoderBlock2D builds it: F.pad (0, 1, 0, 1) then a
    stride-2, pad-0 3x3 conv (named `conv` in the checkpoint)."""

    def __init__(self, C):
        super().__init__()
        self.conv = nn.Conv2d(C, C, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class DownEncoderBlock2D(nn.Module):
    def __init__(self, num_layers, c_in, c_out, add_downsample, groups, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c_in if i == 0 else c_out, c_out, groups, eps) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(c_out)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class Encoder(nn.Module):
    """module/diffusers_vae/vae.py:46-182 (double_z, mid-block attention)."""

    def __init__(self, cfg: "VaeConfig"):
        super().__init__()
        ch, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i in range(len(ch)):
            prev, out = out, ch[i]
            self.down_blocks.append(DownEncoderBlock2D(cfg.layers_per_block, prev, out, i != len(ch) - 1, g))
        self.mid_block = UNetMidBlock2D(ch[-1], g)
        self.conv_norm_out = nn.GroupNorm(g, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg.latent_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def gaussian_sample(moments, noise=None):
    """DiagonalGaussianDistribution (module/diffusers_vae/vae.py): mean + exp(0.5 clamp(logvar, -30, 20)) * noise;
    noise None = mode()."""
    mean, logvar = moments.chunk(2, dim=1)
    if noise is None:
        return mean
    return mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise


class Decoder(nn.Module):
    """module/diffusers_vae/vae.py:185-350 (norm_type 'group', no latent_embeds)."""

    def __init__(self, cfg: VaeConfig):
        super().__init__()
        ch, g = cfg.block_out_channels, cfg.norm_num_groups
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[-1], 3, padding=1)
        self.mid_block = UNetMidBlock2D(ch[-1], g)
        rev = list(reversed(ch))
        self.up_blocks = nn.ModuleList()
        out = rev[0]
        for i in range(len(ch)):
            prev, out = out, rev[i]
            self.up_blocks.append(UpDecoderBlock2D(cfg.layers_per_block + 1, prev, out, i != len(ch) - 1, g))
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKLDecoder(nn.Module):
    """the decode half of AutoencoderKL: state-dict keys 'post_quant_conv.*', 'decoder.*' as in the checkpoint."""

    def __init__(self, cfg: VaeConfig):
        super().__init__()
        self.cfg = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = Decoder(cfg)
        self._init_encoder(cfg)

    def _init_encoder(self, cfg):
        pass

    def decode(self, z):
        """module/diffusers_vae/autoencoder_kl.py:270-281"""
        return self.decoder(self.post_quant_conv(z))


class AutoencoderKL(AutoencoderKLDecoder):
    """encode + decode: adds 'encoder.*' and 'quant_conv.*' (module/diffusers_vae/autoencoder_kl.py:92-113)."""

    def _init_encoder(self, cfg):
        self.encoder = Encoder(cfg)
        self.quant_conv = nn.Conv2d(2 * cfg.latent_channels, 2 * cfg.latent_channels, 1)

    def encode_moments(self, x):
        """module/diffusers_vae/autoencoder_kl.py:236-268: encoder then quant_conv -> (mean | logvar)"""
        return self.quant_conv(self.encoder(x))


def image_to_latents(vae: "AutoencoderKL", image: torch.Tensor, noise=None) -> torch.Tensor:
    """pipelines/sdxl_instantir.py:1375-1376: vae.encode(image).latent_dist.sample() * scaling_factor"""
    return gaussian_sample(vae.encode_moments(image), noise) * vae.cfg.scaling_factor


def latents_to_image(vae: AutoencoderKLDecoder, latents: torch.Tensor) -> torch.Tensor:
    """pipelines/sdxl_instantir.py:1689-1704 with output_type='pt': latents / scaling_factor -> decode ->
    VaeImageProcessor.postprocess = (x / 2 + 0.5).clamp(0, 1)."""
    img = vae.decode(latents / vae.cfg.scaling_factor)
    return (img / 2 + 0.5).clamp(0, 1)


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    """PSNR in dB of images in [0, 1]."""
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else -10.0 * torch.log10(torch.tensor(mse)).item()
