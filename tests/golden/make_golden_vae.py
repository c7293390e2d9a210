"""Generate tests/golden/vae_decoder.pt and vae_encoder.pt by running the REFERENCE's vendored ``Decoder`` /
``Encoder`` / ``DiagonalGaussianDistribution`` (module/diffusers_vae/vae.py:46-350) verbatim in the authoring container (needs /root/reference).

    python tests/golden/make_golden_vae.py [--ref /root/reference]

Runs verbatim from the reference: ``Decoder.__init__/forward`` (conv_in, the fp32 upcast, mid -> up blocks
-> conv_norm_out -> SiLU -> conv_out) and, for the mid-block attention, the reference's own ``AttnProcessor2_0``
(module/ip_adapter/attention_processor.py:337-414: 4-D input, group_norm, residual_connection, rescale).
Supplied by the stub because ``diffusers`` is absent: ``UNetMidBlock2D`` / ``get_up_block`` -> the oracle's
restated ResnetBlock2D / Upsample2D blocks (oracle/vae.py) — their arithmetic is therefore NOT pinned.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import install_stub  # noqa: E402
from oracle import vae as ov  # noqa: E402
from seeding import checksum, rnd, seeded_init  # noqa: E402


def install_vae_stub():
    import module.ip_adapter.attention_processor as rap

    ut = sys.modules["diffusers.utils"]
    ut.is_torch_version = lambda op, v: True

    def mod(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    act = mod("diffusers.models.activations")
    act.get_activation = lambda name: {"silu": nn.SiLU(), "swish": nn.SiLU()}[name]
    sys.modules["diffusers.models.attention_processor"].SpatialNorm = type("SpatialNorm", (), {})
    blk = mod("diffusers.models.unet_2d_blocks")

    class RefProcAttention(ov.VaeAttention):
        """the oracle's duck-typed attention module driven by the reference processor"""

        def forward(self, x, temb=None):
            return rap.AttnProcessor2_0()(self, x, temb=temb)

    class UNetMidBlock2D(ov.UNetMidBlock2D):
        def __init__(self, in_channels, resnet_eps, resnet_act_fn, output_scale_factor, resnet_time_scale_shift,
                     attention_head_dim, resnet_groups, temb_channels, add_attention=True):
            assert temb_channels is None and add_attention and output_scale_factor == 1 and attention_head_dim == in_channels
            super().__init__(in_channels, resnet_groups, resnet_eps, attention_cls=RefProcAttention)

    def get_up_block(up_block_type, num_layers, in_channels, out_channels, prev_output_channel, add_upsample,
                     resnet_eps, resnet_act_fn, resnet_groups, attention_head_dim, temb_channels, resnet_time_scale_shift):
        assert up_block_type == "UpDecoderBlock2D" and temb_channels is None
        return ov.UpDecoderBlock2D(num_layers, in_channels, out_channels, add_upsample, resnet_groups, resnet_eps)

    def get_down_block(down_block_type, num_layers, in_channels, out_channels, add_downsample, resnet_eps,
                       downsample_padding, resnet_act_fn, resnet_groups, attention_head_dim, temb_channels):
        assert down_block_type == "DownEncoderBlock2D" and temb_channels is None and downsample_padding == 0
        return ov.DownEncoderBlock2D(num_layers, in_channels, out_channels, add_downsample, resnet_groups, resnet_eps)

    blk.UNetMidBlock2D, blk.get_up_block, blk.get_down_block = UNetMidBlock2D, get_up_block, get_down_block
    blk.AutoencoderTinyBlock = type("AutoencoderTinyBlock", (), {})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    install_stub(args.ref)
    install_vae_stub()
    torch.set_grad_enabled(False)
    import importlib.util

    spec = importlib.util.spec_from_file_location("ref_vae", os.path.join(args.ref, "module", "diffusers_vae", "vae.py"))
    rv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rv)

    cfg = ov.tiny_vae()
    dec = rv.Decoder(in_channels=cfg.latent_channels, out_channels=cfg.out_channels,
                     up_block_types=("UpDecoderBlock2D",) * len(cfg.block_out_channels),
                     block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                     norm_num_groups=cfg.norm_num_groups, act_fn="silu")
    seeded_init(dec, 61)
    z = rnd(2, cfg.latent_channels, 16, 16, seed=62)
    out = dec(z)
    torch.save({"cfg": cfg.to_dict(), "seed": 61, "checksum": checksum(dec), "names": sorted(k for k, _ in dec.named_parameters()),
                "z": z, "out": out}, os.path.join(HERE, "vae_decoder.pt"))
    print("vae_decoder.pt", tuple(out.shape), float(out.abs().mean()))

    enc = rv.Encoder(in_channels=cfg.in_channels, out_channels=cfg.latent_channels,
                     down_block_types=("DownEncoderBlock2D",) * len(cfg.block_out_channels),
                     block_out_channels=cfg.block_out_channels, layers_per_block=cfg.layers_per_block,
                     norm_num_groups=cfg.norm_num_groups, act_fn="silu", double_z=True)
    seeded_init(enc, 63)
    x = rnd(2, cfg.in_channels, 64, 48, seed=64)
    h = enc(x)
    moments = rnd(2, 2 * cfg.latent_channels, 6, 6, seed=65) * 3.0
    noise = rnd(2, cfg.latent_channels, 6, 6, seed=66)
    dist = rv.DiagonalGaussianDistribution(moments)
    dist_sample = dist.mean + dist.std * noise  # == DiagonalGaussianDistribution.sample() with this noise drawn
    torch.save({"cfg": cfg.to_dict(), "seed": 63, "checksum": checksum(enc), "names": sorted(k for k, _ in enc.named_parameters()),
                "x": x, "out": h, "moments": moments, "noise": noise, "sample": dist_sample, "mode": dist.mode()},
               os.path.join(HERE, "vae_encoder.pt"))
    print("vae_encoder.pt", tuple(h.shape), float(h.abs().mean()))


if __name__ == "__main__":
    main()
