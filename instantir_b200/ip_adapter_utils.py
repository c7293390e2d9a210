"""What ``load_adapter_to_pipe`` (module/ip_adapter/utils.py:73-161) does to the UNet: install the
attention processors (``init_attn_proc``), the Resampler as ``unet.encoder_hid_proj`` and flip
``unet.config.encoder_hid_dim_type`` to "ip_image_proj" (:160)."""
from __future__ import annotations

from .attention_processor import init_attn_proc
from .resampler import MultiIPAdapterImageProjection, Resampler

RESAMPLER_PREFIX = "encoder_hid_proj.image_projection_layers.0"


def load_adapter_to_unet(unet):
    unet.set_attn_processor(init_attn_proc(unet, ip_adapter_tokens=unet.cfg.num_ip_tokens, use_lcm=False, use_adaln=True))
    unet.encoder_hid_proj = MultiIPAdapterImageProjection([Resampler(unet.rt, unet.source, RESAMPLER_PREFIX, unet.cfg)])
    unet.config.encoder_hid_dim_type = "ip_image_proj"
    return unet


def load_adapter_to_pipe(pipe, *args, **kwargs):
    """Reference signature kept; the adapter weights come from the UNet's weight source (the state
    dict handed to UNet2DConditionModel already contains the '<attn2>.processor.*' and
    'encoder_hid_proj.*' keys)."""
    return load_adapter_to_unet(pipe.unet)
