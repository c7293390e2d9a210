"""Generate tests/golden/*.pt by running the REFERENCE's own code, verbatim, in the authoring
container (needs /root/reference; the GPU box only reads the committed vectors).

    python tests/golden/make_golden.py [--ref /root/reference]

What runs verbatim from the reference (imported, not copied):
  module/ip_adapter/attention_processor.py  AttnProcessor2_0, TA_IPAttnProcessor2_0, AdaLayerNorm
  module/ip_adapter/resampler.py            Resampler
  module/ip_adapter/ip_adapter.py           MultiIPAdapterImageProjection
  schedulers/lcm_single_step_scheduler.py   LCMSingleStepScheduler (behind the stub below)
  module/min_sdxl.py                        ResnetBlock2D, Transformer2DModel, down/up/mid blocks and
                                            UNet2DConditionModel.forward, assembled at small widths
  module/aggregator.py                      Aggregator.__init__/from_unet/forward, SFT (behind the
                                            stub; the diffusers blocks it asks for are the oracle's)

`diffusers`/`peft` are not installed and cannot be (no network), so a stub package supplies only the
plumbing names those files import (ConfigMixin, register_to_config, BaseOutput, logging, ...).
No arithmetic lives in the stub except where stated (the Aggregator's diffusers blocks).
"""
from __future__ import annotations

import argparse
import dataclasses
import functools
import inspect
import os
import sys
import types

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import config as ocfg  # noqa: E402
from oracle import model as om  # noqa: E402


# ------------------------------------------------------------------------------ diffusers stub
def install_stub(ref_root: str):
    def mod(name):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
        return m

    d = mod("diffusers")
    cu = mod("diffusers.configuration_utils")

    class _Config(dict):
        __getattr__ = dict.get

    class ConfigMixin:
        config_name = "config.json"

        def register_to_config(self, **kw):
            if not hasattr(self, "_cfg"):
                object.__setattr__(self, "_cfg", _Config())
            self._cfg.update(kw)

        @property
        def config(self):
            return self._cfg

        @classmethod
        def from_config(cls, config, **kw):
            cfg = dict(config)
            cfg.update(kw)
            names = set(inspect.signature(cls.__init__).parameters)
            return cls(**{k: v for k, v in cfg.items() if k in names})

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            vals = {k: v for k, v in bound.arguments.items() if k != "self"}
            ConfigMixin.register_to_config(self, **vals)
            init(self, *args, **kwargs)

        return wrapper

    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config

    ut = mod("diffusers.utils")

    class BaseOutput(dict):
        def __post_init__(self):
            for f in dataclasses.fields(self):
                self[f.name] = getattr(self, f.name)

    class _Logger:
        def __getattr__(self, k):
            return lambda *a, **kw: None

    ut.BaseOutput = BaseOutput
    ut.logging = types.SimpleNamespace(get_logger=lambda name=None: _Logger())
    tu = mod("diffusers.utils.torch_utils")
    tu.randn_tensor = lambda shape, generator=None, device=None, dtype=None, layout=None: torch.randn(
        shape, generator=generator, dtype=dtype)
    mod("diffusers.schedulers")
    su = mod("diffusers.schedulers.scheduling_utils")
    su.SchedulerMixin = type("SchedulerMixin", (), {})

    mod("diffusers.loaders")
    sf = mod("diffusers.loaders.single_file_model")
    sf.FromOriginalModelMixin = type("FromOriginalModelMixin", (), {})
    mod("diffusers.models")
    mu = mod("diffusers.models.modeling_utils")
    mu.ModelMixin = type("ModelMixin", (nn.Module,), {})

    # reference processors stand in for diffusers.models.attention_processor
    sys.path.insert(0, ref_root)
    import module.ip_adapter.attention_processor as rap

    ap = mod("diffusers.models.attention_processor")
    ap.AttnProcessor, ap.AttnProcessor2_0 = rap.AttnProcessor, rap.AttnProcessor2_0
    ap.ADDED_KV_ATTENTION_PROCESSORS = ap.CROSS_ATTENTION_PROCESSORS = ()
    ap.AttentionProcessor = object
    ap.AttnAddedKVProcessor = type("AttnAddedKVProcessor", (), {})

    # --- arithmetic supplied by the stub: diffusers blocks requested by module/aggregator.py -> oracle's
    emb = mod("diffusers.models.embeddings")

    class Timesteps(om.Timesteps):
        def __init__(self, num_channels, flip_sin_to_cos=True, downscale_freq_shift=0):
            assert flip_sin_to_cos and downscale_freq_shift == 0
            super().__init__(num_channels)

    class TimestepEmbedding(om.TimestepEmbedding):
        def __init__(self, in_channels, time_embed_dim, act_fn="silu"):
            super().__init__(in_channels, time_embed_dim)

    emb.Timesteps, emb.TimestepEmbedding = Timesteps, TimestepEmbedding
    emb.TextImageProjection = emb.TextImageTimeEmbedding = emb.TextTimeEmbedding = type("_Unused", (), {})
    mod("diffusers.models.unets")
    blk = mod("diffusers.models.unets.unet_2d_blocks")

    def _scfg(temb_channels, cross_attention_dim, layers, eps, groups):
        return ocfg.StepConfig(time_embed_dim=temb_channels, cross_attention_dim=cross_attention_dim,
                               layers_per_block=layers, norm_eps=eps, norm_num_groups=groups)

    class _Down(om.DownBlock):
        def forward(self, hidden_states, temb=None, encoder_hidden_states=None, cross_attention_kwargs=None):
            x, outs = super().forward(hidden_states, temb, encoder_hidden_states, cross_attention_kwargs)
            return x, tuple(outs)

    class DownBlock2D(_Down):
        pass

    class CrossAttnDownBlock2D(_Down):
        pass

    def get_down_block(down_block_type, num_layers, in_channels, out_channels, temb_channels, add_downsample,
                       resnet_eps, resnet_act_fn, transformer_layers_per_block=1, num_attention_heads=None,
                       resnet_groups=None, cross_attention_dim=None, **unused):
        cfg = _scfg(temb_channels, cross_attention_dim, num_layers, resnet_eps, resnet_groups)
        cls = CrossAttnDownBlock2D if down_block_type == "CrossAttnDownBlock2D" else DownBlock2D
        return cls(cfg, in_channels, out_channels, num_attention_heads, transformer_layers_per_block,
                   down_block_type == "CrossAttnDownBlock2D", add_downsample)

    class UNetMidBlock2DCrossAttn(om.MidBlock):
        def __init__(self, transformer_layers_per_block, in_channels, temb_channels, resnet_eps, resnet_groups,
                     cross_attention_dim, num_attention_heads, **unused):
            super().__init__(_scfg(temb_channels, cross_attention_dim, 2, resnet_eps, resnet_groups), in_channels,
                             num_attention_heads, transformer_layers_per_block)

        def forward(self, hidden_states, temb=None, encoder_hidden_states=None, cross_attention_kwargs=None):
            return super().forward(hidden_states, temb, encoder_hidden_states, cross_attention_kwargs)

    blk.CrossAttnDownBlock2D, blk.DownBlock2D = CrossAttnDownBlock2D, DownBlock2D
    blk.UNetMidBlock2D = type("UNetMidBlock2D", (), {})
    blk.UNetMidBlock2DCrossAttn, blk.get_down_block = UNetMidBlock2DCrossAttn, get_down_block
    uc = mod("diffusers.models.unets.unet_2d_condition")
    uc.UNet2DConditionModel = om.UNet2DConditionModel
    d.__version__ = "stub"


from seeding import checksum, rnd, seeded_init  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    install_stub(args.ref)
    torch.manual_seed(0)
    torch.set_grad_enabled(False)

    # 1 ------------------------------------------------------------ attention processors
    import module.ip_adapter.attention_processor as rap

    C, heads, xdim, tdim, ntok = 128, 2, 96, 160, 8
    attn = om.Attention(C, heads, xdim)  # duck-typed `attn` (fields listed in SURVEY §8b)
    seeded_init(attn, 11)
    attn.prepare_attention_mask = None
    proc = rap.TA_IPAttnProcessor2_0(C, xdim, time_embedding_dim=tdim, scale=0.8, num_tokens=ntok)
    seeded_init(proc, 12)
    hs, text, ip, temb = rnd(2, 64, C, seed=13), rnd(2, 20, xdim, seed=14), rnd(2, ntok, xdim, seed=15), rnd(2, tdim, seed=16)
    out_tuple = proc(attn, hs, encoder_hidden_states=(text, [ip]), temb=temb)
    out_concat = proc(attn, hs, encoder_hidden_states=torch.cat([text, ip], 1), temb=temb)
    self_attn = om.Attention(C, heads)
    seeded_init(self_attn, 17)
    out_self = rap.AttnProcessor2_0()(self_attn, hs, temb=temb)
    ada = rap.AdaLayerNorm(C, tdim)
    seeded_init(ada, 18)
    out_ada = ada(hs, temb)
    torch.save({"seeds": dict(attn=11, proc=12, self_attn=17, ada=18),
                "checksums": dict(attn=checksum(attn), proc=checksum(proc), self_attn=checksum(self_attn), ada=checksum(ada)),
                "hs": hs, "text": text, "ip": ip, "temb": temb,
                "out_tuple": out_tuple, "out_concat": out_concat, "out_self": out_self, "out_ada": out_ada,
                "dims": dict(C=C, heads=heads, xdim=xdim, tdim=tdim, ntok=ntok, scale=0.8)},
               os.path.join(HERE, "processors.pt"))

    # 2 ------------------------------------------------------------------------ resampler
    from module.ip_adapter.ip_adapter import MultiIPAdapterImageProjection
    from module.ip_adapter.resampler import Resampler

    rs = Resampler(dim=128, depth=2, dim_head=64, heads=2, num_queries=16, embedding_dim=64, output_dim=256, ff_mult=4)
    seeded_init(rs, 21)
    x = rnd(3, 1, 33, 64, seed=22)
    out = MultiIPAdapterImageProjection([rs])([x])[0]
    torch.save({"seed": 21, "checksum": checksum(rs), "x": x, "out": out}, os.path.join(HERE, "resampler.pt"))

    # 3 ------------------------------------------------------------------- LCM scheduler
    from schedulers.lcm_single_step_scheduler import LCMSingleStepScheduler

    lcm = LCMSingleStepScheduler(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                                 num_train_timesteps=1000, timestep_spacing="leading", steps_offset=1)
    eps, xs = rnd(2, 4, 8, 8, seed=31), rnd(2, 4, 8, 8, seed=32)
    steps = {}
    for t in (958, 501, 34, 1, 0):
        steps[t] = lcm.step(eps, torch.tensor(t, dtype=torch.int64), xs, return_dict=False)[0]
    noisy = lcm.add_noise(xs, eps, torch.tensor([958, 1]))
    torch.save({"eps": eps, "x": xs, "steps": steps, "noisy": noisy,
                "alphas_cumprod": lcm.alphas_cumprod.clone()}, os.path.join(HERE, "lcm_scheduler.pt"))

    # 4 ------------------------------------------------------------ min_sdxl blocks / UNet
    import module.min_sdxl as ms

    ms.Attention.forward = ms.Attention.orig_forward  # the reference's own text-only path (:359-390)
    ch = (64, 128, 256)
    u = ms.UNet2DConditionModel.__new__(ms.UNet2DConditionModel)
    nn.Module.__init__(u)
    u.conv_in = nn.Conv2d(4, ch[0], 3, padding=1)
    u.time_proj = ms.Timesteps(ch[0])
    u.time_embedding = ms.TimestepEmbedding(ch[0], 1280)
    u.add_time_proj = ms.Timesteps(32)
    u.add_embedding = ms.TimestepEmbedding(64 + 6 * 32, 1280)
    u.down_blocks = nn.ModuleList([ms.DownBlock2D(ch[0], ch[0]),
                                   ms.CrossAttnDownBlock2D(ch[0], ch[1], n_layers=1),
                                   ms.CrossAttnDownBlock2D(ch[1], ch[2], n_layers=2, has_downsamplers=False)])
    u.up_blocks = nn.ModuleList([ms.CrossAttnUpBlock2D(ch[1], ch[2], ch[2], n_layers=2),
                                 ms.CrossAttnUpBlock2D(ch[0], ch[1], ch[2], n_layers=1),
                                 ms.UpBlock2D(ch[0], ch[0], ch[1])])
    u.mid_block = ms.UNetMidBlock2DCrossAttn(ch[2])
    u.mid_block.attentions = nn.ModuleList([ms.Transformer2DModel(ch[2], ch[2], n_layers=2)])
    u.conv_norm_out = nn.GroupNorm(32, ch[0], eps=1e-5)
    u.conv_act = nn.SiLU()
    u.conv_out = nn.Conv2d(ch[0], 4, 3, padding=1)
    seeded_init(u, 41)
    sample, text = rnd(2, 4, 16, 16, seed=42), rnd(2, 12, 2048, seed=43)
    pooled, tids = rnd(2, 64, seed=44), torch.tensor([[256., 256., 0., 0., 256., 256.]] * 2)
    out = u(sample, torch.tensor(501), text, {"text_embeds": pooled, "time_ids": tids})
    if isinstance(out, (tuple, list)):
        out = out[0]
    out = getattr(out, "sample", out)
    res = ms.ResnetBlock2D(96, 64)
    seeded_init(res, 45)
    rx, rt = rnd(2, 96, 8, 8, seed=46), rnd(2, 1280, seed=47)
    t2d = ms.Transformer2DModel(128, 128, n_layers=1)
    seeded_init(t2d, 48)
    tx = rnd(2, 128, 8, 8, seed=49)
    torch.save({"seeds": dict(unet=41, res=45, t2d=48), "n_params": sum(p.numel() for p in u.parameters()),
                "checksums": dict(unet=checksum(u), res=checksum(res), t2d=checksum(t2d)),
                "names": sorted(k for k, _ in u.named_parameters()), "sample": sample, "text": text, "pooled": pooled, "time_ids": tids,
                "t": 501, "unet_out": out, "res_x": rx, "res_temb": rt,
                "res_out": res(rx, rt), "t2d_x": tx, "t2d_text": text,
                "t2d_out": t2d(tx, text)}, os.path.join(HERE, "min_sdxl.pt"))

    # 5 ------------------------------------------------------------------------ aggregator
    import module.aggregator as ragg

    cfg = ocfg.tiny()
    cfg.transformer_layers_per_block = (1, 1, 1)
    ounet = om.UNet2DConditionModel(cfg)
    class _NS(types.SimpleNamespace):
        def __contains__(self, k):
            return hasattr(self, k)

    ounet.config = _NS(
        transformer_layers_per_block=cfg.transformer_layers_per_block, encoder_hid_dim=None, encoder_hid_dim_type=None,
        addition_embed_type="text_time", addition_time_embed_dim=cfg.addition_time_embed_dim,
        in_channels=4, flip_sin_to_cos=True, freq_shift=0, down_block_types=cfg.down_block_types,
        only_cross_attention=False, block_out_channels=cfg.block_out_channels, layers_per_block=2,
        downsample_padding=1, mid_block_scale_factor=1, act_fn="silu", norm_num_groups=32, norm_eps=1e-5,
        cross_attention_dim=cfg.cross_attention_dim, attention_head_dim=cfg.num_attention_heads,
        num_attention_heads=None, use_linear_projection=True, class_embed_type=None, num_class_embeds=None,
        upcast_attention=False, resnet_time_scale_shift="default",
        projection_class_embeddings_input_dim=cfg.projection_class_embeddings_input_dim,
        mid_block_type="UNetMidBlock2DCrossAttn")
    seeded_init(ounet, 51)
    agg = ragg.Aggregator.from_unet(ounet)
    agg.encoder_hid_proj = None
    from pipelines_stub import remove_attn2  # noqa  (defined below via exec of the reference function)

    remove_attn2(agg)
    zero_out = agg(rnd(2, 4, 16, 16, seed=52), torch.tensor(501), rnd(2, 5, cfg.cross_attention_dim, seed=53),
                   controlnet_cond=rnd(2, 4, 16, 16, seed=54),
                   added_cond_kwargs={"text_embeds": rnd(2, cfg.pooled_dim, seed=55),
                                      "time_ids": torch.tensor([[256., 256., 0., 0., 256., 256.]] * 2)},
                   return_dict=False)
    assert all(float(t.abs().max()) == 0.0 for t in zero_out[0]) and float(zero_out[1].abs().max()) == 0.0, \
        "from_unet must give exactly-zero residuals (zero 1x1 convs)"
    seeded_init(agg, 56)  # randomise everything, zero convs included
    a_in = dict(sample=rnd(2, 4, 16, 16, seed=52), cond=rnd(2, 4, 16, 16, seed=54), pooled=rnd(2, cfg.pooled_dim, seed=55),
                time_ids=torch.tensor([[256., 256., 0., 0., 256., 256.]] * 2), t=501)
    down, mid = agg(a_in["sample"], torch.tensor(a_in["t"]), rnd(2, 5, cfg.cross_attention_dim, seed=53),
                    controlnet_cond=a_in["cond"],
                    added_cond_kwargs={"text_embeds": a_in["pooled"], "time_ids": a_in["time_ids"]},
                    conditioning_scale=1.0, return_dict=False)
    torch.save({"cfg": cfg.to_dict(), "seed": 56, "checksum": checksum(agg),
                "names": sorted(k for k, _ in agg.named_parameters()), "inputs": a_in, "down": list(down), "mid": mid},
               os.path.join(HERE, "aggregator.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


def _install_remove_attn2(ref_root):
    """exec only the reference's remove_attn2 (pipelines/sdxl_instantir.py:165-177): the pipeline file
    itself needs real diffusers/peft/transformers imports and cannot be imported here."""
    src = open(os.path.join(ref_root, "pipelines", "sdxl_instantir.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if l.startswith("def remove_attn2"))
    end = next(i for i in range(start + 1, len(src)) if src[i] and not src[i].startswith((" ", "\t")))
    m = types.ModuleType("pipelines_stub")
    exec("\n".join(src[start:end]), m.__dict__)
    sys.modules["pipelines_stub"] = m


if __name__ == "__main__":
    _install_remove_attn2("/root/reference" if "--ref" not in sys.argv else sys.argv[sys.argv.index("--ref") + 1])
    main()
