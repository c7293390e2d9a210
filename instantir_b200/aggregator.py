"""InstantIR Aggregator on the sm_100a kernels — ``Aggregator.forward(sample, timestep,
encoder_hidden_states, controlnet_cond, cat_dim=-2, conditioning_scale=1.0, ..., added_cond_kwargs,
cross_attention_kwargs, return_dict)`` as in module/aggregator.py:758-977.

The SDXL down+mid blocks (cross-attention removed, pipelines/sdxl_instantir.py:165-177) run on one
NHWC canvas of height 2h: conv_in(LQ latent) fills rows [0,h), ref_conv_in(preview latent) rows
[h,2h) (module/aggregator.py:889-902) — written there directly by the two input convolutions, no
torch.cat.  3x3 convs and downsamplers see the seam like the reference (one image of height 2h).
Each of the 9+1 heads is SFT (module/aggregator.py:70-90) + a zero-initialised 1x1 conv; the SFT
modulation h*(gamma+1)+beta is the epilogue of ONE implicit-GEMM that computes gamma and beta
together (weights pair-packed per N tile).
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from . import ops
from .attention_processor import silu_of
from .config import ModelConfig
from .nn import Conv3x3, FMap, Linear, Runtime, _bias, _conv_to_gemm, _load_w, _pack_pairs, _ShortcutSrc
from .unet import DownBlock, MidBlock, _EmbeddingMixin


class SFTHead:
    """nn.Sequential(SFT(C, C), zero_module(Conv2d(C, C, 1))) — module/aggregator.py:414-417."""

    def __init__(self, rt: Runtime, src, p, C, hidden):
        self.rt, self.C, self.hidden = rt, C, hidden
        self.mlp_shared = Conv3x3(rt, src, p + ".0.mlp_shared.0")
        self.bn = ops.default_bn(2 * C, pair=True)
        wm = _conv_to_gemm(src.get(p + ".0.mul.weight"))
        wa = _conv_to_gemm(src.get(p + ".0.add.weight"))
        self.w_ga = _pack_pairs(wm, wa, self.bn).to(rt.w_dtype).contiguous()
        self.b_ga = _pack_pairs(src.get(p + ".0.mul.bias"), src.get(p + ".0.add.bias"), self.bn).contiguous()
        self.zero_conv = Linear(rt, _ShortcutSrc(src), p + ".1")

    def __call__(self, canvas: FMap, out=None) -> torch.Tensor:
        """canvas [n, 2H, W, C] fp32 -> residual [n, C, H, W] view (NHWC memory, activation dtype), written into the
        caller's [n*H*W, C] buffer `out` when given."""
        rt, C = self.rt, self.C
        n, H, W = canvas.n, canvas.H // 2, canvas.W
        M, half = n * H * W, H * W * C
        c = rt.empty(M, C)                                             # cond half  (rows [:H])
        h = torch.empty(M, C, device=rt.device, dtype=torch.float32)   # ref half   (rows [-H:])
        ops.cast2d(canvas.t, 2 * half, c, half, rows=n, cols=half)
        ops.cast2d(canvas.t.view(-1)[half:], 2 * half, h, half, rows=n, cols=half)
        actv = self.mlp_shared(FMap(c, n, H, W, C), act=ops.ACT_SILU)
        sft = rt.empty(M, C)
        ops.gemm(actv.t, self.w_ga, sft, M=M, N=2 * C, K=9 * self.hidden, bias=self.b_ga, pair=ops.PAIR_SFT, aux=h,
                 bn=self.bn, conv=dict(n_img=n, H=H, W=W, Cin=self.hidden), tc=rt.tc)
        out = self.zero_conv(sft, M, out=out)
        return FMap(out, n, H, W, C).nchw()


class Aggregator(_EmbeddingMixin):
    def __init__(self, cfg: ModelConfig, source, device="cuda", precision="bf16"):
        self.cfg, self.source = cfg, source
        self.rt = rt = Runtime(device, precision)
        ch = cfg.block_out_channels
        self.config = SimpleNamespace(controlnet_conditioning_channel_order="rgb", addition_embed_type="text_time",
                                      global_pool_conditions=False, class_embed_type=None, block_out_channels=ch)
        self.conv_in_w = source.get("conv_in.weight").permute(0, 2, 3, 1).contiguous()
        self.conv_in_b = source.get("conv_in.bias").contiguous()
        self.ref_conv_in_w = source.get("ref_conv_in.weight").permute(0, 2, 3, 1).contiguous()
        self.ref_conv_in_b = source.get("ref_conv_in.bias").contiguous()
        self._init_embeddings(rt, source, cfg)
        self.down_blocks, self.controlnet_down_blocks = [], []
        idx = 0
        self.controlnet_down_blocks.append(SFTHead(rt, source, f"controlnet_down_blocks.{idx}", ch[0], cfg.sft_hidden))
        out = ch[0]
        for i, t in enumerate(cfg.down_block_types):
            inp, out = out, ch[i]
            last = i == len(ch) - 1
            self.down_blocks.append(DownBlock(rt, source, f"down_blocks.{i}", cfg, inp, out, cfg.num_attention_heads[i],
                                              cfg.transformer_layers_per_block[i], t == "CrossAttnDownBlock2D", not last,
                                              cross=False))
            for _ in range(cfg.layers_per_block + (0 if last else 1)):
                idx += 1
                self.controlnet_down_blocks.append(SFTHead(rt, source, f"controlnet_down_blocks.{idx}", out, cfg.sft_hidden))
        self.controlnet_mid_block = SFTHead(rt, source, "controlnet_mid_block", ch[-1], cfg.sft_hidden)
        self.mid_block = MidBlock(rt, source, "mid_block", cfg, ch[-1], cfg.num_attention_heads[-1],
                                  cfg.transformer_layers_per_block[-1], cross=False)

    @classmethod
    def from_unet(cls, unet, **kw):
        raise NotImplementedError(
            "Aggregator.from_unet builds an untrained aggregator whose outputs are exactly zero "
            "(module/aggregator.py:414-417,503-578); construct Aggregator(cfg, source) from aggregator.pt weights")

    def __call__(self, *a, **kw):
        return self.forward(*a, **kw)

    def forward(self, sample, timestep, encoder_hidden_states=None, controlnet_cond=None, cat_dim=-2,
                conditioning_scale=1.0, class_labels=None, timestep_cond=None, attention_mask=None,
                added_cond_kwargs=None, cross_attention_kwargs=None, return_dict=False, head_stream=None,
                out_buffers=None):
        """Extensions.  `head_stream`: the SFT heads of every down block but the last are enqueued on that stream as
        soon as the block's skip tensors exist, so they overlap the rest of the trunk; the CALLER joins
        `head_stream` before reading the returned residuals.  `out_buffers` = (list of 9 [M_i, C_i] tensors, [M, C]
        tensor): the heads write the residuals there (the pipeline ping-pongs two such sets to run the aggregator
        of the NEXT step beside the UNet of the current one)."""
        if self.config.controlnet_conditioning_channel_order != "rgb":
            raise ValueError(f"unknown `controlnet_conditioning_channel_order`: {self.config.controlnet_conditioning_channel_order}")
        if cat_dim not in (-2, 2):
            raise ValueError(f"Aggregator shall concat along spatial dimension H (cat_dim=-2), but is asked to concat dim: {cat_dim}.")
        if conditioning_scale != 1.0:
            raise NotImplementedError("conditioning_scale != 1 (the pipeline never passes it, pipelines/sdxl_instantir.py:1596)")
        rt, cfg = self.rt, self.cfg
        rt.new_forward()
        n, _, H, W = sample.shape
        emb = self._emb(sample, timestep, added_cond_kwargs)
        temb_act = silu_of(rt, emb)
        ch0 = cfg.block_out_channels[0]
        canvas = torch.empty(n * 2 * H * W, ch0, device=rt.device, dtype=torch.float32)
        ops.conv3x3_direct(sample.contiguous(), self.conv_in_w, self.conv_in_b, canvas, in_nchw=True, out_nchw=False,
                           n_img=n, H=H, W=W, Cin=cfg.in_channels, Cout=ch0, out_H=2 * H, out_row_off=0)
        cond = controlnet_cond.to(dtype=torch.float32).contiguous()
        ops.conv3x3_direct(cond, self.ref_conv_in_w, self.ref_conv_in_b, canvas, in_nchw=True, out_nchw=False,
                           n_img=n, H=H, W=W, Cin=cfg.in_channels, Cout=ch0, out_H=2 * H, out_row_off=H)
        x = FMap(canvas, n, 2 * H, W, ch0)
        skips = [x]
        kw = dict(cross_attention_kwargs or {})
        down = [None] * len(self.controlnet_down_blocks)
        ob_down, ob_mid = out_buffers if out_buffers is not None else ([None] * len(down), None)
        done = 0
        for bi, blk in enumerate(self.down_blocks):
            x, outs = blk(x, temb_act, None, kw)
            skips += outs
            if head_stream is not None and bi < len(self.down_blocks) - 1:
                ev = torch.cuda.Event()
                ev.record()
                head_stream.wait_event(ev)
                with torch.cuda.stream(head_stream):
                    for i in range(done, len(skips)):
                        down[i] = self.controlnet_down_blocks[i](skips[i], ob_down[i])
                done = len(skips)
        x = self.mid_block(x, temb_act, None, kw)
        for i in range(done, len(skips)):
            down[i] = self.controlnet_down_blocks[i](skips[i], ob_down[i])
        mid = self.controlnet_mid_block(x, ob_mid)
        if not return_dict:
            return (down, mid)
        return SimpleNamespace(down_block_res_samples=down, mid_block_res_sample=mid)
