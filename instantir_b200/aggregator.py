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


class _FromUNetSource:
    """weight source of Aggregator.from_unet (module/aggregator.py:564-576): trunk tensors come from the UNet's own
    source, SFT-head convolutions get torch's default Conv2d initialisation (kaiming_uniform(a=sqrt 5) weights,
    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) biases; seeded), the closing 1x1 convolutions are zero (zero_module, :980-983)."""

    def __init__(self, cfg, unet_source, device, seed=0):
        from .weights import aggregator_param_shapes

        self.shapes, self.src, self.device, self.seed = aggregator_param_shapes(cfg), unet_source, device, seed

    def has(self, key):
        return key in self.shapes

    def get(self, key):
        if key not in self.shapes:
            raise KeyError(f"missing weight '{key}'")
        if key.startswith("ref_conv_in."):
            return self.src.get("conv_in." + key.split(".", 1)[1])
        if not key.startswith("controlnet_"):
            return self.src.get(key)
        shape = self.shapes[key]
        head_out = key.startswith("controlnet_mid_block.1.") or (key.startswith("controlnet_down_blocks.") and key.split(".")[2] == "1")
        if head_out:
            return torch.zeros(shape, device=self.device, dtype=torch.float32)
        import hashlib

        wshape = self.shapes[key.rsplit(".", 1)[0] + ".weight"]
        fan_in = wshape[1] * wshape[2] * wshape[3]
        bound = fan_in ** -0.5  # kaiming_uniform(a=sqrt(5)): gain sqrt(2/6) * sqrt(3/fan_in) = 1/sqrt(fan_in); same for the bias
        h = int.from_bytes(hashlib.sha256(f"{self.seed}:{key}".encode()).digest()[:6], "little")
        g = torch.Generator(device="cpu").manual_seed(h)
        return ((torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound).to(self.device)

    def get_lora(self, module):
        return None


class _OverlaySource:
    """load_state_dict(strict=False): keys absent from the new state dict keep their current values"""

    def __init__(self, sd, base, device):
        self.sd, self.base, self.device = sd, base, device

    def has(self, key):
        return key in self.sd or self.base.has(key)

    def get(self, key):
        if key in self.sd:
            return self.sd[key].detach().to(device=self.device, dtype=torch.float32)
        return self.base.get(key)

    def get_lora(self, module):
        return None


class Aggregator(_EmbeddingMixin):
    weights_version = 0

    def __init__(self, cfg: ModelConfig, source, device="cuda", precision="fp16"):
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
    def from_unet(cls, unet, controlnet_conditioning_channel_order="rgb", conditioning_embedding_out_channels=None,
                  load_weights_from_unet=True, conditioning_channels=3, seed=0, precision=None):
        """module/aggregator.py:503-578, as the pipeline uses it (pipelines/sdxl_instantir.py:320-322: from_unet followed
        by remove_attn2): conv_in and ref_conv_in both start from the UNet's conv_in, time / add embeddings, down
        blocks and mid block are copies of the UNet's (minus attn2 / norm2), the SFT heads are freshly initialised
        (torch's default Conv2d init, seeded) and every head ends in a zero 1x1 convolution — so the residuals of a
        from_unet aggregator are exactly zero until `load_state_dict(aggregator.pt)` (infer.py:142-144)."""
        if not load_weights_from_unet:
            raise NotImplementedError("from_unet(load_weights_from_unet=False): random trunk weights have no use at inference")
        src = _FromUNetSource(unet.cfg, unet.source, unet.rt.device, seed)
        agg = cls(unet.cfg, src, unet.rt.device, precision or unet.rt.precision)
        agg.config.controlnet_conditioning_channel_order = controlnet_conditioning_channel_order
        return agg

    def state_dict_keys(self):
        """key -> shape of the state dict this module loads (module/aggregator.py:414-471 after remove_attn2)"""
        from .weights import aggregator_param_shapes

        return aggregator_param_shapes(self.cfg)

    def load_state_dict(self, state_dict, strict=True):
        """infer.py:142-144 / gradio_demo/app.py:81: (re)pack every layer from `state_dict` (aggregator.pt key layout,
        SURVEY Appendix D).  Returns (missing_keys, unexpected_keys) like torch; strict=True raises on either.
        Pipelines notice the new weights through `weights_version` and re-capture their CUDA graphs."""
        from types import SimpleNamespace as NS

        from .weights import StateDictSource

        want = self.state_dict_keys()
        missing = [k for k in want if k not in state_dict]
        unexpected = [k for k in state_dict if k not in want]
        bad = [k for k in want if k in state_dict and tuple(state_dict[k].shape) != tuple(want[k])]
        if bad:
            raise RuntimeError("size mismatch for " + ", ".join(f"{k}: {tuple(state_dict[k].shape)} vs {tuple(want[k])}" for k in bad[:8]))
        if strict and (missing or unexpected):
            raise RuntimeError(f"Error(s) in loading state_dict for Aggregator: missing {missing[:8]}{'...' if len(missing) > 8 else ''}, "
                               f"unexpected {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}")
        src = StateDictSource(state_dict, self.rt.device) if not missing else _OverlaySource(state_dict, self.source, self.rt.device)
        version = self.weights_version + 1
        order = self.config.controlnet_conditioning_channel_order
        self.__init__(self.cfg, src, self.rt.device, self.rt.precision)
        self.weights_version = version
        self.config.controlnet_conditioning_channel_order = order
        return NS(missing_keys=missing, unexpected_keys=unexpected)

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
        if conditioning_scale != 1.0:
            # module/aggregator.py:963-964: every residual times conditioning_scale (the pipeline itself always passes
            # 1.0 and scales by cond_scale inside the UNet's fused concat, pipelines/sdxl_instantir.py:1596,1602-1603)
            if head_stream is not None:
                torch.cuda.current_stream().wait_stream(head_stream)
            for r in down + [mid]:
                flat = r.permute(0, 2, 3, 1)  # NHWC memory behind the NCHW view
                ops.scale(flat, flat, float(conditioning_scale))
        if not return_dict:
            return (down, mid)
        return SimpleNamespace(down_block_res_samples=down, mid_block_res_sample=mid)
