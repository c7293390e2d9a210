"""SDXL UNet forward on the sm_100a kernels, behind the call surface the reference pipeline uses
(pipelines/sdxl_instantir.py:1516-1529,1546-1554,1606-1616): ``unet(sample, t,
encoder_hidden_states=, cross_attention_kwargs={"temb":..}, down_block_additional_residuals=,
mid_block_additional_residual=, added_cond_kwargs={"text_embeds","time_ids","image_embeds"},
return_dict=False)[0]`` plus ``get_time_embed / time_embedding / get_aug_embed / time_embed_act /
enable_adapters / disable_adapters / set_attn_processor / attn_processors / config``.

Structure follows diffusers' UNet2DConditionModel as restated by the reference's
module/min_sdxl.py:789-914 (blocks) and module/unet/unet_2d_ZeroSFT.py:998-1122,1184-1397 (forward
plumbing); residual injection is the stock ControlNet add, fused here into
the concat kernel of the up path.
"""
from __future__ import annotations

import os
from types import SimpleNamespace
from typing import List, Optional

import torch

from . import ops
from .attention_processor import Attention, AttnProcessor2_0, CtxCache, silu_of
from .config import ModelConfig
from .nn import (Conv3x3, Downsample2D, FMap, GroupNorm, LayerNorm, Linear, LNStream, ResnetBlock2D, Runtime, SmallLinear,
                 Upsample2D, _bias, _load_w, _pack_pairs, _Packed, fmap_from_nchw)
from .resampler import MultiIPAdapterImageProjection, Resampler


class GEGLUFeedForward:
    """FeedForward(GEGLU) (module/min_sdxl.py:502-528): proj GEMM with the x1*gelu(x2) product fused
    into its epilogue (weights pair-packed per N tile), then the output GEMM with bias+residual."""

    def __init__(self, rt, src, p, C, fold=None):
        self.rt, self.C = rt, C
        self.bn = ops.default_bn(8 * C, pair=True)
        bn = self.bn

        def pack(w):
            return _pack_pairs(w[: 4 * C], w[4 * C:], bn)

        self.b1 = pack(src.get(p + ".net.0.proj.bias")).contiguous()
        self.w1 = _load_w(rt, src, p + ".net.0.proj", pack, fold=fold, bias=self.b1)
        self.folded = fold is not None
        self.out = Linear(rt, src, p + ".net.2")

    def __call__(self, a, M, residual, out_dtype=None):
        """a: normalised activations, or the block's LNStream when norm3 is folded (then `residual` is it too)."""
        rt, C = self.rt, self.C
        g = rt.empty(M, 4 * C)
        if self.folded:
            st = a
            ops.gemm(st.h16, self.w1.get(), g, M=M, N=8 * C, K=C, bias=self.w1.bias.get(), pair=ops.PAIR_GEGLU, bn=self.bn,
                     tc=True, ln_in=(st, self.w1.colsum.get(), st.eps))
            if out_dtype is None:  # in place on the fp32 stream, refreshing its 16-bit copy + row statistics
                return self.out(g, M, out=st.h, residual=st.h, ln_out=st)
            return self.out(g, M, out_dtype=out_dtype, residual=st.h)
        ops.gemm(a, self.w1.get(), g, M=M, N=8 * C, K=C, bias=self.b1, pair=ops.PAIR_GEGLU, bn=self.bn, tc=rt.tc)
        if out_dtype is None:  # in-place on the fp32 stream
            return self.out(g, M, out=residual, residual=residual)
        return self.out(g, M, out_dtype=out_dtype, residual=residual)


class BasicTransformerBlock:
    """module/min_sdxl.py:531-562; attn2/norm2 absent in the Aggregator (remove_attn2)."""

    def __init__(self, rt, src, p, C, heads, cross_dim):
        self.rt, self.C = rt, C
        # tcgen05 path: the three LayerNorms are folded into the GEMMs on either side of them (nn.LNStream,
        # DESIGN.md §3.2); the fp32 check mode runs them as kernels
        self.fold_ln = rt.tc and os.environ.get("IIR_LN_FOLD", "1") != "0"
        self.norm1 = LayerNorm(rt, src, p + ".norm1", C)
        self.attn1 = Attention(rt, src, p + ".attn1", C, heads, fold=self._fold(self.norm1))
        self.attn2 = None
        if cross_dim is not None:
            self.norm2 = LayerNorm(rt, src, p + ".norm2", C)
            self.attn2 = Attention(rt, src, p + ".attn2", C, heads, cross_dim, fold=self._fold(self.norm2))
        self.norm3 = LayerNorm(rt, src, p + ".norm3", C)
        self.ff = GEGLUFeedForward(rt, src, p + ".ff", C, fold=self._fold(self.norm3))

    def _fold(self, norm):
        return (norm.g, norm.b) if self.fold_ln else None

    def __call__(self, h, B, n, encoder_hidden_states, kw, last):
        """h: fp32 stream [B*n, C] (or its LNStream on the folded path), updated in place; returns an act-dtype
        tensor when `last`."""
        M, C = B * n, self.C
        if self.fold_ln:
            st = h
            self.attn1(st, encoder_hidden_states=None, residual=st, **kw)
            if self.attn2 is not None:
                self.attn2(st, encoder_hidden_states=encoder_hidden_states, residual=st, **kw)
            return self.ff(st, M, residual=st, out_dtype=self.rt.act_dtype if last else None)
        h3 = h.view(B, n, C)
        self.attn1(self.norm1(h, M).view(B, n, C), encoder_hidden_states=None, residual=h3, **kw)
        if self.attn2 is not None:
            self.attn2(self.norm2(h, M).view(B, n, C), encoder_hidden_states=encoder_hidden_states, residual=h3, **kw)
        return self.ff(self.norm3(h, M), M, residual=h, out_dtype=self.rt.act_dtype if last else None)


class Transformer2DModel:
    """module/min_sdxl.py:565-595 (GN eps 1e-6, linear proj_in/out; NHWC makes the permutes no-ops)."""

    def __init__(self, rt, src, p, cfg, C, heads, n_layers, cross):
        self.rt, self.C = rt, C
        self.norm = GroupNorm(rt, src, p + ".norm", C, cfg.norm_num_groups, 1e-6)
        self.proj_in = Linear(rt, src, p + ".proj_in")
        self.transformer_blocks = [
            BasicTransformerBlock(rt, src, f"{p}.transformer_blocks.{k}", C, heads, cfg.cross_attention_dim if cross else None)
            for k in range(n_layers)]
        self.proj_out = Linear(rt, src, p + ".proj_out")

    def __call__(self, x: FMap, encoder_hidden_states, kw) -> FMap:
        n_tok = x.H * x.W
        a = self.norm(x, silu=False)
        if self.transformer_blocks[0].fold_ln:
            h = LNStream(self.rt, self.rt.stream(x.M, self.C), x.n, n_tok, self.C, self.transformer_blocks[0].norm1.eps)
            self.proj_in(a.t, x.M, out=h.h, ln_out=h)
        else:
            h = self.proj_in(a.t, x.M, out_dtype=torch.float32)
        y = h
        for k, blk in enumerate(self.transformer_blocks):
            y = blk(h, x.n, n_tok, encoder_hidden_states, kw, last=k == len(self.transformer_blocks) - 1)
        # opt-in (IIR_GN_FUSE): the block output is normalised next by a GroupNorm of the same group count (the next
        # resnet's norm1) unless a concat comes first — its statistics ride on this epilogue
        gn = None
        if self.rt.gn_fuse and ops.gn_eligible(N=self.C, groups=self.norm.groups, rows_per_sample=n_tok, residual=x.t):
            gn = self.rt.gn_site(x.n, self.norm.groups)
        out = self.proj_out(y, x.M, out_dtype=torch.float32, residual=x.t, gn=gn, rows_per_sample=n_tok if gn is not None else 0)
        return FMap(out, x.n, x.H, x.W, x.C, gn)


class DownBlock:
    def __init__(self, rt, src, p, cfg, c_in, c_out, heads, n_tx, has_attn, add_down, cross=True):
        self.resnets = [ResnetBlock2D(rt, src, f"{p}.resnets.{j}", cfg, c_in if j == 0 else c_out, c_out)
                        for j in range(cfg.layers_per_block)]
        self.attentions = [Transformer2DModel(rt, src, f"{p}.attentions.{j}", cfg, c_out, heads, n_tx, cross)
                           for j in range(cfg.layers_per_block)] if has_attn else None
        self.downsamplers = [Downsample2D(rt, src, f"{p}.downsamplers.0")] if add_down else None

    def __call__(self, x, temb_act, ehs, kw):
        outs = []
        for j, r in enumerate(self.resnets):
            x = r(x, temb_act)
            if self.attentions is not None:
                x = self.attentions[j](x, ehs, kw)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x, gn_groups=self.resnets[0].norm1.groups)  # the next block's norm1 reads it
            outs.append(x)
        return x, outs


class MidBlock:
    def __init__(self, rt, src, p, cfg, C, heads, n_tx, cross=True):
        self.resnets = [ResnetBlock2D(rt, src, f"{p}.resnets.{j}", cfg, C, C) for j in range(2)]
        self.attentions = [Transformer2DModel(rt, src, f"{p}.attentions.0", cfg, C, heads, n_tx, cross)]

    def __call__(self, x, temb_act, ehs, kw):
        x = self.resnets[0](x, temb_act)
        x = self.attentions[0](x, ehs, kw)
        return self.resnets[1](x, temb_act)


class UpBlock:
    def __init__(self, rt, src, p, cfg, c_in, c_out, c_prev, heads, n_tx, has_attn, add_up):
        self.rt = rt
        n = cfg.layers_per_block + 1
        self.resnets = []
        for j in range(n):
            skip = c_in if j == n - 1 else c_out
            self.resnets.append(ResnetBlock2D(rt, src, f"{p}.resnets.{j}", cfg, (c_prev if j == 0 else c_out) + skip, c_out))
        self.attentions = [Transformer2DModel(rt, src, f"{p}.attentions.{j}", cfg, c_out, heads, n_tx, True)
                           for j in range(n)] if has_attn else None
        self.upsamplers = [Upsample2D(rt, src, f"{p}.upsamplers.0")] if add_up else None

    def __call__(self, x: FMap, skips: List[FMap], residuals: List[Optional[FMap]], mid_res, cond_scale, temb_act, ehs, kw):
        rt = self.rt
        for j, r in enumerate(self.resnets):
            skip, res = skips.pop(), residuals.pop()
            cat = rt.empty(x.M, x.C + skip.C)
            # torch.cat([hidden, skip], 1) fused with skip += cond_scale*res (and mid += cond_scale*mid_res)
            ops.concat_inject(x.t, x.C, skip.t, skip.C, cat, M=x.M, rh=None if mid_res is None else mid_res.t,
                              rs=None if res is None else res.t, cond_scale=cond_scale,
                              rows_per_sample=x.H * x.W)
            mid_res = None
            x = r(FMap(cat, x.n, x.H, x.W, x.C + skip.C), temb_act)
            if self.attentions is not None:
                x = self.attentions[j](x, ehs, kw)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class TimestepEmbedding:
    """Linear -> SiLU -> Linear on [n, d] fp32 vectors (module/min_sdxl.py:227-239)."""

    def __init__(self, rt, src, p):
        self.linear_1 = SmallLinear(rt, src, p + ".linear_1")
        self.linear_2 = SmallLinear(rt, src, p + ".linear_2")
        self.in_features = self.linear_1.K

    def __call__(self, sample, condition=None):
        return self.linear_2(self.linear_1(sample, act=ops.ACT_SILU))


class _EmbeddingMixin:
    """time / text_time embeddings shared by the UNet and the Aggregator
    (module/unet/unet_2d_ZeroSFT.py:998-1072; module/aggregator.py:824-882)."""

    def _init_embeddings(self, rt, src, cfg):
        self.time_embedding = TimestepEmbedding(rt, src, "time_embedding")
        self.add_embedding = TimestepEmbedding(rt, src, "add_embedding")
        self.time_embed_act = None

    def _timestep_tensor(self, timestep, n):
        if torch.is_tensor(timestep):
            t = timestep.to(device=self.rt.device, dtype=torch.float32).reshape(-1)
        else:
            t = torch.tensor([float(timestep)], device=self.rt.device, dtype=torch.float32)
        return t.expand(n).contiguous() if t.numel() == 1 else t.contiguous()

    def get_time_embed(self, sample, timestep):
        n = sample.shape[0]
        out = torch.empty(n, self.cfg.block_out_channels[0], device=self.rt.device, dtype=torch.float32)
        return ops.timestep_embedding(self._timestep_tensor(timestep, n), out.shape[1], out)

    def get_aug_embed(self, emb, encoder_hidden_states, added_cond_kwargs):
        if "text_embeds" not in added_cond_kwargs or "time_ids" not in added_cond_kwargs:
            raise ValueError("addition_embed_type 'text_time' requires `text_embeds` and `time_ids` in `added_cond_kwargs`")
        text_embeds = added_cond_kwargs["text_embeds"].to(device=self.rt.device, dtype=torch.float32)
        time_ids = added_cond_kwargs["time_ids"].to(device=self.rt.device, dtype=torch.float32)
        n, d = text_embeds.shape[0], self.cfg.addition_time_embed_dim
        add = torch.empty(n, self.cfg.pooled_dim + 6 * d, device=self.rt.device, dtype=torch.float32)
        tid = torch.empty(n * 6, d, device=self.rt.device, dtype=torch.float32)
        ops.timestep_embedding(time_ids.reshape(-1).contiguous(), d, tid)
        ops.cast2d(text_embeds.contiguous(), self.cfg.pooled_dim, add, add.shape[1], rows=n, cols=self.cfg.pooled_dim)
        ops.cast2d(tid.view(n, 6 * d), 6 * d, add[:, self.cfg.pooled_dim:], add.shape[1], rows=n, cols=6 * d)
        return self.add_embedding(add)

    def _emb(self, sample, timestep, added_cond_kwargs):
        emb = self.time_embedding(self.get_time_embed(sample, timestep))
        aug = self.get_aug_embed(emb, None, added_cond_kwargs)
        out = torch.empty_like(emb)
        return ops.add(emb, aug, out)


class UNet2DConditionModel(_EmbeddingMixin):
    def __init__(self, cfg: ModelConfig, source, device="cuda", precision="fp16", adapter: bool = True):
        self.cfg, self.source = cfg, source
        self.rt = rt = Runtime(device, precision)
        ch = cfg.block_out_channels
        self.config = SimpleNamespace(
            in_channels=cfg.in_channels, time_cond_proj_dim=None, addition_time_embed_dim=cfg.addition_time_embed_dim,
            cross_attention_dim=cfg.cross_attention_dim, block_out_channels=ch, encoder_hid_dim_type=None,
            addition_embed_type="text_time", down_block_types=cfg.down_block_types,
            transformer_layers_per_block=cfg.transformer_layers_per_block, layers_per_block=cfg.layers_per_block)
        self.conv_in_w = source.get("conv_in.weight").permute(0, 2, 3, 1).contiguous()
        self.conv_in_b = source.get("conv_in.bias").contiguous()
        self._init_embeddings(rt, source, cfg)
        self.down_blocks = []
        out = ch[0]
        for i, t in enumerate(cfg.down_block_types):
            inp, out = out, ch[i]
            self.down_blocks.append(DownBlock(rt, source, f"down_blocks.{i}", cfg, inp, out, cfg.num_attention_heads[i],
                                              cfg.transformer_layers_per_block[i], t == "CrossAttnDownBlock2D",
                                              i != len(ch) - 1))
        self.mid_block = MidBlock(rt, source, "mid_block", cfg, ch[-1], cfg.num_attention_heads[-1],
                                  cfg.transformer_layers_per_block[-1])
        rch, rheads = list(reversed(ch)), list(reversed(cfg.num_attention_heads))
        rtx, rtypes = list(reversed(cfg.transformer_layers_per_block)), list(reversed(cfg.down_block_types))
        self.up_blocks = []
        out = rch[0]
        for i in range(len(ch)):
            prev, out = out, rch[i]
            inp = rch[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(UpBlock(rt, source, f"up_blocks.{i}", cfg, inp, out, prev, rheads[i], rtx[i],
                                          rtypes[i] == "CrossAttnDownBlock2D", i != len(ch) - 1))
        self.conv_norm_out = GroupNorm(rt, source, "conv_norm_out", ch[0], cfg.norm_num_groups, cfg.norm_eps)
        self.conv_out_w = source.get("conv_out.weight").permute(0, 2, 3, 1).contiguous()
        self.conv_out_b = source.get("conv_out.bias").contiguous()
        self.encoder_hid_proj = None
        self._ip_cache = CtxCache()
        if adapter:
            from .ip_adapter_utils import load_adapter_to_unet

            load_adapter_to_unet(self)

    # ----------------------------------------------------------------- reference-facing helpers
    def attention_modules(self):
        """(name, Attention) in diffusers key order."""
        out = []

        def walk_t2d(p, t2d):
            for k, blk in enumerate(t2d.transformer_blocks):
                out.append((f"{p}.transformer_blocks.{k}.attn1", blk.attn1))
                if blk.attn2 is not None:
                    out.append((f"{p}.transformer_blocks.{k}.attn2", blk.attn2))

        for i, b in enumerate(self.down_blocks):
            for j, a in enumerate(b.attentions or []):
                walk_t2d(f"down_blocks.{i}.attentions.{j}", a)
        walk_t2d("mid_block.attentions.0", self.mid_block.attentions[0])
        for i, b in enumerate(self.up_blocks):
            for j, a in enumerate(b.attentions or []):
                walk_t2d(f"up_blocks.{i}.attentions.{j}", a)
        return out

    @property
    def attn_processors(self):
        return {name + ".processor": a.processor for name, a in self.attention_modules()}

    def set_attn_processor(self, processor):
        from .attention_processor import AdaLNKVBatch, TA_IPAttnProcessor2_0
        from .nn import SmallLinearBank

        for name, a in self.attention_modules():
            a.set_processor(processor[name + ".processor"] if isinstance(processor, dict) else processor)
        pairs = [(a.processor, a) for _, a in self.attention_modules() if isinstance(a.processor, TA_IPAttnProcessor2_0)]
        self.rt.adaln_bank = SmallLinearBank(self.rt)
        self.adaln_batch = AdaLNKVBatch(self.rt, pairs) if pairs and len({id(p) for p, _ in pairs}) == len(pairs) else None

    def enable_adapters(self):
        """previewer LoRA on (peft enable_adapters, pipelines/sdxl_instantir.py:1545): selects the
        LoRA-merged weight set W + (alpha/r) B A."""
        self.rt.lora_enabled = True

    def disable_adapters(self):
        self.rt.lora_enabled = False

    def process_encoder_hidden_states(self, encoder_hidden_states, added_cond_kwargs):
        if self.encoder_hid_proj is not None and self.config.encoder_hid_dim_type == "ip_image_proj":
            if "image_embeds" not in added_cond_kwargs:
                raise ValueError("encoder_hid_dim_type 'ip_image_proj' requires `image_embeds` in `added_cond_kwargs`")
            image_embeds = added_cond_kwargs["image_embeds"]
            key_t = image_embeds[0] if isinstance(image_embeds, list) else image_embeds
            tag = "L" if self.rt.lora_enabled else "B"
            toks = self._ip_cache.get("ip" + tag, key_t, lambda out: self.encoder_hid_proj(image_embeds, out=out))
            encoder_hidden_states = (encoder_hidden_states, toks)
        return encoder_hidden_states

    def refresh_context(self, encoder_hidden_states, added_cond_kwargs, temb):
        """(Re)compute every step-invariant tensor for both LoRA states, eagerly and in place: Resampler
        tokens, text K/V and pre-adaLN image K/V of all cross-attention layers.  The pipeline calls this
        once per image before replaying captured CUDA graphs."""
        was = self.rt.lora_enabled
        for state in (False, True):
            self.rt.lora_enabled = state
            ehs = self.process_encoder_hidden_states(encoder_hidden_states, added_cond_kwargs)
            for _, a in self.attention_modules():
                if a.cross_dim is not None and hasattr(a.processor, "prefetch"):
                    a.processor.prefetch(a, ehs)
        self.rt.lora_enabled = was

    # ---------------------------------------------------------------------------------- forward
    def __call__(self, *a, **kw):
        return self.forward(*a, **kw)

    def forward(self, sample, timestep, encoder_hidden_states, timestep_cond=None, cross_attention_kwargs=None,
                added_cond_kwargs=None, down_block_additional_residuals=None, mid_block_additional_residual=None,
                return_dict=False, additional_residual_scale=None):
        """sample [n,4,h,w] fp32 NCHW -> (eps [n,4,h,w] fp32,).  `additional_residual_scale` [n] (fp32)
        is an extension: residuals are multiplied by it inside the fused concat kernel instead of by
        separate elementwise kernels (pipelines/sdxl_instantir.py:1602-1603)."""
        state = self.forward_down_mid(sample, timestep, encoder_hidden_states, cross_attention_kwargs=cross_attention_kwargs,
                                      added_cond_kwargs=added_cond_kwargs)
        return self.forward_up(state, down_block_additional_residuals, mid_block_additional_residual,
                               additional_residual_scale)

    def forward_down_mid(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None, added_cond_kwargs=None):
        """First half of forward(): embeddings, conv_in, down blocks and the mid block.  None of it reads the
        ControlNet-style residuals (they enter the skip connections and the mid output in forward_up), so the
        pipeline runs this half on a second stream WHILE the Aggregator computes them."""
        rt, cfg = self.rt, self.cfg
        rt.new_forward()
        n, _, H, W = sample.shape
        kw = dict(cross_attention_kwargs or {})
        emb = self._emb(sample, timestep, added_cond_kwargs)
        if "temb" not in kw:
            kw["temb"] = emb
        temb_act = silu_of(rt, emb)
        ehs = self.process_encoder_hidden_states(encoder_hidden_states, added_cond_kwargs)
        ch0 = cfg.block_out_channels[0]
        x = torch.empty(n * H * W, ch0, device=rt.device, dtype=torch.float32)
        ops.conv3x3_direct(sample.contiguous(), self.conv_in_w, self.conv_in_b, x, in_nchw=True, out_nchw=False,
                           n_img=n, H=H, W=W, Cin=cfg.in_channels, Cout=ch0)
        x = FMap(x, n, H, W, ch0)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb_act, ehs, kw)
            skips += outs
        x = self.mid_block(x, temb_act, ehs, kw)
        return SimpleNamespace(x=x, skips=skips, temb_act=temb_act, ehs=ehs, kw=kw, n=n, H=H, W=W)

    def forward_up(self, state, down_block_additional_residuals=None, mid_block_additional_residual=None,
                   additional_residual_scale=None):
        """Second half of forward(): residual injection (fused into the up path's concat), up blocks, conv_out."""
        rt, cfg = self.rt, self.cfg
        x, skips, temb_act, ehs, kw = state.x, list(state.skips), state.temb_act, state.ehs, state.kw
        n, H, W = state.n, state.H, state.W
        ch0 = cfg.block_out_channels[0]
        is_controlnet = mid_block_additional_residual is not None and down_block_additional_residuals is not None
        residuals = [None] * len(skips)
        mid_res = None
        cond_scale = None
        if is_controlnet:
            residuals = [fmap_from_nchw(r) for r in down_block_additional_residuals]
            mid_res = fmap_from_nchw(mid_block_additional_residual)
            if additional_residual_scale is not None:
                cond_scale = additional_residual_scale.to(device=rt.device, dtype=torch.float32).reshape(-1).contiguous()
        for blk in self.up_blocks:
            x = blk(x, skips, residuals, mid_res, cond_scale, temb_act, ehs, kw)
            mid_res = None
        a = self.conv_norm_out(x, silu=True)
        eps = torch.empty(n, cfg.out_channels, H, W, device=rt.device, dtype=torch.float32)
        ops.conv3x3_direct(a.t, self.conv_out_w, self.conv_out_b, eps, in_nchw=False, out_nchw=True, n_img=n, H=H, W=W,
                           Cin=ch0, Cout=cfg.out_channels)
        return (eps,)
