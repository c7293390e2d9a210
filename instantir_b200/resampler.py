"""Image-token projector: Perceiver ``Resampler`` (module/ip_adapter/resampler.py:81-147) and
``MultiIPAdapterImageProjection`` (module/ip_adapter/ip_adapter.py:63-90) on the sm_100a kernels.
DINOv2 tokens [B, 257, 1024] -> 64 image tokens [B, 64, 2048].  Step-invariant: the UNet caches the
result per LoRA state (the previewer LoRA targets to_q/to_kv/to_out/proj_in/proj_out here)."""
from __future__ import annotations

import torch

from . import ops
from .nn import LayerNorm, Linear, Runtime


class PerceiverAttention:
    """module/ip_adapter/resampler.py:34-78: q from the latents, k/v from cat(x, latents); q and k are
    each scaled by dim_head^-1/4 (== one softmax scale of dim_head^-1/2), softmax in fp32."""

    def __init__(self, rt, src, p, dim, dim_head, heads):
        self.rt, self.dim, self.heads, self.inner = rt, dim, heads, dim_head * heads
        self.norm1 = LayerNorm(rt, src, p + ".norm1", dim)
        self.norm2 = LayerNorm(rt, src, p + ".norm2", dim)
        self.to_q = Linear(rt, src, p + ".to_q", bias=False)
        self.to_kv = Linear(rt, src, p + ".to_kv", bias=False)
        self.to_out = Linear(rt, src, p + ".to_out", bias=False)
        self.scale = float(dim_head) ** -0.5

    def __call__(self, x, latents, B, S, nq):
        """x: act [B*S, dim]; latents: fp32 [B*nq, dim], updated in place (attn(x, latents) + latents)."""
        rt, dim, inner = self.rt, self.dim, self.inner
        xn = self.norm1(x, B * S)
        ln = self.norm2(latents, B * nq)
        q = self.to_q(ln, B * nq)
        kv_in = rt.empty(B * (S + nq), dim)
        ops.cast2d(xn, S * dim, kv_in, (S + nq) * dim, rows=B, cols=S * dim)
        ops.cast2d(ln, nq * dim, kv_in.view(-1)[S * dim:], (S + nq) * dim, rows=B, cols=nq * dim)
        kv = self.to_kv(kv_in, B * (S + nq))
        o = rt.empty(B * nq, inner)
        ops.attention(q, 0, inner, [kv], [0], [2 * inner], [kv], [inner], [2 * inner], [S + nq], [1.0], o, 0, inner,
                      B=B, heads=self.heads, n_q=nq, softmax_scale=self.scale, tc=rt.tc)
        self.to_out(o, B * nq, out=latents, residual=latents)


class _FeedForward:
    """LayerNorm -> Linear(no bias) -> GELU -> Linear(no bias) (module/ip_adapter/resampler.py:13-21)."""

    def __init__(self, rt, src, p, dim):
        self.norm = LayerNorm(rt, src, p + ".0", dim)
        self.fc1 = Linear(rt, src, p + ".1", bias=False)
        self.fc2 = Linear(rt, src, p + ".3", bias=False)

    def __call__(self, latents, rows):
        h = self.fc1(self.norm(latents, rows), rows, act=ops.ACT_GELU)
        self.fc2(h, rows, out=latents, residual=latents)


class Resampler:
    def __init__(self, rt: Runtime, src, p, cfg):
        self.rt, self.cfg, self.dim = rt, cfg, cfg.resampler_dim
        self.latents = src.get(p + ".latents").contiguous()  # [1, nq, dim] fp32
        self.num_queries = self.latents.shape[1]
        self.proj_in = Linear(rt, src, p + ".proj_in")
        self.proj_out = Linear(rt, src, p + ".proj_out")
        self.norm_out = LayerNorm(rt, src, p + ".norm_out", cfg.cross_attention_dim)
        self.layers = [(PerceiverAttention(rt, src, f"{p}.layers.{i}.0", self.dim, cfg.resampler_dim_head, cfg.resampler_heads),
                        _FeedForward(rt, src, f"{p}.layers.{i}.1", self.dim)) for i in range(cfg.resampler_depth)]

    def __call__(self, x, out=None):
        """x [B, S, D] -> [B, nq, cross_attention_dim] in the runtime's activation dtype."""
        rt, dim, nq = self.rt, self.dim, self.num_queries
        B, S, D = x.shape
        xa = rt.empty(B * S, D)
        ops.cast2d(x.reshape(B * S, D).contiguous(), D, xa, D, rows=B * S, cols=D)
        xp = self.proj_in(xa, B * S)
        latents = torch.empty(B * nq, dim, device=rt.device, dtype=torch.float32)
        ops.cast2d(self.latents.view(1, nq * dim).expand(B, nq * dim), 0, latents, nq * dim, rows=B, cols=nq * dim)
        for attn, ff in self.layers:
            attn(xp, latents, B, S, nq)
            ff(latents, B * nq)
        la = rt.empty(B * nq, dim)
        ops.cast2d(latents, dim, la, dim, rows=B * nq, cols=dim)
        y = self.proj_out(la, B * nq, out_dtype=torch.float32)
        if out is None:
            out = rt.empty(B, nq, self.cfg.cross_attention_dim)
        ops.layernorm(y, self.norm_out.g, self.norm_out.b, out, rows=B * nq, C=self.cfg.cross_attention_dim, eps=1e-5)
        return out


class MultiIPAdapterImageProjection:
    """module/ip_adapter/ip_adapter.py:63-90 (single image-prompt adapter, the live configuration)."""

    def __init__(self, IPAdapterImageProjectionLayers):
        self.image_projection_layers = list(IPAdapterImageProjectionLayers)

    def __call__(self, image_embeds, out=None):
        if not isinstance(image_embeds, list):
            image_embeds = [image_embeds.unsqueeze(1)]
        if len(image_embeds) != len(self.image_projection_layers):
            raise ValueError(
                f"image_embeds must have the same length as image_projection_layers, got {len(image_embeds)} and {len(self.image_projection_layers)}")
        outs = []
        for e, layer in zip(image_embeds, self.image_projection_layers):
            b, n = e.shape[0], e.shape[1]
            outs.append(layer(e.reshape((b * n,) + tuple(e.shape[2:])), out=out[0] if out is not None else None))
        return outs
