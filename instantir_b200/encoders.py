"""Once-per-image encoders on the sm_100a kernels: the two CLIP text encoders and the DINOv2 image
encoder the reference pipeline calls through ``transformers`` (pipelines/sdxl_instantir.py:522,580 ``encode_prompt``;
:643-667 ``encode_image``; loaded at module/ip_adapter/utils.py:106-118).

``transformers==4.36.2`` (requirements.txt:13) is a third-party dependency absent from /root/reference; what is restated
here is its published algorithm (``modeling_clip.py`` CLIPTextTransformer / CLIPTextModelWithProjection,
``modeling_dinov2.py`` Dinov2Model), with the state-dict key layout of those classes so real checkpoints load as is:

* CLIPTextModel: token + position embeddings -> L x [LN1 -> causal self-attention (q/k/v/out with bias, head_dim 64) -> +res;
  LN2 -> fc1 -> quick_gelu | gelu -> fc2 -> +res] -> final LN; pooled = final-LN row of the EOS token; the
  ``WithProjection`` variant returns ``text_embeds = text_projection(pooled)``.  ``hidden_states[k]`` is the residual
  stream entering layer k (the pipeline reads ``hidden_states[-2]``).
* Dinov2Model: 14x14 patch embedding (a GEMM over patch rows) + cls token + bicubically interpolated position
  embeddings -> L x [LN1 -> self-attention -> layer_scale1 -> +res; LN2 -> fc1 -> gelu -> fc2 -> layer_scale2 -> +res]
  -> final LN (eps 1e-6).  LayerScale is folded into the out-projection / fc2 weights at pack time.

Same kernels as the denoising step: tcgen05 GEMMs with bias / activation / fp32-residual epilogues, the flash attention
kernel (``causal=1`` for CLIP), the LayerNorm kernel; fp32 residual stream, 16-bit GEMM operands (fp32 check mode:
everything fp32 on the SIMT kernels).
"""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional

import torch

from . import ops
from .nn import Runtime


# ------------------------------------------------------------------------------------------ configs
@dataclass
class CLIPTextConfig:
    """transformers CLIPTextConfig fields that shape the arithmetic.  Defaults = SDXL's ``text_encoder`` (CLIP ViT-L/14)."""
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5
    projection_dim: int = 768
    eos_token_id: int = 2  # SDXL's config.json files keep the legacy value: the pooled row is input_ids.argmax(-1)

    def __post_init__(self):
        if self.hidden_size != 64 * self.num_attention_heads:
            raise ValueError("head_dim must be 64 (the sm_100a attention kernel)")
        if self.hidden_act not in ("quick_gelu", "gelu"):
            raise ValueError(f"unsupported hidden_act {self.hidden_act!r}")


def clip_l() -> CLIPTextConfig:
    return CLIPTextConfig()


def clip_bigg() -> CLIPTextConfig:
    """SDXL's ``text_encoder_2`` (OpenCLIP ViT-bigG/14 text tower)"""
    return CLIPTextConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=20,
                          hidden_act="gelu", projection_dim=1280)


@dataclass
class Dinov2Config:
    """transformers Dinov2Config.  Defaults = facebook/dinov2-large (module/ip_adapter/utils.py:106-118)."""
    hidden_size: int = 1024
    num_hidden_layers: int = 24
    num_attention_heads: int = 16
    mlp_ratio: int = 4
    image_size: int = 518   # the position table covers (image_size / patch_size)^2 patches
    patch_size: int = 14
    num_channels: int = 3
    layer_norm_eps: float = 1e-6

    def __post_init__(self):
        if self.hidden_size != 64 * self.num_attention_heads:
            raise ValueError("head_dim must be 64 (the sm_100a attention kernel)")


# ------------------------------------------------------------------------------------ parameter shapes
def clip_text_param_shapes(cfg: CLIPTextConfig, with_projection: bool):
    d, f = cfg.hidden_size, cfg.intermediate_size
    s = {"text_model.embeddings.token_embedding.weight": (cfg.vocab_size, d),
         "text_model.embeddings.position_embedding.weight": (cfg.max_position_embeddings, d)}
    for i in range(cfg.num_hidden_layers):
        p = f"text_model.encoder.layers.{i}."
        for n in ("layer_norm1", "layer_norm2"):
            s[p + n + ".weight"], s[p + n + ".bias"] = (d,), (d,)
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"], s[p + f"self_attn.{n}.bias"] = (d, d), (d,)
        s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (f, d), (f,)
        s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (d, f), (d,)
    s["text_model.final_layer_norm.weight"], s["text_model.final_layer_norm.bias"] = (d,), (d,)
    if with_projection:
        s["text_projection.weight"] = (cfg.projection_dim, d)
    return s


def dinov2_param_shapes(cfg: Dinov2Config):
    d, f, ps = cfg.hidden_size, cfg.hidden_size * cfg.mlp_ratio, cfg.patch_size
    n_pos = (cfg.image_size // ps) ** 2 + 1
    s = {"embeddings.cls_token": (1, 1, d), "embeddings.mask_token": (1, d), "embeddings.position_embeddings": (1, n_pos, d),
         "embeddings.patch_embeddings.projection.weight": (d, cfg.num_channels, ps, ps),
         "embeddings.patch_embeddings.projection.bias": (d,)}
    for i in range(cfg.num_hidden_layers):
        p = f"encoder.layer.{i}."
        for n in ("norm1", "norm2"):
            s[p + n + ".weight"], s[p + n + ".bias"] = (d,), (d,)
        for n in ("query", "key", "value"):
            s[p + f"attention.attention.{n}.weight"], s[p + f"attention.attention.{n}.bias"] = (d, d), (d,)
        s[p + "attention.output.dense.weight"], s[p + "attention.output.dense.bias"] = (d, d), (d,)
        s[p + "layer_scale1.lambda1"], s[p + "layer_scale2.lambda1"] = (d,), (d,)
        s[p + "mlp.fc1.weight"], s[p + "mlp.fc1.bias"] = (f, d), (f,)
        s[p + "mlp.fc2.weight"], s[p + "mlp.fc2.bias"] = (d, f), (d,)
    s["layernorm.weight"], s["layernorm.bias"] = (d,), (d,)
    return s


# ---------------------------------------------------------------------------------- shared block
class _Lin:
    """weight [N, K] in the operand dtype + fp32 bias; optional per-output-channel scale folded in (LayerScale)"""

    def __init__(self, rt: Runtime, w, b=None, scale=None):
        if scale is not None:
            w = w * scale[:, None]
            b = None if b is None else b * scale
        self.rt, self.w = rt, w.to(rt.w_dtype).contiguous()
        self.b = None if b is None else b.float().contiguous()
        self.N, self.K = self.w.shape

    def __call__(self, a, M, out=None, out_dtype=None, residual=None, act=ops.ACT_NONE):
        if out is None:
            out = torch.empty(M, self.N, device=self.rt.device, dtype=out_dtype or self.rt.act_dtype)
        return ops.gemm(a, self.w, out, M=M, N=self.N, K=self.K, lda=a.stride(0), bias=self.b, residual=residual, act=act, tc=self.rt.tc)


class _EncoderLayer:
    """pre-LN transformer layer shared by CLIP's text encoder and DINOv2 (names differ, arithmetic does not)"""

    def __init__(self, rt, src, names, d, heads, eps, act, causal, ls1=None, ls2=None):
        g = src.get
        self.rt, self.d, self.heads, self.eps, self.act, self.causal = rt, d, heads, eps, act, causal
        self.ln1 = (g(names["ln1"] + ".weight").contiguous(), g(names["ln1"] + ".bias").contiguous())
        self.ln2 = (g(names["ln2"] + ".weight").contiguous(), g(names["ln2"] + ".bias").contiguous())
        self.qkv = _Lin(rt, torch.cat([g(names[k] + ".weight") for k in ("q", "k", "v")], 0),
                        torch.cat([g(names[k] + ".bias") for k in ("q", "k", "v")], 0))
        self.out = _Lin(rt, g(names["o"] + ".weight"), g(names["o"] + ".bias"), scale=None if ls1 is None else g(ls1))
        self.fc1 = _Lin(rt, g(names["fc1"] + ".weight"), g(names["fc1"] + ".bias"))
        self.fc2 = _Lin(rt, g(names["fc2"] + ".weight"), g(names["fc2"] + ".bias"), scale=None if ls2 is None else g(ls2))

    def __call__(self, h, B, n):
        """h: fp32 residual stream [B * n, d], updated in place"""
        rt, d, M = self.rt, self.d, B * n
        a = rt.empty(M, d)
        ops.layernorm(h, self.ln1[0], self.ln1[1], a, rows=M, C=d, eps=self.eps)
        qkv = self.qkv(a, M)
        o = rt.empty(M, d)
        ops.attention(qkv, 0, 3 * d, [qkv], [d], [3 * d], [qkv], [2 * d], [3 * d], [n], [1.0], o, 0, d, B=B, heads=self.heads,
                      n_q=n, softmax_scale=0.125, tc=rt.tc, scratch_owner=id(rt), causal=self.causal)
        self.out(o, M, out=h, residual=h)
        ops.layernorm(h, self.ln2[0], self.ln2[1], a, rows=M, C=d, eps=self.eps)
        f = self.fc1(a, M, act=self.act)
        self.fc2(f, M, out=h, residual=h)
        return h


# ------------------------------------------------------------------------------------------- CLIP
class CLIPTextModel:
    """transformers CLIPTextModel / CLIPTextModelWithProjection (``with_projection=True``) as encode_prompt uses them:
    ``enc(input_ids, output_hidden_states=True)`` -> ``[0]`` (last_hidden_state, or text_embeds with the projection),
    ``.pooler_output``, ``.hidden_states`` (tuple of L + 1 fp32 tensors [B, S, d])."""

    def __init__(self, cfg: CLIPTextConfig, source, device="cuda", precision="fp16", with_projection=False):
        self.config, self.rt = cfg, Runtime(device, precision)
        self.dtype = self.rt.act_dtype
        rt, g = self.rt, source.get
        self.tok = g("text_model.embeddings.token_embedding.weight").contiguous()
        self.pos = g("text_model.embeddings.position_embedding.weight").contiguous()
        act = ops.ACT_QUICK_GELU if cfg.hidden_act == "quick_gelu" else ops.ACT_GELU
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"text_model.encoder.layers.{i}."
            names = dict(ln1=p + "layer_norm1", ln2=p + "layer_norm2", q=p + "self_attn.q_proj", k=p + "self_attn.k_proj",
                         v=p + "self_attn.v_proj", o=p + "self_attn.out_proj", fc1=p + "mlp.fc1", fc2=p + "mlp.fc2")
            self.layers.append(_EncoderLayer(rt, source, names, cfg.hidden_size, cfg.num_attention_heads, cfg.layer_norm_eps, act, True))
        self.final_ln = (g("text_model.final_layer_norm.weight").contiguous(), g("text_model.final_layer_norm.bias").contiguous())
        self.text_projection = _Lin(rt, g("text_projection.weight")) if with_projection else None

    def __call__(self, input_ids, output_hidden_states=True, **unused):
        cfg, rt = self.config, self.rt
        ids = torch.as_tensor(input_ids)
        if ids.ndim != 2 or ids.shape[1] > cfg.max_position_embeddings:
            raise ValueError(f"input_ids must be [batch, <= {cfg.max_position_embeddings}] token ids")
        B, S = ids.shape
        d, M = cfg.hidden_size, B * S
        ids_dev = ids.to(device=rt.device, dtype=torch.int64).contiguous()
        h = torch.empty(M, d, device=rt.device, dtype=torch.float32)
        ops.embed_tokens(ids_dev.view(-1), self.tok, self.pos, h, seq_len=S)
        hidden = [h.clone().view(B, S, d)] if output_hidden_states else None
        for layer in self.layers:
            layer(h, B, S)
            if output_hidden_states:
                hidden.append(h.clone().view(B, S, d))
        last = torch.empty(M, d, device=rt.device, dtype=torch.float32)
        ops.layernorm(h, self.final_ln[0], self.final_ln[1], last, rows=M, C=d, eps=cfg.layer_norm_eps)
        # pooled row: the EOS token (legacy configs, eos_token_id == 2: the highest id in the sequence)
        ids_cpu = ids.to("cpu", torch.int64)
        eos = ids_cpu.argmax(-1) if cfg.eos_token_id == 2 else (ids_cpu == cfg.eos_token_id).int().argmax(-1)
        pooled = torch.empty(B, d, device=rt.device, dtype=torch.float32)
        for b in range(B):
            ops.cast2d(last[b * S + int(eos[b])], d, pooled[b], d, rows=1, cols=d)
        out = SimpleNamespace(last_hidden_state=last.view(B, S, d), pooler_output=pooled,
                              hidden_states=tuple(hidden) if output_hidden_states else None)
        if self.text_projection is not None:
            a = pooled
            if rt.tc:
                a = rt.empty(B, d)
                ops.cast2d(pooled, d, a, d, rows=B, cols=d)
            out.text_embeds = self.text_projection(a, B, out_dtype=torch.float32)
            out.first = out.text_embeds
        else:
            out.first = out.last_hidden_state
        return _Indexable(out)


class _Indexable(SimpleNamespace):
    """ModelOutput-style ``out[0]`` access (encode_prompt reads ``prompt_embeds[0]``, pipelines/sdxl_instantir.py:526)"""

    def __init__(self, ns):
        super().__init__(**vars(ns))

    def __getitem__(self, i):
        if i == 0:
            return self.first
        raise IndexError("only [0] (text_embeds / last_hidden_state) is provided; use the named fields")


# ----------------------------------------------------------------------------------------- DINOv2
class Dinov2Model:
    """transformers Dinov2Model: ``model(pixel_values).last_hidden_state`` [B, 1 + (H/14)(W/14), 1024]
    (encode_image, pipelines/sdxl_instantir.py:659-667)."""

    def __init__(self, cfg: Dinov2Config, source, device="cuda", precision="fp16"):
        self.config, self.rt = cfg, Runtime(device, precision)
        self.dtype = self.rt.act_dtype
        rt, g = self.rt, source.get
        d, ps = cfg.hidden_size, cfg.patch_size
        self.kk = cfg.num_channels * ps * ps
        self.kpad = (self.kk + 7) // 8 * 8   # 588 -> 592: the GEMM wants K % 8 == 0; the extra columns are zero on both sides
        w = g("embeddings.patch_embeddings.projection.weight").reshape(d, self.kk)
        wp = torch.zeros(d, self.kpad, device=w.device, dtype=torch.float32)
        wp[:, :self.kk] = w
        self.patch = _Lin(rt, wp, g("embeddings.patch_embeddings.projection.bias"))
        self.cls = g("embeddings.cls_token").reshape(d).contiguous()
        self.pos_table = g("embeddings.position_embeddings").reshape(-1, d).contiguous()
        self._pos_cache = {}
        self.layers = []
        for i in range(cfg.num_hidden_layers):
            p = f"encoder.layer.{i}."
            names = dict(ln1=p + "norm1", ln2=p + "norm2", q=p + "attention.attention.query", k=p + "attention.attention.key",
                         v=p + "attention.attention.value", o=p + "attention.output.dense", fc1=p + "mlp.fc1", fc2=p + "mlp.fc2")
            self.layers.append(_EncoderLayer(rt, source, names, d, cfg.num_attention_heads, cfg.layer_norm_eps, ops.ACT_GELU, False,
                                             ls1=p + "layer_scale1.lambda1", ls2=p + "layer_scale2.lambda1"))
        self.final_ln = (g("layernorm.weight").contiguous(), g("layernorm.bias").contiguous())

    def parameters(self):  # encode_image reads next(image_encoder.parameters()).dtype
        yield self.patch.w

    def position_embeddings(self, gh: int, gw: int, mode: str = "size"):
        """Dinov2Embeddings.interpolate_pos_encoding: the (image_size/patch)^2 table bicubically resampled to gh x gw.  A
        function of the weights and the input resolution only, so it is computed once per resolution at load time with
        torch's interpolate (weight preparation, like packing); ``mode='size'`` = transformers >= 4.38 (target size),
        ``'scale_0.1'`` = the 4.36.2 pinned by the reference (scale_factor with the +0.1 offset of the original DINOv2 code)."""
        key = (gh, gw, mode)
        if key not in self._pos_cache:
            n_pos = self.pos_table.shape[0] - 1
            side = int(round(n_pos ** 0.5))
            if gh * gw == n_pos and gh == gw:
                pos = self.pos_table
            else:
                grid = self.pos_table[1:].reshape(1, side, side, -1).permute(0, 3, 1, 2).float()
                if mode == "size":
                    grid = torch.nn.functional.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
                else:
                    grid = torch.nn.functional.interpolate(grid, scale_factor=((gh + 0.1) / side, (gw + 0.1) / side), mode="bicubic",
                                                           align_corners=False)
                    if grid.shape[-2:] != (gh, gw):
                        raise ValueError("position-embedding interpolation produced the wrong grid")
                pos = torch.cat([self.pos_table[:1], grid.permute(0, 2, 3, 1).reshape(gh * gw, -1)], 0)
            self._pos_cache[key] = pos.contiguous()
        return self._pos_cache[key]

    def __call__(self, pixel_values, output_hidden_states=False, pos_mode: str = "size", **unused):
        cfg, rt = self.config, self.rt
        x = pixel_values.to(device=rt.device, dtype=torch.float32).contiguous()
        B, C, H, W = x.shape
        ps, d = cfg.patch_size, cfg.hidden_size
        if C != cfg.num_channels or H % ps or W % ps:
            raise ValueError(f"pixel_values must be [B, {cfg.num_channels}, H, W] with H, W multiples of {ps}")
        gh, gw = H // ps, W // ps
        P, n = gh * gw, gh * gw + 1
        rows = torch.empty(B * P, self.kpad, device=rt.device, dtype=rt.act_dtype)
        ops.patchify(x, rows, patch=ps)
        emb = self.patch(rows, B * P, out_dtype=torch.float32)
        h = torch.empty(B * n, d, device=rt.device, dtype=torch.float32)
        ops.vit_assemble(emb, self.cls, self.position_embeddings(gh, gw, pos_mode), h, n_img=B, P=P)
        hidden = [h.clone().view(B, n, d)] if output_hidden_states else None
        for layer in self.layers:
            layer(h, B, n)
            if output_hidden_states:
                hidden.append(h.clone().view(B, n, d))
        last = torch.empty(B * n, d, device=rt.device, dtype=torch.float32)
        ops.layernorm(h, self.final_ln[0], self.final_ln[1], last, rows=B * n, C=d, eps=cfg.layer_norm_eps)
        last = last.view(B, n, d)
        return SimpleNamespace(last_hidden_state=last, pooler_output=last[:, 0], hidden_states=tuple(hidden) if hidden else None)


# --------------------------------------------------------------------------- image preprocessing
DINOV2_MEAN = (0.485, 0.456, 0.406)
DINOV2_STD = (0.229, 0.224, 0.225)


def dinov2_preprocess(images01: torch.Tensor, shortest_edge: int = 256, crop: int = 224) -> torch.Tensor:
    """The AutoImageProcessor of facebook/dinov2-large (BitImageProcessor: resize shortest edge to 256 with bicubic
    resampling, center-crop 224, rescale, ImageNet normalise; module/ip_adapter/utils.py:113-118) for images already
    in [0, 1] as a [B, 3, H, W] tensor.  Host-side glue in torch (once per image, not a kernel of the hot path); PIL
    inputs can instead go through a user-supplied ``feature_extractor`` exactly as in the reference."""
    x = images01.float()
    B, C, H, W = x.shape
    s = shortest_edge / min(H, W)
    nh, nw = max(crop, int(round(H * s))), max(crop, int(round(W * s)))
    x = torch.nn.functional.interpolate(x, size=(nh, nw), mode="bicubic", align_corners=False, antialias=True).clamp(0, 1)
    top, left = (nh - crop) // 2, (nw - crop) // 2
    x = x[:, :, top:top + crop, left:left + crop]
    mean = torch.tensor(DINOV2_MEAN, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(DINOV2_STD, device=x.device).view(1, 3, 1, 1)
    return ((x - mean) / std).contiguous()
