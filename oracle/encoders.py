"""CPU fp32 restatement of the once-per-image encoders — TEST INFRASTRUCTURE, never on the product path.

The arithmetic lives in ``transformers==4.36.2`` (reference requirements.txt:13), a third-party dependency absent from
/root/reference; call sites: pipelines/sdxl_instantir.py:522,580 (CLIPTextModel / CLIPTextModelWithProjection inside
``encode_prompt``) and :643-667 (Dinov2Model inside ``encode_image``), loaded at module/ip_adapter/utils.py:106-118.  This file
restates the published algorithm of ``modeling_clip.py`` (CLIPTextTransformer) and ``modeling_dinov2.py`` (Dinov2Model) with
their parameter names.  PINNED: tests/golden/make_golden_encoders.py runs the transformers build installed in the authoring
container (5.5.0; same arithmetic as 4.36.2 except the position-embedding interpolation of DINOv2, see ``pos_mode``) on
seeded small configurations and commits inputs / outputs (tests/golden/encoders.pt); tests/test_oracle_golden.py replays them.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _attn(x, wq, wk, wv, wo, heads, causal):
    B, n, d = x.shape
    q, k, v = (lin(x).view(B, n, heads, d // heads).transpose(1, 2) for lin in (wq, wk, wv))
    s = q @ k.transpose(-1, -2) / math.sqrt(d // heads)
    if causal:  # CLIPTextTransformer builds a causal mask (_create_4d_causal_attention_mask)
        s = s + torch.full((n, n), float("-inf")).triu(1)
    return wo((s.softmax(-1) @ v).transpose(1, 2).reshape(B, n, d))


class CLIPLayer(nn.Module):
    def __init__(self, d, f, heads, act, eps):
        super().__init__()
        self.heads, self.act = heads, act
        self.layer_norm1, self.layer_norm2 = nn.LayerNorm(d, eps=eps), nn.LayerNorm(d, eps=eps)
        self.self_attn = nn.ModuleDict(dict(q_proj=nn.Linear(d, d), k_proj=nn.Linear(d, d), v_proj=nn.Linear(d, d), out_proj=nn.Linear(d, d)))
        self.mlp = nn.ModuleDict(dict(fc1=nn.Linear(d, f), fc2=nn.Linear(f, d)))

    def forward(self, h):
        a = self.self_attn
        h = h + _attn(self.layer_norm1(h), a["q_proj"], a["k_proj"], a["v_proj"], a["out_proj"], self.heads, True)
        y = self.mlp["fc1"](self.layer_norm2(h))
        y = y * torch.sigmoid(1.702 * y) if self.act == "quick_gelu" else F.gelu(y)
        return h + self.mlp["fc2"](y)


class CLIPTextModel(nn.Module):
    """CLIPTextModel (with_projection=False) / CLIPTextModelWithProjection; state-dict keys = transformers'."""

    def __init__(self, vocab_size, hidden_size, intermediate_size, num_hidden_layers, num_attention_heads, max_position_embeddings=77,
                 hidden_act="quick_gelu", layer_norm_eps=1e-5, projection_dim=None, eos_token_id=2):
        super().__init__()
        self.eos_token_id = eos_token_id
        tm = nn.Module()
        tm.embeddings = nn.Module()
        tm.embeddings.token_embedding = nn.Embedding(vocab_size, hidden_size)
        tm.embeddings.position_embedding = nn.Embedding(max_position_embeddings, hidden_size)
        tm.encoder = nn.Module()
        tm.encoder.layers = nn.ModuleList([CLIPLayer(hidden_size, intermediate_size, num_attention_heads, hidden_act, layer_norm_eps)
                                           for _ in range(num_hidden_layers)])
        tm.final_layer_norm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)
        self.text_model = tm
        self.text_projection = nn.Linear(hidden_size, projection_dim, bias=False) if projection_dim else None

    def forward(self, input_ids):
        tm = self.text_model
        B, S = input_ids.shape
        h = tm.embeddings.token_embedding(input_ids) + tm.embeddings.position_embedding(torch.arange(S))[None]
        hidden = [h]
        for layer in tm.encoder.layers:
            h = layer(h)
            hidden.append(h)
        last = tm.final_layer_norm(h)
        eos = input_ids.argmax(-1) if self.eos_token_id == 2 else (input_ids == self.eos_token_id).int().argmax(-1)
        pooled = last[torch.arange(B), eos]
        out = dict(last_hidden_state=last, pooler_output=pooled, hidden_states=tuple(hidden))
        if self.text_projection is not None:
            out["text_embeds"] = self.text_projection(pooled)
        return out


class Dinov2Layer(nn.Module):
    def __init__(self, d, f, heads, eps):
        super().__init__()
        self.heads = heads
        self.norm1, self.norm2 = nn.LayerNorm(d, eps=eps), nn.LayerNorm(d, eps=eps)
        att = nn.Module()
        att.attention = nn.ModuleDict(dict(query=nn.Linear(d, d), key=nn.Linear(d, d), value=nn.Linear(d, d)))
        att.output = nn.ModuleDict(dict(dense=nn.Linear(d, d)))
        self.attention = att
        self.layer_scale1, self.layer_scale2 = nn.Module(), nn.Module()
        self.layer_scale1.lambda1, self.layer_scale2.lambda1 = nn.Parameter(torch.ones(d)), nn.Parameter(torch.ones(d))
        self.mlp = nn.ModuleDict(dict(fc1=nn.Linear(d, f), fc2=nn.Linear(f, d)))

    def forward(self, h):
        a = self.attention
        y = _attn(self.norm1(h), a.attention["query"], a.attention["key"], a.attention["value"], a.output["dense"], self.heads, False)
        h = h + y * self.layer_scale1.lambda1
        y = self.mlp["fc2"](F.gelu(self.mlp["fc1"](self.norm2(h))))
        return h + y * self.layer_scale2.lambda1


class Dinov2Model(nn.Module):
    def __init__(self, hidden_size, num_hidden_layers, num_attention_heads, mlp_ratio=4, image_size=518, patch_size=14, num_channels=3,
                 layer_norm_eps=1e-6):
        super().__init__()
        self.patch_size = patch_size
        n_pos = (image_size // patch_size) ** 2 + 1
        e = nn.Module()
        e.cls_token = nn.Parameter(torch.randn(1, 1, hidden_size))
        e.mask_token = nn.Parameter(torch.zeros(1, hidden_size))
        e.position_embeddings = nn.Parameter(torch.randn(1, n_pos, hidden_size))
        e.patch_embeddings = nn.Module()
        e.patch_embeddings.projection = nn.Conv2d(num_channels, hidden_size, patch_size, stride=patch_size)
        self.embeddings = e
        self.encoder = nn.Module()
        self.encoder.layer = nn.ModuleList([Dinov2Layer(hidden_size, hidden_size * mlp_ratio, num_attention_heads, layer_norm_eps)
                                            for _ in range(num_hidden_layers)])
        self.layernorm = nn.LayerNorm(hidden_size, eps=layer_norm_eps)

    def pos(self, gh, gw, pos_mode="size"):
        table = self.embeddings.position_embeddings
        n_pos = table.shape[1] - 1
        side = int(round(n_pos ** 0.5))
        if gh * gw == n_pos and gh == gw:
            return table
        grid = table[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2)
        if pos_mode == "size":  # transformers >= 4.38
            grid = F.interpolate(grid, size=(gh, gw), mode="bicubic", align_corners=False)
        else:                   # transformers 4.36.2 (the reference's pin): scale factors with the +0.1 offset
            grid = F.interpolate(grid, scale_factor=((gh + 0.1) / side, (gw + 0.1) / side), mode="bicubic", align_corners=False)
        return torch.cat([table[:, :1], grid.permute(0, 2, 3, 1).reshape(1, gh * gw, -1)], 1)

    def forward(self, pixel_values, pos_mode="size"):
        e = self.embeddings
        B, _, H, W = pixel_values.shape
        x = e.patch_embeddings.projection(pixel_values).flatten(2).transpose(1, 2)
        h = torch.cat([e.cls_token.expand(B, -1, -1), x], 1) + self.pos(H // self.patch_size, W // self.patch_size, pos_mode)
        for layer in self.encoder.layer:
            h = layer(h)
        return self.layernorm(h)
