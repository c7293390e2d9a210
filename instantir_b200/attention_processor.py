"""Attention modules and processors with the reference's processor API
(module/ip_adapter/attention_processor.py): ``proc(attn, hidden_states, encoder_hidden_states=None,
attention_mask=None, external_kv=None, temb=None) -> Tensor`` on ``[B, n, C]`` CUDA tensors.

Live processors only: ``AttnProcessor2_0`` (:323-414) and ``TA_IPAttnProcessor2_0``
(:1063-1207) with ``AdaLayerNorm`` (:6-26); factory ``init_attn_proc`` (:1364-1415).
One extension: an optional ``residual=`` keyword fuses the block's residual add into the
``to_out`` GEMM epilogue (the reference adds it in BasicTransformerBlock, module/min_sdxl.py:546-552).

Step-invariant tensors (text K/V, pre-adaLN image K/V) are cached per module and recomputed only
when the conditioning tensors change (``invalidate()``), instead of 140x per step as in the
reference.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .nn import Linear, LNStream, Runtime, SmallLinear, _Packed, _bias, _load_w

HEAD_DIM = 64
SOFTMAX_SCALE = HEAD_DIM ** -0.5


def _key(t: torch.Tensor):
    return (t.data_ptr(), tuple(t.shape), t.dtype, t._version)


class CtxCache:
    """slot -> tensor cache keyed by identity+version of the conditioning tensor.  The key tensor is
    kept alive (so its address cannot be recycled for different data) and output buffers are reused
    in place, so CUDA graphs that captured their addresses stay valid after a refresh."""

    def __init__(self):
        self.buf, self.valid, self.keep = {}, {}, {}

    def get(self, slot, tensor, make):
        key = _key(tensor)
        if self.valid.get(slot) == key:
            return self.buf[slot]
        out = make(self.buf.get((slot, key[1])))  # reuse a buffer only for the same conditioning shape
        # the kernels wrote `out` through raw pointers: tell torch, so that caches keyed on `out`
        # downstream (e.g. image K/V keyed on the Resampler tokens) see a new version
        for t in (out if isinstance(out, (list, tuple)) else [out]):
            torch.autograd.graph.increment_version(t)
        self.buf[(slot, key[1])] = out
        self.buf[slot], self.valid[slot], self.keep[slot] = out, key, tensor
        return out

    def invalidate(self):
        self.valid.clear()
        self.keep.clear()


def silu_of(rt: Runtime, temb: torch.Tensor) -> torch.Tensor:
    """silu(temb) computed once per temb tensor (resnets and adaLN all consume it)."""
    key = _key(temb)
    hit = getattr(rt, "_silu_cache", None)
    if hit is not None and hit[0] == key:
        return hit[1]
    out = torch.empty_like(temb, dtype=torch.float32)
    ops.silu(temb, out)
    rt._silu_cache = (key, out, temb)  # holding temb keeps its address from being recycled
    return out


class _Fused:
    """several bias-free nn.Linear weights stacked along N (fused QKV / KV projections)."""

    def __init__(self, rt, src, names, fold=None):
        parts = [_load_w(rt, src, n, fold=fold) for n in names]
        any_lora = any(p.lora is not None for p in parts)

        def cat(get):
            base = torch.cat([get(p).base for p in parts], 0).contiguous()
            lora = None
            if any_lora:
                lora = torch.cat([get(p).lora if get(p).lora is not None else get(p).base for p in parts], 0).contiguous()
            return _Packed(rt, base, lora)

        self.rt, self.w = rt, cat(lambda p: p)
        self.folded = fold is not None
        if self.folded:
            self.colsum, self.bias = cat(lambda p: p.colsum), cat(lambda p: p.bias)
        self.N, self.K = self.w.base.shape

    def __call__(self, a, M, out=None, out_dtype=None):
        if out is None:
            out = torch.empty(M, self.N, device=self.rt.device, dtype=out_dtype or self.rt.act_dtype)
        if isinstance(a, LNStream) != self.folded:
            raise ops._lib.IIRError("a LayerNorm-folded projection takes the LNStream of its transformer block (and only it)")
        if self.folded:
            ops.gemm(a.h16, self.w.get(), out, M=M, N=self.N, K=self.K, bias=self.bias.get(), tc=True,
                     ln_in=(a, self.colsum.get(), a.eps))
        else:
            ops.gemm(a, self.w.get(), out, M=M, N=self.N, K=self.K, tc=self.rt.tc)
        return out


class Attention:
    """diffusers Attention as configured for SDXL (q/k/v no bias, out bias, head_dim 64) exposing
    the fields the reference processors read (.to_q/.to_k/.to_v/.to_out/.heads/...)."""

    def __init__(self, rt: Runtime, src, p: str, C: int, heads: int, cross_dim: Optional[int] = None, fold=None):
        """fold = (gamma, beta) of the block's LayerNorm in front of this attention: its query-side projection
        then takes the block's LNStream instead of a normalised tensor (nn.LNStream)."""
        self.rt, self.C, self.heads, self.cross_dim = rt, C, heads, cross_dim
        if cross_dim is None:
            self.to_qkv = _Fused(rt, src, [p + ".to_q", p + ".to_k", p + ".to_v"], fold=fold)
        else:
            self.to_q = Linear(rt, src, p + ".to_q", bias=False, fold=fold)
            self.to_kv = _Fused(rt, src, [p + ".to_k", p + ".to_v"])
        self.to_out = [Linear(rt, src, p + ".to_out.0"), None]
        self.processor = AttnProcessor2_0()
        self.ctx = CtxCache()
        self.spatial_norm = self.group_norm = None
        self.norm_cross = False
        self.residual_connection = False
        self.rescale_output_factor = 1.0

    def set_processor(self, processor):
        self.processor = processor

    def get_processor(self, return_deprecated_lora=False):
        return self.processor

    def __call__(self, hidden_states, encoder_hidden_states=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, **kw)

    # -- helpers shared by the processors
    def _to_act(self, x3d):
        if isinstance(x3d, LNStream):
            return x3d
        B, n, C = x3d.shape
        x = x3d.reshape(B * n, C)
        if x.dtype != self.rt.act_dtype or not x.is_contiguous():
            y = self.rt.empty(B * n, C)
            ops.cast2d(x.contiguous(), C, y, C, rows=B * n, cols=C)
            x = y
        return x

    def text_kv(self, text):
        B, nt, D = text.shape
        return self.ctx.get("text_kv", text, lambda out: self.to_kv(self._to_act(text), B * nt, out=out))

    def project_out(self, o, M, residual):
        if isinstance(residual, LNStream):  # in place on the fp32 stream; refreshes its 16-bit copy + row statistics
            self.to_out[0](o, M, out=residual.h, residual=residual.h, ln_out=residual)
            return residual.h
        if residual is not None:
            r2 = residual.reshape(M, self.C)
            return self.to_out[0](o, M, out=r2, residual=r2)
        return self.to_out[0](o, M)


class AttnProcessor2_0:
    """module/ip_adapter/attention_processor.py:323-414 (self-attention in the UNet's attn1 and, in
    a UNet without the adapter, plain text cross-attention)."""

    def __init__(self, hidden_size=None, cross_attention_dim=None):
        pass

    def __call__(self, attn: Attention, hidden_states, encoder_hidden_states=None, attention_mask=None,
                 external_kv=None, temb=None, residual=None):
        if attention_mask is not None or external_kv:
            raise NotImplementedError("attention_mask / external_kv are unused on the live path (SURVEY §8 a5)")
        rt = attn.rt
        B, n, C = hidden_states.shape
        M = B * n
        hs = attn._to_act(hidden_states)
        o = rt.empty(M, C)
        if encoder_hidden_states is None:
            qkv = attn.to_qkv(hs, M)
            ops.attention(qkv, 0, 3 * C, [qkv], [C], [3 * C], [qkv], [2 * C], [3 * C], [n], [1.0], o, 0, C,
                          B=B, heads=attn.heads, n_q=n, softmax_scale=SOFTMAX_SCALE, tc=rt.tc, scratch_owner=id(rt))
        else:
            if isinstance(encoder_hidden_states, tuple):
                encoder_hidden_states = encoder_hidden_states[0]
            q = attn.to_q(hs, M)
            kv = attn.text_kv(encoder_hidden_states)
            nt = encoder_hidden_states.shape[1]
            ops.attention(q, 0, C, [kv], [0], [2 * C], [kv], [C], [2 * C], [nt], [1.0], o, 0, C,
                          B=B, heads=attn.heads, n_q=n, softmax_scale=SOFTMAX_SCALE, tc=rt.tc, scratch_owner=id(rt))
        out = attn.project_out(o, M, residual)
        return out.view(B, n, C)


class AdaLayerNorm:
    """module/ip_adapter/attention_processor.py:6-26: LN(eps 1e-6, no affine)(x)*(1+scale)+shift with
    (shift, scale) = linear(silu(temb)).chunk(2)."""

    def __init__(self, rt, src, p, C):
        self.rt, self.C = rt, C
        self.linear = SmallLinear(rt, src, p + ".linear")

    def __call__(self, x, timestep_embedding, rows_per_sample=None):
        rt = self.rt
        B = timestep_embedding.shape[0]
        x2 = x.reshape(-1, self.C)
        rows = x2.shape[0]
        mod = self.linear(silu_of(rt, timestep_embedding))
        out = rt.empty(rows, self.C)
        ops.layernorm(x2, None, None, out, rows=rows, C=self.C, eps=1e-6, mod=mod,
                      rows_per_sample=rows_per_sample or rows // B)
        return out


class TA_IPAttnProcessor2_0:
    """Decoupled text + image cross-attention with time-aware adaLN on the image K/V
    (module/ip_adapter/attention_processor.py:1063-1207).  Both softmaxes run in ONE launch of the
    two-segment attention kernel: out = SDPA(Q,K_t,V_t) + scale * SDPA(Q,K_i,V_i)."""

    def __init__(self, rt, src, p, hidden_size, cross_attention_dim=None, time_embedding_dim=None, scale=1.0,
                 num_tokens=4):
        self.rt = rt
        self.hidden_size, self.cross_attention_dim = hidden_size, cross_attention_dim
        self.scale, self.num_tokens = scale, num_tokens
        self.to_k_ip = Linear(rt, src, p + ".to_k_ip", bias=False)
        self.to_v_ip = Linear(rt, src, p + ".to_v_ip", bias=False)
        self.ln_k_ip = AdaLayerNorm(rt, src, p + ".ln_k_ip", hidden_size)
        self.ln_v_ip = AdaLayerNorm(rt, src, p + ".ln_v_ip", hidden_size)
        self.batch = None  # AdaLNKVBatch once installed on a UNet (unet.set_attn_processor)

    def __call__(self, attn: Attention, hidden_states, encoder_hidden_states=None, attention_mask=None,
                 external_kv=None, temb=None, residual=None):
        assert temb is not None, "Timestep embedding is needed for a time-aware attention processor."
        if attention_mask is not None or external_kv:
            raise NotImplementedError("attention_mask / external_kv are unused on the live path (SURVEY §8 a6)")
        rt = attn.rt
        if not isinstance(encoder_hidden_states, tuple):
            end_pos = encoder_hidden_states.shape[1] - self.num_tokens
            text = encoder_hidden_states[:, :end_pos, :].contiguous()
            ip = encoder_hidden_states[:, end_pos:, :].contiguous()
        else:
            ip = encoder_hidden_states[1][0]
            text = encoder_hidden_states[0]
        B, n, C = hidden_states.shape
        M = B * n
        nt, ni = text.shape[1], ip.shape[1]
        q = attn.to_q(attn._to_act(hidden_states), M)
        kv_t = attn.text_kv(text)
        if self.batch is not None and B <= 16:
            # every layer's adaLN'd image K/V were produced by ONE launch at the first layer of this forward
            k_i, v_i = self.batch.outputs(self, ip, temb)
        else:
            k_pre, v_pre = self._pre_kv(attn, ip)
            k_i = self.ln_k_ip(k_pre, temb, rows_per_sample=ni)
            v_i = self.ln_v_ip(v_pre, temb, rows_per_sample=ni)
        o = rt.empty(M, C)
        ops.attention(q, 0, C, [kv_t, k_i], [0, 0], [2 * C, C], [kv_t, v_i], [C, 0], [2 * C, C], [nt, ni],
                      [1.0, float(self.scale)], o, 0, C, B=B, heads=attn.heads, n_q=n,
                      softmax_scale=SOFTMAX_SCALE, tc=rt.tc)
        out = attn.project_out(o, M, residual)
        return out.view(B, n, C)


def _ta_pre_kv(self, attn, ip):
    """to_k_ip / to_v_ip of the image tokens before adaLN: step-invariant, cached per LoRA state."""
    rt = attn.rt
    B, ni, _ = ip.shape
    C = self.hidden_size
    tag = "L" if rt.lora_enabled else "B"
    ip_act = []

    def pre(lin, which):
        def make(out):
            if not ip_act:
                ip_act.append(attn._to_act(ip))
            if self.batch is not None and B <= 16:
                out = self.batch.pre_buffer(self, which, tag, B * ni)   # pooled: stable address for the batched launch
            elif out is None:
                out = torch.empty(B * ni, C, device=rt.device, dtype=torch.float32)
            return lin(ip_act[0], B * ni, out=out)
        return make

    return attn.ctx.get("ipk" + tag, ip, pre(self.to_k_ip, 0)), attn.ctx.get("ipv" + tag, ip, pre(self.to_v_ip, 1))


class AdaLNKVBatch:
    """The time-aware image K/V of ALL cross-attention layers in one launch per forward.

    K_i = adaLN_k(to_k_ip(img), temb), V_i = adaLN_v(to_v_ip(img), temb)
    (module/ip_adapter/attention_processor.py:1173-1178): the projections are step-invariant (cached), the
    modulation depends only on temb, so nothing here depends on the layer's hidden states.  The 140 adaLN
    linears are one banked launch (nn.SmallLinearBank) and the 140 LayerNorm+modulate passes one
    `iir_adaln_batched` launch over a device table of (pre-projection, output, mod offset, C)."""

    def __init__(self, rt: Runtime, pairs):
        self.rt = rt
        self.procs = [p for p, _ in pairs]
        self.attn = {id(p): a for p, a in pairs}
        self.idx = {id(p): k for k, p in enumerate(self.procs)}
        self.state = {}
        self.bank = rt.adaln_bank
        for p in self.procs:
            self.bank.add(p.ln_k_ip.linear)
            self.bank.add(p.ln_v_ip.linear)
            p.batch = self
        self.bank.listeners.append(self._invalidate)  # new modulation vectors -> every state is stale

    def _invalidate(self, _out=None):
        for st in self.state.values():
            st["key"] = None

    def _st(self, tag, rows):
        st = self.state.get((tag, rows))
        if st is None:
            rt = self.rt
            self.bank.build()
            pre = [[torch.empty(rows, p.hidden_size, device=rt.device, dtype=torch.float32) for _ in range(2)] for p in self.procs]
            out = [[rt.empty(rows, p.hidden_size) for _ in range(2)] for p in self.procs]
            entries = []
            for k, p in enumerate(self.procs):
                entries.append((pre[k][0], out[k][0], p.ln_k_ip.linear.off, p.hidden_size))
                entries.append((pre[k][1], out[k][1], p.ln_v_ip.linear.off, p.hidden_size))
            nbytes = sum(rows * c * (4 + out[0][0].element_size()) for _, _, _, c in entries)
            st = dict(pre=pre, out=out, table=ops.adaln_items(entries, rt.device), n=len(entries), key=None, bytes=float(nbytes))
            self.state[(tag, rows)] = st
        return st

    def pre_buffer(self, proc, which, tag, rows):
        return self._st(tag, rows)["pre"][self.idx[id(proc)]][which]

    def outputs(self, proc, ip, temb):
        rt = self.rt
        B, ni, _ = ip.shape
        tag = "L" if rt.lora_enabled else "B"
        st = self._st(tag, B * ni)
        mod = self.bank.result(silu_of(rt, temb))
        key = _key(ip)
        if st["key"] != key:
            for p in self.procs:  # make sure every layer's cached pre-projection is current (dict hits when it is)
                p._pre_kv(self.attn[id(p)], ip)
            ops.adaln_batched(st["table"], st["n"], mod, rt.act_dtype, rows=B * ni, rows_per_sample=ni, eps=1e-6,
                              bytes_moved=st["bytes"])
            st["key"] = key
        k = self.idx[id(proc)]
        return st["out"][k][0], st["out"][k][1]


def _ta_prefetch(self, attn, encoder_hidden_states):
    if isinstance(encoder_hidden_states, tuple):
        attn.text_kv(encoder_hidden_states[0])
        self._pre_kv(attn, encoder_hidden_states[1][0])


TA_IPAttnProcessor2_0._pre_kv = _ta_pre_kv
TA_IPAttnProcessor2_0.prefetch = _ta_prefetch


def init_attn_proc(unet, ip_adapter_tokens=16, use_lcm=False, use_adaln=True, use_external_kv=False):
    """module/ip_adapter/attention_processor.py:1364-1415: {name: processor} for unet.set_attn_processor.
    attn1 -> AttnProcessor2_0, attn2 -> TA_IPAttnProcessor2_0 (weights under '<attn2>.processor.*'; when
    the source has no such keys, to_k_ip/to_v_ip start as copies of to_k/to_v like the reference)."""
    if use_external_kv or not use_adaln:
        raise NotImplementedError("only the live configuration (use_adaln=True, no external kv) is built")
    procs = {}
    for name, attn in unet.attention_modules():
        if name.endswith("attn1"):
            procs[name + ".processor"] = AttnProcessor2_0()
        else:
            procs[name + ".processor"] = TA_IPAttnProcessor2_0(
                unet.rt, _ProcSrc(unet.source, name), name + ".processor", attn.C, unet.cfg.cross_attention_dim,
                time_embedding_dim=unet.cfg.time_embed_dim, scale=unet.cfg.ip_scale, num_tokens=ip_adapter_tokens)
    return procs


class _ProcSrc:
    """falls back to the attention's own to_k/to_v (and zero adaLN) when the processor keys are
    absent — the reference's initialisation (attention_processor.py:14-16,1395-1411)."""

    def __init__(self, src, attn_name):
        self.src, self.attn_name = src, attn_name

    def has(self, k):
        return self.src.has(k) or k.endswith((".to_k_ip.weight", ".to_v_ip.weight", ".linear.weight", ".linear.bias"))

    def get(self, k):
        if self.src.has(k):
            return self.src.get(k)
        if k.endswith(".to_k_ip.weight"):
            return self.src.get(self.attn_name + ".to_k.weight")
        if k.endswith(".to_v_ip.weight"):
            return self.src.get(self.attn_name + ".to_v.weight")
        ref = self.src.get(self.attn_name + ".to_q.weight")
        T = self.src.get("time_embedding.linear_2.weight").shape[0]
        C = ref.shape[0]
        if k.endswith(".linear.weight"):
            return torch.zeros(2 * C, T, device=ref.device)
        if k.endswith(".linear.bias"):
            return torch.zeros(2 * C, device=ref.device)
        raise KeyError(k)

    def get_lora(self, m):
        return self.src.get_lora(m)
