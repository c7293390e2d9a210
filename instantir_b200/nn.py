"""Building blocks of the host-side executor: light Python objects that mirror the reference's
module tree, hold weights pre-packed for the sm_100a kernels, and launch those kernels.

Data layout: activations are NHWC / token-major ``[n_img*H*W, C]``.
In ``bf16`` precision (``fp16`` is identical with IEEE-half operands: the reference's own inference
precision, 8x finer mantissa, same tensor-core rate) GEMM operands (outputs of norms, attention, GEGLU) are bf16 while the
*residual stream* (resnet outputs, transformer hidden states, skips) stays fp32 so that rounding
does not accumulate over the 200+ residual adds of a forward; ``fp32`` precision is the north
star's check mode (everything fp32, SIMT kernels).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import ops
from .weights import merge_lora


class Runtime:
    """Execution mode shared by every layer of a model."""

    def __init__(self, device, precision: str = "fp16"):
        if precision not in ("bf16", "fp16", "fp32"):
            raise ValueError("precision must be 'bf16' / 'fp16' (tcgen05 path) or 'fp32' (check mode)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ops._lib.IIRError("instantir_b200 runs on CUDA only; there is no CPU fallback")
        self.precision = precision
        self.tc = precision in ("bf16", "fp16")
        self.act_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[precision]
        self.w_dtype = self.act_dtype
        self.lora_enabled = False
        self.temb_bank = SmallLinearBank(self)    # resnet time_emb_proj (input: silu(emb))
        self.adaln_bank = SmallLinearBank(self)   # adaLN linears of the IP-adapter processors (input: silu(temb))
        # OPT-IN (IIR_GN_FUSE=1, tcgen05 path only): GroupNorm statistics are accumulated by the epilogue of
        # the GEMM / conv that PRODUCES the tensor, so GroupNorm is one pass (iir_groupnorm_apply_sums) instead of two.
        # Not yet verified on a GPU: off by default.
        self.gn_fuse = self.tc and os.environ.get("IIR_GN_FUSE", "0") == "1"
        self._gn_arena, self._gn_used = None, 0

    GN_ARENA_WORDS = 1 << 16  # int64 words: 67 GroupNorm sites x (CFG batch x 32 groups x 2) fits 8 images per forward

    def new_forward(self):
        """start of a model forward: per-forward caches are dropped"""
        self._silu_cache = None
        self.temb_bank.reset()
        self.adaln_bank.reset()
        if self.gn_fuse:
            if self._gn_arena is None:
                self._gn_arena = torch.zeros(self.GN_ARENA_WORDS, device=self.device, dtype=torch.int64)
            # every accumulator of the forward is cleared by ONE memset node at its start
            ops.memset_zero(self._gn_arena)
            self._gn_used = 0

    def gn_site(self, n_img: int, groups: int):
        """a zeroed [n_img, groups, 2] int64 accumulator for one GroupNorm input of this forward (None when the arena is
        exhausted: the caller falls back to the two-kernel GroupNorm)"""
        if not self.gn_fuse or self._gn_arena is None:
            return None
        n = n_img * groups * 2
        if self._gn_used + n > self._gn_arena.numel():
            return None
        t = self._gn_arena[self._gn_used:self._gn_used + n].view(n_img, groups, 2)
        self._gn_used += n
        return t

    def empty(self, *shape, dtype=None):
        return torch.empty(*shape, device=self.device, dtype=dtype or self.act_dtype)

    def stream(self, *shape):
        return torch.empty(*shape, device=self.device, dtype=torch.float32)


@dataclass
class FMap:
    """NHWC feature map: t is [n*H*W, C] contiguous."""

    t: torch.Tensor
    n: int
    H: int
    W: int
    C: int
    gn: Optional[torch.Tensor] = None  # GroupNorm (sum, sum of squares) accumulated by the kernel that produced `t` (opt-in)

    @property
    def M(self):
        return self.n * self.H * self.W

    def nchw(self):
        """zero-copy logical NCHW view (channels_last memory) for reference-shaped returns."""
        return self.t.view(self.n, self.H, self.W, self.C).permute(0, 3, 1, 2)


def fmap_from_nchw(x: torch.Tensor, dtype=None) -> FMap:
    """accept a reference-style [n,C,H,W] tensor; zero-copy when it is already channels_last."""
    n, c, h, w = x.shape
    t = x.permute(0, 2, 3, 1)
    if not t.is_contiguous():
        t = t.contiguous()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return FMap(t.reshape(n * h * w, c), n, h, w, c)


class _Packed:
    """weight (+ optional LoRA-merged twin) selected by rt.lora_enabled."""

    def __init__(self, rt: Runtime, base: torch.Tensor, lora: Optional[torch.Tensor] = None):
        self.rt, self.base, self.lora = rt, base, lora

    def get(self):
        return self.lora if (self.rt.lora_enabled and self.lora is not None) else self.base


def _pack_pairs(w1, w2, bn):
    """interleave two [N,K] (or [N]) tensors per bn-wide output tile: [w1 half | w2 half]."""
    half = bn // 2
    n = w1.shape[0]
    assert n % half == 0
    a = w1.reshape(n // half, half, *w1.shape[1:])
    b = w2.reshape(n // half, half, *w2.shape[1:])
    return torch.cat([a, b], dim=1).reshape(2 * n, *w1.shape[1:]).contiguous()


def _conv_to_gemm(w):  # [Co,Ci,k,k] -> [Co, k*k*Ci] tap-major
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


def _load_w(rt, src, name, transform=None, lora_ok=True, fold=None, bias=None):
    """returns _Packed of the (transformed) weight in rt.w_dtype; merges LoRA when the source has it.

    fold = (gamma, beta) of the LayerNorm in front of this linear (folded-LN GEMM, include/instantir_b200.h):
    the packed weight becomes W' = W∘gamma and the result carries ``.colsum`` = Σ_k W'[n,k] (of the ROUNDED
    weight, so that the mean term cancels exactly what the tensor cores accumulate) and ``.bias`` = W·beta + b,
    each with a LoRA-merged twin when the weight has one."""
    w = src.get(name + ".weight")

    def pack(w_):
        t = w_.float() * fold[0].float()[None, :] if fold is not None else w_
        t = transform(t) if transform else t
        return t.to(rt.w_dtype).contiguous()

    def extras(w_, packed):
        t = transform(w_.float()) if transform else w_.float()
        b = t @ fold[1].float()
        if bias is not None:
            b = b + bias.float()
        return packed.float().sum(1).contiguous(), b.contiguous()

    base = pack(w)
    lw = None
    lo = src.get_lora(name) if lora_ok else None
    m = None
    if lo is not None:
        m = merge_lora(w, lo)
        lw = pack(m)
    out = _Packed(rt, base, lw)
    if fold is not None:
        cs, bf = extras(w, base)
        cs_l = bf_l = None
        if m is not None:
            cs_l, bf_l = extras(m, lw)
        out.colsum, out.bias = _Packed(rt, cs, cs_l), _Packed(rt, bf, bf_l)
    return out


class LNStream:
    """The fp32 residual stream of a transformer stack plus what the folded-LayerNorm GEMMs exchange: a 16-bit
    copy (the consumers' A operand) and per-row fixed-point (sum, sum of squares) accumulated by whichever GEMM
    last updated the stream.  Quacks like the [B, n, C] hidden-states tensor the processors expect."""

    def __init__(self, rt, h, B, n, C, eps):
        self.rt, self.h, self.B, self.n, self.C, self.eps = rt, h, B, n, C, eps
        self.h16 = rt.empty(B * n, C)
        self.acc = torch.zeros(2, B * n, 2, device=rt.device, dtype=torch.int64)  # alternating row-sum accumulators
        self.cur = 0
        self.shape = (B, n, C)


def _bias(src, name):
    return src.get(name + ".bias").contiguous() if src.has(name + ".bias") else None


class Linear:
    """nn.Linear on [M,K] activations (module/min_sdxl.py:301-307 etc.) with fused epilogues."""

    def __init__(self, rt, src, name, bias=True, fold=None):
        self.rt = rt
        self.b = _bias(src, name) if bias else None
        self.w = _load_w(rt, src, name, fold=fold, bias=self.b)
        self.folded = fold is not None
        self.N, self.K = self.w.base.shape

    def __call__(self, a, M, out=None, out_dtype=None, residual=None, act=ops.ACT_NONE, ln_out=None, gn=None,
                 rows_per_sample=0):
        rt = self.rt
        if out is None:
            out = torch.empty(M, self.N, device=rt.device, dtype=out_dtype or rt.act_dtype)
        if isinstance(a, LNStream) != self.folded:
            raise ops._lib.IIRError("a LayerNorm-folded linear takes the LNStream of its transformer block (and only it)")
        if self.folded:
            ops.gemm(a.h16, self.w.get(), out, M=M, N=self.N, K=self.K, bias=self.w.bias.get(), residual=residual, act=act,
                     tc=True, ln_in=(a, self.w.colsum.get(), a.eps))
        else:
            ops.gemm(a, self.w.get(), out, M=M, N=self.N, K=self.K, bias=self.b, residual=residual, act=act, tc=rt.tc,
                     ln_out=ln_out, gn=gn, rows_per_sample=rows_per_sample)
        return out


class SmallLinear:
    """M <= 16 linear on fp32 vectors (time/add embeddings, time_emb_proj, adaLN).  When it belongs
    to a SmallLinearBank, the first call of a forward with a given input computes EVERY member of the
    bank in one launch and each member returns its column slice."""

    def __init__(self, rt, src, name, bank: "Optional[SmallLinearBank]" = None):
        self.rt = rt
        self.w = _load_w(rt, src, name)
        self.b = _bias(src, name)
        self.N, self.K = self.w.base.shape
        self.in_features, self.out_features = self.K, self.N  # nn.Linear's names (pipelines/sdxl_instantir.py:973)
        self.bank, self.off = None, 0
        if bank is not None:
            bank.add(self)

    def __call__(self, x, act=ops.ACT_NONE):
        M = x.shape[0]
        if self.bank is not None and act == ops.ACT_NONE and M <= 16:
            return self.bank.result(x)[:, self.off:self.off + self.N]
        out = torch.empty(M, self.N, device=self.rt.device, dtype=torch.float32)
        for m0 in range(0, M, 16):
            m1 = min(M, m0 + 16)
            ops.linear_small(x[m0:m1], self.w.get(), self.b, out[m0:m1], M=m1 - m0, N=self.N, K=self.K, act=act)
        return out


class SmallLinearBank:
    """All small linears of a model that consume the same vector (silu(temb): 17+8 resnet
    time_emb_proj, 140 adaLN linears) stacked along N: one HBM-bound launch per forward instead of
    one per layer.  Members' weights become row views of the stacked tensor (no copy is kept)."""

    def __init__(self, rt):
        self.rt, self.items, self.w, self.b = rt, [], None, None
        self.key, self.out, self.keep = None, None, None
        self.listeners = []  # called after each recompute with (bank output)

    def add(self, lin: SmallLinear):
        assert self.w is None, "bank already built"
        lin.bank = self
        self.items.append(lin)

    def build(self):
        if self.w is not None or not self.items:
            return self
        K = self.items[0].K
        assert all(l.K == K for l in self.items)
        off = 0
        for l in self.items:
            l.off = off
            off += l.N
        self.N, self.K = off, K
        dev = self.rt.device
        base = torch.empty(off, K, device=dev, dtype=self.rt.w_dtype)
        has_lora = any(l.w.lora is not None for l in self.items)
        lora = torch.empty(off, K, device=dev, dtype=self.rt.w_dtype) if has_lora else None
        bias = torch.zeros(off, device=dev, dtype=torch.float32)
        for l in self.items:
            sl = slice(l.off, l.off + l.N)
            base[sl] = l.w.base
            if lora is not None:
                lora[sl] = l.w.lora if l.w.lora is not None else l.w.base
            if l.b is not None:
                bias[sl] = l.b
            l.w = _Packed(self.rt, base[sl], None if lora is None else lora[sl])
        self.w, self.b = _Packed(self.rt, base, lora), bias
        return self

    def reset(self):
        self.key = None

    def result(self, x):
        if self.w is None:
            self.build()
        key = (x.data_ptr(), tuple(x.shape), x._version, self.rt.lora_enabled)
        if self.key != key:
            M = x.shape[0]
            if self.out is None or self.out.shape[0] != M:
                self.out = torch.empty(M, self.N, device=self.rt.device, dtype=torch.float32)
            ops.linear_small(x, self.w.get(), self.b, self.out, M=M, N=self.N, K=self.K)
            self.key, self.keep = key, x
            for fn in self.listeners:
                fn(self.out)
        return self.out


class Conv3x3:
    """3x3 pad-1 convolution as implicit GEMM (module/min_sdxl.py:246-260,601-618)."""

    def __init__(self, rt, src, name, stride=1, asym=False):
        """asym (stride 2 only): pad bottom/right only — diffusers Downsample2D(padding=0) of the VAE encoder"""
        self.rt, self.stride, self.asym = rt, stride, asym
        self.w = _load_w(rt, src, name, _conv_to_gemm)
        self.b = _bias(src, name)
        self.Cout = self.w.base.shape[0]
        self.Cin = self.w.base.shape[1] // 9

    def __call__(self, x: FMap, out_dtype=None, rowvec=None, residual=None, act=ops.ACT_NONE, up2=False,
                 gn_groups: int = 0) -> FMap:
        """gn_groups > 0 (opt-in path): the epilogue also accumulates the GroupNorm statistics of the output, returned as
        FMap.gn, when the launch qualifies (stride 1, tcgen05 path)"""
        rt = self.rt
        H, W = (2 * x.H, 2 * x.W) if up2 else (x.H, x.W)
        Ho, Wo = (H - 1) // self.stride + 1, (W - 1) // self.stride + 1
        M = x.n * Ho * Wo
        out = torch.empty(M, self.Cout, device=rt.device, dtype=out_dtype or rt.act_dtype)
        kw = dict(M=M, N=self.Cout, K=9 * self.Cin, bias=self.b, rowvec=rowvec, rows_per_sample=Ho * Wo,
                  residual=residual, act=act)
        if not rt.tc:
            ops.gemm(x.t, self.w.get(), out, conv=dict(n_img=x.n, H=H, W=W, Cin=self.Cin, stride=self.stride, up2=int(up2),
                                                       asym=int(self.asym)),
                     tc=False, **kw)
        else:
            src = x
            if up2:  # nearest-2x then conv (module/min_sdxl.py:617-618)
                up = rt.empty(x.n * H * W, x.C)
                ops.upsample2x(x.t, up, n_img=x.n, H=x.H, W=x.W, C=x.C)
                src = FMap(up, x.n, H, W, x.C)
            elif x.t.dtype != rt.act_dtype or self.stride != 1:
                pass
            if self.stride == 2:
                cols = rt.empty(M, 9 * self.Cin)
                ops.im2col3x3_s2(src.t, cols, n_img=x.n, H=H, W=W, C=self.Cin, asym=self.asym)
                gn = None
                if gn_groups and rt.gn_fuse and act == ops.ACT_NONE and ops.gn_eligible(
                        N=self.Cout, groups=gn_groups, rows_per_sample=Ho * Wo, residual=residual):
                    gn = rt.gn_site(x.n, gn_groups)
                ops.gemm(cols, self.w.get(), out, tc=True, gn=gn, **kw)
                return FMap(out, x.n, Ho, Wo, self.Cout, gn)
            else:
                a = src.t
                if a.dtype != rt.act_dtype:
                    a = rt.empty(src.M, src.C)
                    ops.cast2d(src.t, src.C, a, src.C, rows=src.M, cols=src.C)
                cv = dict(n_img=x.n, H=H, W=W, Cin=self.Cin)
                gn = None
                if gn_groups and rt.gn_fuse and act == ops.ACT_NONE and ops.gn_eligible(
                        N=self.Cout, groups=gn_groups, rows_per_sample=Ho * Wo, conv=cv, residual=residual):
                    gn = rt.gn_site(x.n, gn_groups)
                ops.gemm(a, self.w.get(), out, conv=cv, tc=True, gn=gn, **kw)
                return FMap(out, x.n, Ho, Wo, self.Cout, gn)
        return FMap(out, x.n, Ho, Wo, self.Cout)


class GroupNorm:
    def __init__(self, rt, src, name, C, groups, eps):
        self.rt, self.C, self.groups, self.eps = rt, C, groups, eps
        self.g, self.b = src.get(name + ".weight").contiguous(), src.get(name + ".bias").contiguous()

    def __call__(self, x: FMap, silu: bool) -> FMap:
        out = self.rt.empty(x.M, x.C)
        if x.gn is not None and tuple(x.gn.shape) == (x.n, self.groups, 2):
            # the producer of x accumulated the statistics in its epilogue: one pass over x (opt-in path)
            ops.groupnorm_apply_sums(x.t, self.g, self.b, x.gn, out, n_img=x.n, HW=x.H * x.W, C=x.C, groups=self.groups,
                                     eps=self.eps, silu=silu)
        else:
            ops.groupnorm(x.t, self.g, self.b, out, n_img=x.n, HW=x.H * x.W, C=x.C, groups=self.groups, eps=self.eps,
                          silu=silu, scratch_owner=id(self.rt))
        return FMap(out, x.n, x.H, x.W, x.C)


class LayerNorm:
    def __init__(self, rt, src, name, C, eps=1e-5):
        self.rt, self.C, self.eps = rt, C, eps
        self.g, self.b = src.get(name + ".weight").contiguous(), src.get(name + ".bias").contiguous()

    def __call__(self, x, rows, out_dtype=None):
        out = torch.empty(rows, self.C, device=self.rt.device, dtype=out_dtype or self.rt.act_dtype)
        ops.layernorm(x, self.g, self.b, out, rows=rows, C=self.C, eps=self.eps)
        return out


class ResnetBlock2D:
    """module/min_sdxl.py:242-283.  Kernel plan: GN+SiLU -> conv1 (+bias +temb row-vector in the
    epilogue) -> GN+SiLU -> conv2 (+bias +residual in the epilogue); the 1x1 shortcut is a GEMM
    whose fp32 output is that residual."""

    def __init__(self, rt, src, p, cfg, c_in, c_out):
        self.rt, self.c_in, self.c_out = rt, c_in, c_out
        self.norm1 = GroupNorm(rt, src, p + ".norm1", c_in, cfg.norm_num_groups, cfg.norm_eps)
        self.conv1 = Conv3x3(rt, src, p + ".conv1")
        self.time_emb_proj = SmallLinear(rt, src, p + ".time_emb_proj", bank=rt.temb_bank)
        self.norm2 = GroupNorm(rt, src, p + ".norm2", c_out, cfg.norm_num_groups, cfg.norm_eps)
        self.conv2 = Conv3x3(rt, src, p + ".conv2")
        self.conv_shortcut = None
        if c_in != c_out:
            self.conv_shortcut = Linear(rt, _ShortcutSrc(src), p + ".conv_shortcut")

    def __call__(self, x: FMap, temb_act: torch.Tensor) -> FMap:
        """x: fp32 stream (or act-dtype concat); temb_act = silu(emb) [n, T] fp32."""
        rt = self.rt
        G = self.norm2.groups
        h = self.conv1(self.norm1(x, silu=True), rowvec=self.time_emb_proj(temb_act), gn_groups=G)
        h = self.norm2(h, silu=True)
        if self.conv_shortcut is not None:
            a = x.t
            if a.dtype != rt.act_dtype:
                a = rt.empty(x.M, x.C)
                ops.cast2d(x.t, x.C, a, x.C, rows=x.M, cols=x.C)
            res = self.conv_shortcut(a, x.M, out_dtype=torch.float32)
        else:
            res = x.t
        # the block output is normalised next by a GroupNorm of the same group count (the next resnet's norm1, the
        # Transformer2DModel's norm, conv_norm_out) unless it goes into a concat first: offering the statistics costs a
        # few shuffles per epilogue chunk
        return self.conv2(h, out_dtype=torch.float32, residual=res, gn_groups=G)


class _ShortcutSrc:
    """view of a source that hands a 1x1 conv weight [Co,Ci,1,1] out as a linear weight [Co,Ci]."""

    def __init__(self, src):
        self.src = src

    def has(self, k):
        return self.src.has(k)

    def get(self, k):
        t = self.src.get(k)
        return t.reshape(t.shape[0], -1) if t.ndim == 4 else t

    def get_lora(self, m):
        lo = self.src.get_lora(m)
        if lo is None:
            return None
        a, b, s = lo
        return a.reshape(a.shape[0], -1), b.reshape(b.shape[0], -1), s


class Downsample2D:
    def __init__(self, rt, src, p):
        self.conv = Conv3x3(rt, src, p + ".conv", stride=2)

    def __call__(self, x: FMap, gn_groups: int = 0) -> FMap:
        return self.conv(x, out_dtype=torch.float32, gn_groups=gn_groups)


class Upsample2D:
    def __init__(self, rt, src, p):
        self.conv = Conv3x3(rt, src, p + ".conv")

    def __call__(self, x: FMap) -> FMap:
        return self.conv(x, out_dtype=torch.float32, up2=True)
