"""Torch (CPU) emulations of the C-ABI entry points the UNet / Aggregator / Resampler host code launches, written from the
CONTRACTS in include/instantir_b200.h and instantir_b200/ops.py — not from the kernels.  With them installed (and nn.Runtime
allowed on the CPU) the product's model classes run on a box without a GPU, so their host logic — weight packing (tap-major convs,
per-tile pair interleave, fused QKV, LoRA merge), the folded-LayerNorm producer / consumer chain, the time-embedding banks, the
step-invariant context caches, residual injection inside the concat, the opt-in GroupNorm-statistics plumbing — can be checked
against the oracle.  The kernels themselves are the GPU tests' job; tolerances here are those of fp32 / 16-bit STORAGE only.

Pointer + leading-dimension arguments are honoured through as_strided views of the tensors' storage, outputs are written in place
with the dtype of `out`, and fixed-point accumulators use the library's scales (csrc: LN 2^32 / 2^24, GN 2^24 / 2^26)."""
import math
import os

import torch
import torch.nn.functional as F

from instantir_b200 import nn, ops

LN_S1, LN_S2 = 2.0 ** 32, 2.0 ** 24
GN_S1, GN_S2 = 2.0 ** 24, 2.0 ** 26
ACT_NONE, ACT_SILU, ACT_GELU, ACT_QUICK_GELU = 0, 1, 2, 3
PAIR_NONE, PAIR_GEGLU, PAIR_SFT = 0, 1, 2


def _rows(t, nrows, ncols, ld, off=0):
    """[nrows, ncols] view of t's storage starting `off` elements after t's first element, row stride ld"""
    return t.as_strided((nrows, ncols), (ld, 1), t.storage_offset() + off)


def _act(x, act):
    if act == ACT_SILU:
        return F.silu(x)
    if act == ACT_GELU:
        return F.gelu(x)
    if act == ACT_QUICK_GELU:
        return x * torch.sigmoid(1.702 * x)
    return x


def _unpair(t, bn):
    """packed [.., N] with per-bn-tile [first half | second half] -> (first [.., N/2], second [.., N/2])"""
    half = bn // 2
    v = t.reshape(*t.shape[:-1], t.shape[-1] // bn, 2, half)
    return v[..., 0, :].reshape(*t.shape[:-1], -1), v[..., 1, :].reshape(*t.shape[:-1], -1)


class Emulation:
    def __init__(self):
        self.calls = []

    # ---------------------------------------------------------------------------------------------- GEMM / conv
    def gemm(self, a, w, out, *, M, N, K, lda=None, bias=None, rowvec=None, rows_per_sample=0, residual=None, ld_res=None,
             aux=None, ld_aux=None, ld_out=None, act=ACT_NONE, pair=PAIR_NONE, bn=None, conv=None, tc=True, cluster=None,
             ln_out=None, ln_in=None, gn=None):
        self.calls.append(("gemm", dict(M=M, N=N, K=K, conv=conv is not None, pair=pair, tc=tc, ln_out=ln_out is not None,
                                        ln_in=ln_in is not None, gn=gn is not None)))
        n_out = N // 2 if pair else N
        if tc:
            assert a.dtype in (torch.float16, torch.bfloat16) and w.dtype == a.dtype, "tcgen05 path: 16-bit operands"
            assert not (conv is not None and (conv.get("stride", 1) != 1 or conv.get("up2", 0))), "tc conv: 3x3 stride 1 only"
        if ln_in is not None or ln_out is not None or gn is not None:
            assert tc, "folded LayerNorm / GroupNorm statistics are tcgen05-path features"
        W = w.float().reshape(N, K)
        if conv is not None:
            n, H, Wd, Cin = conv["n_img"], conv["H"], conv["W"], conv["Cin"]
            stride, up2, asym = conv.get("stride", 1), conv.get("up2", 0), conv.get("asym", 0)
            assert K == 9 * Cin
            if up2:
                x = a.float().reshape(n, H // 2, Wd // 2, Cin).permute(0, 3, 1, 2)
                x = F.interpolate(x, scale_factor=2, mode="nearest")
            else:
                x = a.float().reshape(n, H, Wd, Cin).permute(0, 3, 1, 2)
            wk = W.reshape(N, 3, 3, Cin).permute(0, 3, 1, 2)
            if stride == 2 and asym:
                acc = F.conv2d(F.pad(x, (0, 1, 0, 1)), wk, None, stride=2)
            else:
                acc = F.conv2d(x, wk, None, stride=stride, padding=1)
            acc = acc.permute(0, 2, 3, 1).reshape(-1, N)
            assert acc.shape[0] == M, (acc.shape, M)
        else:
            A = _rows(a, M, K, lda if lda is not None else K).float()
            acc = A @ W.t()
        if ln_in is not None:
            st, colsum, eps = ln_in
            assert a is st.h16, "a LayerNorm-folded GEMM reads the stream's 16-bit copy"
            sums = st.acc[st.cur]
            mean = sums[:, 0].double() / LN_S1 / K
            var = (sums[:, 1].double() / LN_S2 / K - mean * mean).clamp_min(0)
            rstd = (1.0 / torch.sqrt(var + eps)).float()[:, None]
            acc = rstd * acc + (-mean.float()[:, None] * rstd) * colsum.float()[None, :]
            st.acc[st.cur ^ 1].zero_()
            st.cur ^= 1
        if bias is not None:
            acc = acc + bias.float()[None, :]
        if rowvec is not None:
            assert rows_per_sample > 0
            acc = acc + rowvec.float()[:, :N].repeat_interleave(rows_per_sample, 0)[:M]
        if pair == PAIR_NONE:
            val = _act(acc, act)
        else:
            assert bn and bn % 64 == 0 and N % bn == 0
            first, second = _unpair(acc, bn)
            if pair == PAIR_GEGLU:
                val = first * F.gelu(second)
            else:
                h = _rows(aux, M, n_out, ld_aux if ld_aux is not None else n_out).float()
                val = h * (first + 1.0) + second
        if residual is not None:
            val = val + _rows(residual, M, n_out, ld_res if ld_res is not None else n_out).float()
        if ln_out is not None:
            assert pair == PAIR_NONE and conv is None
            acc_t = ln_out.acc[ln_out.cur]
            acc_t[:, 0] += torch.round(val.double().sum(1) * LN_S1).long()
            acc_t[:, 1] += torch.round((val.double() ** 2).sum(1) * LN_S2).long()
            ln_out.h16.copy_(val.to(ln_out.h16.dtype))
        if gn is not None:
            assert int(gn.abs().sum()) == 0, "gn_sums must be zero on entry"
            n_s, groups = gn.shape[0], gn.shape[1]
            assert pair == PAIR_NONE and ln_out is None and rows_per_sample * n_s == M
            assert ops.gn_eligible(N=N, groups=groups, rows_per_sample=rows_per_sample, conv=conv, residual=residual)
            t = val.double().view(n_s, M // n_s, groups, N // groups)
            gn[..., 0] += torch.round(t.sum((1, 3)) * GN_S1).long()
            gn[..., 1] += torch.round((t * t).sum((1, 3)) * GN_S2).long()
        _rows(out, M, n_out, ld_out if ld_out is not None else n_out).copy_(val.to(out.dtype))
        return out

    def conv3x3_direct(self, x, w, bias, out, *, in_nchw, out_nchw, n_img, H, W, Cin, Cout, out_H=None, out_row_off=0):
        self.calls.append(("conv3x3_direct", {}))
        xi = x.float().reshape(n_img, Cin, H, W) if in_nchw else x.float().reshape(n_img, H, W, Cin).permute(0, 3, 1, 2)
        y = F.conv2d(xi, w.float().reshape(Cout, 3, 3, Cin).permute(0, 3, 1, 2), bias.float(), padding=1)
        if out_nchw:
            out.reshape(n_img, Cout, H, W).copy_(y.to(out.dtype))
        else:
            oh = out_H if out_H is not None else H
            out.reshape(n_img, oh, W, Cout)[:, out_row_off:out_row_off + H].copy_(y.permute(0, 2, 3, 1).to(out.dtype))
        return out

    def linear_small(self, x, w, bias, out, *, M, N, K, act=ACT_NONE):
        self.calls.append(("linear_small", dict(N=N)))
        y = x.float().reshape(M, K) @ w.float().reshape(N, K).t()
        if bias is not None:
            y = y + bias.float()
        out.copy_(_act(y, act).to(out.dtype))
        return out

    # ---------------------------------------------------------------------------------------------- attention
    def attention(self, q, q_off, ldq, ks, k_offs, ldks, vs, v_offs, ldvs, kv_lens, seg_scales, out, out_off, ldo, *, B, heads,
                  n_q, softmax_scale, tc=True, scratch_owner=None, causal=False):
        self.calls.append(("attention", dict(n_q=n_q, kv=list(kv_lens))))
        C = heads * 64
        Q = _rows(q, B * n_q, C, ldq, q_off).float().view(B, n_q, heads, 64).transpose(1, 2)
        acc = 0.0
        for s in range(len(ks)):
            n_k = kv_lens[s]
            Kt = _rows(ks[s], B * n_k, C, ldks[s], k_offs[s]).float().view(B, n_k, heads, 64).transpose(1, 2)
            Vt = _rows(vs[s], B * n_k, C, ldvs[s], v_offs[s]).float().view(B, n_k, heads, 64).transpose(1, 2)
            sc = Q @ Kt.transpose(-1, -2) * softmax_scale
            if causal:
                assert n_q == n_k and len(ks) == 1
                sc = sc.masked_fill(torch.ones(n_q, n_k, dtype=torch.bool).triu(1), float("-inf"))
            acc = acc + seg_scales[s] * (torch.softmax(sc, -1) @ Vt)
        _rows(out, B * n_q, C, ldo, out_off).copy_(acc.transpose(1, 2).reshape(B * n_q, C).to(out.dtype))
        return out

    # ---------------------------------------------------------------------------------------------- norms
    def groupnorm(self, x, gamma, beta, out, *, n_img, HW, C, groups=32, eps=1e-5, silu=False, scratch_owner=None):
        self.calls.append(("groupnorm", {}))
        y = F.group_norm(x.float().reshape(n_img, HW, C).permute(0, 2, 1), groups, gamma, beta, eps).permute(0, 2, 1).reshape(n_img * HW, C)
        out.reshape(n_img * HW, C).copy_((F.silu(y) if silu else y).to(out.dtype))
        return out

    def groupnorm_apply_sums(self, x, gamma, beta, sums, out, *, n_img, HW, C, groups=32, eps=1e-5, silu=False):
        self.calls.append(("groupnorm_apply_sums", {}))
        assert tuple(sums.shape) == (n_img, groups, 2) and sums.dtype == torch.int64
        count = HW * (C // groups)
        mean = sums[..., 0].double() / GN_S1 / count
        var = (sums[..., 1].double() / GN_S2 / count - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        xs = x.double().reshape(n_img, HW, groups, C // groups)
        y = ((xs - mean[:, None, :, None]) * rstd[:, None, :, None]).reshape(n_img * HW, C) * gamma.double() + beta.double()
        y = y.float()
        out.reshape(n_img * HW, C).copy_((F.silu(y) if silu else y).to(out.dtype))
        return out

    def memset_zero(self, t):
        self.calls.append(("memset_zero", {}))
        return t.zero_()

    def layernorm(self, x, gamma, beta, out, *, rows, C, eps=1e-5, mod=None, rows_per_sample=0):
        self.calls.append(("layernorm", {}))
        y = F.layer_norm(x.float().reshape(rows, C), (C,), gamma, beta, eps)
        if mod is not None:
            m = mod.float().repeat_interleave(rows_per_sample, 0)[:rows]
            y = y * (1.0 + m[:, C:2 * C]) + m[:, :C]
        out.reshape(rows, C).copy_(y.to(out.dtype))
        return out

    def adaln_items(self, entries, device):
        return list(entries)

    def adaln_batched(self, table, n_items, mod, out_dtype, *, rows, rows_per_sample, eps=1e-6, bytes_moved=0.0):
        self.calls.append(("adaln_batched", dict(n=n_items)))
        assert len(table) == n_items
        for x, out, off, C in table:
            assert out.dtype == out_dtype
            y = F.layer_norm(x.float().reshape(rows, C), (C,), None, None, eps)
            m = mod.float().repeat_interleave(rows_per_sample, 0)[:rows]
            out.reshape(rows, C).copy_((y * (1.0 + m[:, off + C:off + 2 * C]) + m[:, off:off + C]).to(out.dtype))

    # ---------------------------------------------------------------------------------------------- data movement
    def concat_inject(self, h, C1, skip, C2, out, *, M, rh=None, rs=None, cond_scale=None, rows_per_sample=0):
        self.calls.append(("concat_inject", {}))
        s = 1.0 if cond_scale is None else cond_scale.float().repeat_interleave(rows_per_sample)[:M, None]
        a = h.float().reshape(M, C1)
        if rh is not None:
            a = a + s * rh.float().reshape(M, C1)
        parts = [a]
        if C2:
            b = skip.float().reshape(M, C2)
            if rs is not None:
                b = b + s * rs.float().reshape(M, C2)
            parts.append(b)
        out.reshape(M, C1 + C2).copy_(torch.cat(parts, 1).to(out.dtype))
        return out

    def upsample2x(self, x, out, *, n_img, H, W, C):
        self.calls.append(("upsample2x", {}))
        y = x.float().reshape(n_img, H, W, C).repeat_interleave(2, 1).repeat_interleave(2, 2)
        out.reshape(n_img, 2 * H, 2 * W, C).copy_(y.to(out.dtype))
        return out

    def im2col3x3_s2(self, x, out, *, n_img, H, W, C, asym=False):
        self.calls.append(("im2col3x3_s2", {}))
        xi = x.float().reshape(n_img, H, W, C).permute(0, 3, 1, 2)
        xp = F.pad(xi, (0, 1, 0, 1)) if asym else F.pad(xi, (1, 1, 1, 1))
        cols = F.unfold(xp, 3, stride=2)  # [n, C*9, L], channel-major
        L = cols.shape[-1]
        out.reshape(n_img * L, 9 * C).copy_(cols.view(n_img, C, 9, L).permute(0, 3, 2, 1).reshape(n_img * L, 9 * C).to(out.dtype))
        return out

    def cast2d(self, x, ld_in, out, ld_out, *, rows, cols):
        self.calls.append(("cast2d", {}))
        _rows(out, rows, cols, ld_out).copy_(_rows(x, rows, cols, ld_in).to(out.dtype))
        return out

    def silu(self, x, out):
        out.copy_(F.silu(x.float()).to(out.dtype))
        return out

    def add(self, a, b, out):
        out.copy_((a.float() + b.float()).to(out.dtype))
        return out

    def scale(self, x, out, alpha):
        out.copy_((x.float() * alpha).to(out.dtype))
        return out

    def timestep_embedding(self, t, dim, out):
        half = dim // 2
        freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
        arg = t.float().reshape(-1, 1) * freq[None, :]
        out.copy_(torch.cat([torch.cos(arg), torch.sin(arg)], 1).to(out.dtype))
        return out

    # ---------------------------------------------------------------------------------------------- rows f1 / f2
    def softmax_rows(self, x, out, *, scale):
        self.calls.append(("softmax_rows", {}))
        out.copy_(torch.softmax(x.float() * scale, -1).to(out.dtype))
        return out

    def gaussian_sample(self, moments, noise, out, *, scale=1.0):
        self.calls.append(("gaussian_sample", {}))
        mean, logvar = moments.float().chunk(2, dim=1)
        std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
        out.copy_((mean + (std * noise if noise is not None else 0.0)) * scale)
        return out

    def embed_tokens(self, ids, token_embedding, position_embedding, out, *, seq_len):
        self.calls.append(("embed_tokens", {}))
        r = torch.arange(ids.numel())
        out.reshape(ids.numel(), -1).copy_(token_embedding[ids.reshape(-1)] + position_embedding[r % seq_len])
        return out

    def patchify(self, img, out, *, patch):
        self.calls.append(("patchify", {}))
        n, c, h, w = img.shape
        cols = F.unfold(img.float(), patch, stride=patch)          # [n, c*p*p, gh*gw], (c, ky, kx) order
        rows = cols.transpose(1, 2).reshape(-1, c * patch * patch)
        out.zero_()
        out[:, :rows.shape[1]].copy_(rows.to(out.dtype))
        return out

    def vit_assemble(self, patches, cls, pos, out, *, n_img, P):
        self.calls.append(("vit_assemble", {}))
        D = patches.shape[-1]
        o = out.reshape(n_img, P + 1, D)
        o[:, 0] = cls.reshape(1, D) + pos.reshape(P + 1, D)[0]
        o[:, 1:] = patches.reshape(n_img, P, D) + pos.reshape(P + 1, D)[1:]
        return out

    NAMES = ("softmax_rows", "gaussian_sample", "embed_tokens", "patchify", "vit_assemble", "gemm", "conv3x3_direct", "linear_small", "attention", "groupnorm", "groupnorm_apply_sums", "memset_zero", "layernorm",
             "adaln_items", "adaln_batched", "concat_inject", "upsample2x", "im2col3x3_s2", "cast2d", "silu", "add", "scale",
             "timestep_embedding")


def install(setattr_fn, fuse_gn=False):
    """replace the ops above and let nn.Runtime live on the CPU (its constructor refuses non-CUDA devices: the product has no CPU
    path — this is a test harness)"""
    emu = Emulation()
    for name in Emulation.NAMES:
        setattr_fn(ops, name, getattr(emu, name))
    orig = nn.Runtime.__init__

    def init(self, device, precision="fp16"):
        orig(self, "cuda", precision)
        self.device = torch.device("cpu")
        self.gn_fuse = bool(fuse_gn) and self.tc

    setattr_fn(nn.Runtime, "__init__", init)
    from _cpu_loop import NullStream

    setattr_fn(torch.cuda, "Stream", NullStream)
    os.environ.pop("IIR_GN_FUSE", None)
    return emu
