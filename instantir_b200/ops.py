"""Thin torch-tensor wrappers over the C ABI (one Python function per entry point).

PyTorch is used for device memory and streams only; every function here launches hand-written
sm_100a kernels from ``libinstantir_b200.so`` on torch's current CUDA stream and raises on any
failure.  Nothing in this module computes with torch ops.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_SILU, BF16, F16, F32, PAIR_GEGLU, PAIR_NONE, PAIR_SFT  # noqa: F401


# ---- optional per-launch profiling (bench.py's roofline leg): when PROFILE is a list every op
# appends (kernel name, work dict, start event, end event); events are recorded on the launching
# stream around the C-ABI call.
PROFILE = None


class _Prof:
    def __init__(self, name, **work):
        self.name, self.work = name, work

    def __enter__(self):
        if PROFILE is not None:
            # under CUDA-graph capture the events become event-record NODES of the graph (external=True):
            # after a replay they hold the in-graph start/end of this launch — no host launch gap inside
            ext = torch.cuda.is_current_stream_capturing()
            self.work["in_graph"] = ext
            self.e0 = torch.cuda.Event(enable_timing=True, external=ext)
            self.e1 = torch.cuda.Event(enable_timing=True, external=ext)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and exc[0] is None:
            self.e1.record()
            PROFILE.append((self.name, self.work, self.e0, self.e1))
        return False


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    raise TypeError(f"unsupported dtype {t.dtype} (fp32, bf16 or fp16 only)")


def _L(*tensors):
    """the library build matching the 16-bit dtype of the call's tensors (the default build for pure-fp32 calls)"""
    for t in tensors:
        if t is not None and t.dtype == torch.float16:
            return _lib.load(h16=F16)
        if t is not None and t.dtype == torch.bfloat16:
            return _lib.load(h16=BF16)
    return _lib.load()


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.IIRError("instantir_b200 kernels need CUDA tensors (there is no CPU fallback)")
    return t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: Optional[torch.Tensor], name: str):
    if t is not None and (t.dtype != torch.float32 or not t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous fp32 tensor")
    return t


def _f32rows(t: Optional[torch.Tensor], name: str):
    """fp32 [rows, n] tensor whose rows are contiguous (a column slice of a wider buffer is fine);
    returns (tensor, row stride in elements; 0 when None)."""
    if t is None:
        return None, 0
    if t.dtype != torch.float32 or t.ndim != 2 or t.stride(1) != 1:
        raise TypeError(f"{name} must be an fp32 [rows, n] tensor with unit column stride")
    return t, t.stride(0)


_N_SM = {}


def n_sm() -> int:
    """SM count of the current CUDA device (148 on a B200; also the answer on a GPU-less build box)"""
    if not torch.cuda.is_available():
        return 148
    d = torch.cuda.current_device()
    if d not in _N_SM:
        _N_SM[d] = torch.cuda.get_device_properties(d).multi_processor_count
    return _N_SM[d]


def default_bn(n: int, pair: bool = False) -> int:
    """pack-time N-tile width for paired (GEGLU/SFT) weights: the largest multiple of 64 <= 256 that
    divides N (their row interleave is fixed when the weights are packed)."""
    step = 64 if pair else 32
    for bn in range(256, step - 1, -step):
        if n % bn == 0:
            return bn
    return 128


def choose_bn(M: int, N: int, K: int) -> int:
    """Wave-aware N-tile width for the persistent tcgen05 GEMM (one CTA per SM, static round-robin
    over 128 x BN tiles).  Cost model per tile: K/16 UMMA steps of max(BN/2, 32 + BN/4) cycles (tensor
    pipe vs. shared-memory operand reads at 128 B/clk) + pipeline fill/drain and the exposed epilogue;
    total = waves x tile cost.  A narrower tile often wins when 128 x 256 tiles would leave SMs idle
    (e.g. M=2048, N=1280: 80 tiles on 148 SMs -> BN=160 gives 128 tiles of 0.625x the length)."""
    tm = (M + 127) // 128
    best, best_cost = 256, None
    for bn in range(32, 257, 32):
        tn = (N + bn - 1) // bn
        waves = (tm * tn + n_sm() - 1) // n_sm()
        # per 16-deep UMMA step: tensor pipe, smem operand reads (128 B/clk), L2->SM fill (~52 B/clk/SM;
        # the weight tile is split over a 2-CTA cluster and multicast, so each CTA pulls A + B/2)
        cyc = max(bn / 2.0, 32.0 + bn / 4.0, (4096.0 + bn * 16.0) / 52.0)
        cost = waves * ((K / 16.0) * cyc + 400.0 + 3.0 * bn)
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = bn, cost
    return best


# ---- measured tile choices ------------------------------------------------------------------
# tools/autotune.py sweeps (tile width x {one CTA, CTA pair}) for every GEMM / conv shape a step
# launches on a B200 and writes tuning_b200.json; unseen shapes fall back to the analytic model.
_TUNING = None


def _tuning():
    global _TUNING
    if _TUNING is None:
        import json
        import os

        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tuning_b200.json")
        _TUNING = {}
        if os.path.exists(path) and os.environ.get("IIR_NO_TUNING") != "1":
            with open(path) as f:
                _TUNING = json.load(f).get("gemm", {})
    return _TUNING


def gemm_key(M, N, K, conv, pair, epi=0):
    """tuning-table key; epi = 0: 16-bit output, 1: fp32 output, 2: fp32 output + residual (the exposed
    epilogue of a one-wave launch differs enough between them to change the best tile)"""
    g = f"c{conv['n_img']}x{conv['H']}x{conv['W']}" if conv is not None else "lin"
    return f"{g}:{M}:{N}:{K}:{int(pair)}:{int(epi)}"


def gemm(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int,
         lda: Optional[int] = None, bias=None, rowvec=None, rows_per_sample: int = 0,
         residual=None, ld_res: Optional[int] = None, aux=None, ld_aux: Optional[int] = None,
         ld_out: Optional[int] = None, act: int = ACT_NONE, pair: int = PAIR_NONE,
         bn: Optional[int] = None, conv: Optional[dict] = None, tc: bool = True,
         cluster: Optional[int] = None, ln_out=None, ln_in=None, gn=None):
    """out = epilogue(A · Wᵀ).  conv = dict(n_img, H, W, Cin, stride=1, up2=0) for 3x3 pad-1.

    Folded LayerNorm (tc only, include/instantir_b200.h): ``ln_out`` = object with ``.acc`` ([2, M, 2] int64: two
    alternating row-sum accumulators, zero-initialised), ``.cur`` (index of the live one) and ``.h16`` ([M, N]
    16-bit): this GEMM adds the sums of its rows into ``acc[cur]`` and writes ``h16`` next to ``out``.  ``ln_in`` =
    (such an object, colsum [N] fp32, eps): A must be its ``.h16``; clears ``acc[cur ^ 1]`` and flips ``cur``.

    GroupNorm statistics from the epilogue (tc only, opt-in): ``gn`` = zeroed int64 tensor [n_samples, groups, 2]; the launch
    adds the fixed-point (sum, sum of squares) of its final output values per (sample, group) into it
    (iir_gemm_args.gn_sums; consumer: ``groupnorm_apply_sums``).  ``gn_eligible`` tells whether a launch qualifies."""
    lib = _L(a, w, out)
    n_out = N // 2 if pair else N
    g = _lib.GemmArgs()
    g.a, g.w = _p(a), _p(w)
    g.a_dtype, g.w_dtype = _dt(a), _dt(w)
    g.M, g.N, g.K = M, N, K
    g.lda = lda if lda is not None else K
    if conv is not None:
        g.conv = 3
        g.n_img, g.H, g.W, g.Cin = conv["n_img"], conv["H"], conv["W"], conv["Cin"]
        g.stride = conv.get("stride", 1)
        g.up2 = conv.get("up2", 0)
        g.conv_asym = int(conv.get("asym", 0))
    g.bias = _p(_f32c(bias, "bias"))
    rowvec, g.ld_rowvec = _f32rows(rowvec, "rowvec")
    g.rowvec = _p(rowvec)
    g.rows_per_sample = rows_per_sample
    if residual is not None:
        g.residual, g.res_dtype = _p(residual), _dt(residual)
        g.ld_res = ld_res if ld_res is not None else n_out
    if aux is not None:
        g.aux, g.aux_dtype = _p(aux), _dt(aux)
        g.ld_aux = ld_aux if ld_aux is not None else n_out
    g.out, g.out_dtype = _p(out), _dt(out)
    g.ld_out = ld_out if ld_out is not None else n_out
    g.act, g.pair = act, pair
    epi = (2 if residual is not None else 1) if out.dtype == torch.float32 else 0
    key = gemm_key(M, N, K, conv, pair, epi)
    tuned = _tuning().get(key) if (tc and (bn is None or cluster is None)) else None
    if bn is None:
        bn = default_bn(N, True) if pair else (tuned["bn"] if tuned else choose_bn(M, N, K))
    if cluster is None:
        cluster = tuned["cluster"] if tuned else 0
    g.bn, g.cluster = bn, cluster
    if ln_out is not None:
        g.ln_stats_out, g.ln_out16, g.ld_ln_out16 = _p(ln_out.acc[ln_out.cur]), _p(ln_out.h16), ln_out.h16.stride(0)
    if ln_in is not None:
        st, colsum, eps = ln_in
        g.ln_stats_in, g.ln_stats_zero = _p(st.acc[st.cur]), _p(st.acc[st.cur ^ 1])
        g.ln_colsum, g.ln_eps = _p(_f32c(colsum, "ln colsum")), eps
        st.cur ^= 1  # the next producer adds into the accumulator this launch clears
    if gn is not None:
        if gn.dtype != torch.int64 or not gn.is_contiguous() or gn.ndim != 3 or gn.shape[2] != 2 or N % gn.shape[1]:
            raise TypeError("gn must be a contiguous int64 [n_samples, groups, 2] tensor with N % groups == 0")
        g.gn_sums, g.gn_groups, g.gn_cpg = _p(gn), gn.shape[1], N // gn.shape[1]
    fn = lib.iir_gemm_tc if tc else lib.iir_gemm_simt
    name = ("conv3x3_" if conv is not None else "gemm_") + ("tc" if tc else "simt")
    with _Prof(name, flops=2.0 * M * N * K, M=M, N=N, K=K, pair=int(pair), key=key, epi=epi,
               conv=None if conv is None else (conv["n_img"], conv["H"], conv["W"], conv["Cin"])):
        _lib.check(fn(C.byref(g), _stream()), "iir_gemm_tc" if tc else "iir_gemm_simt", lib)
    return out


def gn_eligible(*, N: int, groups: int, rows_per_sample: int, conv: Optional[dict] = None, residual=None, pair: int = PAIR_NONE) -> bool:
    """can this tcgen05 launch accumulate GroupNorm statistics in its epilogue?  Mirrors the checks of iir_gemm_tc
    (plain epilogue on the direct-store path, even channels per group, the 32 rows of a warp inside one sample)."""
    if pair != PAIR_NONE or groups <= 0 or N % groups or (N // groups) % 2:
        return False
    if residual is not None and residual.dtype != torch.float32:
        return False
    if os.environ.get("IIR_GEMM_DIRECT", "2") != "2" or os.environ.get("IIR_GEMM_CLUSTER", "0") not in ("0", "1", "22"):
        return False
    if conv is not None:
        return conv.get("stride", 1) == 1 and conv["W"] % 8 == 0 and conv["H"] >= 4
    return rows_per_sample > 0 and rows_per_sample % 32 == 0


def groupnorm_apply_sums(x, gamma, beta, sums, out, *, n_img: int, HW: int, C: int, groups: int = 32, eps: float = 1e-5,
                         silu: bool = False):
    """GroupNorm's second half alone: (mean, rstd) come from the fixed-point sums the PRODUCER of x accumulated (gemm(gn=...))"""
    lib = _L(x, out)
    if sums.dtype != torch.int64 or not sums.is_contiguous() or tuple(sums.shape) != (n_img, groups, 2):
        raise TypeError("sums must be the contiguous int64 [n_img, groups, 2] tensor the producing GEMM accumulated into")
    with _Prof("groupnorm_apply", bytes=float(n_img) * HW * C * (x.element_size() + out.element_size())):
        _lib.check(lib.iir_groupnorm_apply_sums(_p(x), _dt(x), _p(_f32c(gamma, "gamma")), _p(_f32c(beta, "beta")), _p(sums),
                                                _p(out), _dt(out), n_img, HW, C, groups, eps, int(silu), _stream()),
                   "iir_groupnorm_apply_sums", lib)
    return out


def memset_zero(t: torch.Tensor):
    """cudaMemsetAsync on the current stream (a memset node inside a captured forward)"""
    lib = _L(t)
    if not t.is_contiguous():
        raise TypeError("memset_zero needs a contiguous tensor")
    _lib.check(lib.iir_memset_zero(_p(t), t.numel() * t.element_size(), _stream()), "iir_memset_zero", lib)
    return t


def conv3x3_direct(x, w, bias, out, *, in_nchw: bool, out_nchw: bool, n_img: int, H: int, W: int,
                   Cin: int, Cout: int, out_H: Optional[int] = None, out_row_off: int = 0):
    lib = _L(x, out)
    _f32c(w, "w")
    _f32c(bias, "bias")
    with _Prof("conv3x3_direct", bytes=float(n_img) * H * W * (Cin * x.element_size() + Cout * out.element_size())):
        _lib.check(lib.iir_conv3x3_direct(_p(x), _dt(x), int(in_nchw), _p(w), _p(bias), _p(out), _dt(out),
                                          int(out_nchw), n_img, H, W, Cin, Cout,
                                          out_H if out_H is not None else H, out_row_off, _stream()),
                   "iir_conv3x3_direct", lib)
    return out


_ATTN_WS = {}


def _attn_workspace(lib, device, B: int, heads: int, n_q: int, owner=None) -> torch.Tensor:
    """Zero-initialised scratch of the self-attention kernel (partial tiles + tickets; every launch leaves the tickets
    zero, launches are stream-ordered, so one buffer per (device, owner) serves every attention of a model).  `owner`
    separates models that may run CONCURRENTLY on different streams (UNet down path || Aggregator)."""
    need = int(lib.iir_attn_workspace_bytes(B, heads, n_q))
    key = (torch.device(device).index, owner)
    bufs = _ATTN_WS.setdefault(key, [])
    if not bufs or bufs[-1].numel() < need:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.IIRError("attention workspace must be created before CUDA-graph capture: run the op once eagerly first")
        # outgrown buffers stay alive: CUDA graphs captured earlier hold their addresses
        bufs.append(torch.zeros(max(need, 32 << 20), dtype=torch.uint8, device=device))
    return bufs[-1]


def attention(q, q_off: int, ldq: int, ks: Sequence[torch.Tensor], k_offs: Sequence[int],
              ldks: Sequence[int], vs: Sequence[torch.Tensor], v_offs: Sequence[int],
              ldvs: Sequence[int], kv_lens: Sequence[int], seg_scales: Sequence[float], out,
              out_off: int, ldo: int, *, B: int, heads: int, n_q: int, softmax_scale: float,
              tc: bool = True, scratch_owner=None, causal: bool = False):
    lib = _L(q, out)
    a = _lib.AttnArgs()
    a.causal = int(causal)
    if tc and len(ks) == 1:
        ws = _attn_workspace(lib, q.device, B, heads, n_q, scratch_owner)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.q, a.ldq, a.q_off = _p(q), ldq, q_off
    a.n_seg = len(ks)
    for s in range(len(ks)):
        a.k[s], a.ldk[s], a.k_off[s] = _p(ks[s]), ldks[s], k_offs[s]
        a.v[s], a.ldv[s], a.v_off[s] = _p(vs[s]), ldvs[s], v_offs[s]
        a.kv_len[s] = kv_lens[s]
        a.seg_scale[s] = seg_scales[s]
    a.out, a.ldo, a.out_off = _p(out), ldo, out_off
    a.dtype = _dt(q)
    a.B, a.heads, a.n_q = B, heads, n_q
    a.softmax_scale = softmax_scale
    fn = lib.iir_attn_tc if tc else lib.iir_attn_simt
    with _Prof("attn_tc" if tc else "attn_simt", flops=4.0 * B * heads * n_q * sum(kv_lens) * 64, n_q=n_q, n_kv=sum(kv_lens)):
        _lib.check(fn(C.byref(a), _stream()), "iir_attn_tc" if tc else "iir_attn_simt", lib)
    return out


_GN_SCRATCH = {}


def _gn_partials(device, n_img: int, groups: int, owner=None) -> torch.Tensor:
    """GroupNorm scratch (chunk partials, stats, per-image ticket counters).  The ticket words have to be
    zero before the first call and every call leaves them zero, so one zero-initialised buffer per
    (device, n_img, groups, owner) is kept for the life of the process; launches are stream-ordered, so reuse
    by consecutive GroupNorms on a stream is safe.  `owner` separates models that may run CONCURRENTLY on
    different streams (the pipeline overlaps the UNet's down path with the Aggregator)."""
    key = (torch.device(device).index, n_img, groups, owner)
    buf = _GN_SCRATCH.get(key)
    if buf is None:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.IIRError("GroupNorm scratch must be created before CUDA-graph capture: run the op once eagerly first")
        need = int(_lib.load().iir_groupnorm_scratch_floats(n_img, groups))
        buf = _GN_SCRATCH[key] = torch.zeros(need, dtype=torch.float32, device=device)
    return buf


def groupnorm(x, gamma, beta, out, *, n_img: int, HW: int, C: int, groups: int = 32,
              eps: float = 1e-5, silu: bool = False, scratch_owner=None):
    lib = _L(x, out)
    part = _gn_partials(x.device, n_img, groups, scratch_owner)
    with _Prof("groupnorm", bytes=float(n_img) * HW * C * (2 * x.element_size() + out.element_size())):
        _lib.check(lib.iir_groupnorm(_p(x), _dt(x), _p(_f32c(gamma, "gamma")), _p(_f32c(beta, "beta")),
                                     _p(out), _dt(out), n_img, HW, C, groups, eps, int(silu), _p(part),
                                     _stream()), "iir_groupnorm", lib)
    return out


def layernorm(x, gamma, beta, out, *, rows: int, C: int, eps: float = 1e-5, mod=None,
              rows_per_sample: int = 0):
    lib = _L(x, out)
    mod, mod_ld = _f32rows(mod, "mod")
    with _Prof("layernorm", bytes=float(rows) * C * (x.element_size() + out.element_size())):
        _lib.check(lib.iir_layernorm(_p(x), _dt(x), _p(_f32c(gamma, "gamma")), _p(_f32c(beta, "beta")),
                                     _p(mod), mod_ld, rows_per_sample, _p(out), _dt(out), rows, C,
                                     eps, _stream()), "iir_layernorm", lib)
    return out


def adaln_items(entries, device) -> torch.Tensor:
    """device table for adaln_batched: entries = [(x fp32 [rows,C], out [rows,C], mod column offset, C)]."""
    arr = (_lib.AdaLNItem * len(entries))()
    for k, (x, out, off, c) in enumerate(entries):
        if x.dtype != torch.float32 or not x.is_contiguous() or not out.is_contiguous():
            raise TypeError("adaln_items: x must be contiguous fp32, out contiguous")
        arr[k].x, arr[k].out, arr[k].mod_off, arr[k].C = _p(x), _p(out), off, c
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return raw.to(device)


def adaln_batched(table: torch.Tensor, n_items: int, mod, out_dtype, *, rows: int, rows_per_sample: int,
                  eps: float = 1e-6, bytes_moved: float = 0.0):
    """every item's LN(x)*(1+scale)+shift in one launch (table from adaln_items)."""
    lib = _lib.load(h16=F16) if out_dtype == torch.float16 else _lib.load(h16=BF16) if out_dtype == torch.bfloat16 else _lib.load()
    mod, mod_ld = _f32rows(mod, "mod")
    dt = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}[out_dtype]
    with _Prof("adaln_batched", bytes=bytes_moved):
        _lib.check(lib.iir_adaln_batched(_p(table), n_items, rows, rows_per_sample, _p(mod), mod_ld, eps, dt,
                                         _stream()), "iir_adaln_batched", lib)


def concat_inject(h, C1: int, skip, C2: int, out, *, M: int, rh=None, rs=None, cond_scale=None,
                  rows_per_sample: int = 0):
    lib = _L(h, skip, rh, rs, out)
    with _Prof("concat_inject", bytes=float(M) * (C1 + C2) * (h.element_size() + out.element_size())):
        _lib.check(lib.iir_concat_inject(_p(h), _dt(h), C1, _p(rh), _dt(rh) if rh is not None else 0,
                                         _p(skip), _dt(skip) if skip is not None else 0, C2, _p(rs),
                                         _dt(rs) if rs is not None else 0, _p(_f32c(cond_scale, "cond_scale")),
                                         rows_per_sample, _p(out), _dt(out), M, _stream()),
                   "iir_concat_inject", lib)
    return out


def upsample2x(x, out, *, n_img: int, H: int, W: int, C: int):
    lib = _L(x, out)
    with _Prof("upsample2x", bytes=float(n_img) * H * W * C * (x.element_size() + 4 * out.element_size())):
        _lib.check(lib.iir_upsample2x(_p(x), _dt(x), _p(out), _dt(out), n_img, H, W, C, _stream()),
                   "iir_upsample2x", lib)
    return out


def im2col3x3_s2(x, out, *, n_img: int, H: int, W: int, C: int, asym: bool = False):
    """asym: pad bottom/right only (the VAE encoder's Downsample2D(padding=0)) instead of pad 1 on every side"""
    lib = _L(x, out)
    with _Prof("im2col3x3_s2", bytes=float(n_img) * H * W * C * (x.element_size() + 2.25 * out.element_size())):
        _lib.check(lib.iir_im2col3x3_s2(_p(x), _dt(x), _p(out), _dt(out), n_img, H, W, C, int(asym), _stream()),
                   "iir_im2col3x3_s2", lib)
    return out


def cast2d(x, ld_in: int, out, ld_out: int, *, rows: int, cols: int):
    lib = _L(x, out)
    with _Prof("cast2d", bytes=float(rows) * cols * (x.element_size() + out.element_size())):
        _lib.check(lib.iir_cast2d(_p(x), _dt(x), ld_in, _p(out), _dt(out), ld_out, rows, cols, _stream()),
                   "iir_cast2d", lib)
    return out


def silu(x, out):
    lib = _L(x, out)
    _lib.check(lib.iir_silu(_p(x), _dt(x), _p(out), _dt(out), x.numel(), _stream()), "iir_silu", lib)
    return out


def add(a, b, out):
    lib = _L(a, b, out)
    _lib.check(lib.iir_add(_p(a), _dt(a), _p(b), _dt(b), _p(out), _dt(out), out.numel(), _stream()),
               "iir_add", lib)
    return out


def scale(x, out, alpha: float):
    """out = alpha * x elementwise"""
    lib = _L(x, out)
    _lib.check(lib.iir_scale(_p(x), _dt(x), _p(out), _dt(out), out.numel(), float(alpha), _stream()), "iir_scale", lib)
    return out


def step_prologue(latents, x_in, n_rep: int, *, t: float, t_dev, cond_scale: float, cond_scale_dev):
    """x_in = cat([latents] * n_rep); t_dev[0] = t; cond_scale_dev[:] = cond_scale — one launch at the top of a step"""
    lib = _L()
    _f32c(latents, "latents"), _f32c(x_in, "x_in"), _f32c(t_dev, "t_dev"), _f32c(cond_scale_dev, "cond_scale_dev")
    if x_in.numel() != n_rep * latents.numel():
        raise ValueError("step_prologue: x_in must hold n_rep copies of latents")
    _lib.check(lib.iir_step_prologue(_p(latents), latents.numel(), n_rep, _p(x_in), float(t), _p(t_dev), float(cond_scale),
                                     _p(cond_scale_dev), 0 if cond_scale_dev is None else cond_scale_dev.numel(), _stream()),
               "iir_step_prologue", lib)
    return x_in


def embed_tokens(ids, token_embedding, position_embedding, out, *, seq_len: int):
    """out[r] = token_embedding[ids[r]] + position_embedding[r % seq_len] (fp32 tables, fp32 out [n_tokens, dim])"""
    lib = _L()
    if ids.dtype != torch.int64 or not ids.is_contiguous():
        raise TypeError("embed_tokens: ids must be a contiguous int64 tensor")
    _f32c(token_embedding, "token_embedding"), _f32c(position_embedding, "position_embedding"), _f32c(out, "out")
    vocab, dim = token_embedding.shape
    _lib.check(lib.iir_embed_tokens(_p(ids), ids.numel(), seq_len, _p(token_embedding), vocab, _p(position_embedding), dim, _p(out),
                                    _stream()), "iir_embed_tokens", lib)
    return out


def patchify(img, out, *, patch: int):
    """NCHW fp32 image -> [n_img * gh * gw, ld] rows of patch x patch windows in (c, ky, kx) order, zero-padded columns"""
    lib = _L(out)
    _f32c(img, "img")
    n, c, h, w = img.shape
    _lib.check(lib.iir_patchify(_p(img), n, c, h, w, patch, _p(out), _dt(out), out.stride(0), _stream()), "iir_patchify", lib)
    return out


def vit_assemble(patches, cls, pos, out, *, n_img: int, P: int):
    """out[b, 0] = cls + pos[0]; out[b, 1 + p] = patches[b * P + p] + pos[1 + p]"""
    lib = _L()
    for t, n in ((patches, "patches"), (cls, "cls"), (pos, "pos"), (out, "out")):
        _f32c(t, n)
    _lib.check(lib.iir_vit_assemble(_p(patches), _p(cls), _p(pos), _p(out), n_img, P, patches.shape[-1], _stream()), "iir_vit_assemble", lib)
    return out


def adastep_update(preview, pred_x0, previewer_mean, preview_factor, cond_scale, *, n_rep: int, next_scale: float, next_keep: float):
    """adastep_restore bookkeeping of one step (see iir_adastep_update); every tensor fp32, updated in place"""
    lib = _L()
    for t, n in ((preview, "preview"), (pred_x0, "pred_x0"), (previewer_mean, "previewer_mean"), (preview_factor, "preview_factor"),
                 (cond_scale, "cond_scale")):
        _f32c(t, n)
    B = pred_x0.shape[0]
    _lib.check(lib.iir_adastep_update(_p(preview), _p(pred_x0), _p(previewer_mean), _p(preview_factor), _p(cond_scale), B, n_rep,
                                      pred_x0[0].numel(), float(next_scale), float(next_keep), _stream()), "iir_adastep_update", lib)


def timestep_embedding(t, dim: int, out):
    lib = _L(out)
    _f32c(t, "t")
    _lib.check(lib.iir_timestep_embedding(_p(t), t.numel(), dim, _p(out), _dt(out), _stream()),
               "iir_timestep_embedding", lib)
    return out


def linear_small(x, w, bias, out, *, M: int, N: int, K: int, act: int = ACT_NONE):
    lib = _L(w, out)
    with _Prof("linear_small", bytes=float(N) * K * w.element_size()):
        _lib.check(lib.iir_linear_small(_p(x), _dt(x), _p(w), _dt(w), _p(_f32c(bias, "bias")), _p(out),
                                        _dt(out), M, N, K, act, _stream()), "iir_linear_small", lib)
    return out


def lcm_step(eps, x, out, *, alpha_prod_t: float, c_skip: float, c_out: float):
    lib = _L(eps)
    _f32c(x, "x")
    _f32c(out, "out")
    _lib.check(lib.iir_lcm_step(_p(eps), _dt(eps), _p(x), _p(out), x.numel(), alpha_prod_t, c_skip,
                                c_out, _stream()), "iir_lcm_step", lib)
    return out


def cfg_ddpm_step(eps_uncond, eps_cond, x, noise, prev, pred_x0, *, guidance: float,
                  alpha_prod_t: float, c_x0: float, c_xt: float, sigma: float):
    lib = _L(eps_uncond)
    _f32c(x, "x")
    _f32c(noise, "noise")
    _f32c(prev, "prev")
    _f32c(pred_x0, "pred_x0")
    _lib.check(lib.iir_cfg_ddpm_step(_p(eps_uncond), _p(eps_cond), _dt(eps_uncond), _p(x), _p(noise),
                                     _p(prev), _p(pred_x0), x.numel(), guidance, alpha_prod_t, c_x0,
                                     c_xt, sigma, _stream()), "iir_cfg_ddpm_step", lib)
    return prev


def add_noise(x0, noise, out, *, alpha_prod_t: float):
    lib = _L()
    _lib.check(lib.iir_add_noise(_p(_f32c(x0, "x0")), _p(_f32c(noise, "noise")), _p(_f32c(out, "out")),
                                 x0.numel(), alpha_prod_t, _stream()), "iir_add_noise", lib)
    return out


def softmax_rows(x: torch.Tensor, out: torch.Tensor, *, scale: float):
    """out[r, :] = softmax(x[r, :] * scale); x fp32 [rows, n] (row-contiguous), out fp32 or 16-bit."""
    lib = _L(x, out)
    if x.dtype != torch.float32 or x.ndim != 2 or out.ndim != 2 or x.stride(1) != 1 or out.stride(1) != 1 or x.shape != out.shape:
        raise TypeError("softmax_rows: x fp32 [rows, n] and out [rows, n] with unit column stride")
    rows, n = x.shape
    with _Prof("softmax_rows", bytes=float(rows) * n * (3 * 4 + out.element_size())):
        _lib.check(lib.iir_softmax_rows(_p(x), x.stride(0), _p(out), _dt(out), out.stride(0), rows, n, float(scale), _stream()),
                   "iir_softmax_rows", lib)
    return out


def gaussian_sample(moments: torch.Tensor, noise, out: torch.Tensor, *, scale: float = 1.0):
    """moments [B, 2L, h, w] fp32 (mean | logvar) -> out [B, L, h, w] = (mean + exp(0.5 clamp(logvar)) * noise) * scale;
    noise None = the mode."""
    lib = _L()
    B = moments.shape[0]
    half = moments[0].numel() // 2
    _f32c(moments, "moments"), _f32c(out, "out")
    if noise is not None:
        _f32c(noise, "noise")
    _lib.check(lib.iir_gaussian_sample(_p(moments), _p(noise), _p(out), B, half, float(scale), _stream()),
               "iir_gaussian_sample", lib)
    return out


def cfg_rescale(eps_uncond, eps_cond, out, *, guidance: float, rescale: float):
    """out = rescale_noise_cfg(e_u + g (e_c - e_u), e_c, rescale) per sample; fp32 [B, ...] tensors."""
    lib = _L()
    for t, n in ((eps_uncond, "eps_uncond"), (eps_cond, "eps_cond"), (out, "out")):
        _f32c(t, n)
    B = eps_cond.shape[0]
    _lib.check(lib.iir_cfg_rescale(_p(eps_uncond), _p(eps_cond), _p(out), B, eps_cond[0].numel(), float(guidance), float(rescale),
                                   _stream()), "iir_cfg_rescale", lib)
    return out
