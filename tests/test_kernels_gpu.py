"""Per-kernel numerics: every C-ABI entry point against a plain PyTorch fp32 reference of the same
op (computed on the GPU in fp32 with TF32 disabled).  Tolerances: fp32 kernels 2e-5 relative L2,
bf16 tensor-core kernels 6e-3 relative L2 (bf16 operand rounding, fp32 accumulation)."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from instantir_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
DEV = "cuda"
# 16-bit operand type of the tensor-core tests in this file: fp16 = the headline precision (the reference's own,
# infer.py:119); tests/test_kernels_bf16_gpu.py re-runs the same tests on the bf16 build.  Tolerances are the bf16
# ones (6e-3 / 8e-3 relative L2: 8-bit-mantissa operands, fp32 accumulation); fp16 passes them with a wide margin.
H16 = torch.bfloat16 if os.environ.get("IIR_TEST_H16", "fp16") == "bf16" else torch.float16


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


def pack_pairs(w, b, bn):
    """[first | second] halves interleaved per bn-wide tile (GEGLU / SFT weight packing)."""
    n2 = w.shape[0] // 2
    half = bn // 2
    w1, w2 = w[:n2], w[n2:]
    b1, b2 = b[:n2], b[n2:]
    ws, bs = [], []
    for t in range(n2 // half):
        ws += [w1[t * half:(t + 1) * half], w2[t * half:(t + 1) * half]]
        bs += [b1[t * half:(t + 1) * half], b2[t * half:(t + 1) * half]]
    return torch.cat(ws).contiguous(), torch.cat(bs).contiguous()


# ------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (300, 192, 320, 64), (1024, 640, 640, 128),
                                      (2048, 1280, 1280, 256), (515, 320, 2560, 160), (1024, 3840, 640, 224),
                                      (640, 1280, 320, 96), (2048, 1280, 1280, None), (4096, 640, 2560, None)])
def test_gemm_linear(tc, M, N, K, bn):
    dt = H16 if tc else torch.float32
    a = rnd(M, K, seed=1, dtype=dt)
    w = rnd(N, K, seed=2, scale=K ** -0.5, dtype=dt)
    bias = rnd(N, seed=3)
    res = rnd(M, N, seed=4)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float32)
    ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, residual=res, bn=bn, tc=tc)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias + res
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < (3e-3 if tc else 2e-5)


@pytest.mark.parametrize("tc", [True, False])
def test_gemm_bf16_out_silu_rowvec(tc):
    M, N, K, rps = 512, 256, 192, 128
    dt = H16 if tc else torch.float32
    a = rnd(M, K, seed=1, dtype=dt)
    w = rnd(N, K, seed=2, scale=K ** -0.5, dtype=dt)
    bias = rnd(N, seed=3)
    rv = rnd(M // rps, N, seed=5)
    out = torch.empty(M, N, device=DEV, dtype=dt)
    ops.gemm(a, w, out, M=M, N=N, K=K, bias=bias, rowvec=rv, rows_per_sample=rps, act=ops.ACT_SILU,
             bn=128, tc=tc)
    torch.cuda.synchronize()
    ref = F.silu(a.float() @ w.float().t() + bias + rv.repeat_interleave(rps, 0))
    assert rel_l2(out, ref) < (6e-3 if tc else 2e-5)


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("bn", [64, 128, 256])
def test_gemm_geglu(tc, bn):
    M, C, K = 384, 256, 128  # proj: K -> 2*C, out C
    dt = H16 if tc else torch.float32
    a = rnd(M, K, seed=1, dtype=dt)
    w = rnd(2 * C, K, seed=2, scale=K ** -0.5, dtype=dt)
    b = rnd(2 * C, seed=3)
    wp, bp = pack_pairs(w, b, bn)
    out = torch.empty(M, C, device=DEV, dtype=dt)
    ops.gemm(a, wp, out, M=M, N=2 * C, K=K, bias=bp, pair=ops.PAIR_GEGLU, bn=bn, tc=tc)
    torch.cuda.synchronize()
    y = a.float() @ w.float().t() + b
    ref = y[:, :C] * F.gelu(y[:, C:])
    assert rel_l2(out, ref) < (8e-3 if tc else 2e-5)


@pytest.mark.parametrize("tc", [True, False])
def test_gemm_sft_pair(tc):
    M, C, K, bn = 256, 128, 192, 128
    dt = H16 if tc else torch.float32
    a = rnd(M, K, seed=1, dtype=dt)
    w = rnd(2 * C, K, seed=2, scale=K ** -0.5, dtype=dt)  # rows [gamma | beta]
    b = rnd(2 * C, seed=3)
    h = rnd(M, C, seed=6)
    wp, bp = pack_pairs(w, b, bn)
    out = torch.empty(M, C, device=DEV, dtype=dt)
    ops.gemm(a, wp, out, M=M, N=2 * C, K=K, bias=bp, pair=ops.PAIR_SFT, aux=h, bn=bn, tc=tc)
    torch.cuda.synchronize()
    y = a.float() @ w.float().t() + b
    ref = h * (y[:, :C] + 1) + y[:, C:]
    assert rel_l2(out, ref) < (8e-3 if tc else 2e-5)


def _conv_ref(x_nhwc, w_packed, bias, stride=1, up2=False):
    n, h, w_, c = x_nhwc.shape
    co = w_packed.shape[0]
    x = x_nhwc.float().permute(0, 3, 1, 2)
    if up2:
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
    wt = w_packed.float().view(co, 3, 3, c).permute(0, 3, 1, 2)
    y = F.conv2d(x, wt, bias, stride=stride, padding=1)
    return y.permute(0, 2, 3, 1).reshape(-1, co)


# ------------------------------------------------------------------ LayerNorm folded into GEMMs
class _Stream:
    def __init__(self, M, C):
        self.h16 = torch.full((M, C), float("nan"), device=DEV, dtype=H16)
        self.acc = torch.zeros(2, M, 2, device=DEV, dtype=torch.int64)
        self.acc[1] = 12345  # the consumer must clear it for the next producer
        self.cur = 0


@pytest.mark.parametrize("M,C,N2,bn_prod,bn_cons,residual,cluster", [
    (2048, 1280, 3840, None, None, True, None), (300, 320, 320, 96, 64, True, 1), (1024, 640, 1920, 256, 160, False, 2),
    (515, 1280, 1280, 192, 224, True, None), (256, 64, 64, 32, 32, True, 1)])
def test_gemm_folded_layernorm(M, C, N2, bn_prod, bn_cons, residual, cluster):
    """producer GEMM writes the fp32 stream + its 16-bit copy + per-row partial sums; consumer GEMM on that copy
    with W' = W*gamma, colsum, b' equals Linear(LayerNorm(stream)) (include/instantir_b200.h, folded LN)."""
    K0, eps = 256, 1e-5
    a = rnd(M, K0, seed=1, dtype=H16)
    w0 = rnd(C, K0, seed=2, scale=K0 ** -0.5, dtype=H16)
    b0 = rnd(C, seed=3)
    res = rnd(M, C, seed=4, scale=2.0) + 0.5 if residual else None
    st = _Stream(M, C)
    h = res.clone() if residual else torch.empty(M, C, device=DEV)
    ops.gemm(a, w0, h, M=M, N=C, K=K0, bias=b0, residual=h if residual else None, bn=bn_prod, cluster=cluster, ln_out=st)
    torch.cuda.synchronize()
    h_ref = a.float() @ w0.float().t() + b0 + (res if residual else 0)
    assert rel_l2(h, h_ref) < 3e-3
    assert torch.equal(st.h16, h.to(H16))
    assert torch.allclose(st.acc[0, :, 0].double() / 2.0 ** 32, h.double().sum(1), rtol=1e-5, atol=1e-3)
    assert torch.allclose(st.acc[0, :, 1].double() / 2.0 ** 24, (h.double() ** 2).sum(1), rtol=1e-5, atol=1e-3)
    gamma, beta = 1.0 + 0.3 * rnd(C, seed=5), 0.2 * rnd(C, seed=6)
    w = rnd(N2, C, seed=7, scale=C ** -0.5)
    b = rnd(N2, seed=8)
    wf = (w * gamma[None, :]).to(H16)
    colsum = wf.float().sum(1).contiguous()
    bf = (w @ beta + b).contiguous()
    out = torch.full((M, N2), float("nan"), device=DEV, dtype=H16)
    ops.gemm(st.h16, wf, out, M=M, N=N2, K=C, bias=bf, bn=bn_cons, cluster=cluster, ln_in=(st, colsum, eps))
    torch.cuda.synchronize()
    ref = F.layer_norm(h, (C,), gamma, beta, eps) @ w.t() + b
    assert rel_l2(out, ref) < 8e-3
    assert st.cur == 1 and int(st.acc[1].abs().max()) == 0


def test_gemm_folded_layernorm_geglu_consumer():
    M, C, eps, bn = 640, 320, 1e-5, 256
    h = rnd(M, C, seed=1, scale=1.5) + 0.3
    st = _Stream(M, C)
    # producer: identity-free path, write stats through a 1-tile GEMM with zero weights + residual
    z = torch.zeros(M, 64, device=DEV, dtype=H16)
    ops.gemm(z, torch.zeros(C, 64, device=DEV, dtype=H16), h, M=M, N=C, K=64, residual=h, ln_out=st)
    gamma, beta = 1.0 + 0.3 * rnd(C, seed=5), 0.2 * rnd(C, seed=6)
    w = rnd(8 * C, C, seed=7, scale=C ** -0.5)
    b = rnd(8 * C, seed=8)
    wp, bp = pack_pairs(w * gamma[None, :], w @ beta + b, bn)
    wp16 = wp.to(H16)
    out = torch.empty(M, 4 * C, device=DEV, dtype=H16)
    ops.gemm(st.h16, wp16, out, M=M, N=8 * C, K=C, bias=bp.contiguous(), pair=ops.PAIR_GEGLU, bn=bn,
             ln_in=(st, wp16.float().sum(1).contiguous(), eps))
    torch.cuda.synchronize()
    y = F.layer_norm(h, (C,), gamma, beta, eps) @ w.t() + b
    ref = y[:, :4 * C] * F.gelu(y[:, 4 * C:])
    assert rel_l2(out, ref) < 8e-3


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("n,H,W,Cin,Cout,bn", [(2, 16, 16, 64, 128, 128), (1, 32, 32, 128, 64, 64), (2, 32, 32, 64, 320, None),
                                               (4, 8, 8, 64, 64, 64), (2, 64, 32, 64, 192, 64),
                                               (1, 8, 128, 64, 64, 64), (3, 8, 8, 128, 128, 128),
                                               (1, 12, 24, 64, 64, 64)])
def test_conv3x3(tc, n, H, W, Cin, Cout, bn):
    dt = H16 if tc else torch.float32
    x = rnd(n, H, W, Cin, seed=1, dtype=dt)
    w = rnd(Cout, 9 * Cin, seed=2, scale=(9 * Cin) ** -0.5, dtype=dt)
    bias = rnd(Cout, seed=3)
    rv = rnd(n, Cout, seed=4)
    M = n * H * W
    out = torch.full((M, Cout), float("nan"), device=DEV, dtype=torch.float32)
    ops.gemm(x, w, out, M=M, N=Cout, K=9 * Cin, bias=bias, rowvec=rv, rows_per_sample=H * W, bn=bn,
             conv=dict(n_img=n, H=H, W=W, Cin=Cin), tc=tc)
    torch.cuda.synchronize()
    ref = _conv_ref(x, w, bias) + rv.repeat_interleave(H * W, 0)
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) < (3e-3 if tc else 2e-5)


@pytest.mark.parametrize("stride,up2", [(2, 0), (1, 1)])
def test_conv3x3_simt_stride_upsample(stride, up2):
    n, H, W, Cin, Cout = 2, 8, 12, 32, 48
    x = rnd(n, H, W, Cin, seed=1)
    w = rnd(Cout, 9 * Cin, seed=2, scale=0.05)
    bias = rnd(Cout, seed=3)
    Hin, Win = (2 * H, 2 * W) if up2 else (H, W)
    Ho, Wo = (Hin - 1) // stride + 1, (Win - 1) // stride + 1
    M = n * Ho * Wo
    out = torch.empty(M, Cout, device=DEV)
    ops.gemm(x, w, out, M=M, N=Cout, K=9 * Cin, bias=bias,
             conv=dict(n_img=n, H=Hin, W=Win, Cin=Cin, stride=stride, up2=up2), tc=False)
    torch.cuda.synchronize()
    assert rel_l2(out, _conv_ref(x, w, bias, stride=stride, up2=bool(up2))) < 2e-5


def test_im2col_s2_matches_strided_conv():
    n, H, W, C, Cout = 2, 16, 16, 64, 64
    x = rnd(n, H, W, C, seed=1)
    w = rnd(Cout, 9 * C, seed=2, scale=0.04, dtype=H16)
    cols = torch.empty(n * (H // 2) * (W // 2), 9 * C, device=DEV, dtype=H16)
    ops.im2col3x3_s2(x, cols, n_img=n, H=H, W=W, C=C)
    out = torch.empty(cols.shape[0], Cout, device=DEV)
    ops.gemm(cols, w, out, M=cols.shape[0], N=Cout, K=9 * C, bn=64, tc=True)
    torch.cuda.synchronize()
    ref = _conv_ref(x.to(H16), w, None, stride=2)
    assert rel_l2(out, ref) < 3e-3


@pytest.mark.parametrize("tc", [True, False])
def test_stride2_conv_padded_bottom_right_only(tc):
    """the VAE encoder's Downsample2D(padding=0): F.pad (0, 1, 0, 1) then a pad-0 stride-2 conv; tcgen05 path via the
    im2col gather, fp32 check mode inside the SIMT conv"""
    n, H, W, C, Cout = 2, 16, 24, 64, 128
    dt = H16 if tc else torch.float32
    x = rnd(n, H, W, C, seed=1)
    w = rnd(Cout, 9 * C, seed=2, scale=0.04, dtype=dt)
    M = n * (H // 2) * (W // 2)
    out = torch.full((M, Cout), float("nan"), device=DEV)
    if tc:
        cols = torch.empty(M, 9 * C, device=DEV, dtype=dt)
        ops.im2col3x3_s2(x, cols, n_img=n, H=H, W=W, C=C, asym=True)
        ops.gemm(cols, w, out, M=M, N=Cout, K=9 * C, tc=True)
    else:
        ops.gemm(x.reshape(-1, C), w, out, M=M, N=Cout, K=9 * C, conv=dict(n_img=n, H=H, W=W, Cin=C, stride=2, asym=1), tc=False)
    torch.cuda.synchronize()
    xx = x.to(dt).float().permute(0, 3, 1, 2)
    w4 = w.float().reshape(Cout, 3, 3, C).permute(0, 3, 1, 2)
    ref = F.conv2d(F.pad(xx, (0, 1, 0, 1)), w4, stride=2).permute(0, 2, 3, 1).reshape(M, Cout)
    assert rel_l2(out, ref) < (3e-3 if tc else 2e-5)
    with pytest.raises(Exception):
        ops.im2col3x3_s2(x[:, :15].contiguous(), torch.empty(8, 9 * C, device=DEV, dtype=H16), n_img=n, H=15, W=W, C=C, asym=True)


def test_conv3x3_direct_layouts():
    n, H, W, Cin, Cout = 2, 16, 8, 4, 64
    x = rnd(n, Cin, H, W, seed=1)
    w = rnd(Cout, 3, 3, Cin, seed=2, scale=0.2)
    b = rnd(Cout, seed=3)
    canvas = torch.zeros(n, 2 * H, W, Cout, device=DEV)
    ops.conv3x3_direct(x, w, b, canvas, in_nchw=True, out_nchw=False, n_img=n, H=H, W=W, Cin=Cin,
                       Cout=Cout, out_H=2 * H, out_row_off=H)
    ref = F.conv2d(x, w.permute(0, 3, 1, 2), b, padding=1).permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert rel_l2(canvas[:, H:], ref) < 2e-5 and float(canvas[:, :H].abs().max()) == 0.0
    # NHWC bf16 in -> NCHW fp32 out (conv_out)
    xh = rnd(n, H, W, 64, seed=5, dtype=H16)
    w2 = rnd(4, 3, 3, 64, seed=6, scale=0.05)
    b2 = rnd(4, seed=7)
    o2 = torch.empty(n, 4, H, W, device=DEV)
    ops.conv3x3_direct(xh, w2, b2, o2, in_nchw=False, out_nchw=True, n_img=n, H=H, W=W, Cin=64, Cout=4)
    torch.cuda.synchronize()
    ref2 = F.conv2d(xh.float().permute(0, 3, 1, 2), w2.permute(0, 3, 1, 2), b2, padding=1)
    assert rel_l2(o2, ref2) < 2e-5


# ------------------------------------------------------------------------------- attention
def _sdpa_ref(q, k, v, heads, scale):
    B, n, C = q.shape
    d = C // heads
    qh = q.float().view(B, n, heads, d).transpose(1, 2)
    kh = k.float().view(B, -1, heads, d).transpose(1, 2)
    vh = v.float().view(B, -1, heads, d).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, n, C)


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("B,heads,n", [(1, 1, 128), (2, 2, 256), (2, 4, 320), (1, 2, 1024), (2, 1, 64)])
def test_self_attention_fused_qkv(tc, B, heads, n):
    C = heads * 64
    dt = H16 if tc else torch.float32
    qkv = rnd(B, n, 3 * C, seed=1, dtype=dt)
    out = torch.full((B, n, C), float("nan"), device=DEV, dtype=dt)
    ops.attention(qkv, 0, 3 * C, [qkv], [C], [3 * C], [qkv], [2 * C], [3 * C], [n], [1.0], out, 0, C,
                  B=B, heads=heads, n_q=n, softmax_scale=0.125, tc=tc)
    torch.cuda.synchronize()
    ref = _sdpa_ref(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, 0.125)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < (8e-3 if tc else 2e-5)


@pytest.mark.parametrize("B,heads,n", [(2, 20, 1024), (2, 10, 4096), (3, 20, 1000), (2, 20, 2048)])
def test_self_attention_persistent_many_tiles(B, heads, n):
    """more (batch, head, query-block) tiles than the 2 x 148 resident CTAs: every CTA of the persistent kernel
    walks 2-3 tiles (barrier phases, K/V ring and TMEM carried across tiles); n = 1000 adds ragged last blocks.
    Rows are compared per batch against fp32 SDPA; a second launch must give bit-identical results."""
    C = heads * 64
    qkv = rnd(B, n, 3 * C, seed=11, dtype=H16)
    out = torch.full((B, n, C), float("nan"), device=DEV, dtype=H16)
    args = (qkv, 0, 3 * C, [qkv], [C], [3 * C], [qkv], [2 * C], [3 * C], [n], [1.0])
    ops.attention(*args, out, 0, C, B=B, heads=heads, n_q=n, softmax_scale=0.125)
    out2 = torch.empty_like(out)
    ops.attention(*args, out2, 0, C, B=B, heads=heads, n_q=n, softmax_scale=0.125)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all() and torch.equal(out, out2)
    for b in range(B):
        ref = _sdpa_ref(qkv[b:b + 1, :, :C], qkv[b:b + 1, :, C:2 * C], qkv[b:b + 1, :, 2 * C:], heads, 0.125)
        assert rel_l2(out[b:b + 1], ref) < 8e-3, b


def _sdpa_ref_chunked(q, k, v, heads, scale, chunk=4096):
    """fp32 softmax(QK^T)V one (batch, head, query chunk) at a time: bounded memory at 32 768 keys"""
    B, n, C = q.shape
    d = C // heads
    out = torch.empty(B, n, C, device=q.device, dtype=torch.float32)
    for b in range(B):
        for h in range(heads):
            kh = k[b, :, h * d:(h + 1) * d].float()
            vh = v[b, :, h * d:(h + 1) * d].float()
            for q0 in range(0, n, chunk):
                qh = q[b, q0:q0 + chunk, h * d:(h + 1) * d].float()
                out[b, q0:q0 + chunk, h * d:(h + 1) * d] = torch.softmax(qh @ kh.t() * scale, dim=-1) @ vh
    return out


@pytest.mark.parametrize("B,heads,n", [(2, 10, 8192), (1, 3, 16384), (1, 2, 32768), (1, 1, 32768 - 77)])
def test_self_attention_long_sequences(B, heads, n):
    """the Aggregator's 1024² self-attention (8192 tokens, 10 heads, CFG batch 2) and BASELINE config 5's shapes
    (2048²: 16 384 tokens in the UNet, 32 768 in the Aggregator; 64 / 128 / 256 key blocks per query tile), plus a
    ragged 32 691-token case, against an fp32 softmax reference"""
    C = heads * 64
    qkv = rnd(B, n, 3 * C, seed=21, dtype=H16)
    out = torch.full((B, n, C), float("nan"), device=DEV, dtype=H16)
    ops.attention(qkv, 0, 3 * C, [qkv], [C], [3 * C], [qkv], [2 * C], [3 * C], [n], [1.0], out, 0, C,
                  B=B, heads=heads, n_q=n, softmax_scale=0.125)
    torch.cuda.synchronize()
    ref = _sdpa_ref_chunked(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, 0.125)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 8e-3


def test_cross_attention_persistent_many_tiles():
    """the one-block-per-segment kernel (text 77 + image 64 keys) with 320 tiles on 296 resident CTAs"""
    B, heads, n, nt, ni, scale_ip = 2, 20, 1024, 77, 64, 0.7
    C = heads * 64
    q = rnd(B, n, C, seed=1, dtype=H16)
    kvt = rnd(B, nt, 2 * C, seed=2, dtype=H16)
    ki = rnd(B, ni, C, seed=3, dtype=H16)
    vi = rnd(B, ni, C, seed=4, dtype=H16)
    out = torch.full((B, n, C), float("nan"), device=DEV, dtype=H16)
    ops.attention(q, 0, C, [kvt, ki], [0, 0], [2 * C, C], [kvt, vi], [C, 0], [2 * C, C], [nt, ni],
                  [1.0, scale_ip], out, 0, C, B=B, heads=heads, n_q=n, softmax_scale=0.125)
    torch.cuda.synchronize()
    ref = _sdpa_ref(q, kvt[..., :C], kvt[..., C:], heads, 0.125) + scale_ip * _sdpa_ref(q, ki, vi, heads, 0.125)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < 8e-3


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("nt,ni", [(77, 64), (200, 64), (128, 1)])
def test_decoupled_cross_attention_two_segments(tc, nt, ni):
    """(77, 64): both segments fit one key block -> the normalise-in-registers kernel; (200, 64): the
    multi-block two-segment kernel; (128, 1): a full block next to a single key"""
    B, heads, n, scale_ip = 2, 2, 256, 0.7
    C = heads * 64
    dt = H16 if tc else torch.float32
    q = rnd(B, n, C, seed=1, dtype=dt)
    kvt = rnd(B, nt, 2 * C, seed=2, dtype=dt)
    ki = rnd(B, ni, C, seed=3, dtype=dt)
    vi = rnd(B, ni, C, seed=4, dtype=dt)
    out = torch.empty(B, n, C, device=DEV, dtype=dt)
    ops.attention(q, 0, C, [kvt, ki], [0, 0], [2 * C, C], [kvt, vi], [C, 0], [2 * C, C], [nt, ni],
                  [1.0, scale_ip], out, 0, C, B=B, heads=heads, n_q=n, softmax_scale=0.125, tc=tc)
    torch.cuda.synchronize()
    ref = _sdpa_ref(q, kvt[..., :C], kvt[..., C:], heads, 0.125) + scale_ip * _sdpa_ref(q, ki, vi, heads, 0.125)
    assert rel_l2(out, ref) < (8e-3 if tc else 2e-5)


# ----------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("xdt,odt", [(torch.float32, torch.float32), (torch.float32, H16),
                                     (H16, H16)])
@pytest.mark.parametrize("n,HW,C,silu", [(2, 256, 64, True), (2, 1024, 320, True), (1, 4096, 1920, False),
                                         (3, 64, 2560, True)])
def test_groupnorm(xdt, odt, n, HW, C, silu):
    x = (rnd(n, HW, C, seed=1) * 2 + 0.5).to(xdt)
    g, b = rnd(C, seed=2), rnd(C, seed=3)
    out = torch.empty(n, HW, C, device=DEV, dtype=odt)
    ops.groupnorm(x, g, b, out, n_img=n, HW=HW, C=C, eps=1e-5, silu=silu)
    torch.cuda.synchronize()
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, g, b, 1e-5).permute(0, 2, 1)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(out, ref) < (5e-3 if odt == H16 else 2e-5)


@pytest.mark.parametrize("C", [64, 640, 1280, 2048])
def test_layernorm_affine_and_adaln(C):
    rows, rps = 96, 48
    x = rnd(rows, C, seed=1) * 3 + 1
    g, b = rnd(C, seed=2), rnd(C, seed=3)
    out = torch.empty(rows, C, device=DEV)
    ops.layernorm(x, g, b, out, rows=rows, C=C, eps=1e-5)
    torch.cuda.synchronize()
    assert rel_l2(out, F.layer_norm(x, (C,), g, b, 1e-5)) < 2e-5
    mod = rnd(rows // rps, 2 * C, seed=4) * 0.3
    ops.layernorm(x, None, None, out, rows=rows, C=C, eps=1e-6, mod=mod, rows_per_sample=rps)
    torch.cuda.synchronize()
    m = mod.repeat_interleave(rps, 0)
    ref = F.layer_norm(x, (C,), None, None, 1e-6) * (1 + m[:, C:]) + m[:, :C]
    assert rel_l2(out, ref) < 2e-5


def test_adaln_batched_and_strided_mod():
    """one launch over several tensors of different width == per-tensor adaLN; mod is a column slice"""
    rows, rps, Cs = 64, 32, [128, 256, 128]
    total = sum(2 * c for c in Cs) + 64
    mod_all = rnd(rows // rps, total, seed=9) * 0.3
    xs = [rnd(rows, c, seed=10 + k) * 2 + 0.5 for k, c in enumerate(Cs)]
    outs = [torch.empty(rows, c, device=DEV, dtype=H16) for c in Cs]
    offs, o = [], 64
    for c in Cs:
        offs.append(o)
        o += 2 * c
    table = ops.adaln_items([(x, out, off, c) for x, out, off, c in zip(xs, outs, offs, Cs)], DEV)
    ops.adaln_batched(table, len(Cs), mod_all, H16, rows=rows, rows_per_sample=rps, eps=1e-6)
    torch.cuda.synchronize()
    for x, out, off, c in zip(xs, outs, offs, Cs):
        m = mod_all[:, off:off + 2 * c].repeat_interleave(rps, 0)
        ref = F.layer_norm(x, (c,), None, None, 1e-6) * (1 + m[:, c:]) + m[:, :c]
        assert rel_l2(out, ref) < 4e-3
        single = torch.empty(rows, c, device=DEV)
        ops.layernorm(x, None, None, single, rows=rows, C=c, eps=1e-6, mod=mod_all[:, off:off + 2 * c], rows_per_sample=rps)
        torch.cuda.synchronize()
        assert rel_l2(single, ref) < 2e-5


def test_gemm_rowvec_column_slice():
    """rowvec taken as a column slice of a wider buffer (banked time_emb_proj outputs)"""
    M, N, K, rps = 256, 128, 64, 128
    a, w = rnd(M, K, seed=1, dtype=H16), rnd(N, K, seed=2, dtype=H16)
    wide = rnd(M // rps, 3 * N, seed=3)
    for tc in (True, False):
        out = torch.empty(M, N, device=DEV)
        ops.gemm(a, w, out, M=M, N=N, K=K, rowvec=wide[:, N:2 * N], rows_per_sample=rps, tc=tc)
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t() + wide[:, N:2 * N].repeat_interleave(rps, 0)
        assert rel_l2(out, ref) < 1e-5


# ----------------------------------------------------------------------- movement / small
def test_concat_inject_and_plain_add():
    M, C1, C2, rps = 512, 128, 64, 256
    h, sk = rnd(M, C1, seed=1), rnd(M, C2, seed=2)
    rh, rs = rnd(M, C1, seed=3, dtype=H16), rnd(M, C2, seed=4, dtype=H16)
    cs = torch.tensor([0.0, 1.0], device=DEV)
    out = torch.empty(M, C1 + C2, device=DEV, dtype=H16)
    ops.concat_inject(h, C1, sk, C2, out, M=M, rh=rh, rs=rs, cond_scale=cs, rows_per_sample=rps)
    torch.cuda.synchronize()
    s = cs.repeat_interleave(rps)[:, None]
    ref = torch.cat([h + s * rh.float(), sk + s * rs.float()], 1)
    assert rel_l2(out, ref) < 4e-3
    out2 = torch.empty(M, C1, device=DEV)
    ops.concat_inject(h, C1, None, 0, out2, M=M, rh=rh, cond_scale=cs, rows_per_sample=rps)
    torch.cuda.synchronize()
    assert rel_l2(out2, h + s * rh.float()) < 1e-6


def test_upsample_cast_silu_add_timestep():
    n, H, W, C = 2, 8, 4, 64
    x = rnd(n, H, W, C, seed=1)
    up = torch.empty(n, 2 * H, 2 * W, C, device=DEV, dtype=H16)
    ops.upsample2x(x, up, n_img=n, H=H, W=W, C=C)
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    torch.cuda.synchronize()
    assert rel_l2(up, ref) < 4e-3
    src = rnd(64, 256, seed=2)
    dst = torch.zeros(64, 128, device=DEV, dtype=H16)
    ops.cast2d(src[:, 64:], 256, dst, 128, rows=64, cols=128)
    torch.cuda.synchronize()
    assert rel_l2(dst, src[:, 64:192]) < 4e-3
    y = torch.empty_like(src)
    ops.silu(src, y)
    z = torch.empty_like(src)
    ops.add(src, y, z)
    torch.cuda.synchronize()
    assert rel_l2(y, F.silu(src)) < 1e-6 and rel_l2(z, src + F.silu(src)) < 1e-6
    t = torch.tensor([958.0, 501.0, 1.0, 1024.0], device=DEV)
    emb = torch.empty(4, 320, device=DEV)
    ops.timestep_embedding(t, 320, emb)
    torch.cuda.synchronize()
    half = 160
    freq = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=DEV) / half)
    arg = t[:, None] * freq[None]
    assert rel_l2(emb, torch.cat([arg.cos(), arg.sin()], -1)) < 1e-5


@pytest.mark.parametrize("wdt", [torch.float32, H16])
def test_linear_small(wdt):
    M, N, K = 2, 2560, 1280
    x = rnd(M, K, seed=1)
    w = rnd(N, K, seed=2, scale=K ** -0.5, dtype=wdt)
    b = rnd(N, seed=3)
    out = torch.empty(M, N, device=DEV)
    ops.linear_small(x, w, b, out, M=M, N=N, K=K, act=ops.ACT_SILU)
    torch.cuda.synchronize()
    assert rel_l2(out, F.silu(x @ w.float().t() + b)) < 2e-5


# ------------------------------------------------------------------------------- scheduler
def test_scheduler_kernels():
    n = 2 * 4 * 32 * 32
    eps_u, eps_c, x, z = (rnd(n, seed=s) for s in (1, 2, 3, 4))
    abar, c_skip, c_out = 0.0075347, 2.7e-9, 0.99999
    out = torch.empty(n, device=DEV)
    ops.lcm_step(eps_c, x, out, alpha_prod_t=abar, c_skip=c_skip, c_out=c_out)
    x0 = (x - math.sqrt(1 - abar) * eps_c) / math.sqrt(abar)
    torch.cuda.synchronize()
    assert rel_l2(out, c_out * x0 + c_skip * x) < 1e-6
    prev, px0 = torch.empty(n, device=DEV), torch.empty(n, device=DEV)
    g, c0, c1, sig = 7.0, 0.031, 0.84, 0.52
    ops.cfg_ddpm_step(eps_u, eps_c, x, z, prev, px0, guidance=g, alpha_prod_t=abar, c_x0=c0, c_xt=c1,
                      sigma=sig)
    e = eps_u + g * (eps_c - eps_u)
    x0 = (x - math.sqrt(1 - abar) * e) / math.sqrt(abar)
    torch.cuda.synchronize()
    assert rel_l2(px0, x0) < 1e-6 and rel_l2(prev, c0 * x0 + c1 * x + sig * z) < 1e-6
    ops.add_noise(x, z, out, alpha_prod_t=abar)
    torch.cuda.synchronize()
    assert rel_l2(out, math.sqrt(abar) * x + math.sqrt(1 - abar) * z) < 1e-6


def test_no_cpu_fallback():
    from instantir_b200._lib import IIRError

    with pytest.raises(IIRError):
        ops.silu(torch.zeros(8), torch.zeros(8))


# ---------------------------------------------------------------------- fp16 library build
def test_fp16_build_gemm_conv_attention():
    """libinstantir_b200_fp16.so: same kernels with IEEE-half operands (8x finer mantissa than bf16)."""
    M, N, K = 512, 320, 640
    a = rnd(M, K, seed=1, dtype=torch.float16)
    w = rnd(N, K, seed=2, scale=K ** -0.5, dtype=torch.float16)
    out = torch.empty(M, N, device=DEV, dtype=torch.float16)
    ops.gemm(a, w, out, M=M, N=N, K=K)
    torch.cuda.synchronize()
    assert rel_l2(out, a.float() @ w.float().t()) < 6e-4
    x = rnd(2, 16, 16, 64, seed=3, dtype=torch.float16)
    wc = rnd(128, 9 * 64, seed=4, scale=(9 * 64) ** -0.5, dtype=torch.float16)
    oc = torch.empty(2 * 256, 128, device=DEV)
    ops.gemm(x, wc, oc, M=512, N=128, K=9 * 64, conv=dict(n_img=2, H=16, W=16, Cin=64))
    torch.cuda.synchronize()
    assert rel_l2(oc, _conv_ref(x, wc, None)) < 6e-4
    B, heads, n = 2, 2, 256
    C = heads * 64
    qkv = rnd(B, n, 3 * C, seed=5, dtype=torch.float16)
    o = torch.empty(B, n, C, device=DEV, dtype=torch.float16)
    ops.attention(qkv, 0, 3 * C, [qkv], [C], [3 * C], [qkv], [2 * C], [3 * C], [n], [1.0], o, 0, C,
                  B=B, heads=heads, n_q=n, softmax_scale=0.125)
    torch.cuda.synchronize()
    assert rel_l2(o, _sdpa_ref(qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:], heads, 0.125)) < 1.5e-3
    ln = torch.empty(64, 640, device=DEV, dtype=torch.float16)
    xs = rnd(64, 640, seed=6)
    ops.layernorm(xs, None, None, ln, rows=64, C=640)
    torch.cuda.synchronize()
    assert rel_l2(ln, F.layer_norm(xs, (640,))) < 6e-4


def test_wrong_16bit_dtype_is_rejected_by_each_build():
    import ctypes as C

    from instantir_b200 import _lib

    for h16, other in ((_lib.BF16, _lib.F16), (_lib.F16, _lib.BF16)):
        lib = _lib.load(h16=h16)
        assert lib.iir_h16_dtype() == h16
        g = _lib.GemmArgs()
        g.a_dtype = g.w_dtype = other
        g.out_dtype = _lib.F32
        g.M = g.N = g.K = 64
        g.bn = 64
        assert lib.iir_gemm_tc(C.byref(g), None) == -1


@pytest.mark.parametrize("odt", [torch.float32, H16])
@pytest.mark.parametrize("rows,n", [(5, 256), (300, 1024), (64, 16384), (3, 4104)])
def test_softmax_rows(odt, rows, n):
    x = rnd(rows, n, seed=3, scale=4.0)
    x[0, 7] = 60.0  # a dominant logit
    out = torch.full((rows, n), float("nan"), device=DEV, dtype=odt)
    ops.softmax_rows(x, out, scale=0.37)
    torch.cuda.synchronize()
    ref = torch.softmax(x.double() * 0.37, dim=-1)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out, ref) < (2e-6 if odt == torch.float32 else 4e-3)
    assert torch.allclose(out.float().sum(-1), torch.ones(rows, device=DEV), atol=1e-5 if odt == torch.float32 else 2e-2)


def test_cfg_rescale_matches_rescale_noise_cfg():
    """iir_cfg_rescale vs the reference's rescale_noise_cfg arithmetic (pipelines/sdxl_instantir.py:181-192) in fp64"""
    B, shape, g, phi = 3, (4, 32, 48), 7.0, 0.7
    eu, ec = rnd(B, *shape, seed=1), rnd(B, *shape, seed=2) * 1.3 + 0.2
    out = torch.full((B, *shape), float("nan"), device=DEV)
    ops.cfg_rescale(eu, ec, out, guidance=g, rescale=phi)
    torch.cuda.synchronize()
    u, c = eu.double(), ec.double()
    cfg = u + g * (c - u)
    dims = [1, 2, 3]
    ref = phi * cfg * (c.std(dim=dims, keepdim=True) / cfg.std(dim=dims, keepdim=True)) + (1 - phi) * cfg
    assert rel_l2(out, ref) < 1e-6


def test_adastep_update():
    """iir_adastep_update vs the reference arithmetic (pipelines/sdxl_instantir.py:1636-1644, :1538-1540)"""
    B, n_rep = 3, 2
    pv, x0, pm = rnd(B, 4, 16, 24, seed=1), rnd(B, 4, 16, 24, seed=2), rnd(B, 4, 16, 24, seed=3)
    pm0 = pm.clone()
    factor = torch.full((B,), float("nan"), device=DEV)
    cs = torch.full((n_rep * B,), float("nan"), device=DEV)
    ops.adastep_update(pv, x0, pm, factor, cs, n_rep=n_rep, next_scale=0.8, next_keep=1.0)
    torch.cuda.synchronize()
    # the kernel sums in fp64 and divides the two rounded fp32 sums
    want = ((pv - x0).double().pow(2).sum((1, 2, 3)).float() / (pv - pm0).double().pow(2).sum((1, 2, 3)).float())
    assert rel_l2(factor, want) < 1e-6
    assert torch.equal(pm, pv)
    assert rel_l2(cs, want.clamp(0.0, 0.8).repeat(n_rep)) < 1e-6
    ops.adastep_update(pv, x0, pm, factor, cs, n_rep=n_rep, next_scale=0.8, next_keep=0.0)   # previewer_mean == preview now: x / 0
    torch.cuda.synchronize()
    assert bool(torch.isinf(factor).all()) and float(cs.abs().sum()) == 0.0
