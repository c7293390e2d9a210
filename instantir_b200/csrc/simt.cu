// fp32 SIMT kernels: the north star's "fp32 check mode" (<=1e-4 vs the oracle) and the on-GPU
// cross-check of the tcgen05 kernels, plus the tiny-channel direct convolution (conv_in/conv_out)
// and the small-M linear used by the time-embedding / adaLN projections.
#include "common.cuh"

namespace iir {
namespace {

// ------------------------------------------------------------------------------ SIMT GEMM
struct GemmSimtParams {
  const void* a;
  const void* w;
  int M, N, K;
  long long lda;
  int conv, n_img, H, W, Cin, stride, up2, Ho, Wo, lead;
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  int rows_per_sample;
  const void* residual; int res_bf16; long long ld_res;
  const void* aux; int aux_bf16; long long ld_aux;
  void* out; int out_bf16; long long ld_out;
  int act, bn;
};

template <typename TA>
__device__ __forceinline__ float load_a(const GemmSimtParams& p, int m, int k) {
  const TA* A = reinterpret_cast<const TA*>(p.a);
  if (!p.conv) return ld_f(A + static_cast<long long>(m) * p.lda + k);
  int tap = k / p.Cin;
  int c = k - tap * p.Cin;
  int ky = tap / 3, kx = tap - ky * 3;
  int hw = p.Ho * p.Wo;
  int n = m / hw;
  int r = m - n * hw;
  int oy = r / p.Wo, ox = r - oy * p.Wo;
  int y = oy * p.stride + ky - p.lead, x = ox * p.stride + kx - p.lead;
  if (y < 0 || y >= p.H || x < 0 || x >= p.W) return 0.0f;
  int sh = p.H, sw = p.W;
  if (p.up2) { y >>= 1; x >>= 1; sh >>= 1; sw >>= 1; }
  return ld_f(A + ((static_cast<long long>(n) * sh + y) * sw + x) * p.Cin + c);
}

constexpr int ST = 64;   // tile edge
constexpr int SK = 16;   // k step

template <typename TA, typename TW, int PAIR>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmSimtParams p) {
  __shared__ float As[SK][ST + 1];
  __shared__ float Ws[SK][ST + 1];
  __shared__ float Ws2[PAIR ? SK : 1][ST + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * ST;
  const int o0 = blockIdx.x * ST;  // first OUTPUT column of the tile
  const int n_out = PAIR ? p.N / 2 : p.N;
  const int half = p.bn / 2;
  const TW* Wp = reinterpret_cast<const TW*>(p.w);
  float acc[4][4] = {};
  float acc2[4][4] = {};
  for (int k0 = 0; k0 < p.K; k0 += SK) {
    for (int i = threadIdx.x; i < ST * SK; i += 256) {
      int r = i / SK, kk = i - r * SK;
      int m = m0 + r, k = k0 + kk;
      As[kk][r] = (m < p.M && k < p.K) ? load_a<TA>(p, m, k) : 0.0f;
      int on = o0 + r;
      float w1 = 0.f, w2 = 0.f;
      if (on < n_out && k < p.K) {
        if (PAIR) {
          int t = on / half, rr = on - t * half;
          long long pn = static_cast<long long>(t) * p.bn + rr;
          w1 = ld_f(Wp + pn * p.K + k);
          w2 = ld_f(Wp + (pn + half) * p.K + k);
        } else {
          w1 = ld_f(Wp + static_cast<long long>(on) * p.K + k);
        }
      }
      Ws[kk][r] = w1;
      if (PAIR) Ws2[kk][r] = w2;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SK; ++kk) {
      float av[4], wv[4], wv2[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = As[kk][ty * 4 + i];
        wv[i] = Ws[kk][tx * 4 + i];
        if (PAIR) wv2[i] = Ws2[kk][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
          if (PAIR) acc2[i][j] = fmaf(av[i], wv2[j], acc2[i][j]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    int sample = m / p.rows_per_sample;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int on = o0 + tx * 4 + j;
      if (on >= n_out) continue;
      long long pn = on, pn2 = 0;
      if (PAIR) {
        int t = on / half, rr = on - t * half;
        pn = static_cast<long long>(t) * p.bn + rr;
        pn2 = pn + half;
      }
      float v = acc[i][j];
      if (p.bias) v += p.bias[pn];
      if (p.rowvec) v += p.rowvec[static_cast<long long>(sample) * p.ld_rowvec + pn];
      if (p.act == IIR_ACT_SILU) v = silu_f(v);
      else if (p.act == IIR_ACT_GELU) v = gelu_erf_f(v);
      else if (p.act == IIR_ACT_QUICK_GELU) v = quick_gelu_f(v);
      if (PAIR) {
        float g = acc2[i][j];
        if (p.bias) g += p.bias[pn2];
        if (PAIR == IIR_PAIR_GEGLU) {
          v = v * gelu_erf_f(g);
        } else {
          float h = p.aux_bf16
              ? ld_f(reinterpret_cast<const h16*>(p.aux) + m * p.ld_aux + on)
              : ld_f(reinterpret_cast<const float*>(p.aux) + m * p.ld_aux + on);
          v = h * (v + 1.0f) + g;
        }
      }
      if (p.residual) {
        v += p.res_bf16
            ? ld_f(reinterpret_cast<const h16*>(p.residual) + m * p.ld_res + on)
            : ld_f(reinterpret_cast<const float*>(p.residual) + m * p.ld_res + on);
      }
      if (p.out_bf16) st_f(reinterpret_cast<h16*>(p.out) + m * p.ld_out + on, v);
      else st_f(reinterpret_cast<float*>(p.out) + m * p.ld_out + on, v);
    }
  }
}

template <typename TA, typename TW>
void launch_gemm_simt(const GemmSimtParams& p, int pair, cudaStream_t st) {
  int n_out = pair ? p.N / 2 : p.N;
  dim3 grid((n_out + ST - 1) / ST, (p.M + ST - 1) / ST);
  if (pair == IIR_PAIR_NONE) gemm_simt_kernel<TA, TW, 0><<<grid, 256, 0, st>>>(p);
  else if (pair == IIR_PAIR_GEGLU) gemm_simt_kernel<TA, TW, 1><<<grid, 256, 0, st>>>(p);
  else gemm_simt_kernel<TA, TW, 2><<<grid, 256, 0, st>>>(p);
}

template <typename TW>
__device__ __forceinline__ void ld8(const TW* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<h16>(const h16* p, float (&v)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  const h162* h = reinterpret_cast<const h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = h162_to_ff(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}

// ------------------------------------------------------------------- direct 3x3 conv (tiny C)
// one thread per (pixel, cout); weights [Cout,3,3,Cin] fp32.
template <typename TI, typename TO>
__global__ void conv3x3_direct_kernel(const TI* __restrict__ in, int in_nchw,
                                      const float* __restrict__ w, const float* __restrict__ bias,
                                      TO* __restrict__ out, int out_nchw, int n_img, int H, int W,
                                      int Cin, int Cout, int out_H, int out_row_off) {
  long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  long long total = static_cast<long long>(n_img) * H * W * Cout;
  if (idx >= total) return;
  int co = static_cast<int>(idx % Cout);
  long long pix = idx / Cout;
  int x = static_cast<int>(pix % W);
  int y = static_cast<int>((pix / W) % H);
  int n = static_cast<int>(pix / (static_cast<long long>(W) * H));
  float acc = bias ? bias[co] : 0.0f;
  for (int ky = 0; ky < 3; ++ky) {
    int yy = y + ky - 1;
    if (yy < 0 || yy >= H) continue;
    for (int kx = 0; kx < 3; ++kx) {
      int xx = x + kx - 1;
      if (xx < 0 || xx >= W) continue;
      const float* wp = w + ((co * 3 + ky) * 3 + kx) * Cin;
      if (in_nchw) {
        for (int c = 0; c < Cin; ++c)
          acc = fmaf(ld_f(in + ((static_cast<long long>(n) * Cin + c) * H + yy) * W + xx), wp[c], acc);
      } else {
        const TI* ip = in + ((static_cast<long long>(n) * H + yy) * W + xx) * Cin;
        for (int c = 0; c < Cin; ++c) acc = fmaf(ld_f(ip + c), wp[c], acc);
      }
    }
  }
  if (out_nchw) {
    st_f(out + ((static_cast<long long>(n) * Cout + co) * H + y) * W + x, acc);
  } else {
    st_f(out + ((static_cast<long long>(n) * out_H + out_row_off + y) * W + x) * Cout + co, acc);
  }
}

// conv_in-type: tiny Cin (<= 8), many output channels.  Weights are staged once per block in shared
// memory transposed to [tap*Cin + c][Cout] so that consecutive threads (consecutive co) read
// consecutive words; each block produces PIX pixels x Cout outputs with coalesced stores.
constexpr int SC_PIX = 128;  // pixels per block
// conv_in-type: tiny Cin (<= 8), many output channels, NHWC output.  FMA-bound when register-blocked: a thread
// owns 4 pixels x 8 channels (channels c4..c4+3 and Cout/2 + c4..c4+3 so that both weight loads of a warp are
// conflict-free float4 rows); weights [k][Cout] and the 128-pixel patch matrix [k][pixel] sit in shared memory,
// so every FMA costs 3/32 of a 16-byte shared load (the first version did 2 scalar loads per FMA: LDS-bound, and
// re-staged the 46 KB weight matrix for every 32 pixels).  Requires Cout % 8 == 0.
template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
conv3x3_small_cin_kernel(const TI* __restrict__ in, int in_nchw, const float* __restrict__ w,
                         const float* __restrict__ bias, TO* __restrict__ out, int n_img, int H, int W,
                         int Cin, int Cout, int out_H, int out_row_off) {
  extern __shared__ __align__(16) float ws[];  // [9*Cin][Cout] weights, then [9*Cin][SC_PIX] input patches
  const int kk = 9 * Cin;
  float* xs = ws + kk * Cout;
  for (int i = threadIdx.x; i < kk * Cout; i += 256) {
    int co = i / kk, k = i - co * kk;  // source order [co][tap][c]
    ws[k * Cout + co] = w[i];
  }
  const long long pix0 = blockIdx.x * static_cast<long long>(SC_PIX);
  const long long npix = static_cast<long long>(n_img) * H * W;
  for (int i = threadIdx.x; i < SC_PIX * kk; i += 256) {
    int k = i / SC_PIX, pp = i - k * SC_PIX;  // consecutive threads -> consecutive pixels (coalesced for NCHW input)
    long long pix = pix0 + pp;
    float v = 0.f;
    if (pix < npix) {
      int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
      long long n = pix / (static_cast<long long>(W) * H);
      int tap = k / Cin, c = k - tap * Cin;
      int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W)
        v = in_nchw ? ld_f(in + ((n * Cin + c) * H + yy) * W + xx) : ld_f(in + ((n * H + yy) * W + xx) * Cin + c);
    }
    xs[k * SC_PIX + pp] = v;
  }
  __syncthreads();
  const int half = Cout >> 1;
  const int ncg = half >> 2;                 // channel groups: 4 + 4 channels each
  const int ntile = (SC_PIX / 4) * ncg;      // thread tiles of 4 pixels x 8 channels
  for (int t = threadIdx.x; t < ntile; t += 256) {
    const int cg = t % ncg, pg = t / ncg;    // consecutive lanes -> consecutive channel groups of one pixel group
    const int c0 = cg * 4, c1 = half + cg * 4, p0 = pg * 4;
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[p][j] = 0.f;
    for (int k = 0; k < kk; ++k) {
      const float4 xv = *reinterpret_cast<const float4*>(xs + k * SC_PIX + p0);
      const float4 wa = *reinterpret_cast<const float4*>(ws + k * Cout + c0);
      const float4 wb = *reinterpret_cast<const float4*>(ws + k * Cout + c1);
      const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        acc[p][0] = fmaf(xa[p], wa.x, acc[p][0]); acc[p][1] = fmaf(xa[p], wa.y, acc[p][1]);
        acc[p][2] = fmaf(xa[p], wa.z, acc[p][2]); acc[p][3] = fmaf(xa[p], wa.w, acc[p][3]);
        acc[p][4] = fmaf(xa[p], wb.x, acc[p][4]); acc[p][5] = fmaf(xa[p], wb.y, acc[p][5]);
        acc[p][6] = fmaf(xa[p], wb.z, acc[p][6]); acc[p][7] = fmaf(xa[p], wb.w, acc[p][7]);
      }
    }
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;
    if (bias) { ba = *reinterpret_cast<const float4*>(bias + c0); bb = *reinterpret_cast<const float4*>(bias + c1); }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const long long pix = pix0 + p0 + p;
      if (pix >= npix) break;
      int x = static_cast<int>(pix % W), y = static_cast<int>((pix / W) % H);
      long long n = pix / (static_cast<long long>(W) * H);
      TO* o = out + ((n * out_H + out_row_off + y) * W + x) * Cout;
      st4(o + c0, make_float4(acc[p][0] + ba.x, acc[p][1] + ba.y, acc[p][2] + ba.z, acc[p][3] + ba.w));
      st4(o + c1, make_float4(acc[p][4] + bb.x, acc[p][5] + bb.y, acc[p][6] + bb.z, acc[p][7] + bb.w));
    }
  }
}

// conv_out-type: NHWC input with many channels, tiny Cout (<= 8), NCHW or NHWC output.
// One warp per 4 consecutive pixels of a row; lanes split the (tap, channel) reduction with 8-wide
// vector loads, so each weight vector loaded is applied to 4 pixels.
template <typename TI, typename TO, int CO>
__global__ void __launch_bounds__(256)
conv3x3_small_cout_kernel(const TI* __restrict__ in, const float* __restrict__ w,
                          const float* __restrict__ bias, TO* __restrict__ out, int out_nchw, int n_img,
                          int H, int W, int Cin, int out_H, int out_row_off) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = (W + 3) >> 2;
  const long long gid = blockIdx.x * 8LL + warp;
  if (gid >= static_cast<long long>(n_img) * H * wq) return;
  const int xq = static_cast<int>(gid % wq);
  const int y = static_cast<int>((gid / wq) % H);
  const int n = static_cast<int>(gid / (static_cast<long long>(wq) * H));
  const int x0 = xq * 4;
  float acc[4][CO];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < CO; ++c) acc[p][c] = 0.f;
  const int cv = Cin >> 3;
  for (int i = lane; i < 9 * cv; i += 32) {
    const int tap = i / cv, c8 = (i - tap * cv) * 8;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= H) continue;
    float wv[CO][8];
#pragma unroll
    for (int c = 0; c < CO; ++c) ld8<float>(w + (static_cast<long long>(c) * 9 + tap) * Cin + c8, wv[c]);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int xx = x0 + p + kx - 1;
      if (xx < 0 || xx >= W) continue;
      float xv[8];
      ld8<TI>(in + ((static_cast<long long>(n) * H + yy) * W + xx) * Cin + c8, xv);
#pragma unroll
      for (int c = 0; c < CO; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][c] = fmaf(xv[j], wv[c][j], acc[p][c]);
    }
  }
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      float v = acc[p][c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[p][c] = v;
    }
  if (lane < 4 * CO) {
    const int p = lane / CO, c = lane - p * CO;
    const int x = x0 + p;
    if (x < W) {
      float v = 0.f;
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int cc = 0; cc < CO; ++cc)
          if (pp == p && cc == c) v = acc[pp][cc];
      v += bias ? bias[c] : 0.f;
      if (out_nchw) st_f(out + ((static_cast<long long>(n) * CO + c) * H + y) * W + x, v);
      else st_f(out + ((static_cast<long long>(n) * out_H + out_row_off + y) * W + x) * CO + c, v);
    }
  }
}

// ------------------------------------------------------------------------- SIMT attention
// One warp per query row, head_dim 64 (two dims per lane). Online softmax per key segment.
struct AttnSimtParams {
  const void* q; long long ldq; int q_off;
  int n_seg;
  const void* k[2]; long long ldk[2]; int k_off[2];
  const void* v[2]; long long ldv[2]; int v_off[2];
  int kv_len[2];
  float seg_scale[2];
  void* out; long long ldo; int out_off;
  int B, heads, n_q;
  float scale;
  int causal;
};

template <typename T>
__global__ void __launch_bounds__(256) attn_simt_kernel(const AttnSimtParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8LL + warp;  // flattened (b, h, i)
  const long long total = static_cast<long long>(p.B) * p.heads * p.n_q;
  if (row >= total) return;
  const int i = static_cast<int>(row % p.n_q);
  const int h = static_cast<int>((row / p.n_q) % p.heads);
  const int b = static_cast<int>(row / (static_cast<long long>(p.n_q) * p.heads));
  const T* Q = reinterpret_cast<const T*>(p.q) + (static_cast<long long>(b) * p.n_q + i) * p.ldq +
               p.q_off + h * 64;
  const float q0 = ld_f(Q + lane * 2) * p.scale, q1 = ld_f(Q + lane * 2 + 1) * p.scale;
  float o0 = 0.f, o1 = 0.f;
  for (int s = 0; s < p.n_seg; ++s) {
    const T* K = reinterpret_cast<const T*>(p.k[s]) +
                 static_cast<long long>(b) * p.kv_len[s] * p.ldk[s] + p.k_off[s] + h * 64;
    const T* V = reinterpret_cast<const T*>(p.v[s]) +
                 static_cast<long long>(b) * p.kv_len[s] * p.ldv[s] + p.v_off[s] + h * 64;
    float mx = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f;
    const int n_keys = p.causal ? min(p.kv_len[s], i + 1) : p.kv_len[s];
    for (int j = 0; j < n_keys; ++j) {
      const T* kr = K + static_cast<long long>(j) * p.ldk[s];
      float d = q0 * ld_f(kr + lane * 2) + q1 * ld_f(kr + lane * 2 + 1);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      float mn = fmaxf(mx, d);
      float corr = __expf(mx - mn);
      float pj = __expf(d - mn);
      const T* vr = V + static_cast<long long>(j) * p.ldv[s];
      l = l * corr + pj;
      a0 = a0 * corr + pj * ld_f(vr + lane * 2);
      a1 = a1 * corr + pj * ld_f(vr + lane * 2 + 1);
      mx = mn;
    }
    float inv = p.seg_scale[s] / l;
    o0 += a0 * inv;
    o1 += a1 * inv;
  }
  T* O = reinterpret_cast<T*>(p.out) + (static_cast<long long>(b) * p.n_q + i) * p.ldo + p.out_off +
         h * 64;
  st_f(O + lane * 2, o0);
  st_f(O + lane * 2 + 1, o1);
}

// --------------------------------------------------------------------------- small-M linear
// out[m, n] = act(sum_k x[m,k] w[n,k] + bias[n]), M <= 16.  HBM-bound on the weight read: x is staged
// in shared memory once per block, each warp streams one weight row with 16-byte loads.
constexpr int LS_ROWS = 8;  // weight rows (one per warp) per block

template <typename TW, typename TO, int MT>
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x,
                                                           const TW* __restrict__ w,
                                                           const float* __restrict__ bias,
                                                           TO* __restrict__ out, int M, int N, int K,
                                                           int act) {
  extern __shared__ float xs[];  // [M][K]
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x * 4; i < M * K; i += 256 * 4)
    *reinterpret_cast<float4*>(xs + i) = *reinterpret_cast<const float4*>(x + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * LS_ROWS + warp;
  if (n >= N) return;
  float acc[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) acc[m] = 0.f;
  const TW* wr = w + static_cast<long long>(n) * K;
  for (int k = lane * 8; k < K; k += 256) {
    float wv[8];
    ld8<TW>(wr + k, wv);
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        const float* xr = xs + m * K + k;
        float4 a = *reinterpret_cast<const float4*>(xr), b = *reinterpret_cast<const float4*>(xr + 4);
        acc[m] += a.x * wv[0] + a.y * wv[1] + a.z * wv[2] + a.w * wv[3] + b.x * wv[4] + b.y * wv[5] +
                  b.z * wv[6] + b.w * wv[7];
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    if (m < M) {
      float v = acc[m];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) {
        if (bias) v += bias[n];
        if (act == IIR_ACT_SILU) v = silu_f(v);
        else if (act == IIR_ACT_GELU) v = gelu_erf_f(v);
        else if (act == IIR_ACT_QUICK_GELU) v = quick_gelu_f(v);
        st_f(out + static_cast<long long>(m) * N + n, v);
      }
    }
  }
}

}  // namespace
}  // namespace iir

using namespace iir;
typedef h16 bf16;

extern "C" int iir_gemm_simt(const iir_gemm_args* a, void* stream) {
  IIR_REQUIRE(a != nullptr, "iir_gemm_simt: null args");
  IIR_REQUIRE(dtype_ok(a->a_dtype) && dtype_ok(a->w_dtype) && dtype_ok(a->out_dtype) && (!a->residual || dtype_ok(a->res_dtype)) && (!a->aux || dtype_ok(a->aux_dtype)), "iir_gemm_simt: unsupported dtype for this library build (fp32 or %s only)", IIR_H16 == IIR_F16 ? "fp16" : "bf16");
  IIR_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "iir_gemm_simt: empty problem");
  IIR_REQUIRE(!a->ln_stats_out && !a->ln_out16 && !a->ln_stats_in, "iir_gemm_simt: folded LayerNorm is a tcgen05-path feature (the check mode runs iir_layernorm)");
  IIR_REQUIRE(!a->gn_sums, "iir_gemm_simt: gn_sums is a tcgen05-path feature (the check mode runs iir_groupnorm)");
  IIR_REQUIRE(a->pair == IIR_PAIR_NONE || (a->bn > 0 && a->bn % 2 == 0 && a->N % a->bn == 0),
              "iir_gemm_simt: paired epilogue needs N%%bn==0");
  GemmSimtParams p;
  memset(&p, 0, sizeof(p));
  p.a = a->a; p.w = a->w; p.M = a->M; p.N = a->N; p.K = a->K; p.lda = a->lda;
  p.conv = a->conv;
  if (a->conv) {
    IIR_REQUIRE(a->conv == 3 && (a->stride == 1 || a->stride == 2), "iir_gemm_simt: 3x3 s1/s2 only");
    IIR_REQUIRE(a->K == 9 * a->Cin, "iir_gemm_simt: K must be 9*Cin");
    p.n_img = a->n_img; p.H = a->H; p.W = a->W; p.Cin = a->Cin; p.stride = a->stride; p.up2 = a->up2;
    IIR_REQUIRE(!a->conv_asym || (a->stride == 2 && a->H % 2 == 0 && a->W % 2 == 0 && !a->up2),
                "iir_gemm_simt: conv_asym (pad bottom/right only) needs stride 2 and even H, W");
    p.lead = a->conv_asym ? 0 : 1;
    p.Ho = (a->H + 2 - 3) / a->stride + 1;
    p.Wo = (a->W + 2 - 3) / a->stride + 1;
    IIR_REQUIRE(a->M == a->n_img * p.Ho * p.Wo, "iir_gemm_simt: conv M mismatch (M=%d, expect %d)",
                a->M, a->n_img * p.Ho * p.Wo);
  }
  p.bias = a->bias; p.rowvec = a->rowvec; p.ld_rowvec = a->ld_rowvec > 0 ? a->ld_rowvec : a->N;
  p.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : a->M;
  p.residual = a->residual; p.res_bf16 = a->res_dtype == IIR_H16; p.ld_res = a->ld_res;
  p.aux = a->aux; p.aux_bf16 = a->aux_dtype == IIR_H16; p.ld_aux = a->ld_aux;
  p.out = a->out; p.out_bf16 = a->out_dtype == IIR_H16; p.ld_out = a->ld_out;
  p.act = a->act; p.bn = a->bn > 0 ? a->bn : 2;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->a_dtype == IIR_F32 && a->w_dtype == IIR_F32) launch_gemm_simt<float, float>(p, a->pair, st);
  else if (a->a_dtype == IIR_H16 && a->w_dtype == IIR_H16) launch_gemm_simt<bf16, bf16>(p, a->pair, st);
  else if (a->a_dtype == IIR_F32 && a->w_dtype == IIR_H16) launch_gemm_simt<float, bf16>(p, a->pair, st);
  else launch_gemm_simt<bf16, float>(p, a->pair, st);
  count_launch();
  return check_launch("iir_gemm_simt");
}

extern "C" int iir_conv3x3_direct(const void* in, int in_dtype, int in_nchw, const float* w,
                                  const float* bias, void* out, int out_dtype, int out_nchw,
                                  int n_img, int H, int W, int Cin, int Cout, int out_H,
                                  int out_row_off, void* stream) {
  IIR_REQUIRE(in && w && out && n_img > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0,
              "iir_conv3x3_direct: bad args");
  IIR_REQUIRE(out_nchw || (out_H >= H + out_row_off && out_row_off >= 0),
              "iir_conv3x3_direct: output window out of range");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!in_nchw && Cout == 4 && Cin % 8 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
    long long warps = static_cast<long long>(n_img) * H * ((W + 3) / 4);
    int blocks8 = static_cast<int>((warps + 7) / 8);
#define GOS(TI, TO)                                                                                 \
  conv3x3_small_cout_kernel<TI, TO, 4><<<blocks8, 256, 0, st>>>(                                    \
      reinterpret_cast<const TI*>(in), w, bias, reinterpret_cast<TO*>(out), out_nchw, n_img, H, W,  \
      Cin, out_H, out_row_off)
    if (in_dtype == IIR_F32 && out_dtype == IIR_F32) GOS(float, float);
    else if (in_dtype == IIR_F32 && out_dtype == IIR_H16) GOS(float, bf16);
    else if (in_dtype == IIR_H16 && out_dtype == IIR_F32) GOS(bf16, float);
    else GOS(bf16, bf16);
#undef GOS
    count_launch();
    return check_launch("iir_conv3x3_direct");
  }
  if (!out_nchw && Cin <= 8 && Cout % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
      (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0) &&
      (size_t)(9 * Cin) * (Cout + SC_PIX) * sizeof(float) <= 200 * 1024) {
    size_t smem = (size_t)(9 * Cin) * (Cout + SC_PIX) * sizeof(float);
    long long npix = static_cast<long long>(n_img) * H * W;
    int blocksp = static_cast<int>((npix + SC_PIX - 1) / SC_PIX);
    cudaError_t e = cudaSuccess;
#define GOC(TI, TO)                                                                                   \
  do {                                                                                                \
    if (smem > 48 * 1024)                                                                             \
      e = cudaFuncSetAttribute(conv3x3_small_cin_kernel<TI, TO>,                                      \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
    if (e == cudaSuccess)                                                                             \
      conv3x3_small_cin_kernel<TI, TO><<<blocksp, 256, smem, st>>>(                                   \
          reinterpret_cast<const TI*>(in), in_nchw, w, bias, reinterpret_cast<TO*>(out), n_img, H, W, \
          Cin, Cout, out_H, out_row_off);                                                             \
  } while (0)
    if (in_dtype == IIR_F32 && out_dtype == IIR_F32) GOC(float, float);
    else if (in_dtype == IIR_F32 && out_dtype == IIR_H16) GOC(float, bf16);
    else if (in_dtype == IIR_H16 && out_dtype == IIR_F32) GOC(bf16, float);
    else GOC(bf16, bf16);
#undef GOC
    if (e != cudaSuccess) {
      set_error("iir_conv3x3_direct: %s", cudaGetErrorString(e));
      return IIR_ERR_CUDA;
    }
    count_launch();
    return check_launch("iir_conv3x3_direct");
  }
  long long total = static_cast<long long>(n_img) * H * W * Cout;
  int blocks = static_cast<int>((total + 255) / 256);
#define GO(TI, TO)                                                                               \
  conv3x3_direct_kernel<TI, TO><<<blocks, 256, 0, st>>>(                                         \
      reinterpret_cast<const TI*>(in), in_nchw, w, bias, reinterpret_cast<TO*>(out), out_nchw,   \
      n_img, H, W, Cin, Cout, out_H, out_row_off)
  if (in_dtype == IIR_F32 && out_dtype == IIR_F32) GO(float, float);
  else if (in_dtype == IIR_F32 && out_dtype == IIR_H16) GO(float, bf16);
  else if (in_dtype == IIR_H16 && out_dtype == IIR_F32) GO(bf16, float);
  else GO(bf16, bf16);
#undef GO
  count_launch();
  return check_launch("iir_conv3x3_direct");
}

extern "C" int iir_attn_simt(const iir_attn_args* a, void* stream) {
  IIR_REQUIRE(a != nullptr && a->n_seg >= 1 && a->n_seg <= 2, "iir_attn_simt: bad args");
  IIR_REQUIRE(dtype_ok(a->dtype), "iir_attn_simt: unsupported dtype for this library build (fp32 or %s only)", IIR_H16 == IIR_F16 ? "fp16" : "bf16");
  AttnSimtParams p;
  memset(&p, 0, sizeof(p));
  p.q = a->q; p.ldq = a->ldq; p.q_off = a->q_off; p.n_seg = a->n_seg;
  for (int s = 0; s < a->n_seg; ++s) {
    IIR_REQUIRE(a->kv_len[s] > 0, "iir_attn_simt: empty key segment %d", s);
    p.k[s] = a->k[s]; p.ldk[s] = a->ldk[s]; p.k_off[s] = a->k_off[s];
    p.v[s] = a->v[s]; p.ldv[s] = a->ldv[s]; p.v_off[s] = a->v_off[s];
    p.kv_len[s] = a->kv_len[s]; p.seg_scale[s] = a->seg_scale[s];
  }
  p.out = a->out; p.ldo = a->ldo; p.out_off = a->out_off;
  p.B = a->B; p.heads = a->heads; p.n_q = a->n_q; p.scale = a->softmax_scale;
  IIR_REQUIRE(!a->causal || (a->n_seg == 1 && a->kv_len[0] == a->n_q), "iir_attn_simt: causal needs one segment with kv_len == n_q");
  p.causal = a->causal;
  long long rows = static_cast<long long>(a->B) * a->heads * a->n_q;
  int blocks = static_cast<int>((rows + 7) / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (a->dtype == IIR_F32) attn_simt_kernel<float><<<blocks, 256, 0, st>>>(p);
  else attn_simt_kernel<bf16><<<blocks, 256, 0, st>>>(p);
  count_launch();
  return check_launch("iir_attn_simt");
}

extern "C" int iir_linear_small(const void* x, int x_dtype, const void* w, int w_dtype,
                                const float* bias, void* out, int out_dtype, int M, int N, int K,
                                int act, void* stream) {
  IIR_REQUIRE(x && w && out && M >= 1 && M <= 16 && N > 0 && K > 0,
              "iir_linear_small: need 1 <= M <= 16 (M=%d)", M);
  IIR_REQUIRE(x_dtype == IIR_F32, "iir_linear_small: x must be fp32");
  IIR_REQUIRE(K % 8 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(x) & 15) == 0,
              "iir_linear_small: K=%d must be a multiple of 8 and x, w 16-byte aligned", K);
  size_t smem = static_cast<size_t>(M) * K * sizeof(float);
  IIR_REQUIRE(smem <= 200 * 1024, "iir_linear_small: M*K too large for shared memory");
  int blocks = (N + LS_ROWS - 1) / LS_ROWS;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaSuccess;
#define GO2(TW, TO, MT)                                                                            \
  do {                                                                                             \
    if (smem > 48 * 1024)                                                                          \
      e = cudaFuncSetAttribute(linear_small_kernel<TW, TO, MT>,                                    \
                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
    if (e == cudaSuccess)                                                                          \
      e = launch_pdl(linear_small_kernel<TW, TO, MT>, dim3(blocks), dim3(256), smem, st,           \
          reinterpret_cast<const float*>(x), reinterpret_cast<const TW*>(w), bias,                 \
          reinterpret_cast<TO*>(out), M, N, K, act);                                               \
  } while (0)
#define GO(TW, TO)                   \
  do {                               \
    if (M <= 2) GO2(TW, TO, 2);      \
    else if (M <= 4) GO2(TW, TO, 4); \
    else GO2(TW, TO, 16);            \
  } while (0)
  if (w_dtype == IIR_F32 && out_dtype == IIR_F32) GO(float, float);
  else if (w_dtype == IIR_H16 && out_dtype == IIR_F32) GO(bf16, float);
  else if (w_dtype == IIR_F32 && out_dtype == IIR_H16) GO(float, bf16);
  else GO(bf16, bf16);
#undef GO
#undef GO2
  if (e != cudaSuccess) {
    set_error("iir_linear_small: %s", cudaGetErrorString(e));
    return IIR_ERR_CUDA;
  }
  count_launch();
  return check_launch("iir_linear_small");
}
