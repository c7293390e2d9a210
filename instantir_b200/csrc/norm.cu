// GroupNorm(+SiLU) and LayerNorm(+adaLN) over NHWC / token-major activations.
// HBM-bound: every element is read twice (stats, apply) and written once; reads are coalesced
// float4 / 8-byte bf16x4 vectors, partial sums are reduced in a fixed order (deterministic).
#include "common.cuh"

namespace iir {
namespace {

constexpr int GN_MAX_CHUNKS = 256;
constexpr int GN_THREADS = 256;

// 16-byte vectors: 4 fp32 or 8 16-bit channels per thread and access
template <typename T> struct V16 { static constexpr int N = 16 / sizeof(T); };
// raw 16-byte load (kept packed while in flight: 4 registers), expanded to floats only when consumed
template <typename T>
__device__ __forceinline__ uint4 ldraw(const T* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void expand(const uint4& u, float (&v)[4], const float*) {
  v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
}
__device__ __forceinline__ void expand(const uint4& u, float (&v)[8], const h16*) {
  const h162* h = reinterpret_cast<const h162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 f = h162_to_ff(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
template <typename T, int N>
__device__ __forceinline__ void ldv(const T* p, float (&v)[N]) { expand(ldraw(p), v, p); }
template <int N>
__device__ __forceinline__ void stv(float* p, const float (&v)[N]) {
#pragma unroll
  for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
template <int N>
__device__ __forceinline__ void stv(h16* p, const float (&v)[N]) {
  static_assert(N == 4 || N == 8, "4 or 8 channels per thread");
  if constexpr (N == 8) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
    u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
  } else {
    uint2 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
}

// Thread layout shared by both stages: block = TX x TY threads, tx owns ONE 16-byte channel vector per pass
// (pass p: vector p*TX + tx), ty walks the rows of the block's chunk with stride TY, four rows in flight.
// A warp reads 512 contiguous bytes of a row; per-channel state (sums / scale, shift) lives in registers.

// ---- stage 1: per (image, row-chunk) partial sum / sum of squares for every group; the LAST chunk of an
// image to finish (ticket counter) reduces all partials in a fixed order into (mean, rstd) — no separate
// finalize launch, and still deterministic.
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const T* __restrict__ x, float* __restrict__ partials, float* __restrict__ stats,
                unsigned int* __restrict__ tickets, int HW, int C, int groups, int rows_per_chunk,
                int nchunks, double count, float eps, int TX, int TY) {
  constexpr int N = V16<T>::N;
  extern __shared__ float sm[];  // [TY][C] sums, [TY][C] squares
  __shared__ double sh_s[32][8], sh_q[32][8];
  __shared__ int s_last;
  pdl_trigger();
  pdl_wait();
  float* s_sum = sm;
  float* s_sq = sm + TY * C;
  const int chunk = blockIdx.x, img = blockIdx.y;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;  // ty >= TY: surplus thread of the last warp (idle in the data loop)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthreads = blockDim.x;
  const int cv = ty < TY ? C / N : 0;
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(HW, r0 + rows_per_chunk);
  const T* base = x + static_cast<long long>(img) * HW * C;
  for (int v = tx; v < cv; v += TX) {
    float s[N], q[N];
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] = q[j] = 0.f;
    int r = r0 + ty;
    const T* ptr = base + static_cast<long long>(r) * C + v * N;
    const long long step = static_cast<long long>(TY) * C;
    for (; r + 3 * TY < r1; r += 4 * TY, ptr += 4 * step) {  // four independent 16-byte loads in flight
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = ldraw(ptr + u * step);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float a[N];
        expand(raw[u], a, ptr);
#pragma unroll
        for (int j = 0; j < N; ++j) { s[j] += a[j]; q[j] = fmaf(a[j], a[j], q[j]); }
      }
    }
    for (; r < r1; r += TY, ptr += step) {
      float a[N];
      ldv(ptr, a);
#pragma unroll
      for (int j = 0; j < N; ++j) { s[j] += a[j]; q[j] = fmaf(a[j], a[j], q[j]); }
    }
    stv<N>(s_sum + ty * C + v * N, s);
    stv<N>(s_sq + ty * C + v * N, q);
  }
  __syncthreads();
  // one warp per group: lanes stride over the TY x cpg per-channel sums, fixed-order shuffle tree
  const int cpg = C / groups;
  const int nwarps = nthreads >> 5;
  for (int g = warp; g < groups; g += nwarps) {
    float s = 0.f, q = 0.f;
    for (int i = lane; i < TY * cpg; i += 32) {
      const int t2 = i / cpg, c = g * cpg + (i - t2 * cpg);
      s += s_sum[t2 * C + c];
      q += s_sq[t2 * C + c];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if (lane == 0) {
      float* o = partials + ((static_cast<long long>(img) * GN_MAX_CHUNKS + chunk) * groups + g) * 2;
      o[0] = s;
      o[1] = q;
    }
  }
  // ---- last chunk of this image: reduce the chunk partials
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(&tickets[img], 1u);
    s_last = (t == static_cast<unsigned int>(nchunks) - 1u);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // sub-lanes of a group each sum every SUB-th chunk, then the sub-sums are combined in a fixed order:
  // deterministic whichever chunk happens to finish last
  const int SUB = 8;
  const int gpp = nthreads / SUB;  // groups per pass
  const int g = threadIdx.x / SUB, sub = threadIdx.x % SUB;
  for (int g0 = 0; g0 < groups; g0 += gpp) {
    const int gg = g0 + g;
    const bool act = g < gpp && gg < groups && g < 32;
    double s = 0.0, q = 0.0;
    if (act) {
      for (int ch = sub; ch < nchunks; ch += SUB) {
        const float2 pp = __ldcg(reinterpret_cast<const float2*>(
            partials + ((static_cast<long long>(img) * GN_MAX_CHUNKS + ch) * groups + gg) * 2));
        s += pp.x;
        q += pp.y;
      }
      sh_s[g][sub] = s;
      sh_q[g][sub] = q;
    }
    __syncthreads();
    if (act && sub == 0) {
      s = q = 0.0;
      for (int k = 0; k < SUB; ++k) { s += sh_s[g][k]; q += sh_q[g][k]; }
      double mean = s / count;
      double var = q / count - mean * mean;
      if (var < 0.0) var = 0.0;
      float* o = stats + (static_cast<long long>(img) * groups + gg) * 2;
      o[0] = static_cast<float>(mean);
      o[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) tickets[img] = 0u;  // self-resetting: the scratch is reusable by the next call
}

// ---- stage 2: normalise, affine, optional SiLU ---------------------------------------------
// the row loop shared by gn_apply_kernel and gn_apply_sums_kernel: per-channel (scale, shift) are in shared memory
template <typename TI, typename TO>
__device__ __forceinline__ void gn_apply_rows(const TI* __restrict__ x, TO* __restrict__ out, const float* s_scale,
                                              const float* s_shift, int HW, int C, int silu, int rows_per_block, int TX,
                                              int TY) {
  constexpr int N = V16<TI>::N;
  const int img = blockIdx.y;
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int cv = ty < TY ? C / N : 0;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const TI* xb = x + static_cast<long long>(img) * HW * C;
  TO* ob = out + static_cast<long long>(img) * HW * C;
  const long long step = static_cast<long long>(TY) * C;
  for (int v = tx; v < cv; v += TX) {
    float sc[N], sh[N];
#pragma unroll
    for (int j = 0; j < N; j += 4) {
      float4 a = *reinterpret_cast<const float4*>(s_scale + v * N + j);
      float4 b = *reinterpret_cast<const float4*>(s_shift + v * N + j);
      sc[j] = a.x; sc[j + 1] = a.y; sc[j + 2] = a.z; sc[j + 3] = a.w;
      sh[j] = b.x; sh[j + 1] = b.y; sh[j + 2] = b.z; sh[j + 3] = b.w;
    }
    auto apply = [&](float (&a)[N]) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        a[j] = fmaf(a[j], sc[j], sh[j]);
        if (silu) a[j] = silu_fast(a[j]);
      }
    };
    int r = r0 + ty;
    const TI* ptr = xb + static_cast<long long>(r) * C + v * N;
    TO* optr = ob + static_cast<long long>(r) * C + v * N;
    for (; r + 3 * TY < r1; r += 4 * TY, ptr += 4 * step, optr += 4 * step) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = ldraw(ptr + u * step);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float a[N];
        expand(raw[u], a, ptr);
        apply(a);
        stv<N>(optr + u * step, a);
      }
    }
    for (; r < r1; r += TY, ptr += step, optr += step) {
      float a[N];
      ldv(ptr, a);
      apply(a);
      stv<N>(optr, a);
    }
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ stats,
                TO* __restrict__ out, int HW, int C, int groups, int silu, int rows_per_block, int TX, int TY) {
  extern __shared__ float sm[];  // [C] scale, [C] shift
  pdl_trigger();
  pdl_wait();
  float* s_scale = sm;
  float* s_shift = sm + C;
  const int img = blockIdx.y;
  const int cpg = C / groups;
  const float* s_stats = stats + static_cast<long long>(img) * groups * 2;  // (mean, rstd) pairs
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = s_stats[2 * g + 1] * (gamma ? gamma[c] : 1.0f);
    s_scale[c] = sc;
    s_shift[c] = (beta ? beta[c] : 0.0f) - s_stats[2 * g] * sc;
  }
  __syncthreads();
  gn_apply_rows<TI, TO>(x, out, s_scale, s_shift, HW, C, silu, rows_per_block, TX, TY);
}

// The same pass with (mean, rstd) derived in the prologue from the fixed-point (sum, sum of squares) a producing GEMM / conv
// epilogue accumulated per (sample, group) (iir_gemm_args.gn_sums): the statistics pass over x disappears.
template <typename TI, typename TO>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_sums_kernel(const TI* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     const long long* __restrict__ sums, TO* __restrict__ out, int HW, int C, int groups, double inv_count,
                     float eps, int silu, int rows_per_block, int TX, int TY) {
  extern __shared__ float sm[];  // [C] scale, [C] shift
  __shared__ float s_mean[64], s_rstd[64];  // groups <= 64 (checked on the host)
  pdl_trigger();
  pdl_wait();
  float* s_scale = sm;
  float* s_shift = sm + C;
  const int img = blockIdx.y;
  const int cpg = C / groups;
  const long long* s_sums = sums + static_cast<long long>(img) * groups * 2;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {  // one fp64 finalisation per group and block
    const double mean = static_cast<double>(__ldcg(s_sums + 2 * g)) * (1.0 / GN_S1_SCALE) * inv_count;
    double var = static_cast<double>(__ldcg(s_sums + 2 * g + 1)) * (1.0 / GN_S2_SCALE) * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[g] = static_cast<float>(mean);
    s_rstd[g] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float sc = s_rstd[g] * (gamma ? gamma[c] : 1.0f);
    s_scale[c] = sc;
    s_shift[c] = (beta ? beta[c] : 0.0f) - s_mean[g] * sc;
  }
  __syncthreads();
  gn_apply_rows<TI, TO>(x, out, s_scale, s_shift, HW, C, silu, rows_per_block, TX, TY);
}

// ---- LayerNorm: one warp per row, row cached in registers ----------------------------------
constexpr int LN_MAX_V = 16;  // float4 per lane -> C <= 2048

template <typename TI, typename TO, int NV>
__global__ void __launch_bounds__(256)
layernorm_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ mod, long long mod_ld,
                 int rows_per_sample, TO* __restrict__ out, int rows, int C, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = blockIdx.x * 8LL + warp;
  pdl_trigger();
  pdl_wait();
  if (row >= rows) return;
  const int cv = C >> 2;
  const TI* xr = x + row * C;
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    int i = lane + 32 * j;
    if (i < cv) {
      v[j] = ld4(xr + i * 4);
      s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    int i = lane + 32 * j;
    if (i < cv) {
      float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  const float rstd = rsqrtf(q / C + eps);
  const float* shift = nullptr;
  const float* scale = nullptr;
  if (mod) {
    const float* m = mod + (row / rows_per_sample) * mod_ld;
    shift = m;       // emb.chunk(2): shift first, then scale (attention_processor.py:24)
    scale = m + C;
  }
  TO* orow = out + row * C;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    int i = lane + 32 * j;
    if (i < cv) {
      float4 a = v[j];
      a.x = (a.x - mean) * rstd; a.y = (a.y - mean) * rstd;
      a.z = (a.z - mean) * rstd; a.w = (a.w - mean) * rstd;
      if (gamma) {
        float4 g = ld4(gamma + i * 4);
        a.x *= g.x; a.y *= g.y; a.z *= g.z; a.w *= g.w;
      }
      if (beta) {
        float4 b = ld4(beta + i * 4);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
      }
      if (mod) {
        float4 sc = ld4(scale + i * 4), sh = ld4(shift + i * 4);
        a.x = a.x * (1.f + sc.x) + sh.x; a.y = a.y * (1.f + sc.y) + sh.y;
        a.z = a.z * (1.f + sc.z) + sh.z; a.w = a.w * (1.f + sc.w) + sh.w;
      }
      st4(orow + i * 4, a);
    }
  }
}


// ---- row softmax: one CTA per row, three passes over the (L2-resident) row ---------------------------------
// VAE mid-block attention (one head of dim 512): S = Q K^T is a GEMM, this normalises its rows.
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float o = __shfl_xor_sync(0xffffffffu, v, off);
    v = is_max ? fmaxf(v, o) : v + o;
  }
  __syncthreads();  // red may still be read by the previous reduction
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < 8; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
  return r;
}

template <typename TO>
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ x, long long ldx, TO* __restrict__ out, long long ldo, int n,
                    float scale_log2) {
  __shared__ float red[8];
  pdl_trigger();
  pdl_wait();
  const float* xr = x + blockIdx.x * ldx;
  TO* orow = out + blockIdx.x * ldo;
  const int nv = n >> 2;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += 256) {
    const float4 v = ld4(xr + i * 4);
    mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  mx = block_reduce(mx, red, true);
  const float mneg = -mx * scale_log2;
  float sum = 0.f;
  for (int i = threadIdx.x; i < nv; i += 256) {
    const float4 v = ld4(xr + i * 4);
    sum += (exp2f(fmaf(v.x, scale_log2, mneg)) + exp2f(fmaf(v.y, scale_log2, mneg))) +
           (exp2f(fmaf(v.z, scale_log2, mneg)) + exp2f(fmaf(v.w, scale_log2, mneg)));
  }
  sum = block_reduce(sum, red, false);
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < nv; i += 256) {
    float4 v = ld4(xr + i * 4);
    v.x = exp2f(fmaf(v.x, scale_log2, mneg)) * inv; v.y = exp2f(fmaf(v.y, scale_log2, mneg)) * inv;
    v.z = exp2f(fmaf(v.z, scale_log2, mneg)) * inv; v.w = exp2f(fmaf(v.w, scale_log2, mneg)) * inv;
    st4(orow + i * 4, v);
  }
}

// ---- batched adaLN: one warp per row over n_items tensors of `rows` rows each ----------------
template <typename TO>
__global__ void __launch_bounds__(256)
adaln_batched_kernel(const iir_adaln_item* __restrict__ items, int n_items, int rows,
                     int rows_per_sample, const float* __restrict__ mod, long long mod_ld, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long grow = blockIdx.x * 8LL + warp;
  pdl_trigger();
  pdl_wait();
  if (grow >= static_cast<long long>(n_items) * rows) return;
  const int it = static_cast<int>(grow / rows);
  const int row = static_cast<int>(grow - static_cast<long long>(it) * rows);
  const iir_adaln_item item = items[it];
  const int C = item.C, cv = C >> 2;
  const float* xr = item.x + static_cast<long long>(row) * C;
  const float* m = mod + (row / rows_per_sample) * mod_ld + item.mod_off;  // shift | scale
  float4 v[LN_MAX_V];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_V; ++j) {
    int i = lane + 32 * j;
    if (i < cv) {
      v[j] = ld4(xr + i * 4);
      s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < LN_MAX_V; ++j) {
    int i = lane + 32 * j;
    if (i < cv) {
      float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  const float rstd = rsqrtf(q / C + eps);
  TO* orow = reinterpret_cast<TO*>(item.out) + static_cast<long long>(row) * C;
#pragma unroll
  for (int j = 0; j < LN_MAX_V; ++j) {
    int i = lane + 32 * j;
    if (i < cv) {
      float4 a = v[j];
      float4 sh = ld4(m + i * 4), sc = ld4(m + C + i * 4);
      a.x = (a.x - mean) * rstd * (1.f + sc.x) + sh.x;
      a.y = (a.y - mean) * rstd * (1.f + sc.y) + sh.y;
      a.z = (a.z - mean) * rstd * (1.f + sc.z) + sh.z;
      a.w = (a.w - mean) * rstd * (1.f + sc.w) + sh.w;
      st4(orow + i * 4, a);
    }
  }
}

}  // namespace
}  // namespace iir

using namespace iir;
typedef h16 bf16;

extern "C" int64_t iir_groupnorm_scratch_floats(int n_img, int groups) {
  // chunk partials, the finalised (mean, rstd) table, and one ticket counter per image.  The ticket words
  // must be ZERO before the first call; every call leaves them zero again.
  return static_cast<int64_t>(n_img) * GN_MAX_CHUNKS * groups * 2 + static_cast<int64_t>(n_img) * groups * 2 + n_img;
}

extern "C" int iir_groupnorm(const void* x, int x_dtype, const float* gamma, const float* beta,
                             void* out, int out_dtype, int n_img, int HW, int C, int groups,
                             float eps, int silu, float* partials, void* stream) {
  IIR_REQUIRE(x && out && partials, "iir_groupnorm: null pointer");
  IIR_REQUIRE(dtype_ok(x_dtype) && dtype_ok(out_dtype), "iir_groupnorm: unsupported dtype for this library build (fp32 or %s only)", IIR_H16 == IIR_F16 ? "fp16" : "bf16");
  IIR_REQUIRE(n_img > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 64 && C % groups == 0 &&
                  C % 4 == 0,
              "iir_groupnorm: bad shape n=%d HW=%d C=%d G=%d", n_img, HW, C, groups);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // thread layout: TX channel-vector lanes (one 16-byte vector per lane and pass), TY row lanes
  const int vecn = x_dtype == IIR_F32 ? 4 : 8;
  IIR_REQUIRE(C % vecn == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "iir_groupnorm: C=%d must be a multiple of %d and x/out 16-byte aligned", C, vecn);
  const int cv = C / vecn;
  const int passes = (cv + GN_THREADS - 1) / GN_THREADS;
  const int TX = (cv + passes - 1) / passes;
  const int TY = GN_THREADS / TX < 1 ? 1 : GN_THREADS / TX;
  const int threads = (TX * TY + 31) / 32 * 32;  // whole warps (the shuffle reductions need them)
  // rows per block: every thread gets >= 8 rows per pass when the image is big enough for ~6 blocks per SM
  auto rows_for = [&](int max_blocks) {
    int blocks = (6 * sm_count() + n_img - 1) / n_img;
    if (blocks > max_blocks) blocks = max_blocks;
    int r = (HW + blocks - 1) / blocks;
    if (r < 8 * TY) r = 8 * TY;
    if (r > HW) r = HW;
    return r;
  };
  const int rpc = rows_for(GN_MAX_CHUNKS);
  const int nchunks = (HW + rpc - 1) / rpc;
  size_t smem1 = 2 * (size_t)TY * C * sizeof(float);
  IIR_REQUIRE(smem1 <= 200 * 1024, "iir_groupnorm: C=%d too large", C);
  dim3 g1(nchunks, n_img);
  float* stats = partials + static_cast<long long>(n_img) * GN_MAX_CHUNKS * groups * 2;
  unsigned int* tickets = reinterpret_cast<unsigned int*>(stats + static_cast<long long>(n_img) * groups * 2);
  const double count = static_cast<double>(HW) * (C / groups);
  cudaError_t e = cudaSuccess;
  if (x_dtype == IIR_F32) {
    if (smem1 > 40 * 1024)
      e = cudaFuncSetAttribute(gn_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e == cudaSuccess) e = launch_pdl(gn_stats_kernel<float>, g1, dim3(threads), smem1, st, reinterpret_cast<const float*>(x), partials, stats, tickets, HW, C, groups, rpc, nchunks, count, eps, TX, TY);
  } else {
    if (smem1 > 40 * 1024)
      e = cudaFuncSetAttribute(gn_stats_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e == cudaSuccess) e = launch_pdl(gn_stats_kernel<bf16>, g1, dim3(threads), smem1, st, reinterpret_cast<const bf16*>(x), partials, stats, tickets, HW, C, groups, rpc, nchunks, count, eps, TX, TY);
  }
  if (e != cudaSuccess) { set_error("iir_groupnorm: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  count_launch();
  int rc = check_launch("iir_groupnorm(stats)");
  if (rc) return rc;
  const int rpb = rows_for(1 << 30);
  dim3 g2((HW + rpb - 1) / rpb, n_img);
#define GO(TI, TO)                                                                                 \
  e = launch_pdl(gn_apply_kernel<TI, TO>, g2, dim3(threads), 2 * (size_t)C * sizeof(float), st,                             \
      reinterpret_cast<const TI*>(x), gamma, beta, (const float*)stats, reinterpret_cast<TO*>(out), HW, C, \
      groups, silu, rpb, TX, TY)
  if (x_dtype == IIR_F32 && out_dtype == IIR_F32) GO(float, float);
  else if (x_dtype == IIR_F32 && out_dtype == IIR_H16) GO(float, bf16);
  else if (x_dtype == IIR_H16 && out_dtype == IIR_F32) GO(bf16, float);
  else GO(bf16, bf16);
#undef GO
  if (e != cudaSuccess) { set_error("iir_groupnorm: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  count_launch();
  return check_launch("iir_groupnorm(apply)");
}

extern "C" int iir_groupnorm_apply_sums(const void* x, int x_dtype, const float* gamma, const float* beta, const void* gn_sums,
                                        void* out, int out_dtype, int n_img, int HW, int C, int groups, float eps, int silu,
                                        void* stream) {
  IIR_REQUIRE(x && out && gn_sums, "iir_groupnorm_apply_sums: null pointer");
  IIR_REQUIRE(dtype_ok(x_dtype) && dtype_ok(out_dtype), "iir_groupnorm_apply_sums: unsupported dtype for this library build");
  IIR_REQUIRE(n_img > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 64 && C % groups == 0 && C % 4 == 0,
              "iir_groupnorm_apply_sums: bad shape n=%d HW=%d C=%d G=%d", n_img, HW, C, groups);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int vecn = x_dtype == IIR_F32 ? 4 : 8;
  IIR_REQUIRE(C % vecn == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(gn_sums) & 7) == 0,
              "iir_groupnorm_apply_sums: C=%d must be a multiple of %d, x/out 16-byte and gn_sums 8-byte aligned", C, vecn);
  IIR_REQUIRE(2 * (size_t)C * sizeof(float) <= 48 * 1024, "iir_groupnorm_apply_sums: C=%d too large", C);
  // the thread layout of iir_groupnorm's second kernel
  const int cv = C / vecn;
  const int passes = (cv + GN_THREADS - 1) / GN_THREADS;
  const int TX = (cv + passes - 1) / passes;
  const int TY = GN_THREADS / TX < 1 ? 1 : GN_THREADS / TX;
  const int threads = (TX * TY + 31) / 32 * 32;
  int blocks = (6 * sm_count() + n_img - 1) / n_img;
  int rpb = (HW + blocks - 1) / blocks;
  if (rpb < 8 * TY) rpb = 8 * TY;
  if (rpb > HW) rpb = HW;
  dim3 g2((HW + rpb - 1) / rpb, n_img);
  const double inv_count = 1.0 / (static_cast<double>(HW) * (C / groups));
  cudaError_t e;
#define GO(TI, TO)                                                                                                    \
  e = launch_pdl(gn_apply_sums_kernel<TI, TO>, g2, dim3(threads), 2 * (size_t)C * sizeof(float), st,                  \
                 reinterpret_cast<const TI*>(x), gamma, beta, reinterpret_cast<const long long*>(gn_sums),             \
                 reinterpret_cast<TO*>(out), HW, C, groups, inv_count, eps, silu, rpb, TX, TY)
  if (x_dtype == IIR_F32 && out_dtype == IIR_F32) GO(float, float);
  else if (x_dtype == IIR_F32 && out_dtype == IIR_H16) GO(float, bf16);
  else if (x_dtype == IIR_H16 && out_dtype == IIR_F32) GO(bf16, float);
  else GO(bf16, bf16);
#undef GO
  if (e != cudaSuccess) { set_error("iir_groupnorm_apply_sums: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  count_launch();
  return check_launch("iir_groupnorm_apply_sums");
}

extern "C" int iir_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta,
                             const float* mod, int64_t mod_ld_, int rows_per_sample, void* out,
                             int out_dtype, int rows, int C, float eps, void* stream) {
  const long long mod_ld = mod_ld_ > 0 ? mod_ld_ : 2LL * C;
  IIR_REQUIRE(!mod || (mod_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(mod) & 15) == 0),
              "iir_layernorm: mod / mod_ld must be 16-byte aligned");
  IIR_REQUIRE(x && out && rows > 0 && C > 0 && C % 4 == 0 && C <= LN_MAX_V * 128,
              "iir_layernorm: bad shape rows=%d C=%d (C%%4==0, C<=%d)", rows, C, LN_MAX_V * 128);
  IIR_REQUIRE(!mod || rows_per_sample > 0, "iir_layernorm: mod needs rows_per_sample");
  IIR_REQUIRE(dtype_ok(x_dtype) && dtype_ok(out_dtype), "iir_layernorm: unsupported dtype for this library build");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int blocks = (rows + 7) / 8;
  const int nv = (C / 4 + 31) / 32;
#define GO2(TI, TO, NV)                                                                                \
  launch_pdl(layernorm_kernel<TI, TO, NV>, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const TI*>(x), gamma, beta, mod, mod_ld, \
             rows_per_sample, reinterpret_cast<TO*>(out), rows, C, eps)
#define GO(TI, TO)                     \
  do {                                 \
    if (nv <= 2) GO2(TI, TO, 2);       \
    else if (nv <= 5) GO2(TI, TO, 5);  \
    else if (nv <= 10) GO2(TI, TO, 10);\
    else GO2(TI, TO, 16);              \
  } while (0)
  if (x_dtype == IIR_F32 && out_dtype == IIR_F32) GO(float, float);
  else if (x_dtype == IIR_F32 && out_dtype == IIR_H16) GO(float, bf16);
  else if (x_dtype == IIR_H16 && out_dtype == IIR_F32) GO(bf16, float);
  else GO(bf16, bf16);
#undef GO
#undef GO2
  count_launch();
  return check_launch("iir_layernorm");
}

extern "C" int iir_adaln_batched(const iir_adaln_item* items, int n_items, int rows, int rows_per_sample,
                                 const float* mod, int64_t mod_ld, float eps, int out_dtype, void* stream) {
  IIR_REQUIRE(items && mod && n_items > 0 && rows > 0 && rows_per_sample > 0, "iir_adaln_batched: bad arguments");
  IIR_REQUIRE(mod_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(mod) & 15) == 0,
              "iir_adaln_batched: mod / mod_ld must be 16-byte aligned");
  IIR_REQUIRE(dtype_ok(out_dtype), "iir_adaln_batched: unsupported dtype for this library build");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n_items) * rows;
  const int blocks = static_cast<int>((total + 7) / 8);
  cudaError_t e;
  if (out_dtype == IIR_F32)
    e = launch_pdl(adaln_batched_kernel<float>, dim3(blocks), dim3(256), 0, st, items, n_items, rows, rows_per_sample, mod,
                   static_cast<long long>(mod_ld), eps);
  else
    e = launch_pdl(adaln_batched_kernel<bf16>, dim3(blocks), dim3(256), 0, st, items, n_items, rows, rows_per_sample, mod,
                   static_cast<long long>(mod_ld), eps);
  if (e != cudaSuccess) { set_error("iir_adaln_batched: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  count_launch();
  return check_launch("iir_adaln_batched");
}

extern "C" int iir_softmax_rows(const float* x, int64_t ldx, void* out, int out_dtype, int64_t ldo, int rows, int n,
                                float scale, void* stream) {
  IIR_REQUIRE(x && out && rows > 0 && n > 0 && n % 4 == 0, "iir_softmax_rows: bad shape rows=%d n=%d (n%%4==0)", rows, n);
  IIR_REQUIRE(ldx % 4 == 0 && ldo % (out_dtype == IIR_F32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "iir_softmax_rows: pointers / leading dims must be 16-byte aligned");
  IIR_REQUIRE(dtype_ok(out_dtype), "iir_softmax_rows: unsupported dtype for this library build");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float sl2 = scale * 1.4426950408889634f;
  cudaError_t e;
  if (out_dtype == IIR_F32)
    e = launch_pdl(softmax_rows_kernel<float>, dim3(rows), dim3(256), 0, st, x, static_cast<long long>(ldx),
                   reinterpret_cast<float*>(out), static_cast<long long>(ldo), n, sl2);
  else
    e = launch_pdl(softmax_rows_kernel<bf16>, dim3(rows), dim3(256), 0, st, x, static_cast<long long>(ldx),
                   reinterpret_cast<bf16*>(out), static_cast<long long>(ldo), n, sl2);
  if (e != cudaSuccess) { set_error("iir_softmax_rows: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  count_launch();
  return check_launch("iir_softmax_rows");
}
