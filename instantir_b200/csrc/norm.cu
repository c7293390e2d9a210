// GroupNorm(+SiLU) and LayerNorm(+adaLN) over NHWC / token-major activations.
// HBM-bound: every element is read twice (stats, apply) and written once; reads are coalesced
// float4 / 8-byte bf16x4 vectors, partial sums are reduced in a fixed order (deterministic).
#include "common.cuh"

namespace iir {
namespace {

constexpr int GN_MAX_CHUNKS = 256;
constexpr int GN_THREADS = 256;

// ---- stage 1: per (image, row-chunk) partial sum / sum of squares for every group ------------
// grid (chunks, n_img); thread t: vector lane tv = t%64 walks channel vectors, row lane tr = t/64.
template <typename T>
__global__ void __launch_bounds__(GN_THREADS)
gn_stats_kernel(const T* __restrict__ x, float* __restrict__ partials, int HW, int C, int groups,
                int rows_per_chunk) {
  extern __shared__ float sm[];  // [4][C] sums, [4][C] squares
  pdl_trigger();
  pdl_wait();
  float* s_sum = sm;
  float* s_sq = sm + 4 * C;
  const int chunk = blockIdx.x, img = blockIdx.y;
  const int tv = threadIdx.x & 63, tr = threadIdx.x >> 6;
  const int cv = C >> 2;
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(HW, r0 + rows_per_chunk);
  const T* base = x + static_cast<long long>(img) * HW * C;
  for (int v = tv; v < cv; v += 64) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int r = r0 + tr; r < r1; r += 4) {
      float4 a = ld4(base + static_cast<long long>(r) * C + v * 4);
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      q.x += a.x * a.x; q.y += a.y * a.y; q.z += a.z * a.z; q.w += a.w * a.w;
    }
    float* ps = s_sum + tr * C + v * 4;
    float* pq = s_sq + tr * C + v * 4;
    ps[0] = s.x; ps[1] = s.y; ps[2] = s.z; ps[3] = s.w;
    pq[0] = q.x; pq[1] = q.y; pq[2] = q.z; pq[3] = q.w;
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += GN_THREADS) {
    double s = 0.0, q = 0.0;
    for (int tr2 = 0; tr2 < 4; ++tr2)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        s += s_sum[tr2 * C + c];
        q += s_sq[tr2 * C + c];
      }
    float* o = partials + ((static_cast<long long>(img) * GN_MAX_CHUNKS + chunk) * groups + g) * 2;
    o[0] = static_cast<float>(s);
    o[1] = static_cast<float>(q);
  }
}

// ---- stage 1b: fixed-order reduction of the chunk partials -> (mean, rstd) per (image, group) ----
__global__ void __launch_bounds__(256)
gn_finalize_kernel(const float* __restrict__ partials, float* __restrict__ stats, int groups,
                   int nchunks, double count, float eps) {
  // 8 sub-lanes per group each sum every 8th chunk, then the 8 sub-sums are combined in a fixed
  // order: deterministic, and only nchunks/8 dependent adds deep.
  __shared__ double sh_s[64][8], sh_q[64][8];
  pdl_trigger();
  pdl_wait();
  const int img = blockIdx.x;
  const int g = threadIdx.x >> 3, sub = threadIdx.x & 7;
  for (int g0 = 0; g0 < groups; g0 += 32) {
    const int gg = g0 + g;
    double s = 0.0, q = 0.0;
    if (gg < groups) {
      for (int ch = sub; ch < nchunks; ch += 8) {
        const float2 pp = *reinterpret_cast<const float2*>(
            partials + ((static_cast<long long>(img) * GN_MAX_CHUNKS + ch) * groups + gg) * 2);
        s += pp.x;
        q += pp.y;
      }
      sh_s[g][sub] = s;
      sh_q[g][sub] = q;
    }
    __syncthreads();
    if (gg < groups && sub == 0) {
      s = q = 0.0;
      for (int k = 0; k < 8; ++k) { s += sh_s[g][k]; q += sh_q[g][k]; }
      double mean = s / count;
      double var = q / count - mean * mean;
      if (var < 0.0) var = 0.0;
      float* o = stats + (static_cast<long long>(img) * groups + gg) * 2;
      o[0] = static_cast<float>(mean);
      o[1] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    }
    __syncthreads();
  }
}

// ---- stage 2: normalise, affine, optional SiLU ---------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(const TI* __restrict__ x, const float* __restrict__ gamma,
                const float* __restrict__ beta, const float* __restrict__ partials,
                TO* __restrict__ out, int HW, int C, int groups, int nchunks, float eps, int silu,
                int rows_per_block) {
  extern __shared__ float sm[];  // [C] scale, [C] shift
  pdl_trigger();
  pdl_wait();
  float* s_scale = sm;
  float* s_shift = sm + C;
  const int img = blockIdx.y;
  const int cpg = C / groups;
  const float* s_stats = partials + static_cast<long long>(img) * groups * 2;  // (mean, rstd) pairs
  for (int c = threadIdx.x; c < C; c += GN_THREADS) {
    int g = c / cpg;
    float sc = s_stats[2 * g + 1] * (gamma ? gamma[c] : 1.0f);
    s_scale[c] = sc;
    s_shift[c] = (beta ? beta[c] : 0.0f) - s_stats[2 * g] * sc;
  }
  __syncthreads();
  const int cv = C >> 2;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const long long total = static_cast<long long>(r1 - r0) * cv;
  const TI* xb = x + (static_cast<long long>(img) * HW + r0) * C;
  TO* ob = out + (static_cast<long long>(img) * HW + r0) * C;
  for (long long i = threadIdx.x; i < total; i += GN_THREADS) {
    int v = static_cast<int>(i % cv);
    float4 a = ld4(xb + i * 4);
    const float* sc = s_scale + v * 4;
    const float* sh = s_shift + v * 4;
    a.x = a.x * sc[0] + sh[0];
    a.y = a.y * sc[1] + sh[1];
    a.z = a.z * sc[2] + sh[2];
    a.w = a.w * sc[3] + sh[3];
    if (silu) { a.x = silu_f(a.x); a.y = silu_f(a.y); a.z = silu_f(a.z); a.w = silu_f(a.w); }
    st4(ob + i * 4, a);
  }
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
  // chunk partials followed by the finalised (mean, rstd) table
  return static_cast<int64_t>(n_img) * GN_MAX_CHUNKS * groups * 2 + static_cast<int64_t>(n_img) * groups * 2;
}

extern "C" int iir_groupnorm(const void* x, int x_dtype, const float* gamma, const float* beta,
                             void* out, int out_dtype, int n_img, int HW, int C, int groups,
                             float eps, int silu, float* partials, void* stream) {
  IIR_REQUIRE(x && out && partials, "iir_groupnorm: null pointer");
  IIR_REQUIRE(dtype_ok(x_dtype) && dtype_ok(out_dtype), "iir_groupnorm: unsupported dtype for this library build (fp32 or %s only)", IIR_H16 == IIR_F16 ? "fp16" : "bf16");
  IIR_REQUIRE(n_img > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 64 && C % groups == 0 &&
                  C % 4 == 0,
              "iir_groupnorm: bad shape n=%d HW=%d C=%d G=%d", n_img, HW, C, groups);
  IIR_REQUIRE(8 * C * sizeof(float) <= 200 * 1024, "iir_groupnorm: C=%d too large", C);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // enough chunks to fill the machine, at most GN_MAX_CHUNKS, at least 16 rows each
  int want = (4 * sm_count() + n_img - 1) / n_img;
  int nchunks = want < 1 ? 1 : (want > GN_MAX_CHUNKS ? GN_MAX_CHUNKS : want);
  int rpc = (HW + nchunks - 1) / nchunks;
  if (rpc < 8) rpc = 8;
  nchunks = (HW + rpc - 1) / rpc;
  size_t smem1 = 8 * (size_t)C * sizeof(float);
  dim3 g1(nchunks, n_img);
  cudaError_t e = cudaSuccess;
  if (x_dtype == IIR_F32) {
    if (smem1 > 48 * 1024)
      e = cudaFuncSetAttribute(gn_stats_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e == cudaSuccess) e = launch_pdl(gn_stats_kernel<float>, g1, dim3(GN_THREADS), smem1, st, reinterpret_cast<const float*>(x), partials, HW, C, groups, rpc);
  } else {
    if (smem1 > 48 * 1024)
      e = cudaFuncSetAttribute(gn_stats_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
    if (e == cudaSuccess) e = launch_pdl(gn_stats_kernel<bf16>, g1, dim3(GN_THREADS), smem1, st, reinterpret_cast<const bf16*>(x), partials, HW, C, groups, rpc);
  }
  if (e != cudaSuccess) { set_error("iir_groupnorm: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  count_launch();
  int rc = check_launch("iir_groupnorm(stats)");
  if (rc) return rc;
  float* stats = partials + static_cast<long long>(n_img) * GN_MAX_CHUNKS * groups * 2;
  e = launch_pdl(gn_finalize_kernel, dim3(n_img), dim3(256), 0, st, (const float*)partials, stats, groups, nchunks,
                 static_cast<double>(HW) * (C / groups), eps);
  if (e != cudaSuccess) { set_error("iir_groupnorm: %s", cudaGetErrorString(e)); return IIR_ERR_CUDA; }
  count_launch();
  rc = check_launch("iir_groupnorm(finalize)");
  if (rc) return rc;
  // apply: blocks of >= 8 rows, ~8 CTAs per SM
  int blocks_per_img = (8 * sm_count() + n_img - 1) / n_img;
  int rpb = (HW + blocks_per_img - 1) / blocks_per_img;
  if (rpb < 8) rpb = 8;
  blocks_per_img = (HW + rpb - 1) / rpb;
  dim3 g2(blocks_per_img, n_img);
  size_t smem2 = 2 * (size_t)C * sizeof(float);
#define GO(TI, TO)                                                                                 \
  e = launch_pdl(gn_apply_kernel<TI, TO>, g2, dim3(GN_THREADS), smem2, st,                        \
      reinterpret_cast<const TI*>(x), gamma, beta, (const float*)stats, reinterpret_cast<TO*>(out), HW, C, \
      groups, nchunks, eps, silu, rpb)
  if (x_dtype == IIR_F32 && out_dtype == IIR_F32) GO(float, float);
  else if (x_dtype == IIR_F32 && out_dtype == IIR_H16) GO(float, bf16);
  else if (x_dtype == IIR_H16 && out_dtype == IIR_F32) GO(bf16, float);
  else GO(bf16, bf16);
#undef GO
  count_launch();
  return check_launch("iir_groupnorm(apply)");
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
