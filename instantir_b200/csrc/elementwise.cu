// Data-movement kernels fused with the reference's elementwise steps (HBM-bound, vectorised).
#include "common.cuh"

namespace iir {
namespace {

typedef h16 bf16;

__device__ __forceinline__ float4 ld4_any(const void* p, int is_bf16, long long idx) {
  return is_bf16 ? ld4(reinterpret_cast<const bf16*>(p) + idx)
                 : ld4(reinterpret_cast<const float*>(p) + idx);
}
__device__ __forceinline__ void st4_any(void* p, int is_bf16, long long idx, float4 v) {
  if (is_bf16) st4(reinterpret_cast<bf16*>(p) + idx, v);
  else st4(reinterpret_cast<float*>(p) + idx, v);
}
__device__ __forceinline__ float ld1_any(const void* p, int is_bf16, long long idx) {
  return is_bf16 ? h16_to_f(reinterpret_cast<const bf16*>(p)[idx])
                 : reinterpret_cast<const float*>(p)[idx];
}
__device__ __forceinline__ void st1_any(void* p, int is_bf16, long long idx, float v) {
  if (is_bf16) reinterpret_cast<bf16*>(p)[idx] = f_to_h16(v);
  else reinterpret_cast<float*>(p)[idx] = v;
}

// out[m, :] = cat(h[m] (+ s*rh[m]), skip[m] (+ s*rs[m]))
__global__ void __launch_bounds__(256)
concat_inject_kernel(const void* h, int h_bf, int C1, const void* rh, int rh_bf, const void* skip,
                     int skip_bf, int C2, const void* rs, int rs_bf, const float* cond_scale,
                     int rows_per_sample, void* out, int out_bf, long long M) {
  const int cv1 = C1 >> 2, cv = (C1 + C2) >> 2;
  const long long total = M * cv;
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    long long m = i / cv;
    int v = static_cast<int>(i - m * cv);
    float s = cond_scale ? cond_scale[m / rows_per_sample] : 1.0f;
    float4 a;
    if (v < cv1) {
      a = ld4_any(h, h_bf, m * C1 + v * 4);
      if (rh) {
        float4 r = ld4_any(rh, rh_bf, m * C1 + v * 4);
        a.x += s * r.x; a.y += s * r.y; a.z += s * r.z; a.w += s * r.w;
      }
    } else {
      int v2 = v - cv1;
      a = ld4_any(skip, skip_bf, m * C2 + v2 * 4);
      if (rs) {
        float4 r = ld4_any(rs, rs_bf, m * C2 + v2 * 4);
        a.x += s * r.x; a.y += s * r.y; a.z += s * r.z; a.w += s * r.w;
      }
    }
    st4_any(out, out_bf, m * (C1 + C2) + v * 4, a);
  }
}

__global__ void __launch_bounds__(256)
upsample2x_kernel(const void* x, int x_bf, void* out, int out_bf, int n_img, int H, int W, int C) {
  const int cv = C >> 2;
  const long long total = static_cast<long long>(n_img) * (2 * H) * (2 * W) * cv;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    int v = static_cast<int>(i % cv);
    long long pix = i / cv;
    int ox = static_cast<int>(pix % (2 * W));
    int oy = static_cast<int>((pix / (2 * W)) % (2 * H));
    long long n = pix / (4LL * W * H);
    float4 a = ld4_any(x, x_bf, ((n * H + (oy >> 1)) * W + (ox >> 1)) * C + v * 4);
    st4_any(out, out_bf, pix * C + v * 4, a);
  }
}

__global__ void __launch_bounds__(256)
im2col3x3_s2_kernel(const void* x, int x_bf, void* out, int out_bf, int n_img, int H, int W, int C,
                    int Ho, int Wo, int lead) {
  const int cv = C >> 2;
  const long long total = static_cast<long long>(n_img) * Ho * Wo * 9 * cv;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    int v = static_cast<int>(i % cv);
    long long t = i / cv;
    int tap = static_cast<int>(t % 9);
    long long pix = t / 9;
    int ox = static_cast<int>(pix % Wo);
    int oy = static_cast<int>((pix / Wo) % Ho);
    long long n = pix / (static_cast<long long>(Wo) * Ho);
    int ky = tap / 3, kx = tap - ky * 3;
    int y = oy * 2 + ky - lead, xx = ox * 2 + kx - lead;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y >= 0 && y < H && xx >= 0 && xx < W) a = ld4_any(x, x_bf, ((n * H + y) * W + xx) * C + v * 4);
    st4_any(out, out_bf, (pix * 9 + tap) * C + v * 4, a);
  }
}

__global__ void __launch_bounds__(256)
cast2d_kernel(const void* in, int in_bf, long long ld_in, void* out, int out_bf, long long ld_out,
              long long rows, int cols) {
  const int cv = cols >> 2;
  const long long total = rows * cv;
  pdl_trigger();
  pdl_wait();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    long long r = i / cv;
    int v = static_cast<int>(i - r * cv);
    st4_any(out, out_bf, r * ld_out + v * 4, ld4_any(in, in_bf, r * ld_in + v * 4));
  }
}

__global__ void __launch_bounds__(256)
silu_kernel(const void* x, int x_bf, void* out, int out_bf, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    st1_any(out, out_bf, i, silu_f(ld1_any(x, x_bf, i)));
}

__global__ void __launch_bounds__(256)
add_kernel(const void* a, int a_bf, const void* b, int b_bf, void* out, int out_bf, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    st1_any(out, out_bf, i, ld1_any(a, a_bf, i) + ld1_any(b, b_bf, i));
}

__global__ void __launch_bounds__(256)
scale_kernel(const void* x, int x_bf, void* out, int out_bf, long long n, float alpha) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    st1_any(out, out_bf, i, alpha * ld1_any(x, x_bf, i));
}

// CLIPTextEmbeddings (transformers modeling_clip.py): token + position embedding lookup, fp32
__global__ void __launch_bounds__(256)
embed_tokens_kernel(const long long* __restrict__ ids, int n_tokens, int seq_len, const float* __restrict__ tok, int vocab,
                    const float* __restrict__ pos, int dim, float* __restrict__ out) {
  const int dv = dim >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < static_cast<long long>(n_tokens) * dv; i += gridDim.x * 256LL) {
    const int r = static_cast<int>(i / dv), c = static_cast<int>(i - static_cast<long long>(r) * dv) * 4;
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const float4 a = ld4(tok + id * dim + c), b = ld4(pos + static_cast<long long>(r % seq_len) * dim + c);
    st4(out + static_cast<long long>(r) * dim + c, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
  }
}

// Dinov2PatchEmbeddings' Conv2d(C, D, patch, stride = patch) as a GEMM: one output row per patch, (c, ky, kx) order
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, int n_img, int C, int H, int W, int patch, void* out, int out_bf, int ld_out) {
  const int gh = H / patch, gw = W / patch;
  const long long total = static_cast<long long>(n_img) * gh * gw * ld_out;
  const int kk = C * patch * patch;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int col = static_cast<int>(i % ld_out);
    const long long row = i / ld_out;
    float v = 0.f;
    if (col < kk) {
      const int c = col / (patch * patch), r2 = col - c * patch * patch, ky = r2 / patch, kx = r2 - ky * patch;
      const int px = static_cast<int>(row % gw), py = static_cast<int>((row / gw) % gh), n = static_cast<int>(row / (static_cast<long long>(gw) * gh));
      v = img[((static_cast<long long>(n) * C + c) * H + py * patch + ky) * W + px * patch + kx];
    }
    st1_any(out, out_bf, i, v);
  }
}

// Dinov2Embeddings: [cls | patch embeddings] + (interpolated) position embeddings
__global__ void __launch_bounds__(256)
vit_assemble_kernel(const float* __restrict__ patches, const float* __restrict__ cls, const float* __restrict__ pos,
                    float* __restrict__ out, int n_img, int P, int dim) {
  const int dv = dim >> 2;
  const long long total = static_cast<long long>(n_img) * (P + 1) * dv;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int c = static_cast<int>(i % dv) * 4;
    const long long r = i / dv;
    const int t = static_cast<int>(r % (P + 1)), b = static_cast<int>(r / (P + 1));
    const float4 a = t == 0 ? ld4(cls + c) : ld4(patches + (static_cast<long long>(b) * P + t - 1) * dim + c);
    const float4 q = ld4(pos + static_cast<long long>(t) * dim + c);
    st4(out + r * dim + c, make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w));
  }
}

// [cos | sin] of t * exp(-ln(10000) * k / half), fp32 like the reference (min_sdxl.py:205-224)
__global__ void timestep_embedding_kernel(const float* t, int n, int dim, void* out, int out_bf) {
  const int half = dim >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * half; i += gridDim.x * blockDim.x) {
    int r = i / half, k = i - r * half;
    float freq = expf(-logf(10000.0f) * static_cast<float>(k) / static_cast<float>(half));
    float arg = t[r] * freq;
    st1_any(out, out_bf, static_cast<long long>(r) * dim + k, cosf(arg));
    st1_any(out, out_bf, static_cast<long long>(r) * dim + half + k, sinf(arg));
  }
}

int grid_for(long long work_items) {
  long long b = (work_items + 255) / 256;
  long long cap = 8LL * sm_count();  // grid-stride: 8 resident CTAs of 256 threads per SM
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace
}  // namespace iir

using namespace iir;

extern "C" int iir_concat_inject(const void* h, int h_dtype, int C1, const void* rh, int rh_dtype,
                                 const void* skip, int skip_dtype, int C2, const void* rs,
                                 int rs_dtype, const float* cond_scale, int rows_per_sample,
                                 void* out, int out_dtype, int64_t M, void* stream) {
  IIR_REQUIRE(h && out && C1 > 0 && C1 % 4 == 0 && C2 >= 0 && C2 % 4 == 0 && M > 0,
              "iir_concat_inject: bad shape C1=%d C2=%d", C1, C2);
  IIR_REQUIRE(C2 == 0 || skip, "iir_concat_inject: skip missing");
  IIR_REQUIRE(dtype_ok(h_dtype) && dtype_ok(out_dtype) && (!skip || dtype_ok(skip_dtype)) && (!rh || dtype_ok(rh_dtype)) && (!rs || dtype_ok(rs_dtype)),
              "iir_concat_inject: unsupported dtype for this library build");
  IIR_REQUIRE(!cond_scale || rows_per_sample > 0, "iir_concat_inject: rows_per_sample missing");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long work = M * ((C1 + C2) / 4);
  launch_pdl(concat_inject_kernel, dim3(grid_for(work)), dim3(256), 0, st,
      h, (int)(h_dtype == IIR_H16), C1, rh, (int)(rh_dtype == IIR_H16), skip, (int)(skip_dtype == IIR_H16), C2, rs,
      (int)(rs_dtype == IIR_H16), cond_scale, rows_per_sample > 0 ? rows_per_sample : 1, out,
      (int)(out_dtype == IIR_H16), (long long)M);
  count_launch();
  return check_launch("iir_concat_inject");
}

extern "C" int iir_upsample2x(const void* x, int x_dtype, void* out, int out_dtype, int n_img, int H,
                              int W, int C, void* stream) {
  IIR_REQUIRE(x && out && n_img > 0 && H > 0 && W > 0 && C % 4 == 0, "iir_upsample2x: bad shape");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long work = 4LL * n_img * H * W * (C / 4);
  upsample2x_kernel<<<grid_for(work), 256, 0, st>>>(x, x_dtype == IIR_H16, out,
                                                    out_dtype == IIR_H16, n_img, H, W, C);
  count_launch();
  return check_launch("iir_upsample2x");
}

extern "C" int iir_im2col3x3_s2(const void* x, int x_dtype, void* out, int out_dtype, int n_img,
                                int H, int W, int C, int asym, void* stream) {
  IIR_REQUIRE(x && out && n_img > 0 && H > 0 && W > 0 && C % 4 == 0, "iir_im2col3x3_s2: bad shape");
  IIR_REQUIRE(!asym || (H % 2 == 0 && W % 2 == 0), "iir_im2col3x3_s2: bottom/right-only padding needs even H, W");
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  long long work = 9LL * n_img * Ho * Wo * (C / 4);
  im2col3x3_s2_kernel<<<grid_for(work), 256, 0, st>>>(x, x_dtype == IIR_H16, out,
                                                      out_dtype == IIR_H16, n_img, H, W, C, Ho, Wo, asym ? 0 : 1);
  count_launch();
  return check_launch("iir_im2col3x3_s2");
}

extern "C" int iir_cast2d(const void* in, int in_dtype, int64_t ld_in, void* out, int out_dtype,
                          int64_t ld_out, int64_t rows, int cols, void* stream) {
  IIR_REQUIRE(dtype_ok(in_dtype) && dtype_ok(out_dtype), "iir_cast2d: unsupported dtype for this library build");
  IIR_REQUIRE(in && out && rows > 0 && cols > 0 && cols % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0,
              "iir_cast2d: bad shape rows=%lld cols=%d", (long long)rows, cols);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  launch_pdl(cast2d_kernel, dim3(grid_for(rows * (cols / 4))), dim3(256), 0, st, in, (int)(in_dtype == IIR_H16),
             (long long)ld_in, out, (int)(out_dtype == IIR_H16), (long long)ld_out, (long long)rows, cols);
  count_launch();
  return check_launch("iir_cast2d");
}

extern "C" int iir_silu(const void* x, int x_dtype, void* out, int out_dtype, int64_t n, void* stream) {
  IIR_REQUIRE(x && out && n > 0, "iir_silu: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  silu_kernel<<<grid_for(n), 256, 0, st>>>(x, x_dtype == IIR_H16, out, out_dtype == IIR_H16, n);
  count_launch();
  return check_launch("iir_silu");
}

extern "C" int iir_add(const void* a, int a_dtype, const void* b, int b_dtype, void* out,
                       int out_dtype, int64_t n, void* stream) {
  IIR_REQUIRE(a && b && out && n > 0, "iir_add: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  add_kernel<<<grid_for(n), 256, 0, st>>>(a, a_dtype == IIR_H16, b, b_dtype == IIR_H16, out,
                                          out_dtype == IIR_H16, n);
  count_launch();
  return check_launch("iir_add");
}

extern "C" int iir_scale(const void* x, int x_dtype, void* out, int out_dtype, int64_t n, float alpha, void* stream) {
  IIR_REQUIRE(x && out && n > 0, "iir_scale: bad args");
  IIR_REQUIRE(dtype_ok(x_dtype) && dtype_ok(out_dtype), "iir_scale: unsupported dtype for this library build");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  scale_kernel<<<grid_for(n), 256, 0, st>>>(x, x_dtype == IIR_H16, out, out_dtype == IIR_H16, n, alpha);
  count_launch();
  return check_launch("iir_scale");
}

extern "C" int iir_embed_tokens(const int64_t* ids, int n_tokens, int seq_len, const float* token_embedding, int vocab,
                                const float* position_embedding, int dim, float* out, void* stream) {
  IIR_REQUIRE(ids && token_embedding && position_embedding && out && n_tokens > 0 && seq_len > 0 && vocab > 0 && dim > 0 && dim % 4 == 0,
              "iir_embed_tokens: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  embed_tokens_kernel<<<grid_for(static_cast<long long>(n_tokens) * (dim / 4)), 256, 0, st>>>(
      reinterpret_cast<const long long*>(ids), n_tokens, seq_len, token_embedding, vocab, position_embedding, dim, out);
  count_launch();
  return check_launch("iir_embed_tokens");
}

extern "C" int iir_patchify(const float* img, int n_img, int C, int H, int W, int patch, void* out, int out_dtype, int ld_out,
                            void* stream) {
  IIR_REQUIRE(img && out && n_img > 0 && C > 0 && patch > 0 && H > 0 && W > 0 && H % patch == 0 && W % patch == 0,
              "iir_patchify: H=%d, W=%d must be multiples of patch=%d", H, W, patch);
  IIR_REQUIRE(ld_out >= C * patch * patch && dtype_ok(out_dtype), "iir_patchify: ld_out too small or unsupported dtype");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  patchify_kernel<<<grid_for(static_cast<long long>(n_img) * (H / patch) * (W / patch) * ld_out), 256, 0, st>>>(
      img, n_img, C, H, W, patch, out, out_dtype == IIR_H16, ld_out);
  count_launch();
  return check_launch("iir_patchify");
}

extern "C" int iir_vit_assemble(const float* patches, const float* cls, const float* pos, float* out, int n_img, int P, int dim,
                                void* stream) {
  IIR_REQUIRE(patches && cls && pos && out && n_img > 0 && P > 0 && dim > 0 && dim % 4 == 0, "iir_vit_assemble: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  vit_assemble_kernel<<<grid_for(static_cast<long long>(n_img) * (P + 1) * (dim / 4)), 256, 0, st>>>(patches, cls, pos, out, n_img, P, dim);
  count_launch();
  return check_launch("iir_vit_assemble");
}

extern "C" int iir_timestep_embedding(const float* t, int n, int dim, void* out, int out_dtype,
                                      void* stream) {
  IIR_REQUIRE(t && out && n > 0 && dim > 0 && dim % 2 == 0, "iir_timestep_embedding: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int work = n * dim / 2;
  timestep_embedding_kernel<<<(work + 127) / 128, 128, 0, st>>>(t, n, dim, out, out_dtype == IIR_H16);
  count_launch();
  return check_launch("iir_timestep_embedding");
}
