// Scheduler + guidance kernels: one vectorised pass each instead of the reference's ~10-15
// elementwise torch launches (schedulers/lcm_single_step_scheduler.py:455-484,
// pipelines/sdxl_instantir.py:1619-1633, diffusers DDPMScheduler.step).  fp32 latents.
// HBM roofline: lcm 3 x n x 4 B, cfg+ddpm 6 x n x 4 B.
#include "common.cuh"

namespace iir {
namespace {
typedef h16 bf16;

template <typename TE>
__global__ void __launch_bounds__(256)
lcm_step_kernel(const TE* __restrict__ eps, const float* __restrict__ x, float* __restrict__ out,
                long long n4, float sqrt_beta, float inv_sqrt_alpha, float c_skip, float c_out) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    float4 e = ld4(eps + i * 4);
    float4 s = ld4(x + i * 4);
    float4 o;
    // same operation order as the reference: (sample - sqrt(beta)*eps) / sqrt(alpha)
    o.x = c_out * ((s.x - sqrt_beta * e.x) * inv_sqrt_alpha) + c_skip * s.x;
    o.y = c_out * ((s.y - sqrt_beta * e.y) * inv_sqrt_alpha) + c_skip * s.y;
    o.z = c_out * ((s.z - sqrt_beta * e.z) * inv_sqrt_alpha) + c_skip * s.z;
    o.w = c_out * ((s.w - sqrt_beta * e.w) * inv_sqrt_alpha) + c_skip * s.w;
    st4(out + i * 4, o);
  }
}

template <typename TE>
__global__ void __launch_bounds__(256)
cfg_ddpm_kernel(const TE* __restrict__ eu, const TE* __restrict__ ec, const float* __restrict__ x,
                const float* __restrict__ noise, float* __restrict__ prev,
                float* __restrict__ pred_x0, long long n4, float g, float sqrt_beta,
                float inv_sqrt_alpha, float c_x0, float c_xt, float sigma) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    float4 u = ld4(eu + i * 4);
    float4 c = ec ? ld4(ec + i * 4) : u;
    float4 s = ld4(x + i * 4);
    float4 z = noise ? ld4(noise + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float e0 = ec ? u.x + g * (c.x - u.x) : u.x;
    float e1 = ec ? u.y + g * (c.y - u.y) : u.y;
    float e2 = ec ? u.z + g * (c.z - u.z) : u.z;
    float e3 = ec ? u.w + g * (c.w - u.w) : u.w;
    float4 x0, p;
    x0.x = (s.x - sqrt_beta * e0) * inv_sqrt_alpha;
    x0.y = (s.y - sqrt_beta * e1) * inv_sqrt_alpha;
    x0.z = (s.z - sqrt_beta * e2) * inv_sqrt_alpha;
    x0.w = (s.w - sqrt_beta * e3) * inv_sqrt_alpha;
    p.x = c_x0 * x0.x + c_xt * s.x + sigma * z.x;
    p.y = c_x0 * x0.y + c_xt * s.y + sigma * z.y;
    p.z = c_x0 * x0.z + c_xt * s.z + sigma * z.z;
    p.w = c_x0 * x0.w + c_xt * s.w + sigma * z.w;
    st4(prev + i * 4, p);
    if (pred_x0) st4(pred_x0 + i * 4, x0);
  }
}

__global__ void __launch_bounds__(256)
add_noise_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                 float* __restrict__ out, long long n4, float sa, float sb) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    float4 a = ld4(x0 + i * 4), z = ld4(noise + i * 4);
    st4(out + i * 4, make_float4(sa * a.x + sb * z.x, sa * a.y + sb * z.y, sa * a.z + sb * z.z,
                                 sa * a.w + sb * z.w));
  }
}

// DiagonalGaussianDistribution.sample (module/diffusers_vae/vae.py: mean, logvar = moments.chunk(2, dim=1);
// logvar clamped to [-30, 20]; sample = mean + exp(0.5 logvar) * noise), times `scale`; NCHW, per sample the first
// `half` elements of `moments` are the mean and the next `half` the log-variance
__global__ void __launch_bounds__(256)
gaussian_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise, float* __restrict__ out,
                       long long total, long long half, float scale) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const long long b = i / half, r = i - b * half;
    const float mean = moments[b * 2 * half + r];
    const float logvar = fminf(fmaxf(moments[b * 2 * half + half + r], -30.0f), 20.0f);
    out[i] = (mean + expf(0.5f * logvar) * (noise ? noise[i] : 0.0f)) * scale;
  }
}

// CFG combine + rescale_noise_cfg (pipelines/sdxl_instantir.py:181-192, 1619-1625; guidance_rescale > 0): one CTA per
// sample.  cfg = e_u + g (e_c - e_u); out = cfg * (rescale * std(e_c) / std(cfg) + 1 - rescale), std = torch.std over the
// sample (unbiased).  Two passes over the sample (it is read from L2 the second time); sums in fp64.
__global__ void __launch_bounds__(1024)
cfg_rescale_kernel(const float* __restrict__ eu, const float* __restrict__ ec, float* __restrict__ out, long long n_per,
                   float g, float rescale) {
  __shared__ double red[4][32];
  const float* u = eu + blockIdx.x * n_per;
  const float* c = ec + blockIdx.x * n_per;
  float* o = out + blockIdx.x * n_per;
  double s[4] = {0.0, 0.0, 0.0, 0.0};  // sum / sum of squares of e_c, then of cfg
  for (long long i = threadIdx.x; i < n_per; i += 1024) {
    const float cv = c[i], f = u[i] + g * (cv - u[i]);
    s[0] += cv; s[1] += static_cast<double>(cv) * cv; s[2] += f; s[3] += static_cast<double>(f) * f;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
    if (lane == 0) red[k][warp] = s[k];
  }
  __syncthreads();
  double t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    t[k] = 0.0;
    for (int w = 0; w < 32; ++w) t[k] += red[k][w];  // fixed order: every thread gets the same value
  }
  const double n = static_cast<double>(n_per);
  const double var_t = (t[1] - t[0] * t[0] / n) / (n - 1.0), var_c = (t[3] - t[2] * t[2] / n) / (n - 1.0);
  const float factor = rescale * static_cast<float>(sqrt(var_t > 0 ? var_t : 0.0) / sqrt(var_c)) + (1.0f - rescale);
  for (long long i = threadIdx.x; i < n_per; i += 1024) o[i] = (u[i] + g * (c[i] - u[i])) * factor;
}

// Start of one denoising step: latent_model_input = cat([latents] * n_rep) (scale_model_input is the identity for
// DDPM), the timestep as a device scalar and the per-sample cond_scale (pipelines/sdxl_instantir.py:1503-1506,
// 1538-1540) — one launch instead of two fills and n_rep copies, so that the captured graphs of the step read
// static buffers and no torch op runs between them.
__global__ void __launch_bounds__(256)
step_prologue_kernel(const float* __restrict__ latents, long long n4, int n_rep, float* __restrict__ x_in, float t,
                     float* __restrict__ t_dev, float cond_scale, float* __restrict__ cond_scale_dev, int n_cond) {
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0 && t_dev) t_dev[0] = t;
    if (cond_scale_dev)
      for (int i = threadIdx.x; i < n_cond; i += 256) cond_scale_dev[i] = cond_scale;
  }
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    const float4 v = ld4(latents + i * 4);
    for (int r = 0; r < n_rep; ++r) st4(x_in + (r * n4 + i) * 4, v);
  }
}

// adastep_restore (pipelines/sdxl_instantir.py:1636-1644 and :1538-1540 of the NEXT step), one CTA per image:
//   pred_x0_l2 = sum (preview - pred_x0)^2, previewer_l2 = sum (preview - previewer_mean)^2 (fp32 tensors, fp64 sums),
//   preview_factor = pred_x0_l2 / previewer_l2, previewer_mean <- preview,
//   cond_scale[r * n_img + b] = clamp(preview_factor, 0, next_scale) * next_keep for every CFG branch r of the next step.
__global__ void __launch_bounds__(1024)
adastep_kernel(const float* __restrict__ preview, const float* __restrict__ pred_x0, float* __restrict__ previewer_mean,
               float* __restrict__ preview_factor, float* __restrict__ cond_scale, int n_img, int n_rep, long long n_per,
               float next_scale, float next_keep) {
  __shared__ double red[2][32];
  const int b = blockIdx.x;
  const float* pv = preview + b * n_per;
  const float* x0 = pred_x0 + b * n_per;
  float* pm = previewer_mean + b * n_per;
  double s0 = 0.0, s1 = 0.0;
  for (long long i = threadIdx.x; i < n_per; i += 1024) {
    const float p = pv[i], d0 = p - x0[i], d1 = p - pm[i];
    s0 += static_cast<double>(d0) * d0;
    s1 += static_cast<double>(d1) * d1;
    pm[i] = p;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, off);
    s1 += __shfl_xor_sync(0xffffffffu, s1, off);
  }
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < 32; ++w) { t0 += red[0][w]; t1 += red[1][w]; }  // fixed order
    const float f = static_cast<float>(t0) / static_cast<float>(t1);
    preview_factor[b] = f;
    const float cs = fminf(fmaxf(f, 0.0f), next_scale) * next_keep;
    for (int r = 0; r < n_rep; ++r) cond_scale[r * n_img + b] = cs;
  }
}

int grid_for4(long long n4) {
  long long b = (n4 + 255) / 256;
  long long cap = 8LL * sm_count();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}
}  // namespace
}  // namespace iir

using namespace iir;

extern "C" int iir_lcm_step(const void* eps, int eps_dtype, const float* x, float* out, int64_t n,
                            float alpha_prod_t, float c_skip, float c_out, void* stream) {
  IIR_REQUIRE(eps && x && out && n > 0 && n % 4 == 0, "iir_lcm_step: n=%lld must be a positive multiple of 4", (long long)n);
  IIR_REQUIRE(alpha_prod_t > 0.f && alpha_prod_t <= 1.f, "iir_lcm_step: alpha_prod_t out of (0,1]");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float sb = sqrtf(1.0f - alpha_prod_t), isa = 1.0f / sqrtf(alpha_prod_t);
  if (eps_dtype == IIR_F32)
    lcm_step_kernel<float><<<grid_for4(n / 4), 256, 0, st>>>(reinterpret_cast<const float*>(eps), x, out, n / 4, sb, isa, c_skip, c_out);
  else
    lcm_step_kernel<bf16><<<grid_for4(n / 4), 256, 0, st>>>(reinterpret_cast<const bf16*>(eps), x, out, n / 4, sb, isa, c_skip, c_out);
  count_launch();
  return check_launch("iir_lcm_step");
}

extern "C" int iir_cfg_ddpm_step(const void* eps_uncond, const void* eps_cond, int eps_dtype,
                                 const float* x, const float* noise, float* prev, float* pred_x0,
                                 int64_t n, float guidance, float alpha_prod_t, float c_x0,
                                 float c_xt, float sigma, void* stream) {
  IIR_REQUIRE(eps_uncond && x && prev && n > 0 && n % 4 == 0, "iir_cfg_ddpm_step: bad args (n=%lld)", (long long)n);
  IIR_REQUIRE(alpha_prod_t > 0.f && alpha_prod_t <= 1.f, "iir_cfg_ddpm_step: alpha_prod_t out of (0,1]");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float sb = sqrtf(1.0f - alpha_prod_t), isa = 1.0f / sqrtf(alpha_prod_t);
  if (eps_dtype == IIR_F32)
    cfg_ddpm_kernel<float><<<grid_for4(n / 4), 256, 0, st>>>(
        reinterpret_cast<const float*>(eps_uncond), reinterpret_cast<const float*>(eps_cond), x,
        noise, prev, pred_x0, n / 4, guidance, sb, isa, c_x0, c_xt, sigma);
  else
    cfg_ddpm_kernel<bf16><<<grid_for4(n / 4), 256, 0, st>>>(
        reinterpret_cast<const bf16*>(eps_uncond), reinterpret_cast<const bf16*>(eps_cond), x,
        noise, prev, pred_x0, n / 4, guidance, sb, isa, c_x0, c_xt, sigma);
  count_launch();
  return check_launch("iir_cfg_ddpm_step");
}

extern "C" int iir_step_prologue(const float* latents, int64_t n, int n_rep, float* x_in, float t, float* t_dev,
                                 float cond_scale, float* cond_scale_dev, int n_cond, void* stream) {
  IIR_REQUIRE(latents && x_in && n > 0 && n % 4 == 0 && n_rep >= 1, "iir_step_prologue: bad args (n=%lld)", (long long)n);
  IIR_REQUIRE(!cond_scale_dev || n_cond > 0, "iir_step_prologue: n_cond must be positive");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  step_prologue_kernel<<<grid_for4(n / 4), 256, 0, st>>>(latents, n / 4, n_rep, x_in, t, t_dev, cond_scale, cond_scale_dev,
                                                          n_cond);
  count_launch();
  return check_launch("iir_step_prologue");
}

extern "C" int iir_adastep_update(const float* preview, const float* pred_x0, float* previewer_mean, float* preview_factor,
                                  float* cond_scale, int n_img, int n_rep, int64_t n_per, float next_scale, float next_keep,
                                  void* stream) {
  IIR_REQUIRE(preview && pred_x0 && previewer_mean && preview_factor && cond_scale && n_img > 0 && n_rep >= 1 && n_per > 0,
              "iir_adastep_update: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  adastep_kernel<<<static_cast<unsigned>(n_img), 1024, 0, st>>>(preview, pred_x0, previewer_mean, preview_factor, cond_scale, n_img,
                                                                 n_rep, static_cast<long long>(n_per), next_scale, next_keep);
  count_launch();
  return check_launch("iir_adastep_update");
}

extern "C" int iir_add_noise(const float* x0, const float* noise, float* out, int64_t n,
                             float alpha_prod_t, void* stream) {
  IIR_REQUIRE(x0 && noise && out && n > 0 && n % 4 == 0, "iir_add_noise: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  add_noise_kernel<<<grid_for4(n / 4), 256, 0, st>>>(x0, noise, out, n / 4, sqrtf(alpha_prod_t),
                                                     sqrtf(1.0f - alpha_prod_t));
  count_launch();
  return check_launch("iir_add_noise");
}

extern "C" int iir_gaussian_sample(const float* moments, const float* noise, float* out, int64_t n_samples,
                                   int64_t half, float scale, void* stream) {
  IIR_REQUIRE(moments && out && n_samples > 0 && half > 0, "iir_gaussian_sample: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = static_cast<long long>(n_samples) * half;
  gaussian_sample_kernel<<<grid_for4(total), 256, 0, st>>>(moments, noise, out, total, static_cast<long long>(half), scale);
  count_launch();
  return check_launch("iir_gaussian_sample");
}

extern "C" int iir_cfg_rescale(const float* eps_uncond, const float* eps_cond, float* out, int64_t n_samples, int64_t n_per,
                               float guidance, float rescale, void* stream) {
  IIR_REQUIRE(eps_uncond && eps_cond && out && n_samples > 0 && n_per > 1, "iir_cfg_rescale: bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cfg_rescale_kernel<<<static_cast<unsigned>(n_samples), 1024, 0, st>>>(eps_uncond, eps_cond, out, static_cast<long long>(n_per),
                                                                         guidance, rescale);
  count_launch();
  return check_launch("iir_cfg_rescale");
}
