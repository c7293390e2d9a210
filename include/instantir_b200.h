/*
 * instantir_b200.h — C ABI of the B200-native InstantIR denoising-step kernels.
 *
 * The reference (rebots-online/InstantIR) has no FFI layer: its hot path is Python calling
 * torch ops (cuDNN conv, cuBLAS GEMM, SDPA, native norms).  Each entry point below replaces the
 * torch op(s) named in its comment, cited as reference file:line.  The Python host
 * (instantir_b200/*.py) binds these through ctypes; see INTEGRATION.md for the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (torch); the library never frees it
 *   - activations are NHWC / token-major: a [B,C,H,W] reference tensor is stored as [B*H*W, C]
 *   - `stream` is a cudaStream_t passed as void*
 *   - return 0 on success, negative iir_status on failure; iir_last_error() gives the text
 *   - no exceptions cross the boundary; no CPU fallback exists: without a CUDA device every
 *     compute entry point returns IIR_ERR_CUDA
 */
#ifndef INSTANTIR_B200_H
#define INSTANTIR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IIR_ABI_VERSION 11

typedef enum {
  IIR_OK = 0,
  IIR_ERR_INVALID = -1,     /* bad argument (shape, alignment, dtype) */
  IIR_ERR_CUDA = -2,        /* CUDA runtime / driver error             */
  IIR_ERR_UNSUPPORTED = -3  /* shape outside what the sm_100a kernel handles */
} iir_status;

/* A library build supports fp32 plus exactly ONE 16-bit operand type: libinstantir_b200.so = bf16,
 * libinstantir_b200_fp16.so (-DIIR_FP16) = IEEE fp16, the reference's own inference precision
 * (infer.py:119).  iir_h16_dtype() tells which; the other 16-bit code is rejected with IIR_ERR_INVALID. */
typedef enum { IIR_F32 = 0, IIR_BF16 = 1, IIR_F16 = 2 } iir_dtype;
/* QUICK_GELU = x * sigmoid(1.702 x): the activation of CLIP-L's MLP (transformers CLIPMLP, hidden_act "quick_gelu") */
typedef enum { IIR_ACT_NONE = 0, IIR_ACT_SILU = 1, IIR_ACT_GELU = 2, IIR_ACT_QUICK_GELU = 3 } iir_act;
/* paired epilogues: weight rows are packed per `bn`-wide tile as [first half | second half]
 *   GEGLU: out = (x1 + b1) * gelu_erf(x2 + b2)      reference module/min_sdxl.py:502-510
 *   SFT  : out = h * (gamma + 1) + beta             reference module/aggregator.py:70-90   */
typedef enum { IIR_PAIR_NONE = 0, IIR_PAIR_GEGLU = 1, IIR_PAIR_SFT = 2 } iir_pair;

int iir_abi_version(void);
int iir_h16_dtype(void);
const char* iir_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
uint64_t iir_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * GEMM / implicit-GEMM convolution:  out = epilogue( A · Wᵀ )
 * Replaces nn.Linear (module/min_sdxl.py:301-307,505-523,569-573) and nn.Conv2d 3x3/1x1
 * (module/min_sdxl.py:246-260,601-618; module/aggregator.py:63-68) + the elementwise ops the
 * reference runs after them (bias, temb add :269-270, residual add :281, GEGLU, SFT).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* a;        /* linear: [M, K] row-major, leading dim lda (elements)
                           conv  : NHWC [n_img, H, W, Cin]                              */
  const void* w;        /* [N, K] row-major (K = taps*Cin, tap-major: k = (ky*3+kx)*Cin + c) */
  int a_dtype;          /* tc: must be IIR_BF16; simt: F32 or BF16                        */
  int w_dtype;
  int M, N, K;          /* conv: M = n_img*Ho*Wo                                          */
  int64_t lda;
  int conv;             /* 0 = linear, 3 = 3x3 pad 1                                      */
  int n_img, H, W, Cin; /* conv input geometry (H,W = input size AFTER optional upsample) */
  int stride;           /* conv stride 1 or 2 (tc: 1 only)                                */
  int up2;              /* simt only: input is nearest-2x upsampled on the fly            */
  const float* bias;    /* [N] or NULL (packed order for paired epilogues)                */
  const float* rowvec;  /* [n_samples, N] added to every row of a sample, or NULL         */
  int rows_per_sample;
  const void* residual; /* [M, N_out] added last, or NULL                                 */
  int res_dtype; int64_t ld_res;
  const void* aux;      /* SFT: h [M, N_out]                                              */
  int aux_dtype; int64_t ld_aux;
  void* out;            /* [M, N_out], N_out = N (or N/2 for paired epilogues)            */
  int out_dtype; int64_t ld_out;
  int act;              /* iir_act applied to (acc + bias + rowvec)                       */
  int pair;             /* iir_pair                                                       */
  int bn;               /* N tile (multiple of 32, <=256; multiple of 64 when paired)     */
  int cluster;          /* tc: 0 = automatic, 1 = one CTA per tile, 2 = CTA pair (cta_group::2,
                           256 x bn tile, each CTA stages half of the weight tile)         */
  int64_t ld_rowvec;    /* row stride of rowvec in floats (0 = N): lets rowvec be a column slice
                           of the banked time-embedding projections (iir_linear_small)    */
  /* ---- LayerNorm folded into the GEMMs around it (tc only; nn.LayerNorm of BasicTransformerBlock,
   * module/min_sdxl.py:546-556).  LN(x)·Wᵀ + b = rstd_m·(x·W'ᵀ − mean_m·colsum(W')) + b' with W' = W∘gamma,
   * b' = W·beta + b, so the GEMM that WRITES the fp32 stream x also emits a 16-bit copy of it and per-row
   * partial (sum, sum of squares), and the GEMM that CONSUMES LN(x) reads that copy and applies mean / rstd in
   * its epilogue: no LayerNorm launch, no extra pass over x.
   * producer: ln_stats_out [M, 2] int64 fixed-point accumulators (sum * 2^32, sum of squares * 2^24) that every N
   *           tile ADDS its rows' partial sums into (integer atomics: order-independent, so results stay
   *           bit-reproducible); must be zero on entry.  ln_out16 [M, N] = 16-bit copy of `out` (optional).
   *           Needs pair == NONE, residual NULL or fp32.
   * consumer: ln_stats_in (the producer's accumulators), ln_colsum [N] = sum_k W'[n,k] (packed order), ln_eps; the
   *           row mean uses K as the LN width.  ln_stats_zero (optional) [M, 2] int64 is cleared for the NEXT
   *           producer (two accumulators alternate along a chain of producer / consumer GEMMs).               */
  void* ln_stats_out;
  void* ln_out16; int64_t ld_ln_out16;
  const void* ln_stats_in;
  void* ln_stats_zero;
  const float* ln_colsum;
  float ln_eps;
  int conv_asym;        /* simt, stride 2 only: pad bottom/right only (diffusers Downsample2D(padding=0) of the VAE
                           encoder: F.pad (0,1,0,1) then a pad-0 stride-2 conv); 0 = symmetric pad 1           */
  /* ---- GroupNorm statistics from the producing epilogue (tc only, OPT-IN, ABI 11; nn.GroupNorm of ResnetBlock2D /
   * Transformer2DModel, module/min_sdxl.py:245,250,568,838).  gn_sums [n_samples, gn_groups, 2] int64 fixed-point
   * accumulators (sum * 2^24, sum of squares * 2^26) that every tile ADDS the partial sums of its final output values into
   * (integer atomics: order-independent, bit-reproducible); must be ZERO on entry (iir_memset_zero).  The consumer is
   * iir_groupnorm_apply_sums: the stand-alone statistics pass over the tensor (iir_groupnorm's first kernel) disappears.
   * Needs pair == NONE, N == gn_groups * gn_cpg with gn_cpg even, the direct-store epilogue (no 16-bit residual), and
   * rows_per_sample such that the 32 rows of a warp belong to one sample (linear: rows_per_sample % 32 == 0; conv: an
   * image of >= 32 pixels per tile, i.e. W % 8 == 0 and H >= 4).                                                          */
  void* gn_sums;
  int gn_cpg, gn_groups;
} iir_gemm_args;

/* tcgen05/TMEM/TMA kernel (bf16 operands, fp32 accumulate) */
int iir_gemm_tc(const iir_gemm_args* args, void* stream);
/* fp32 SIMT kernel: the "fp32 check mode" of the north star and the on-GPU cross-check */
int iir_gemm_simt(const iir_gemm_args* args, void* stream);

/* Direct convolution for tiny channel counts (conv_in 4->C, conv_out C->4):
 * module/min_sdxl.py:803,840; module/aggregator.py:304-306,394-396.
 * in : NCHW (in_nchw=1) or NHWC;  out: NHWC rows [out_row_off + y] of an image of out_H rows
 * (lets the Aggregator write both halves of its 2h x w canvas, module/aggregator.py:889-902),
 * or NCHW when out_nchw=1.  w: [Cout, 3, 3, Cin] fp32, bias [Cout] fp32.                  */
int iir_conv3x3_direct(const void* in, int in_dtype, int in_nchw, const float* w, const float* bias,
                       void* out, int out_dtype, int out_nchw, int n_img, int H, int W, int Cin,
                       int Cout, int out_H, int out_row_off, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention: softmax(Q Kᵀ * scale) V per head (head_dim 64), up to two independent key
 * segments summed with per-segment weights:
 *   out = seg_scale[0]*SDPA(Q,K0,V0) + seg_scale[1]*SDPA(Q,K1,V1)
 * One segment  = AttnProcessor2_0   (module/ip_adapter/attention_processor.py:394-396)
 * Two segments = TA_IPAttnProcessor2_0 text + image (same file :1165-1192)
 * Q/K/V/out are [B, n, ld] row-major bf16 (tc) with head h at column off + 64*h.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const void* q; int64_t ldq; int q_off;
  int n_seg;
  const void* k[2]; int64_t ldk[2]; int k_off[2];
  const void* v[2]; int64_t ldv[2]; int v_off[2];
  int kv_len[2];
  float seg_scale[2];
  void* out; int64_t ldo; int out_off;
  int dtype;            /* IIR_BF16 (tc) or IIR_F32/IIR_BF16 (simt), all tensors           */
  int B, heads, n_q;
  float softmax_scale;  /* 1/sqrt(64) (Resampler: also 1/8 = (64^-1/4)^2, resampler.py:71-72) */
  /* optional (tc, one segment): scratch that lets the kernel balance work at key-block granularity — a query tile cut
   * by a CTA boundary parks its partial (O, max, sum) there and the last part to arrive merges them.  At least
   * iir_attn_workspace_bytes(B, heads, n_q) bytes, 256-byte aligned, ZERO before the first use (every launch leaves its
   * ticket words zero again); launches that may run concurrently need distinct workspaces.  NULL = whole tiles per CTA. */
  void* workspace; int64_t workspace_bytes;
  /* 1 = causal self-attention (one segment, n_q == kv_len): query i attends keys 0..i — the CLIP text encoders
   * (pipelines/sdxl_instantir.py:522,580 call transformers' CLIPTextModel, whose encoder applies a causal mask)          */
  int causal;
} iir_attn_args;

int64_t iir_attn_workspace_bytes(int B, int heads, int n_q);
int iir_attn_tc(const iir_attn_args* args, void* stream);
int iir_attn_simt(const iir_attn_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Normalisation
 * ---------------------------------------------------------------------------------------- */
/* GroupNorm(32 groups) [+ SiLU] over NHWC x [n_img, HW, C]: module/min_sdxl.py:245,250,568,838.
 * `partials` is caller-owned scratch of iir_groupnorm_scratch_floats(...) floats that must be
 * ZERO-FILLED before its first use (it ends with one ticket counter per image: the last row-chunk of
 * an image to finish reduces the partial sums, in a fixed order); every call leaves the counters zero,
 * so the same scratch can be reused by later calls on the same stream.                        */
int64_t iir_groupnorm_scratch_floats(int n_img, int groups);
int iir_groupnorm(const void* x, int x_dtype, const float* gamma, const float* beta, void* out,
                  int out_dtype, int n_img, int HW, int C, int groups, float eps, int silu,
                  float* partials, void* stream);
/* The second half of GroupNorm alone (OPT-IN, ABI 11): normalise + affine [+ SiLU] with (mean, rstd) taken from the
 * fixed-point sums a producing GEMM / conv accumulated (iir_gemm_args.gn_sums) — one pass over x instead of two.    */
int iir_groupnorm_apply_sums(const void* x, int x_dtype, const float* gamma, const float* beta, const void* gn_sums,
                             void* out, int out_dtype, int n_img, int HW, int C, int groups, float eps, int silu,
                             void* stream);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (clears gn_sums arenas inside a captured forward) */
int iir_memset_zero(void* ptr, int64_t bytes, void* stream);
/* LayerNorm over the last dim, optional affine (gamma/beta) and optional adaLN modulation
 * out = LN(x) * (1 + mod[b, C:2C]) + mod[b, 0:C]  (module/ip_adapter/attention_processor.py:18-26;
 * module/min_sdxl.py:534-538).                                                              */
int iir_layernorm(const void* x, int x_dtype, const float* gamma, const float* beta,
                  const float* mod, int64_t mod_ld, int rows_per_sample, void* out, int out_dtype,
                  int rows, int C, float eps, void* stream);   /* mod_ld: row stride of mod (0 = 2C) */
/* Batched adaLN over many small tensors in ONE launch: the time-aware image K/V of every
 * cross-attention layer (attention_processor.py:1173-1178) depend only on temb and on step-invariant
 * pre-projections, so all 2 x 70 of them are produced together.  `items` is a DEVICE array;
 * item i: out_i[r, :] = LN(x_i[r, :], eps) * (1 + mod[r / rows_per_sample, off_i + C_i : off_i + 2 C_i])
 *                       + mod[r / rows_per_sample, off_i : off_i + C_i],  r < rows (same for all items) */
typedef struct {
  const float* x;       /* [rows, C] fp32 */
  void* out;            /* [rows, C] out_dtype */
  int64_t mod_off;      /* column offset of (shift | scale) in a row of mod */
  int C;                /* multiple of 4, <= 2048 */
  int pad_;
} iir_adaln_item;
int iir_adaln_batched(const iir_adaln_item* items, int n_items, int rows, int rows_per_sample,
                      const float* mod, int64_t mod_ld, float eps, int out_dtype, void* stream);
/* Row softmax: out[r, :] = softmax(x[r, :] * scale), x fp32 [rows, n] (leading dim ldx), out fp32 or 16-bit.
 * The VAE decoder's mid-block attention has ONE head of dim 512 (diffusers UNetMidBlock2D as built by
 * module/diffusers_vae/vae.py:239-249): S = Q K^T and O = P V run on the GEMM kernel, this normalises S.      */
int iir_softmax_rows(const float* x, int64_t ldx, void* out, int out_dtype, int64_t ldo, int rows, int n,
                     float scale, void* stream);


/* ------------------------------------------------------------------------------------------
 * Data movement fused with the reference's elementwise steps
 * ---------------------------------------------------------------------------------------- */
/* out[M, C1+C2] = cat( h (+ s[b]*rh), skip (+ s[b]*rs) ) along channels:
 * torch.cat of the up blocks (module/min_sdxl.py:708,745) fused with the ControlNet residual
 * injection skip_i + cond_scale*res_i (pipelines/sdxl_instantir.py:1602-1603).
 * C2 may be 0 (plain scaled add).                                                           */
int iir_concat_inject(const void* h, int h_dtype, int C1, const void* rh, int rh_dtype,
                      const void* skip, int skip_dtype, int C2, const void* rs, int rs_dtype,
                      const float* cond_scale, int rows_per_sample, void* out, int out_dtype,
                      int64_t M, void* stream);
/* nearest 2x upsample NHWC (module/min_sdxl.py:617) with dtype conversion                   */
int iir_upsample2x(const void* x, int x_dtype, void* out, int out_dtype, int n_img, int H, int W,
                   int C, void* stream);
/* 3x3 stride-2 patch gather -> [n_img*Ho*Wo, 9*C] (Downsample2D, module/min_sdxl.py:598-606); asym = 0: pad 1 on
 * every side; asym = 1: pad bottom/right only (the VAE encoder's Downsample2D(padding=0), H and W even)          */
int iir_im2col3x3_s2(const void* x, int x_dtype, void* out, int out_dtype, int n_img, int H, int W,
                     int C, int asym, void* stream);
/* strided 2-D copy/cast: out[r, c] = in[r*ld_in + c]                                        */
int iir_cast2d(const void* in, int in_dtype, int64_t ld_in, void* out, int out_dtype,
               int64_t ld_out, int64_t rows, int cols, void* stream);
/* out = silu(x) elementwise (temb nonlinearity, module/min_sdxl.py:268)                     */
int iir_silu(const void* x, int x_dtype, void* out, int out_dtype, int64_t n, void* stream);
/* out = a + b                                                                               */
int iir_add(const void* a, int a_dtype, const void* b, int b_dtype, void* out, int out_dtype,
            int64_t n, void* stream);

/* out = alpha * x  (Aggregator.forward's `conditioning_scale`, module/aggregator.py:963-964)       */
int iir_scale(const void* x, int x_dtype, void* out, int out_dtype, int64_t n, float alpha, void* stream);

/* ---- once-per-image encoders: transformers' CLIPTextModel(WithProjection) and Dinov2Model as called at
 * pipelines/sdxl_instantir.py:522,580,643-667 ---------------------------------------------------------------------- */
/* CLIPTextEmbeddings: out[r, :] = token_embedding[ids[r], :] + position_embedding[r % seq_len, :]; fp32 tables [*, dim],
 * out fp32 [n_tokens, dim] (the residual stream of the text encoder)                                                */
int iir_embed_tokens(const int64_t* ids, int n_tokens, int seq_len, const float* token_embedding, int vocab,
                     const float* position_embedding, int dim, float* out, void* stream);
/* Dinov2PatchEmbeddings as a GEMM operand: NCHW fp32 image [n_img, C, H, W] -> rows of non-overlapping patch x patch
 * windows [n_img * (H/patch) * (W/patch), ld_out] in (c, ky, kx) order (= Conv2d weight flattening), columns past
 * C*patch*patch zero-filled (ld_out = that width rounded up to a multiple of 8)                                     */
int iir_patchify(const float* img, int n_img, int C, int H, int W, int patch, void* out, int out_dtype, int ld_out,
                 void* stream);
/* Dinov2Embeddings: out[b, 0, :] = cls + pos[0]; out[b, 1 + p, :] = patches[b * P + p, :] + pos[1 + p]; all fp32      */
int iir_vit_assemble(const float* patches, const float* cls, const float* pos, float* out, int n_img, int P, int dim,
                     void* stream);

/* sinusoidal timestep embedding [cos|sin], flip_sin_to_cos=True, freq_shift=0
 * (module/min_sdxl.py:205-224): t [n] fp32 -> out [n, dim]                                  */
int iir_timestep_embedding(const float* t, int n, int dim, void* out, int out_dtype, void* stream);
/* small-M linear (M <= 16): time/add embedding MLPs, time_emb_proj, adaLN linears
 * (module/min_sdxl.py:227-239,249; attention_processor.py:14,23). w [N,K] F32/BF16          */
int iir_linear_small(const void* x, int x_dtype, const void* w, int w_dtype, const float* bias,
                     void* out, int out_dtype, int M, int N, int K, int act, void* stream);

/* ------------------------------------------------------------------------------------------
 * Scheduler / guidance kernels (fp32 latents, NCHW [B,4,h,w] flattened)
 * ---------------------------------------------------------------------------------------- */
/* Start of a denoising step (pipelines/sdxl_instantir.py:1503-1506,1538-1540): x_in = cat([latents] * n_rep)
 * (latent_model_input; scale_model_input is the identity for DDPM), *t_dev = t, cond_scale_dev[0..n_cond) =
 * cond_scale.  latents fp32 [n]; x_in fp32 [n_rep * n]; t_dev / cond_scale_dev may be NULL.                       */
int iir_step_prologue(const float* latents, int64_t n, int n_rep, float* x_in, float t, float* t_dev,
                      float cond_scale, float* cond_scale_dev, int n_cond, void* stream);
/* LCM single-step preview: schedulers/lcm_single_step_scheduler.py:421-489
 *   x0 = (x - sqrt(1-abar) eps)/sqrt(abar);  out = c_out*x0 + c_skip*x                      */
int iir_lcm_step(const void* eps, int eps_dtype, const float* x, float* out, int64_t n,
                 float alpha_prod_t, float c_skip, float c_out, void* stream);
/* CFG combine (pipelines/sdxl_instantir.py:1619-1621) + DDPM ancestral step (diffusers
 * DDPMScheduler.step):
 *   eps = eps_u + g (eps_c - eps_u); x0 = (x - sqrt(1-abar_t) eps)/sqrt(abar_t)
 *   prev = c_x0*x0 + c_xt*x + sigma*noise                                                   */
int iir_cfg_ddpm_step(const void* eps_uncond, const void* eps_cond, int eps_dtype, const float* x,
                      const float* noise, float* prev, float* pred_x0, int64_t n, float guidance,
                      float alpha_prod_t, float c_x0, float c_xt, float sigma, void* stream);
/* CFG combine followed by rescale_noise_cfg (pipelines/sdxl_instantir.py:181-192, 1619-1625; guidance_rescale > 0):
 *   cfg = eps_u + g (eps_c - eps_u);  out = cfg * (rescale * std(eps_c) / std(cfg) + 1 - rescale)
 * with torch.std (unbiased) over each sample of n_per elements; eps fp32 [n_samples, n_per].                      */
int iir_cfg_rescale(const float* eps_uncond, const float* eps_cond, float* out, int64_t n_samples, int64_t n_per,
                    float guidance, float rescale, void* stream);
/* adastep_restore (pipelines/sdxl_instantir.py:1636-1644, then :1538-1540 of the next step).  Per image b (fp32 [n_img, n_per]):
 *   preview_factor[b] = sum (preview - pred_x0)^2 / sum (preview - previewer_mean)^2;  previewer_mean <- preview;
 *   cond_scale[r * n_img + b] = clamp(preview_factor[b], 0, next_scale) * next_keep for the n_rep CFG branches of the next step */
int iir_adastep_update(const float* preview, const float* pred_x0, float* previewer_mean, float* preview_factor,
                       float* cond_scale, int n_img, int n_rep, int64_t n_per, float next_scale, float next_keep, void* stream);
/* add_noise (lcm_single_step_scheduler.py:492-513): out = sqrt(abar)*x0 + sqrt(1-abar)*noise */
int iir_add_noise(const float* x0, const float* noise, float* out, int64_t n, float alpha_prod_t,
                  void* stream);
/* DiagonalGaussianDistribution.sample of the VAE encoder (module/diffusers_vae/vae.py, used at
 * pipelines/sdxl_instantir.py:1375-1376): moments [n_samples, 2*half] = (mean | logvar) per sample,
 * out = (mean + exp(0.5*clamp(logvar, -30, 20)) * noise) * scale; noise NULL = the distribution's mode.        */
int iir_gaussian_sample(const float* moments, const float* noise, float* out, int64_t n_samples, int64_t half,
                        float scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* INSTANTIR_B200_H */
