// Shared helpers for the sm_100a kernels: error plumbing, dtype load/store, and the raw PTX
// wrappers (mbarrier, TMA, tcgen05/TMEM) the tensor-core kernels are built from.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/instantir_b200.h"

// One 16-bit operand type per library build: bf16 (libinstantir_b200.so, the north star's precision) or,
// with -DIIR_FP16, IEEE fp16 (libinstantir_b200_fp16.so — the reference's own inference precision,
// infer.py:119; 8x finer mantissa, same tcgen05 kind::f16 rate).
#if defined(IIR_FP16)
typedef __half h16;
typedef __half2 h162;
#define IIR_H16 IIR_F16
#define IIR_H16_TMA CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define IIR_UMMA_FMT 0u
#else
typedef __nv_bfloat16 h16;
typedef __nv_bfloat162 h162;
#define IIR_H16 IIR_BF16
#define IIR_H16_TMA CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define IIR_UMMA_FMT 1u
#endif

namespace iir {

// GroupNorm statistics accumulated by the producing GEMM / conv epilogue (iir_gemm_args.gn_sums, opt-in): int64 fixed point,
// sum * 2^24 and sum of squares * 2^26 per (sample, group).  Headroom: |sum| < 2^39 and sum of squares < 2^37 per group
// (655 360 elements per group at 128² x 40 channels: rms up to ~450); resolution 6e-8 / 1.5e-8 per 32 x 32 partial, i.e.
// still 0.2 % of a partial whose values are ~1e-4.
constexpr float GN_S1_SCALE = 16777216.f;  // 2^24
constexpr float GN_S2_SCALE = 67108864.f;  // 2^26

__host__ __device__ __forceinline__ bool dtype_ok(int d) { return d == IIR_F32 || d == IIR_H16; }
__device__ __forceinline__ float h16_to_f(h16 v) {
#if defined(IIR_FP16)
  return __half2float(v);
#else
  return __bfloat162float(v);
#endif
}
__device__ __forceinline__ h16 f_to_h16(float v) {
#if defined(IIR_FP16)
  return __float2half_rn(v);
#else
  return __float2bfloat16_rn(v);
#endif
}
__device__ __forceinline__ h162 ff_to_h162(float a, float b) {
#if defined(IIR_FP16)
  return __floats2half2_rn(a, b);
#else
  return __floats2bfloat162_rn(a, b);
#endif
}
__device__ __forceinline__ float2 h162_to_ff(h162 v) {
#if defined(IIR_FP16)
  return __half22float2(v);
#else
  return __bfloat1622float2(v);
#endif
}

// ---------------------------------------------------------------------------- host side
void set_error(const char* fmt, ...);
void count_launch();
int check_launch(const char* what);  // cudaGetLastError -> status
CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      CUtensorMapSwizzle swz);
int sm_count();
bool pdl_enabled();

// Launch with Programmatic Dependent Launch allowed: the kernel may start while its stream predecessor
// is still draining; it must call pdl_wait() before touching memory the predecessor wrote.  Every kernel
// of this library calls pdl_trigger() at entry so that ITS successor can be scheduled early.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_cluster_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                      cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t st, Args&&... args) {
  return launch_cluster_pdl(kernel, grid, block, smem, st, 1, static_cast<Args&&>(args)...);
}

#define IIR_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      iir::set_error(__VA_ARGS__);      \
      return IIR_ERR_INVALID;           \
    }                                   \
  } while (0)

// ---------------------------------------------------------------------------- device side
template <typename T>
__device__ __forceinline__ float ld_f(const T* p);
template <>
__device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ld_f<h16>(const h16* p) {
  return h16_to_f(*p);
}
template <typename T>
__device__ __forceinline__ void st_f(T* p, float v);
template <>
__device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void st_f<h16>(h16* p, float v) {
  *p = f_to_h16(v);
}

// load 4 consecutive elements as float4 (pointer must be 4-element aligned)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const h16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  h162 a = *reinterpret_cast<h162*>(&u.x);
  h162 b = *reinterpret_cast<h162*>(&u.y);
  float2 fa = h162_to_ff(a), fb = h162_to_ff(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(h16* p, float4 v) {
  h162 a = ff_to_h162(v.x, v.y);
  h162 b = ff_to_h162(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// 256-bit global store (sm_100: STG.E.256): one full 32-byte sector per lane per instruction; p must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                              uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {  // packs two h16 (name kept)
  h162 v = ff_to_h162(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 2^x on the SFU (MUFU.EX2), no range fix-up code: inputs here are <= 8 and -inf -> 0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// three-input maximum (sm_100: FMNMX3): halves the instruction count of a row maximum
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// 2^x WITHOUT the SFU: round-to-nearest split x = n + f (magic-number add), degree-3 near-minimax polynomial for
// 2^f on [-0.5, 0.5] (max relative error 7.5e-5, below the 16-bit rounding of P), exponent patched in with one
// integer add.  8 FMA/ALU-pipe instructions that run beside MUFU.EX2: the attention softmax is bound by the 16
// exp2/clk/SM of the SFU, so a fraction of its exponentials is computed this way (the FlashAttention-4 trick).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float r = x + 12582912.0f;          // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float f = x - (r - 12582912.0f);
  float p = fmaf(f, 0.055171653628349304f, 0.2426111251115799f);
  p = fmaf(p, f, 0.6932609677314758f);
  p = fmaf(p, f, 0.9999280571937561f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
// same with the approximate divide (2 ulp): the IEEE divide above costs ~4x the instructions, which made
// the GroupNorm+SiLU pass issue-bound instead of memory-bound
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }
// x * sigmoid(1.702 x) (transformers QuickGELUActivation: CLIP-L's MLP)
__device__ __forceinline__ float quick_gelu_f(float v) { return v / (1.0f + __expf(-1.702f * v)); }
__device__ __forceinline__ float quick_gelu_fast(float v) { return __fdividef(v, 1.0f + __expf(-1.702f * v)); }
// exact (erf) GELU, as torch.nn.functional.gelu default (module/min_sdxl.py:510)
__device__ __forceinline__ float gelu_erf_f(float v) {
  return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
}

// erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): one rcp + one ex2 + a degree-5 Horner chain instead of
// erff's ~35 branchy instructions; used by the 16-bit tensor-core epilogues only (the fp32 check mode keeps erff)
__device__ __forceinline__ float gelu_erf_fast(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, x, 1.0f)));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = 1.0f - p * ex2_approx(-x * x * 1.4426950408889634f);  // erf(|v|/sqrt2)
  return 0.5f * v + 0.5f * fabsf(v) * e;                                // 0.5 v (1 + sign(v) e)
}

// ---- Programmatic Dependent Launch ---------------------------------------------------------
__device__ __forceinline__ void pdl_trigger() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// same, with a suspend-time hint (ns): the waiting thread sleeps in hardware until the phase completes
// or the hint expires instead of re-issuing try_wait — single-thread producer / MMA warps otherwise burn
// the issue slots of the compute warps that share their scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity), "r"(hint_ns)
        : "memory");
  } while (!done);
}

// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- cp.async (LDGSTS): 16-byte global -> shared copies that need no registers -----------------
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- TMA ---------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// ---- thread-block clusters ---------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Execution-only rendezvous: keeps every CTA of the cluster (its shared memory, mbarriers, TMEM) alive until all
// peers are done with them.  No memory ordering is requested, so no MEMBAR.ALL.GPU is paid behind the epilogue's
// global stores (the release form above waits for them to become visible GPU-wide: ~1 us at the tail of every
// clustered launch).
__device__ __forceinline__ void cluster_sync_relaxed() {
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
// 2-D TMA load multicast to every CTA in `mask`: the tile lands at the same smem offset, and
// complete_tx is signalled on the mbarrier at the same offset, in each destination CTA.
__device__ __forceinline__ void tma_load_2d_mcast(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                                  int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// ---- CTA pairs (cta_group::2) --------------------------------------------------------------
// shared::cluster address of the SAME smem offset in the pair's leader CTA (even rank): clear the
// rank bit, as cute::Sm100MmaPeerBitMask does (cute/arch/copy_sm100_tma.hpp:45)
__device__ __forceinline__ uint32_t leader_smem_addr(const void* p) { return smem_u32(p) & 0xFEFFFFFFu; }
// TMA loads issued by either CTA of a pair; complete_tx goes to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_smem_addr(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the mbarrier at this smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      // default semantics (.release.cta), as cutlass::arch::ClusterBarrier::arrive(cta_id): what this arrival orders is
      // TMEM traffic (tcgen05 fences), not global memory; a cluster-scope release compiles to MEMBAR.ALL.GPU and stalls
      // every epilogue thread of the pair's second CTA behind its own output stores, once per tile
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's 128 rows] * B[N/2 rows from each CTA]; leader thread issues
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate; one thread issues.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[TMEM: row m in lane m, K elements packed two per 32-bit column] * B[smem desc]
__device__ __forceinline__ void umma_ts_bf16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// register re-partitioning between warpgroups (all 4 warps of a warpgroup must execute it)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// same, arriving on the mbarrier at this smem offset in every CTA of `mask` (cluster multicast)
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// 32 lanes x 32 consecutive columns (fp32 bits) -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// 32 registers per thread -> 32 lanes x 32 consecutive columns
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
      : "memory");
}
// 16 registers per thread -> 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),
        "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
        "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor for a 128B-swizzled operand tile whose rows are 128 bytes
// (64 bf16).  Fields per cute/arch/mma_sm100_desc.hpp (SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
//   K-major : 8-row groups 1024 B apart (SBO); LBO unused (=1).  Advance K by 16 elems = +32 B.
//   MN-major: 64 contiguous MN elements per 128 B row, 8 K-rows per 1024 B group (SBO);
//             LBO = stride between 64-element MN panels.  Advance K by 16 = +2048 B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                    uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Instruction descriptor (InstrDescriptor in the same header): c_format F32 [4,6)=1,
// a/b_format BF16 [7,10),[10,13)=1, a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29).
__device__ __host__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N, int a_mn_major,
                                                             int b_mn_major) {
  return (1u << 4) | (IIR_UMMA_FMT << 7) | (IIR_UMMA_FMT << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

}  // namespace iir
