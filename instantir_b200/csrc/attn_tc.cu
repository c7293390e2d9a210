// Flash-style attention on tcgen05 / TMEM for sm_100a, head_dim 64, bf16.
//
//   out = sum_s seg_scale[s] * softmax(Q K_sᵀ * scale) V_s         (1 or 2 key segments)
//
// One CTA per (128-query tile, head, batch); TWO CTAs per SM (112 KiB smem, 256 TMEM columns each) so every
// warp scheduler holds two softmax warps and one CTA's MMAs overlap the other's softmax:
//   warp 0     : TMA producer — Q once, then (K_j, V_j) 128-key blocks, double buffered
//   warp 1     : single-thread MMA issuer.  S_j = Q K_jᵀ (UMMA 128x128x16 x4) into the TMEM S buffer;
//                O += P_j V_j (UMMA 128x64x16 x8, V as MN-major B operand) ACCUMULATES in TMEM.
//   warps 2..5 : softmax — one query row per thread (TMEM lane == row, no shuffles).  The whole S row
//                is pulled into registers in one pass and the S buffer is released immediately, so
//                Q K_{j+1}ᵀ runs under the softmax of block j.  exp2 is taken relative to a reference
//                max that is only advanced (and O / l rescaled in TMEM, warp-uniformly) when the row max
//                grows by more than 2^8 — the usual case after the first blocks is "no rescale", which
//                removes the per-block O read-modify-write of a classic flash loop.  P_j -> bf16 into
//                128B-swizzled smem (A operand of the PV MMA).
// Two segments = decoupled text + image cross-attention with independent softmaxes
// (module/ip_adapter/attention_processor.py:1165-1192); one segment = self-attention (:394-396).
#include <cstdlib>

#include "common.cuh"

namespace iir {
namespace {

typedef h16 bf16;
constexpr int ATT_THREADS = 192;
constexpr int TILE_BYTES = 128 * 64 * 2;  // 16 KiB: Q, K_j, V_j tiles; P_j is two of these

struct alignas(64) AttnTcParams {
  CUtensorMap tmQ;
  CUtensorMap tmK[2];
  CUtensorMap tmV[2];
  int n_seg;
  int kv_len[2];
  int nblk[2];
  float seg_scale[2];
  int q_off, k_off[2], v_off[2];
  bf16* out;
  long long ldo;
  int out_off;
  int B, heads, n_q;
  float scale_log2;  // softmax_scale * log2(e)
  int wide_out;      // output rows start on 32-byte boundaries (256-bit stores in the attn_ts epilogue)
};

constexpr int ATT_TMEM_COLS = 256;  // S (128) + 2 x O (64): two CTAs share the SM's 512 columns

template <int NSEG>
__global__ void __launch_bounds__(ATT_THREADS, NSEG == 1 ? 2 : 1)
attn_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];  // 128B-swizzle tiles need 1024-byte alignment
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;       // 2 stages
  uint8_t* sV = sK + 2 * TILE_BYTES;   // 2 stages
  uint8_t* sP = sV + 2 * TILE_BYTES;   // 1 buffer x 32 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // 2
  uint64_t* kv_empty = bars + 3;      // 2
  uint64_t* s_full = bars + 5;        // 1
  uint64_t* s_empty = bars + 6;       // 1
  uint64_t* p_full = bars + 7;        // 1
  uint64_t* p_empty = bars + 8;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
  const int nb_total = p.nblk[0] + (NSEG > 1 ? p.nblk[1] : 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK[0]);
    tma_prefetch_desc(&p.tmV[0]);
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    mbar_init(p_full, 128);
    mbar_init(p_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmem_S = tmem_base;        // 128 columns
  const uint32_t tmem_O = tmem_base + 128;  // 64 columns, accumulated across key blocks

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (whole warp, elected issue)
    if (elect_one()) {
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &p.tmQ, q_full, p.q_off + h * 64, q0, b);
    }
    __syncwarp();
    int jb = 0;
    for (int s = 0; s < NSEG; ++s) {
      for (int jl = 0; jl < p.nblk[s]; ++jl, ++jb) {
        const int st = jb & 1;
        const uint32_t ph = (jb >> 1) & 1;
        mbar_wait(&kv_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&kv_full[st], 2 * TILE_BYTES);
          tma_load_3d(sK + st * TILE_BYTES, &p.tmK[s], &kv_full[st], p.k_off[s] + h * 64, jl * 128, b);
          tma_load_3d(sV + st * TILE_BYTES, &p.tmV[s], &kv_full[st], p.v_off[s] + h * 64, jl * 128, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer (whole warp, elected issue)
    const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);  // B (=V) is MN-major
    const int seg1_first = p.nblk[0];  // first flat block index of the second segment
    auto issue_pv = [&](int j) {
      const int st = j & 1;
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      const uint32_t pa = smem_u32(sP);
      const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV + st * TILE_BYTES), 1024, 1024);
      const bool fresh = (j == 0) || (NSEG > 1 && j == seg1_first);  // first block of a segment
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint64_t adesc = umma_desc_sw128(pa + (k >> 2) * TILE_BYTES + (k & 3) * 32, 16, 1024);
          umma_bf16(tmem_O, adesc, vdesc0 + 128 * k, idesc_pv, (k != 0 || !fresh) ? 1u : 0u);
        }
        umma_commit(&kv_empty[st]);
        umma_commit(p_empty);  // PV_j retired: P buffer free and O stable
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    const uint64_t qdesc0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
    for (int jb = 0; jb < nb_total; ++jb) {
      const int st = jb & 1;
      const uint32_t ph = (jb >> 1) & 1;
      mbar_wait(&kv_full[st], ph);
      mbar_wait(s_empty, (jb & 1) ^ 1);
      tc_fence_after();
      const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK + st * TILE_BYTES), 16, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_S, qdesc0 + 2 * k, kdesc0 + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      if (jb > 0) issue_pv(jb - 1);
    }
    issue_pv(nb_total - 1);
  } else {
    // ---------------------------------------------------------------- softmax warps
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const uint32_t trow = static_cast<uint32_t>(lane_base) << 16;
    const float sl2 = p.scale_log2;
    const float kLazy = 8.0f / sl2;  // advance the reference max only when exp2 arguments would exceed 8
    float out_acc[NSEG > 1 ? 64 : 1];
    if (NSEG > 1) {
#pragma unroll
      for (int d = 0; d < 64; ++d) out_acc[d] = 0.f;
    }
    uint8_t* prow = sP + row * 128;
    const int qi = q0 + row;
    bf16* optr = p.out + (static_cast<long long>(b) * p.n_q + qi) * p.ldo + p.out_off + h * 64;
    int jb = 0;
#pragma unroll 1
    for (int s = 0; s < NSEG; ++s) {
      float m_ref = -INFINITY, l_run = 0.f;
#pragma unroll 1
      for (int jl = 0; jl < p.nblk[s]; ++jl, ++jb) {
        const int valid = min(128, p.kv_len[s] - jl * 128);
        mbar_wait(s_full, jb & 1);
        tc_fence_after();
        uint32_t sr[128];
        {
          uint32_t (&a0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[0]);
          uint32_t (&a1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[32]);
          uint32_t (&a2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[64]);
          uint32_t (&a3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[96]);
          tmem_ld32(tmem_S + trow, a0);
          tmem_ld32(tmem_S + trow + 32, a1);
          tmem_ld32(tmem_S + trow + 64, a2);
          tmem_ld32(tmem_S + trow + 96, a3);
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(s_empty);  // S is in registers: Q K_{j+1}^T may overwrite the TMEM buffer now
        if (valid < 128) {
#pragma unroll
          for (int j = 0; j < 128; ++j)
            if (j >= valid) sr[j] = 0xff800000u;  // -inf
        }
        // 8 independent chains instead of one 128-deep dependent chain
        float mxs[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) mxs[c] = __uint_as_float(sr[c]);
#pragma unroll
        for (int j = 8; j < 128; j += 8)
#pragma unroll
          for (int c = 0; c < 8; ++c) mxs[c] = fmaxf(mxs[c], __uint_as_float(sr[j + c]));
        const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                               fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
        // previous P V must have retired before P is overwritten / O is touched
        mbar_wait(p_empty, (jb & 1) ^ 1);
        tc_fence_after();
        const bool first = (jl == 0);
        const bool grow = !first && (mx > m_ref + kLazy);
        if (first) m_ref = mx;
        if (__any_sync(0xffffffffu, grow)) {
          // rare: rescale O (TMEM) and l by exp2((m_ref - mx) * sl2) for the rows that need it
          const float alpha = grow ? ex2_approx((m_ref - mx) * sl2) : 1.0f;
          if (grow) { l_run *= alpha; m_ref = mx; }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            tmem_ld32(tmem_O + trow + half * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) r[d] = __float_as_uint(__uint_as_float(r[d]) * alpha);
            tmem_st32(tmem_O + trow + half * 32, r);
          }
          tmem_st_wait();
        }
        const float mneg = -m_ref * sl2;
        float sums[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) sums[c] = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          float e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            e[j] = ex2_approx(fmaf(__uint_as_float(sr[g * 8 + j]), sl2, mneg));
            sums[j] += e[j];
          }
          const int panel = g >> 3;  // 64 keys per 128-byte panel row
          const int q8 = g & 7;
          uint4 u;
          u.x = pack_bf16(e[0], e[1]);
          u.y = pack_bf16(e[2], e[3]);
          u.z = pack_bf16(e[4], e[5]);
          u.w = pack_bf16(e[6], e[7]);
          *reinterpret_cast<uint4*>(prow + panel * TILE_BYTES + ((q8 ^ (row & 7)) << 4)) = u;
        }
        const float sum = ((sums[0] + sums[1]) + (sums[2] + sums[3])) + ((sums[4] + sums[5]) + (sums[6] + sums[7]));
        l_run += sum;
        fence_async_smem();
        tc_fence_before();
        mbar_arrive(p_full);
      }
      // segment done: wait for its last P V, then O / l
      mbar_wait(p_empty, (jb - 1) & 1);
      tc_fence_after();
      const float w = p.seg_scale[s] / l_run;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32(tmem_O + trow + half * 32, r);
        tmem_ld_wait();
        if (NSEG > 1) {
#pragma unroll
          for (int d = 0; d < 32; ++d) out_acc[half * 32 + d] = fmaf(__uint_as_float(r[d]), w, out_acc[half * 32 + d]);
        }
        if (s == NSEG - 1 && qi < p.n_q) {
#pragma unroll
          for (int d = 0; d < 32; d += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              v[j] = NSEG > 1 ? out_acc[half * 32 + d + j] : __uint_as_float(r[d + j]) * w;
            uint4 u;
            u.x = pack_bf16(v[0], v[1]);
            u.y = pack_bf16(v[2], v[3]);
            u.z = pack_bf16(v[4], v[5]);
            u.w = pack_bf16(v[6], v[7]);
            *reinterpret_cast<uint4*>(optr + half * 32 + d) = u;
          }
        }
      }
      tc_fence_before();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

// ==================================================================================================
// One-segment (self-attention) kernel, second generation: P never touches shared memory.
//
//   TMEM (256 columns per CTA, two CTAs per SM):  S fp32 [0,128) | P 16-bit packed [128,192) | O fp32 [192,256)
//   warps 0..3 : softmax warpgroup, one query row per thread (TMEM lane == row), 216 registers each
//                (setmaxnreg) so the whole 128-column S row lives in registers without spills
//   warp 4     : TMA producer — Q once, then (K_j, V_j) 128-key blocks in a 3-deep ring
//   warp 5     : single-thread MMA issuer — S_j = Q K_jᵀ (SS: both operands in smem, 128x128x16 x4),
//                O += P_j V_j (TS: A = P_j read from TMEM, B = V_j MN-major in smem, 128x64x16 x8)
//   warps 6..7 : idle (complete the second warpgroup, 40 registers)
// Versus the first generation (P through 128B-swizzled smem): no 32 KiB P write + 32 KiB P read per block
// on the 128 B/clk shared-memory port (it was ~90 % busy with two CTAs per SM), one more K/V stage in the
// freed space, and no local-memory spills in the exp loop.
constexpr int ATS_THREADS = 256;
constexpr int ATS_KV_STAGES = 3;

// PERBLOCK = true: every key segment fits ONE 128-key block (the decoupled text + image cross-attention: 77 and
// 64 keys).  Block j is segment j with its own K/V tensors; its softmax is complete after that block, so P is
// normalised (and weighted by seg_scale) in registers and both segments accumulate straight into the same O:
//   out = sum_s seg_scale[s] * softmax(Q K_s^T) V_s     with no per-segment accumulator and no final rescale.
// element j of a row: SFU exp2, or the FMA-pipe polynomial for every POLY-th element
template <int POLY>
__device__ __forceinline__ float exp2_sel(int j, float x) {
  if (POLY > 0 && (j % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) return ex2_poly(x);
  return ex2_approx(x);
}

// POLY = n > 0: every n-th exponential of a row is computed on the FMA pipe (ex2_poly) instead of the SFU.
template <bool PERBLOCK, int POLY>
__global__ void __launch_bounds__(ATS_THREADS, 2)
attn_ts_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;                      // ATS_KV_STAGES tiles
  uint8_t* sV = sK + ATS_KV_STAGES * TILE_BYTES;      // ATS_KV_STAGES tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATS_KV_STAGES * TILE_BYTES);
  uint64_t* q_full = bars;                             // 1
  uint64_t* kv_full = bars + 1;                        // ATS_KV_STAGES
  uint64_t* kv_empty = kv_full + ATS_KV_STAGES;        // ATS_KV_STAGES
  uint64_t* s_full = kv_empty + ATS_KV_STAGES;         // 1
  uint64_t* s_empty = s_full + 1;                      // 1
  uint64_t* p_full = s_empty + 1;                      // 2: P columns [0,32) / [32,64) written
  uint64_t* p_empty = p_full + 2;                      // 2: the MMAs reading that half of P retired
  uint64_t* q_empty = p_empty + 2;                     // 1: the Q K^T MMAs of a tile retired, sQ may be refilled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: CTA c walks tiles c, c + grid, ... (tile = (batch, head, 128-query block), query block fastest so
  // that co-running CTAs share one head's K/V in L2); barriers, TMEM and the K/V ring live across tiles, and the
  // producer runs ahead into the next tile while the softmax of the current one finishes
  const int nqt = (p.n_q + 127) >> 7;
  const int total_tiles = nqt * p.heads * p.B;
  const int nb = PERBLOCK ? p.n_seg : p.nblk[0];

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK[0]);
    tma_prefetch_desc(&p.tmV[0]);
    if (PERBLOCK && p.n_seg > 1) {
      tma_prefetch_desc(&p.tmK[1]);
      tma_prefetch_desc(&p.tmV[1]);
    }
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(s_empty, 128);
    for (int hh = 0; hh < 2; ++hh) {
      mbar_init(&p_full[hh], 128);
      mbar_init(&p_empty[hh], 1);
    }
    for (int s = 0; s < ATS_KV_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;        // 128 columns fp32
  const uint32_t tmem_P = tmem_base + 128;  // 64 columns: 128 probabilities, two per column
  const uint32_t tmem_O = tmem_base + 192;  // 64 columns fp32, accumulated across key blocks

  if (warp >= 4) {
    setmaxnreg_dec<40>();
    pdl_wait();
    // Both single-issuer roles run their loops with the WHOLE warp (warp-uniform control flow keeps
    // descriptors / coordinates in uniform registers) and elect one lane per instruction group.
    if (warp == 4) {
      // ---------------------------------------------------------------- TMA producer
      int st = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int qt = tile % nqt, hb = tile / nqt;
        const int h = hb % p.heads, b = hb / p.heads, q0 = qt * 128;
        if (it > 0) mbar_wait_sleep(q_empty, (it - 1) & 1, 20000);  // Q K^T of the previous tile has read sQ
        if (elect_one()) {
          mbar_expect_tx(q_full, TILE_BYTES);
          tma_load_3d(sQ, &p.tmQ, q_full, p.q_off + h * 64, q0, b);
        }
        __syncwarp();
        for (int jb = 0; jb < nb; ++jb) {
          mbar_wait_sleep(&kv_empty[st], ph ^ 1, 20000);
          if (elect_one()) {
            mbar_expect_tx(&kv_full[st], 2 * TILE_BYTES);
            const int sg = PERBLOCK ? jb : 0, jr = PERBLOCK ? 0 : jb * 128;
            tma_load_3d(sK + st * TILE_BYTES, &p.tmK[sg], &kv_full[st], p.k_off[sg] + h * 64, jr, b);
            tma_load_3d(sV + st * TILE_BYTES, &p.tmV[sg], &kv_full[st], p.v_off[sg] + h * 64, jr, b);
          }
          __syncwarp();
          if (++st == ATS_KV_STAGES) { st = 0; ph ^= 1; }
        }
      }
    } else if (warp == 5) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      const uint32_t idesc_pv = umma_idesc_bf16(128, 64, 0, 1);  // A (=P) K-major in TMEM, B (=V) MN-major
      // O += P_j V_j in two halves of 64 keys: the first half's MMAs run under the exps of the second
      // g = block counter over all tiles of this CTA (barrier parities), jl = block index inside the tile
      auto issue_pv = [&](uint32_t g, int jl, int st) {
        const uint64_t vdesc0 = umma_desc_sw128(smem_u32(sV + st * TILE_BYTES), 1024, 1024);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          mbar_wait_sleep(&p_full[hh], g & 1, 2000);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int k = hh * 4; k < hh * 4 + 4; ++k)  // +16 keys = +2048 B = +128 in the (addr >> 4) field
              umma_ts_bf16(tmem_O, tmem_P + k * 8, vdesc0 + 128 * k, idesc_pv, (k != 0 || jl != 0) ? 1u : 0u);
            if (hh == 1) umma_commit(&kv_empty[st]);
            umma_commit(&p_empty[hh]);  // this half of P is free; after hh == 1, O is stable
          }
          __syncwarp();
        }
      };
      const uint64_t qdesc0 = umma_desc_sw128(smem_u32(sQ), 16, 1024);
      int st = 0, st_prev = 0;
      uint32_t ph = 0, g = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        mbar_wait_sleep(q_full, it & 1, 2000);
        for (int jb = 0; jb < nb; ++jb, ++g) {
          mbar_wait_sleep(&kv_full[st], ph, 2000);
          mbar_wait_sleep(s_empty, (g & 1) ^ 1, 2000);
          tc_fence_after();
          const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK + st * TILE_BYTES), 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_S, qdesc0 + 2 * k, kdesc0 + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            umma_commit(s_full);
            if (jb == nb - 1) umma_commit(q_empty);  // last read of this tile's Q
          }
          __syncwarp();
          // P V of the previous block; the first P V of a tile overwrites O (accumulate = 0): the softmax warps
          // signal its P only after their epilogue has read the previous tile's O out of TMEM
          if (jb > 0) issue_pv(g - 1, jb - 1, st_prev);
          st_prev = st;
          if (++st == ATS_KV_STAGES) { st = 0; ph ^= 1; }
        }
        issue_pv(g - 1, nb - 1, st_prev);
      }
    }
  } else {
    // ---------------------------------------------------------------- softmax warpgroup
    setmaxnreg_inc<216>();
    pdl_wait();
    const int lane_base = warp * 32;
    const int row = lane_base + lane;
    const uint32_t trow = static_cast<uint32_t>(lane_base) << 16;
    const float sl2 = p.scale_log2;
    const float kLazy = 8.0f / sl2;  // advance the reference max only when exp2 arguments would exceed 8
    uint32_t gb = 0;  // block counter over all tiles of this CTA: every per-block barrier completes once per block
#pragma unroll 1
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
    const int qt = tile % nqt, hb = tile / nqt;
    const int h = hb % p.heads, b = hb / p.heads, q0 = qt * 128;
    const int qi = q0 + row;
    bf16* optr = p.out + (static_cast<long long>(b) * p.n_q + qi) * p.ldo + p.out_off + h * 64;
    float m_ref = -INFINITY, l_run = 0.f;
#pragma unroll 1
    for (int jb = 0; jb < nb; ++jb, ++gb) {
      const int valid = PERBLOCK ? min(128, p.kv_len[jb]) : min(128, p.kv_len[0] - jb * 128);
      mbar_wait_sleep(s_full, gb & 1, 20000);
      tc_fence_after();
      uint32_t sr[128];
      {
        uint32_t (&a0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[0]);
        uint32_t (&a1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[32]);
        uint32_t (&a2)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[64]);
        uint32_t (&a3)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sr[96]);
        tmem_ld32(tmem_S + trow, a0);
        tmem_ld32(tmem_S + trow + 32, a1);
        tmem_ld32(tmem_S + trow + 64, a2);
        tmem_ld32(tmem_S + trow + 96, a3);
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty);  // S is in registers: Q K_{j+1}^T may overwrite the TMEM buffer now
      if (valid < 128) {
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (j >= valid) sr[j] = 0xff800000u;  // -inf
      }
      float mxs[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mxs[c] = __uint_as_float(sr[c]);
#pragma unroll
      for (int j = 8; j < 128; j += 8)
#pragma unroll
        for (int c = 0; c < 8; ++c) mxs[c] = fmaxf(mxs[c], __uint_as_float(sr[j + c]));
      const float mx = fmaxf(fmaxf(fmaxf(mxs[0], mxs[1]), fmaxf(mxs[2], mxs[3])),
                             fmaxf(fmaxf(mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7])));
      if (PERBLOCK) {
        // complete softmax of this segment in registers: e -> sum -> P = e * seg_scale / sum
        const float mneg = -mx * sl2;
        float sums[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) sums[c] = 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) {
          const float e = ex2_approx(fmaf(__uint_as_float(sr[j]), sl2, mneg));
          sums[j & 7] += e;
          sr[j] = __float_as_uint(e);
        }
        const float wseg = p.seg_scale[jb] /
            (((sums[0] + sums[1]) + (sums[2] + sums[3])) + ((sums[4] + sums[5]) + (sums[6] + sums[7])));
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            pk[j >> 1] = pack_bf16(__uint_as_float(sr[g * 32 + j]) * wseg, __uint_as_float(sr[g * 32 + j + 1]) * wseg);
          if (g == 0 || g == 2) {
            mbar_wait(&p_empty[g >> 1], (gb & 1) ^ 1);
            tc_fence_after();
          }
          tmem_st16(tmem_P + trow + g * 16, pk);
          if (g == 1 || g == 3) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[g >> 1]);
          }
        }
        l_run = 1.0f;
        continue;
      }
      const bool first = (jb == 0);
      const bool grow = !first && (mx > m_ref + kLazy);
      const float m_old = m_ref;
      if (first || grow) m_ref = mx;
      const float mneg = -m_ref * sl2;
      float sums[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) sums[c] = 0.f;
      // exps of the first 64 keys are computed BEFORE waiting for P V_{j-1}: its tail is hidden under them
      uint32_t pk0[16], pk1[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float e0 = exp2_sel<POLY>(j, fmaf(__uint_as_float(sr[j]), sl2, mneg));
        const float e1 = exp2_sel<POLY>(j + 1, fmaf(__uint_as_float(sr[j + 1]), sl2, mneg));
        sums[j & 7] += e0;
        sums[(j + 1) & 7] += e1;
        pk0[j >> 1] = pack_bf16(e0, e1);
      }
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float e0 = exp2_sel<POLY>(j, fmaf(__uint_as_float(sr[32 + j]), sl2, mneg));
        const float e1 = exp2_sel<POLY>(j + 1, fmaf(__uint_as_float(sr[32 + j + 1]), sl2, mneg));
        sums[j & 7] += e0;
        sums[(j + 1) & 7] += e1;
        pk1[j >> 1] = pack_bf16(e0, e1);
      }
      if (__any_sync(0xffffffffu, grow)) {
        // rare: rescale O (TMEM) and l by exp2((m_old - mx) * sl2) for the rows that need it; O must be stable
        mbar_wait(&p_empty[1], (gb & 1) ^ 1);
        tc_fence_after();
        const float alpha = grow ? ex2_approx((m_old - mx) * sl2) : 1.0f;
        l_run *= alpha;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld32(tmem_O + trow + half * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int d = 0; d < 32; ++d) r[d] = __float_as_uint(__uint_as_float(r[d]) * alpha);
          tmem_st32(tmem_O + trow + half * 32, r);
        }
        tmem_st_wait();
      }
      mbar_wait(&p_empty[0], (gb & 1) ^ 1);  // first half of P V_{j-1} retired: P columns [0,32) are free
      tc_fence_after();
      tmem_st16(tmem_P + trow, pk0);
      tmem_st16(tmem_P + trow + 16, pk1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[0]);
#pragma unroll
      for (int g = 2; g < 4; ++g) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = exp2_sel<POLY>(j, fmaf(__uint_as_float(sr[g * 32 + j]), sl2, mneg));
          const float e1 = exp2_sel<POLY>(j + 1, fmaf(__uint_as_float(sr[g * 32 + j + 1]), sl2, mneg));
          sums[j & 7] += e0;
          sums[(j + 1) & 7] += e1;
          pk[j >> 1] = pack_bf16(e0, e1);
        }
        if (g == 2) {
          mbar_wait(&p_empty[1], (gb & 1) ^ 1);  // all of P V_{j-1} retired
          tc_fence_after();
        }
        tmem_st16(tmem_P + trow + g * 16, pk);
      }
      l_run += ((sums[0] + sums[1]) + (sums[2] + sums[3])) + ((sums[4] + sums[5]) + (sums[6] + sums[7]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[1]);
    }
    // all key blocks done: wait for the last P V, then O / l
    mbar_wait(&p_empty[1], (gb - 1) & 1);
    tc_fence_after();
    const float w = PERBLOCK ? 1.0f : p.seg_scale[0] / l_run;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t r[32];
      tmem_ld32(tmem_O + trow + half * 32, r);
      tmem_ld_wait();
      if (qi < p.n_q) {
        if (p.wide_out) {  // rows start on 32-byte boundaries: 256-bit stores, one full sector per lane per instruction
#pragma unroll
          for (int d = 0; d < 32; d += 16)
            st_global_256(optr + half * 32 + d,
                          pack_bf16(__uint_as_float(r[d]) * w, __uint_as_float(r[d + 1]) * w),
                          pack_bf16(__uint_as_float(r[d + 2]) * w, __uint_as_float(r[d + 3]) * w),
                          pack_bf16(__uint_as_float(r[d + 4]) * w, __uint_as_float(r[d + 5]) * w),
                          pack_bf16(__uint_as_float(r[d + 6]) * w, __uint_as_float(r[d + 7]) * w),
                          pack_bf16(__uint_as_float(r[d + 8]) * w, __uint_as_float(r[d + 9]) * w),
                          pack_bf16(__uint_as_float(r[d + 10]) * w, __uint_as_float(r[d + 11]) * w),
                          pack_bf16(__uint_as_float(r[d + 12]) * w, __uint_as_float(r[d + 13]) * w),
                          pack_bf16(__uint_as_float(r[d + 14]) * w, __uint_as_float(r[d + 15]) * w));
        } else {
#pragma unroll
          for (int d = 0; d < 32; d += 8) {
            uint4 u;
            u.x = pack_bf16(__uint_as_float(r[d]) * w, __uint_as_float(r[d + 1]) * w);
            u.y = pack_bf16(__uint_as_float(r[d + 2]) * w, __uint_as_float(r[d + 3]) * w);
            u.z = pack_bf16(__uint_as_float(r[d + 4]) * w, __uint_as_float(r[d + 5]) * w);
            u.w = pack_bf16(__uint_as_float(r[d + 6]) * w, __uint_as_float(r[d + 7]) * w);
            *reinterpret_cast<uint4*>(optr + half * 32 + d) = u;
          }
        }
      }
    }
    tc_fence_before();
    }  // tiles
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

int make_map3(CUtensorMap* m, const void* base, long long ld, int rows, int B) {
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)rows, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)rows * ld * 2};
  uint32_t box[3] = {64, 128, 1};
  CUresult cr = encode_tiled(m, IIR_H16_TMA, 3, base, dims, strides, box,
                             CU_TENSOR_MAP_SWIZZLE_128B);
  if (cr != CUDA_SUCCESS) {
    set_error("iir_attn_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return IIR_ERR_CUDA;
  }
  return 0;
}

}  // namespace
}  // namespace iir

extern "C" int iir_attn_tc(const iir_attn_args* a, void* stream) {
  using namespace iir;
  IIR_REQUIRE(a != nullptr, "iir_attn_tc: null args");
  IIR_REQUIRE(a->dtype == IIR_H16, "iir_attn_tc: bf16 only (use iir_attn_simt for fp32)");
  IIR_REQUIRE(a->n_seg == 1 || a->n_seg == 2, "iir_attn_tc: n_seg must be 1 or 2");
  IIR_REQUIRE(a->B > 0 && a->heads > 0 && a->n_q > 0, "iir_attn_tc: empty problem");
  IIR_REQUIRE(a->ldq % 8 == 0 && a->ldo % 8 == 0 && a->q_off % 8 == 0 && a->out_off % 8 == 0,
              "iir_attn_tc: leading dims / offsets must be multiples of 8");
  AttnTcParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_map3(&p.tmQ, a->q, a->ldq, a->n_q, a->B);
  if (rc) return rc;
  p.n_seg = a->n_seg;
  for (int s = 0; s < a->n_seg; ++s) {
    IIR_REQUIRE(a->kv_len[s] > 0, "iir_attn_tc: empty key segment %d", s);
    IIR_REQUIRE(a->ldk[s] % 8 == 0 && a->ldv[s] % 8 == 0 && a->k_off[s] % 8 == 0 && a->v_off[s] % 8 == 0,
                "iir_attn_tc: K/V leading dims / offsets must be multiples of 8");
    rc = make_map3(&p.tmK[s], a->k[s], a->ldk[s], a->kv_len[s], a->B);
    if (rc) return rc;
    rc = make_map3(&p.tmV[s], a->v[s], a->ldv[s], a->kv_len[s], a->B);
    if (rc) return rc;
    p.kv_len[s] = a->kv_len[s];
    p.nblk[s] = (a->kv_len[s] + 127) / 128;
    p.seg_scale[s] = a->seg_scale[s];
    p.k_off[s] = a->k_off[s];
    p.v_off[s] = a->v_off[s];
  }
  p.q_off = a->q_off;
  p.out = reinterpret_cast<bf16*>(a->out);
  p.ldo = a->ldo;
  p.out_off = a->out_off;
  p.B = a->B; p.heads = a->heads; p.n_q = a->n_q;
  p.scale_log2 = a->softmax_scale * 1.4426950408889634f;
  p.wide_out = (reinterpret_cast<uintptr_t>(a->out) % 32 == 0) && (a->ldo * 2) % 32 == 0 && (a->out_off * 2) % 32 == 0;

  const size_t smem = 7 * TILE_BYTES + 256;
  dim3 grid((a->n_q + 127) / 128, a->heads, a->B);
  // attn_ts kernels are persistent: two CTAs per SM walk the (batch, head, query block) tiles
  const long long ats_tiles = static_cast<long long>(grid.x) * grid.y * grid.z;
  const dim3 ats_grid(static_cast<unsigned>(ats_tiles < 2LL * sm_count() ? ats_tiles : 2LL * sm_count()));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  static int v1 = -1;
  if (v1 < 0) {
    const char* ev = getenv("IIR_ATTN_V1");
    v1 = (ev && ev[0] == '1') ? 1 : 0;
  }
  const bool one_block_each = a->n_seg == 2 && a->kv_len[0] <= 128 && a->kv_len[1] <= 128;
  static int poly = -1;
  if (poly < 0) {
    const char* ev = getenv("IIR_ATTN_POLY");  // 0 = all exponentials on the SFU; n = every n-th on the FMA pipe
    poly = ev ? atoi(ev) : 4;
  }
#define ATS_LAUNCH(PB, PL)                                                                                         \
  do {                                                                                                             \
    e = cudaFuncSetAttribute(attn_ts_kernel<PB, PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
    if (e == cudaSuccess) e = launch_pdl(attn_ts_kernel<PB, PL>, ats_grid, dim3(ATS_THREADS), smem, st, p);            \
  } while (0)
  if (a->n_seg == 1 && !v1) {
    if (poly == 2) ATS_LAUNCH(false, 2);
    else if (poly == 3) ATS_LAUNCH(false, 3);
    else if (poly == 4) ATS_LAUNCH(false, 4);
    else if (poly == 8) ATS_LAUNCH(false, 8);
    else ATS_LAUNCH(false, 0);
  } else if (one_block_each && !v1) {
    ATS_LAUNCH(true, 0);
  } else if (a->n_seg == 1) {
    e = cudaFuncSetAttribute(attn_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = launch_pdl(attn_tc_kernel<1>, grid, dim3(ATT_THREADS), smem, st, p);
  } else {
    e = cudaFuncSetAttribute(attn_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = launch_pdl(attn_tc_kernel<2>, grid, dim3(ATT_THREADS), smem, st, p);
  }
  if (e != cudaSuccess) {
    set_error("iir_attn_tc: %s", cudaGetErrorString(e));
    return IIR_ERR_CUDA;
  }
  count_launch();
  return check_launch("iir_attn_tc");
}
