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
  // attn_ts work partition: CTA c owns the (tile, key block) units [c * units_per_cta, (c + 1) * units_per_cta) of the
  // tile-major unit list; a tile cut by a CTA boundary is finished by the last CTA to arrive on its ticket
  int causal;  // query i attends keys 0..i (one segment, n_q == kv_len)
  int units_per_cta;
  long long total_units;
  float* ws_part;          // [grid][2][PART_FLOATS]: unnormalised O rows, reference max, row sum of a partial tile
  unsigned int* ws_ticket; // [tiles], zero between launches
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
constexpr int ATS_PART_FLOATS = 128 * 64 + 256;  // one parked partial tile: O [128][64], m_ref [128], l [128]
constexpr int ATS_MAX_PARTS = 8;                 // a tile is cut into at most this many parts

// Work list of one attn_ts CTA.  Two partitions of the (tile, key block) units:
//   whole tiles, round robin (units_per_cta == 0): CTA c takes tiles c, c + grid, ...;
//   contiguous unit ranges (units_per_cta = L > 0): CTA c takes units [c * L, (c + 1) * L) of the tile-major list, so a
//   tile may be cut at a key-block boundary (SEGMENTS of one tile in neighbouring CTAs, merged by the last to arrive).
struct AtsSegments {
  long long u, u_end;
  int nb, tile_step, total_tiles, L;
  int tile, j0, j1;     // current segment: key blocks [j0, j1) of `tile`
  bool opens_range;     // the segment starts at the first unit of this CTA's range
  __device__ __forceinline__ AtsSegments(const AttnTcParams& p, int nb_, int total_tiles_)
      : nb(nb_), tile_step(gridDim.x), total_tiles(total_tiles_), L(p.units_per_cta) {
    if (L > 0) {
      u = static_cast<long long>(blockIdx.x) * L;
      u_end = u + L < p.total_units ? u + L : p.total_units;
    } else {
      u = blockIdx.x;  // tile index
      u_end = total_tiles_;
    }
    tile = j0 = j1 = 0;
    opens_range = false;
  }
  __device__ __forceinline__ bool next() {
    if (u >= u_end) return false;
    if (L > 0) {
      tile = static_cast<int>(u / nb);
      j0 = static_cast<int>(u - static_cast<long long>(tile) * nb);
      j1 = static_cast<long long>(nb - j0) < u_end - u ? nb : j0 + static_cast<int>(u_end - u);
      opens_range = (u == static_cast<long long>(blockIdx.x) * L);
      u += j1 - j0;
    } else {
      tile = static_cast<int>(u);
      j0 = 0;
      j1 = nb;
      u += tile_step;
    }
    return true;
  }
};

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
  uint64_t* q_empty = p_empty + 2;                     // 1: the Q K^T MMAs of a segment retired, sQ may be refilled
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);
  volatile unsigned int* s_ticket = reinterpret_cast<volatile unsigned int*>(tmem_slot + 1);
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  pdl_trigger();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent: the (tile, key block) units — tile = (batch, head, 128-query block), query block fastest so that
  // co-running CTAs share one head's K/V in L2 — are listed tile-major and CTA c owns the contiguous range
  // [c * L, (c + 1) * L).  With 320 tiles of 8 blocks on 296 resident CTAs a tile-granular walk needs two waves (the
  // second one 8 % full); the unit-granular ranges give every CTA 8.65 blocks.  A range is processed as SEGMENTS
  // (tile, first block, last block); a segment that covers only part of its tile leaves (O, m, l) in a workspace slot and
  // takes a ticket on the tile: whoever arrives last merges the parts IN CTA ORDER (bit-deterministic) and writes the
  // tile's output.  Barriers, TMEM and the K/V ring live across segments, and the producer runs ahead into the next
  // segment while the softmax of the current one finishes.
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
      for (AtsSegments sg(p, nb, total_tiles); sg.next(); ++it) {
        const int tile = sg.tile, j0 = sg.j0, j1 = sg.j1;
        const int qt = tile % nqt, hb = tile / nqt;
        const int h = hb % p.heads, b = hb / p.heads, q0 = qt * 128;
        if (it > 0) mbar_wait_sleep(q_empty, (it - 1) & 1, 20000);  // Q K^T of the previous segment has read sQ
        if (elect_one()) {
          mbar_expect_tx(q_full, TILE_BYTES);
          tma_load_3d(sQ, &p.tmQ, q_full, p.q_off + h * 64, q0, b);
        }
        __syncwarp();
        for (int jb = j0; jb < j1; ++jb) {
          mbar_wait_sleep(&kv_empty[st], ph ^ 1, 20000);
          if (elect_one()) {
            mbar_expect_tx(&kv_full[st], 2 * TILE_BYTES);
            const int ks = PERBLOCK ? jb : 0, jr = PERBLOCK ? 0 : jb * 128;
            tma_load_3d(sK + st * TILE_BYTES, &p.tmK[ks], &kv_full[st], p.k_off[ks] + h * 64, jr, b);
            tma_load_3d(sV + st * TILE_BYTES, &p.tmV[ks], &kv_full[st], p.v_off[ks] + h * 64, jr, b);
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
      for (AtsSegments sg(p, nb, total_tiles); sg.next(); ++it) {
        const int nbs = sg.j1 - sg.j0;  // blocks of this segment; jb below counts from the segment's first block
        mbar_wait_sleep(q_full, it & 1, 2000);
        for (int jb = 0; jb < nbs; ++jb, ++g) {
          mbar_wait_sleep(&kv_full[st], ph, 2000);
          mbar_wait_sleep(s_empty, (g & 1) ^ 1, 2000);
          tc_fence_after();
          const uint64_t kdesc0 = umma_desc_sw128(smem_u32(sK + st * TILE_BYTES), 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_S, qdesc0 + 2 * k, kdesc0 + 2 * k, idesc_s, k != 0 ? 1u : 0u);
            umma_commit(s_full);
            if (jb == nbs - 1) umma_commit(q_empty);  // last read of this segment's Q
          }
          __syncwarp();
          // P V of the previous block; the first P V of a tile overwrites O (accumulate = 0): the softmax warps
          // signal its P only after their epilogue has read the previous tile's O out of TMEM
          if (jb > 0) issue_pv(g - 1, jb - 1, st_prev);
          st_prev = st;
          if (++st == ATS_KV_STAGES) { st = 0; ph ^= 1; }
        }
        issue_pv(g - 1, nbs - 1, st_prev);
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
    uint32_t gb = 0;  // block counter over all segments of this CTA: every per-block barrier completes once per block
#pragma unroll 1
    for (AtsSegments sg(p, nb, total_tiles); sg.next();) {
    const int tile = sg.tile, j0 = sg.j0, j1 = sg.j1;
    const bool seg_at_range_start = sg.opens_range;
    const int qt = tile % nqt, hb = tile / nqt;
    const int h = hb % p.heads, b = hb / p.heads, q0 = qt * 128;
    const int qi = q0 + row;
    bf16* optr = p.out + (static_cast<long long>(b) * p.n_q + qi) * p.ldo + p.out_off + h * 64;
    float m_ref = -INFINITY, l_run = 0.f;
#pragma unroll 1
    for (int jb = j0; jb < j1; ++jb, ++gb) {
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
      if (!PERBLOCK && p.causal && jb * 128 + 127 > qi) {  // causal mask: keys past this row's own position
        const int lim = qi - jb * 128;  // last visible key of this block (may be negative: nothing visible)
#pragma unroll
        for (int j = 0; j < 128; ++j)
          if (j > lim) sr[j] = 0xff800000u;
      }
      // row maximum with three-input max (FMNMX3): 8 independent chains of 8 instructions instead of 16
      float mxs[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) mxs[c] = fmax3(__uint_as_float(sr[c]), __uint_as_float(sr[8 + c]), __uint_as_float(sr[16 + c]));
#pragma unroll
      for (int j = 24; j < 120; j += 16)
#pragma unroll
        for (int c = 0; c < 8; ++c) mxs[c] = fmax3(mxs[c], __uint_as_float(sr[j + c]), __uint_as_float(sr[j + 8 + c]));
#pragma unroll
      for (int c = 0; c < 8; ++c) mxs[c] = fmaxf(mxs[c], __uint_as_float(sr[120 + c]));
      const float mx = fmax3(fmax3(mxs[0], mxs[1], mxs[2]), fmax3(mxs[3], mxs[4], mxs[5]), fmaxf(mxs[6], mxs[7]));
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
      const bool first = (jb == j0);
      const bool grow = !first && (mx > m_ref + kLazy);
      const float m_old = m_ref;
      if (first || grow) m_ref = mx;
      const float mneg = -m_ref * sl2;
      // 32 scores -> 16 packed probabilities: one SFU exp2 (or FMA-pipe polynomial) per element, fp32 row sums.
      // (ex2.approx.f16x2 was tried for the fp16 build: ptxas lowers it to two scalar MUFU.EX2.F16 plus a repack — no
      // packed SFU exponential exists on sm_100 — so it saves nothing.)
      float sums[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) sums[c] = 0.f;
      auto exp32 = [&](const uint32_t* s32, uint32_t (&pk)[16]) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = exp2_sel<POLY>(j, fmaf(__uint_as_float(s32[j]), sl2, mneg));
          const float e1 = exp2_sel<POLY>(j + 1, fmaf(__uint_as_float(s32[j + 1]), sl2, mneg));
          sums[j & 7] += e0;
          sums[(j + 1) & 7] += e1;
          pk[j >> 1] = pack_bf16(e0, e1);
        }
      };
      // exps of the first 64 keys are computed BEFORE waiting for P V_{j-1}: its tail is hidden under them
      uint32_t pk0[16], pk1[16];
      exp32(&sr[0], pk0);
      exp32(&sr[32], pk1);
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
        exp32(&sr[g * 32], pk);
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
    // all key blocks of the segment done: wait for the last P V, then O / l
    mbar_wait(&p_empty[1], (gb - 1) & 1);
    tc_fence_after();
    if (!PERBLOCK && (j0 != 0 || j1 != nb)) {
      // ---- partial tile: park (O, m_ref, l) in this CTA's workspace slot (0: the segment opens the CTA's range, 1: it
      // closes it), take a ticket on the tile; the last part to arrive merges all parts in CTA order
      float* mine = p.ws_part + (static_cast<long long>(blockIdx.x) * 2 + (seg_at_range_start ? 0 : 1)) * ATS_PART_FLOATS;
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t r[32];
        tmem_ld32(tmem_O + trow + half * 32, r);
        tmem_ld_wait();
        float* dst = mine + row * 64 + half * 32;
#pragma unroll
        for (int d = 0; d < 32; d += 8)
          st_global_256(dst + d, r[d], r[d + 1], r[d + 2], r[d + 3], r[d + 4], r[d + 5], r[d + 6], r[d + 7]);
      }
      mine[128 * 64 + row] = m_ref;
      mine[128 * 64 + 128 + row] = l_run;
      tc_fence_before();
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long t_u0 = static_cast<long long>(tile) * nb;
      const int c_first = static_cast<int>(t_u0 / p.units_per_cta);
      const int c_last = static_cast<int>((t_u0 + nb - 1) / p.units_per_cta);
      if (threadIdx.x == 0) *s_ticket = atomicAdd(&p.ws_ticket[tile], 1u);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const bool last = *s_ticket == static_cast<unsigned int>(c_last - c_first);
      asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone has read the ticket before a later segment overwrites it
      if (last) {
        __threadfence();
        float acc[64];
#pragma unroll
        for (int d = 0; d < 64; ++d) acc[d] = 0.f;
        float m_all = -INFINITY, l_all = 0.f;
#pragma unroll 1
        for (int c = c_first; c <= c_last; ++c) {
          // CTA c's segment on this tile starts at max(c * L, t_u0): it opens c's range (slot 0) iff c * L >= t_u0
          const int slot = static_cast<long long>(c) * p.units_per_cta >= t_u0 ? 0 : 1;
          const float* part = p.ws_part + (static_cast<long long>(c) * 2 + slot) * ATS_PART_FLOATS;
          const float m_p = __ldcg(part + 128 * 64 + row), l_p = __ldcg(part + 128 * 64 + 128 + row);
          const float m_new = fmaxf(m_all, m_p);
          const float a_old = ex2_approx((m_all - m_new) * sl2), a_p = ex2_approx((m_p - m_new) * sl2);
          l_all = l_all * a_old + l_p * a_p;
          m_all = m_new;
          const float4* src = reinterpret_cast<const float4*>(part + row * 64);
#pragma unroll
          for (int d4 = 0; d4 < 16; ++d4) {
            const float4 v = __ldcg(src + d4);
            acc[4 * d4] = acc[4 * d4] * a_old + v.x * a_p;
            acc[4 * d4 + 1] = acc[4 * d4 + 1] * a_old + v.y * a_p;
            acc[4 * d4 + 2] = acc[4 * d4 + 2] * a_old + v.z * a_p;
            acc[4 * d4 + 3] = acc[4 * d4 + 3] * a_old + v.w * a_p;
          }
        }
        if (threadIdx.x == 0) p.ws_ticket[tile] = 0u;  // self-resetting: the workspace is reusable by the next launch
        if (qi < p.n_q) {
          const float w = p.seg_scale[0] / l_all;
#pragma unroll
          for (int d = 0; d < 64; d += 8)
            *reinterpret_cast<uint4*>(optr + d) = make_uint4(pack_bf16(acc[d] * w, acc[d + 1] * w), pack_bf16(acc[d + 2] * w, acc[d + 3] * w),
                                                             pack_bf16(acc[d + 4] * w, acc[d + 5] * w), pack_bf16(acc[d + 6] * w, acc[d + 7] * w));
        }
      }
      continue;
    }
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
    }  // segments
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

extern "C" int64_t iir_attn_workspace_bytes(int B, int heads, int n_q) {
  // tickets (one per 128-query tile) + two parked partial tiles per resident CTA; see attn_ts_kernel
  if (B <= 0 || heads <= 0 || n_q <= 0) return 0;
  const long long tiles = static_cast<long long>((n_q + 127) / 128) * heads * B;
  return 2LL * iir::sm_count() * 2 * iir::ATS_PART_FLOATS * 4 + tiles * 4;
}

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
  IIR_REQUIRE(!a->causal || (a->n_seg == 1 && a->kv_len[0] == a->n_q), "iir_attn_tc: causal needs one segment with kv_len == n_q");
  p.causal = a->causal;
  p.wide_out = (reinterpret_cast<uintptr_t>(a->out) % 32 == 0) && (a->ldo * 2) % 32 == 0 && (a->out_off * 2) % 32 == 0;

  const size_t smem = 7 * TILE_BYTES + 256;
  dim3 grid((a->n_q + 127) / 128, a->heads, a->B);
  // attn_ts kernels are persistent: up to two CTAs per SM, each owning a contiguous range of (tile, key block) units
  const bool one_block_each = a->n_seg == 2 && a->kv_len[0] <= 128 && a->kv_len[1] <= 128;
  const long long ats_tiles = static_cast<long long>(grid.x) * grid.y * grid.z;
  const int ats_nb = one_block_each ? a->n_seg : p.nblk[0];
  const long long slots = 2LL * sm_count();
  p.total_units = ats_tiles * ats_nb;
  static int split_env = -1;
  if (split_env < 0) {
    const char* ev = getenv("IIR_ATTN_SPLIT");  // 0 = tile-granular ranges only (no partial tiles)
    split_env = ev ? atoi(ev) : 1;
  }
  // IIR_ATTN_SPLIT: 0 = never cut tiles; 1 (default) = cut them when whole tiles dealt round robin would leave more than
  // a quarter of the CTA slots idle AND a tile is long enough (>= 32 key blocks) for the balance to outweigh parking and
  // merging the partial tiles (measured, profiles/attn_micro_r02.txt: 4096 tokens x 10 heads, one CFG branch: 87 -> 79 us;
  // 1024 tokens, 8 blocks per tile: 31 -> 40 us); 2 = whenever the unit ranges do not fall on tile boundaries
  const long long rr_waves = (ats_tiles + slots - 1) / slots;
  const bool rr_unbalanced = 4 * ats_tiles < 3 * slots * rr_waves;  // round robin keeps < 75 % of the CTA slots busy
  bool split = split_env && a->n_seg == 1 && a->workspace != nullptr && (split_env == 2 || (ats_nb >= 32 && rr_unbalanced));
  long long L = 0;
  if (split) {
    L = (p.total_units + slots - 1) / slots;
    const long long min_len = (ats_nb + ATS_MAX_PARTS - 2) / (ATS_MAX_PARTS - 1);  // at most ATS_MAX_PARTS parts per tile
    if (L < min_len) L = min_len;
    if (L % ats_nb == 0) split = false;  // the ranges would fall on tile boundaries anyway
  }
  long long ats_ctas;
  if (split) {
    ats_ctas = (p.total_units + L - 1) / L;
    p.units_per_cta = static_cast<int>(L);
    // layout: [2 * sm_count() CTAs][2 slots][ATS_PART_FLOATS] partial tiles at a FIXED offset, then one ticket word per
    // tile — launches of different shapes share the workspace, so the (always-zero-between-launches) tickets must never
    // overlap another launch's partial-tile area
    const long long part_bytes = slots * 2 * ATS_PART_FLOATS * 4;
    const long long need = part_bytes + ats_tiles * 4;
    IIR_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0 && a->workspace_bytes >= need,
                "iir_attn_tc: workspace of %lld bytes (256-byte aligned, zero-initialised) needed, got %lld",
                need, (long long)a->workspace_bytes);
    p.ws_part = reinterpret_cast<float*>(a->workspace);
    p.ws_ticket = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(a->workspace) + part_bytes);
  } else {
    ats_ctas = ats_tiles < slots ? ats_tiles : slots;  // whole tiles, round robin
    p.units_per_cta = 0;
  }
  const dim3 ats_grid(static_cast<unsigned>(ats_ctas));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e;
  static int v1 = -1;
  if (v1 < 0) {
    const char* ev = getenv("IIR_ATTN_V1");
    v1 = (ev && ev[0] == '1') ? 1 : 0;
  }
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
