// tcgen05 GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   out[M, N] = epilogue( A[M, K] · W[N, K]ᵀ )      bf16 operands, fp32 accumulation in TMEM
//
// Design (B200-first):
//   * persistent: one CTA per SM loops over 128 x BN output tiles (BN <= 256, runtime)
//   * warp 0      : TMA producer — A and W tiles land in 128B-swizzled smem stages
//   * warp 1      : single-thread tcgen05.mma issuer (UMMA 128 x BN x 16), accumulators in TMEM,
//                   two accumulator buffers so the epilogue of tile i overlaps the MMAs of tile i+1
//   * warps 2..9  : epilogue — tcgen05.ld rows out of TMEM, fused bias / temb row-vector /
//                   SiLU / GEGLU / SFT / residual add, bf16 or fp32 store.  Two warps per TMEM lane
//                   quarter take alternate 32-column chunks, halving the exposed epilogue of
//                   single-wave problems
//   * Programmatic Dependent Launch: barrier init / TMEM alloc / descriptor prefetch overlap the
//     previous kernel's tail; global memory is touched only after griddepcontrol.wait
//   * clusters    : CL (1/2/4) CTAs that work on the same N tile (consecutive M tiles) form a cluster;
//                   each fetches 1/CL of the weight tile and TMA-multicasts it to its peers, cutting
//                   the L2->SM traffic that bounds a 128 x 256 tile (48 KB per 64-deep k-block) to
//                   16 + 32/CL KB.  MMAs stay cta_group::1; a stage is refilled only after the MMAs of
//                   ALL cluster CTAs retired it (multicast tcgen05.commit onto every CTA's empty barrier)
//   * CTA pairs   : MMA2 = cta_group::2.  A 128 x 256 tile on one SM needs 192 B/clk of shared-memory
//                   traffic (TMA writes + UMMA operand reads share the 128 B/clk port) = 67 % of the tensor
//                   peak — measured 1165 of a predicted 1172 TFLOP/s.  A CTA pair computes a 256 x BN
//                   tile with ONE tcgen05.mma.cta_group::2 stream issued by the leader; each CTA stages
//                   its own 128 A rows and only HALF of the weight tile, i.e. 128 B/clk.  Both CTAs' TMA
//                   loads signal the leader's full barrier; the leader's commits are multicast to both
//                   CTAs' empty / accumulator-full barriers; both epilogues arrive on the leader's
//                   accumulator-empty barrier.
//   * conv mode   : the A tile is a (Nt x Ht x Wt) pixel box of an NHWC tensor fetched with a 4-D
//                   TMA map at the tap offset (kx-1, ky-1); out-of-bounds box elements are
//                   zero-filled by TMA, which *is* the conv's zero padding. No im2col buffer.
//
// Replaces: nn.Linear / nn.Conv2d + following elementwise ops of the reference
// (module/min_sdxl.py:246-283,301-307,502-523,569-573; module/aggregator.py:63-90).
#include <cstdlib>

#include "common.cuh"

namespace iir {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;
// Footprint variants (template parameter EW = epilogue warps):
//   EW = 8: 8 epilogue warps, all 512 TMEM columns (two accumulator buffers), up to 227 KB of shared memory, one CTA/SM;
//   EW = 4 ("half-SM"): 4 epilogue warps, 256 TMEM columns (one accumulator buffer), <= 113 KB, TWO CTAs per SM, so that
//           GEMM CTAs of two streams, or a GEMM CTA and an attention CTA, can share an SM.  Chosen per
//           launch for one-wave problems when IIR_GEMM_HALF=1.
constexpr int ACC_STRIDE = 256;                 // TMEM columns between the two accumulator buffers
constexpr int gemm_threads(int ew) { return 64 + 32 * ew; }  // TMA warp + MMA warp + epilogue warps
constexpr int smem_budget(int ew) { return ew == 4 ? 113 * 1024 : 227 * 1024; }
constexpr int EPI_STAGE_BYTES = 4096;  // per epilogue warp: 32 rows x 32 fp32 columns, XOR-swizzled
constexpr int VEC_BYTES = 4096;        // tile bias + tile column sums, 2 x 256 floats each (double-buffered)
constexpr float LN_S1_SCALE = 4294967296.f;  // 2^32: |row sum| < 2^31
constexpr float LN_S2_SCALE = 16777216.f;    // 2^24: row sum of squares < 2^39

struct alignas(64) GemmTcParams {
  CUtensorMap tmA;
  CUtensorMap tmB;  // box {64, BN / CL}
  int M, N, K, num_kb;
  int conv, n_img, H, W, Cin;
  int Wt, Ht, Nt, tiles_x, tiles_y, tiles_img;
  int tiles_m, tiles_n, BN, stages;
  int epi_slots;   // 4 KiB staging slots per epilogue warp (> 1: the fp32 residual tile is prefetched into them)
  int tiles_m_cl;  // ceil(tiles_m / CL): M super-tiles per N tile
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  int rows_per_sample;
  const void* residual;
  int res_bf16;
  long long ld_res;
  const void* aux;
  int aux_bf16;
  long long ld_aux;
  void* out;
  int out_bf16;
  long long ld_out;
  int act;
  // LayerNorm folded into the GEMMs around it
  long long* ln_stats_out;        // producer: per-row fixed-point (sum, sum of squares) of the final fp32 values
  void* ln_out16;                 // producer: 16-bit copy of the output (the consumer's A operand)
  long long ld_ln16;
  const long long* ln_stats_in;   // consumer: the producer's row sums
  long long* ln_stats_zero;       // consumer: accumulator to clear for the next producer
  const float* ln_colsum;         // consumer: sum_k W'[n, k]
  float ln_inv_k, ln_eps;
  int direct;  // row-per-thread stores straight from registers (no staging transpose)
  int wide_out, wide_ln16;  // rows of out / ln_out16 start on 32-byte boundaries: 256-bit stores
  int probe;  // measurement builds only (-DIIR_GEMM_PROBE): 1 = skip the epilogue body, 2 = skip its global stores,
              // 3 = TMEM loads + row-phase math only, 4 = TMEM loads only
  // GroupNorm statistics of the output accumulated by the epilogue (GNS instantiations only; opt-in)
  unsigned long long* gn_sums;  // [n_samples][gn_groups][2] int64 fixed point (GN_S1_SCALE / GN_S2_SCALE), zero on entry
  int gn_cpg, gn_groups;
};

// One group's partial (sum, sum of squares) of a warp's 32 rows: fixed-order shuffle tree, then two integer atomics by
// lane 0 (integer adds commute: the totals do not depend on the order in which tiles and warps arrive).
__device__ __forceinline__ void gn_flush(float s, float q, unsigned long long* acc, int lane) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if (lane == 0) {
    atomicAdd(acc, static_cast<unsigned long long>(__float2ll_rn(s * GN_S1_SCALE)));
    atomicAdd(acc + 1, static_cast<unsigned long long>(__float2ll_rn(q * GN_S2_SCALE)));
  }
}

struct TileCoord {
  int m0;          // linear: first row
  int n_img0, y0, x0;  // conv: first image / row / column of the pixel box
  int n0;          // first weight row (column of the packed output)
};

// super-tile index -> coordinates of this CTA's 128 x BN tile (CTA `rank` of the cluster takes the
// rank-th M tile of the super-tile; it may lie past the matrix end: TMA zero-fills, stores are masked)
template <int CL>
__device__ __forceinline__ TileCoord tile_coord(const GemmTcParams& p, int stile, int rank) {
  TileCoord c;
  int tn = stile / p.tiles_m_cl;
  int tm = (stile - tn * p.tiles_m_cl) * CL + rank;
  c.n0 = tn * p.BN;
  c.m0 = tm * BM;
  c.n_img0 = c.y0 = c.x0 = 0;
  if (p.conv) {
    int tx = tm % p.tiles_x;
    int t2 = tm / p.tiles_x;
    int ty = t2 % p.tiles_y;
    int ti = t2 / p.tiles_y;
    c.x0 = tx * p.Wt;
    c.y0 = ty * p.Ht;
    c.n_img0 = ti * p.Nt;
  }
  return c;
}

template <int PAIR, int CL, bool MMA2, int EW, bool GNS = false>
__global__ void __launch_bounds__(gemm_threads(EW), EW == 4 ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ GemmTcParams p) {
  static_assert(!GNS || (PAIR == 0 && EW == 8), "GroupNorm statistics: plain epilogue, 8 epilogue warps");
  static_assert(!MMA2 || CL == 2, "cta_group::2 needs a 2-CTA cluster");
  static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
  constexpr int TMEM_COLS = EW == 4 ? 256 : 512;
  constexpr int NBUF = EW == 4 ? 1 : 2;        // accumulator buffers
  constexpr int CH_STRIDE = 32 * (EW / 4);     // column distance between consecutive chunks of one epilogue warp
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  // MMA2: each CTA of the pair stages only its half of the weight tile
  const int stage_bytes = A_STAGE_BYTES + (MMA2 ? p.BN / 2 : p.BN) * BK * 2;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* tfull_bar = empty_bar + p.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_bias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // 2 x 256 floats (tile bias)
  float* s_cs = s_bias + 512;                                                             // 2 x 256 floats (tile column sums)
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(s_bias) + VEC_BYTES;  // 8 epilogue warps x slots x EPI_STAGE_BYTES

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = p.tiles_m_cl * p.tiles_n;            // super-tiles
  const int rank = CL > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int first_tile = blockIdx.x / CL, tile_step = gridDim.x / CL;
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << CL) - 1);
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      // multicast variant: one arrival per cluster CTA whose MMAs read this stage; MMA2: the leader's
      // single commit is multicast to both CTAs
      mbar_init(&empty_bar[s], MMA2 ? 1 : CL);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], (MMA2 ? 2 : 1) * 32 * EW);  // MMA2: both CTAs' epilogues arrive on the leader's
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (MMA2) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (CL > 1) cluster_sync_all();  // peers' barriers are initialised before anyone multicasts / arrives
  pdl_wait();  // everything above overlapped the predecessor; its results are visible from here on

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp runs the loop (warp-uniform control flow keeps coordinates and descriptors in
    // uniform registers: no per-instruction R2UR "waterfall"); one elected lane issues the copies.
    {
      int stage = 0;
      uint32_t phase = 0;
      const int cpb = p.conv ? p.Cin / BK : 1;  // 64-channel blocks per filter tap
      const int b_rows = p.BN / CL;  // weight rows this CTA fetches (and multicasts)
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        TileCoord c = tile_coord<CL>(p, tile, rank);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_sleep(&empty_bar[stage], phase ^ 1, 20000);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + A_STAGE_BYTES;
          int ca0 = kb * BK, ca1 = c.m0, ca2 = 0, ca3 = 0;
          if (p.conv) {
            int tap = kb / cpb;
            int cb = kb - tap * cpb;
            int ky = tap / 3, kx = tap - ky * 3;
            ca0 = cb * BK; ca1 = c.x0 + kx - 1; ca2 = c.y0 + ky - 1; ca3 = c.n_img0;
          }
          if (elect_one()) {
            if (MMA2) {
              // both CTAs' loads complete_tx on the leader's barrier: it expects 2 x stage_bytes
              if (rank == 0) mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(2 * stage_bytes));
              if (p.conv) tma_load_4d_2sm(sa, &p.tmA, &full_bar[stage], ca0, ca1, ca2, ca3);
              else tma_load_2d_2sm(sa, &p.tmA, &full_bar[stage], ca0, ca1);
              tma_load_2d_2sm(sb, &p.tmB, &full_bar[stage], kb * BK, c.n0 + rank * b_rows);
            } else {
              mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(stage_bytes));
              if (p.conv) tma_load_4d(sa, &p.tmA, &full_bar[stage], ca0, ca1, ca2, ca3);
              else tma_load_2d(sa, &p.tmA, &full_bar[stage], ca0, ca1);
              if (CL > 1)
                tma_load_2d_mcast(sb + rank * b_rows * (BK * 2), &p.tmB, &full_bar[stage], kb * BK,
                                  c.n0 + rank * b_rows, kMask);
              else
                tma_load_2d(sb, &p.tmB, &full_bar[stage], kb * BK, c.n0);
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // warp-uniform loop; one elected lane issues the tcgen05.mma / commit instructions
    if (!MMA2 || rank == 0) {
      const uint32_t idesc = umma_idesc_bf16(MMA2 ? 2 * BM : BM, p.BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_i = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++acc_i) {
        const uint32_t buf = acc_i % NBUF;
        const uint32_t acc_phase = (acc_i / NBUF) & 1;
        mbar_wait_sleep(&tempty_bar[buf], acc_phase ^ 1, 20000);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * ACC_STRIDE;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_sleep(&full_bar[stage], phase, 20000);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + A_STAGE_BYTES;
          const uint64_t adesc0 = umma_desc_sw128(sa, 16, 1024);
          const uint64_t bdesc0 = umma_desc_sw128(sb, 16, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advancing K by 16 elements = +32 bytes = +2 in the descriptor's (addr >> 4) field
              if (MMA2) umma2_bf16(tmem_d, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
              else umma_bf16(tmem_d, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // frees the smem stage (in every cluster CTA that multicasts into it) when these MMAs retire
            if (MMA2) umma2_commit_mcast(&empty_bar[stage], kMask);
            else if (CL > 1) umma_commit_mcast(&empty_bar[stage], kMask);
            else umma_commit(&empty_bar[stage]);
            // accumulator complete -> epilogue (of both CTAs when paired)
            if (kb == p.num_kb - 1) {
              if (MMA2) umma2_commit_mcast(&tfull_bar[buf], kMask);
              else umma_commit(&tfull_bar[buf]);
            }
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Two phases per 32-column chunk.  ROW phase: TMEM lane == row, so a thread holds 32 consecutive columns
    // of ONE row (bias, temb row-vector, activation, GEGLU / SFT are applied here).  Writing global memory
    // from that layout touches 32 different lines per instruction, so the chunk is transposed through a
    // private, XOR-swizzled 4 KiB shared-memory tile and the COALESCED phase (8 lanes per row, float4 per
    // lane: four full 128-byte row segments per instruction) adds the residual and stores.
    const int lane_base = (warp & 3) * 32;  // TMEM lane quarter this warp may access
    const int chunk_par = (warp - 2) >> 2;   // which of the two warps of this quarter: odd/even chunks
    const int row = lane_base + lane;
    const int half = p.BN >> 1;
    const int n_out_total = PAIR ? (p.N >> 1) : p.N;
    const int R = p.epi_slots;
    uint8_t* stg = epi_stage + (warp - 2) * R * EPI_STAGE_BYTES;
    // fp32 residual: its tile is fetched into the staging ring with cp.async WHILE the main loop runs, so the
    // epilogue never waits on global-memory latency (a one-wave launch exposed ~10-25 us of it before)
    const bool pre = p.residual != nullptr && !p.res_bf16;
    const int crow = lane >> 3;              // coalesced phase: row within a group of 4
    const int cc4 = lane & 7;                // coalesced phase: float4 column
    uint32_t acc_i = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step, ++acc_i) {
      const uint32_t buf = acc_i % NBUF;
      const uint32_t acc_phase = (acc_i / NBUF) & 1;
      TileCoord c = tile_coord<CL>(p, tile, rank);
      auto row_to_m = [&](int r) -> long long {  // output row of tile row r, or -1 when outside the matrix
        if (p.conv) {
          int xi = r % p.Wt;
          int t2 = r / p.Wt;
          int yi = t2 % p.Ht;
          int ni = t2 / p.Ht;
          int x = c.x0 + xi, y = c.y0 + yi, n = c.n_img0 + ni;
          if (x >= p.W || y >= p.H || n >= p.n_img) return -1;
          return (static_cast<long long>(n) * p.H + y) * p.W + x;
        }
        long long m = c.m0 + r;
        return m < p.M ? m : -1;
      };
      const long long m_own = row_to_m(row);
      const bool valid = m_own >= 0;
      const int sample = valid ? static_cast<int>(m_own / p.rows_per_sample) : 0;
      long long mrow[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mrow[i] = row_to_m(lane_base + i * 4 + crow);
      // the tile's bias vector goes through shared memory once (each row-phase thread needs all of it: 8-16
      // broadcast global loads per chunk before); double-buffered by tile parity, one named barrier per tile
      float* sb = s_bias + (acc_i & 1) * 256;
      float* sc = s_cs + (acc_i & 1) * 256;
      if (p.bias || p.ln_colsum) {
        for (int t = threadIdx.x - 64; t < p.BN; t += 32 * EW) {  // the epilogue warps' threads
          if (p.bias) sb[t] = (c.n0 + t < p.N) ? p.bias[c.n0 + t] : 0.f;
          if (p.ln_colsum) sc[t] = (c.n0 + t < p.N) ? p.ln_colsum[c.n0 + t] : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      }
      // folded LayerNorm, consumer side: y = rstd * (x W'^T) - rstd * mean * colsum(W') + b'  (b' arrives as bias);
      // mean / rstd of this thread's row from the producer's partial sums
      float ln_rstd = 1.f, ln_nm = 0.f;
      if (p.ln_stats_in && valid) {
        const longlong2 t = *reinterpret_cast<const longlong2*>(p.ln_stats_in + m_own * 2);
        const float mean = __ll2float_rn(t.x) * (1.0f / LN_S1_SCALE) * p.ln_inv_k;
        const float var = fmaxf(__ll2float_rn(t.y) * (1.0f / LN_S2_SCALE) * p.ln_inv_k - mean * mean, 0.f);
        ln_rstd = rsqrtf(var + p.ln_eps);
        ln_nm = -mean * ln_rstd;
        // the accumulator the NEXT producer will add into is cleared here, by the tile that owns column 0 of the row
        if (p.ln_stats_zero && c.n0 == 0 && chunk_par == 0)
          *reinterpret_cast<longlong2*>(p.ln_stats_zero + m_own * 2) = make_longlong2(0, 0);
      }
      float ln_s1 = 0.f, ln_s2 = 0.f;  // producer side: this thread's row, this warp's chunks

      const int ncols = PAIR ? half : p.BN;  // accumulator columns that map to output columns
      const int nout0 = PAIR ? (c.n0 >> 1) : c.n0;
      const int nchunks_w = ncols > chunk_par * 32 ? (ncols - chunk_par * 32 + CH_STRIDE - 1) / CH_STRIDE : 0;  // chunks of this warp
      auto prefetch = [&](int k) {
        const int pc = chunk_par * 32 + CH_STRIDE * k;
        const int pcol = nout0 + pc + cc4 * 4;
        const bool ok = (pc + cc4 * 4 < ncols) && (pcol < n_out_total);
        uint8_t* slot = stg + (k % R) * EPI_STAGE_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + crow;
          if (mrow[i] >= 0 && ok)
            cp_async16(slot + rl * 128 + ((cc4 ^ (rl & 7)) << 4),
                       reinterpret_cast<const float*>(p.residual) + mrow[i] * p.ld_res + pcol);
        }
        cp_async_commit();
      };
      if (pre)
        for (int k = 0; k < R && k < nchunks_w; ++k) prefetch(k);

      mbar_wait_sleep(&tfull_bar[buf], acc_phase, 100000);  // a whole main loop: sleep, do not spin (power)
      tc_fence_after();
#ifdef IIR_GEMM_PROBE
      if (p.probe == 1) {  // how long is the kernel without any epilogue work?
        if (pre) cp_async_wait<0>();
        tc_fence_before();
        if (MMA2 && rank != 0) mbar_arrive_remote(&tempty_bar[buf], 0);
        else mbar_arrive(&tempty_bar[buf]);
        continue;
      }
#endif
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_base) << 16) + buf * ACC_STRIDE;
      // Plain epilogues: the TMEM read of chunk k+1 is issued as soon as chunk k has left the registers for the staging
      // tile, so it runs under chunk k's coalesced store phase instead of in front of chunk k+1's math (r is dead by
      // then).  Paired epilogues hold two register tiles per chunk and would spill: they read at the top of the loop.
      constexpr bool EARLY_LD = PAIR == 0;
      uint32_t r[32];
      uint32_t r2[32];
      auto issue_ld = [&](int k) {
        const int c0 = chunk_par * 32 + CH_STRIDE * k;
        tmem_ld32(taddr + c0, r);
        if (PAIR) tmem_ld32(taddr + half + c0, r2);
      };
      if (EARLY_LD && nchunks_w > 0) issue_ld(0);
      for (int kc = 0; kc < nchunks_w; ++kc) {
        const int cc = chunk_par * 32 + CH_STRIDE * kc;
        uint8_t* slot = stg + (kc % R) * EPI_STAGE_BYTES;
        if (!EARLY_LD) issue_ld(kc);
        tmem_ld_wait();
#ifdef IIR_GEMM_PROBE
        if (p.probe == 4) {  // ... with only the TMEM loads?
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
          if (acc == 1.2345e-33f) reinterpret_cast<float*>(p.out)[0] = 0.f;  // keep the load alive
          if (EARLY_LD && kc + 1 < nchunks_w) issue_ld(kc + 1);
          continue;
        }
#endif
        float v[32];
        const int pn = c.n0 + cc;  // packed column of r[0]
        const int on = nout0 + cc;  // output column of v[0]
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.ln_stats_in) {  // rstd * acc + (-mean * rstd) * colsum + b'  (b' is always present on this path)
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 s4 = *reinterpret_cast<const float4*>(sc + cc + j);
            const float4 b = *reinterpret_cast<const float4*>(sb + cc + j);
            v[j] = fmaf(ln_rstd, v[j], fmaf(ln_nm, s4.x, b.x)); v[j + 1] = fmaf(ln_rstd, v[j + 1], fmaf(ln_nm, s4.y, b.y));
            v[j + 2] = fmaf(ln_rstd, v[j + 2], fmaf(ln_nm, s4.z, b.z)); v[j + 3] = fmaf(ln_rstd, v[j + 3], fmaf(ln_nm, s4.w, b.w));
          }
        } else if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 b = *reinterpret_cast<const float4*>(sb + cc + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (p.rowvec) {
          const float* rv = p.rowvec + static_cast<long long>(sample) * p.ld_rowvec + pn;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (pn + j < p.N) {
              float4 b = ld4(rv + j);
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          }
        }
        if (p.act == IIR_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j]);
        } else if (p.act == IIR_ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf_fast(v[j]);
        } else if (p.act == IIR_ACT_QUICK_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = quick_gelu_fast(v[j]);
        }
        if (PAIR) {
          float g[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) g[j] = __uint_as_float(r2[j]);
          if (p.ln_stats_in) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 s4 = *reinterpret_cast<const float4*>(sc + half + cc + j);
              const float4 b = *reinterpret_cast<const float4*>(sb + half + cc + j);
              g[j] = fmaf(ln_rstd, g[j], fmaf(ln_nm, s4.x, b.x)); g[j + 1] = fmaf(ln_rstd, g[j + 1], fmaf(ln_nm, s4.y, b.y));
              g[j + 2] = fmaf(ln_rstd, g[j + 2], fmaf(ln_nm, s4.z, b.z)); g[j + 3] = fmaf(ln_rstd, g[j + 3], fmaf(ln_nm, s4.w, b.w));
            }
          } else if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 b = *reinterpret_cast<const float4*>(sb + half + cc + j);
              g[j] += b.x; g[j + 1] += b.y; g[j + 2] += b.z; g[j + 3] += b.w;
            }
          }
          if (PAIR == IIR_PAIR_GEGLU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] * gelu_erf_fast(g[j]);
          } else {  // SFT: h * (gamma + 1) + beta
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (valid && on + j < n_out_total) {
                float4 h = p.aux_bf16
                    ? ld4(reinterpret_cast<const h16*>(p.aux) + m_own * p.ld_aux + on + j)
                    : ld4(reinterpret_cast<const float*>(p.aux) + m_own * p.ld_aux + on + j);
                v[j] = h.x * (v[j] + 1.0f) + g[j];
                v[j + 1] = h.y * (v[j + 1] + 1.0f) + g[j + 1];
                v[j + 2] = h.z * (v[j + 2] + 1.0f) + g[j + 2];
                v[j + 3] = h.w * (v[j + 3] + 1.0f) + g[j + 3];
              }
            }
          }
        }
#ifdef IIR_GEMM_PROBE
        if (p.probe == 3) {  // ... with the TMEM loads and the row-phase math, but no staging / stores?
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) acc += v[j];
          if (acc == 1.2345e-33f) reinterpret_cast<float*>(p.out)[0] = acc;
          if (EARLY_LD && kc + 1 < nchunks_w) issue_ld(kc + 1);
          continue;
        }
#endif
        // ---- DIRECT stores (default): every thread writes its own row's 32 consecutive outputs straight from registers
        // with 256-bit stores (one full 32-byte sector per lane per instruction: 2 for 16-bit, 4 for fp32 outputs).
        // Measured against the transposed, 4-rows-per-instruction path below (tools/probe_epilogue.py, bench_lnfold.py):
        // QKV 2048x3840x1280 21.1 -> 17.3 us, GEGLU 2048x10240x1280 42.8 -> 39.5 us, 8192x5120x640 64.2 -> 51.9 us — the
        // tile's trip through shared memory (write + syncwarp + read) was the largest part of the exposed epilogue.  An
        // fp32 residual still arrives through the cp.async staging ring (reading it row-per-thread from global memory is
        // latency-bound) and is added in registers.
        if (p.direct) {
          if (pre) {
            const int issued = kc + R < nchunks_w ? kc + R : nchunks_w;
            const int pending = issued - (kc + 1);
            if (pending <= 0) cp_async_wait<0>();
            else if (pending == 1) cp_async_wait<1>();
            else if (pending == 2) cp_async_wait<2>();
            else cp_async_wait<3>();
            __syncwarp();
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 q = *reinterpret_cast<const float4*>(slot + lane * 128 + ((c4 ^ (lane & 7)) << 4));
              v[4 * c4] += q.x; v[4 * c4 + 1] += q.y; v[4 * c4 + 2] += q.z; v[4 * c4 + 3] += q.w;
            }
          }
          if (valid) {
            const int lim = min(ncols - cc, n_out_total - on);  // valid columns of this chunk (a multiple of 8)
            if (p.ln_stats_out) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (j < lim) {
                  ln_s1 += (v[j] + v[j + 1]) + (v[j + 2] + v[j + 3]);
                  ln_s2 += (v[j] * v[j] + v[j + 1] * v[j + 1]) + (v[j + 2] * v[j + 2] + v[j + 3] * v[j + 3]);
                }
              }
            }
            auto st16 = [&](h16* dst, bool wide) {  // 32 outputs as 16-bit: 2 x 32 B (or 4 x 16 B when rows are not 32 B aligned)
              if (wide) {
#pragma unroll
                for (int j = 0; j < 32; j += 16)
                  if (j + 8 < lim)
                    st_global_256(dst + j, pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                                  pack_bf16(v[j + 6], v[j + 7]), pack_bf16(v[j + 8], v[j + 9]), pack_bf16(v[j + 10], v[j + 11]),
                                  pack_bf16(v[j + 12], v[j + 13]), pack_bf16(v[j + 14], v[j + 15]));
                  else if (j < lim)
                    *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                                    pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 8)
                  if (j < lim)
                    *reinterpret_cast<uint4*>(dst + j) = make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                                                                    pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
              }
            };
            if (p.out_bf16) {
              st16(reinterpret_cast<h16*>(p.out) + m_own * p.ld_out + on, p.wide_out);
            } else {
              float* orow = reinterpret_cast<float*>(p.out) + m_own * p.ld_out + on;
              if (p.wide_out) {
#pragma unroll
                for (int j = 0; j < 32; j += 8)
                  if (j + 4 < lim)
                    st_global_256(orow + j, __float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]),
                                  __float_as_uint(v[j + 3]), __float_as_uint(v[j + 4]), __float_as_uint(v[j + 5]),
                                  __float_as_uint(v[j + 6]), __float_as_uint(v[j + 7]));
                  else if (j < lim)
                    *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  if (j < lim) *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
            }
            if (p.ln_out16) st16(reinterpret_cast<h16*>(p.ln_out16) + m_own * p.ld_ln16 + on, p.wide_ln16);
          }
          if constexpr (GNS) {
            // GroupNorm partial sums of this warp's 32 rows x (up to) 32 final output values each.  The host guarantees
            // that the 32 rows belong to ONE sample; rows outside the matrix contribute zeros.  Group boundaries fall on
            // even columns (gn_cpg is even, chunks start on multiples of 32), so the branch below is warp-uniform.
            const int glim = min(ncols - cc, n_out_total - on);  // valid columns of this chunk
            const int smp = __reduce_max_sync(0xffffffffu, valid ? sample : 0);
            unsigned long long* gacc = p.gn_sums + static_cast<long long>(smp) * p.gn_groups * 2;
            int g = on / p.gn_cpg;
            int next = (g + 1) * p.gn_cpg - on;  // first column (relative to `on`) of the next group
            float gs = 0.f, gq = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              if (j == next && j < glim) {  // a group ends in front of column j and valid columns follow
                gn_flush(gs, gq, gacc + 2 * g, lane);
                gs = gq = 0.f;
                ++g;
                next += p.gn_cpg;
              }
              if (valid && j < glim) {
                gs += v[j] + v[j + 1];
                gq = fmaf(v[j], v[j], fmaf(v[j + 1], v[j + 1], gq));
              }
            }
            if (glim > 0) gn_flush(gs, gq, gacc + 2 * g, lane);
          }
          if (EARLY_LD && kc + 1 < nchunks_w) issue_ld(kc + 1);
          if (pre) {
            __syncwarp();  // every lane has read its row of the slot before the next prefetch overwrites it
            if (kc + R < nchunks_w) prefetch(kc + R);
          }
          continue;
        }
        // ---- transpose through the swizzled staging tile: row `lane`, 16-byte unit (c4 ^ (lane & 7))
        if (pre) {
          // groups still allowed in flight: the chunks after this one that were already requested
          const int issued = kc + R < nchunks_w ? kc + R : nchunks_w;
          const int pending = issued - (kc + 1);
          if (pending <= 0) cp_async_wait<0>();
          else if (pending == 1) cp_async_wait<1>();
          else if (pending == 2) cp_async_wait<2>();
          else cp_async_wait<3>();
          __syncwarp();
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4* a4 = reinterpret_cast<float4*>(slot + lane * 128 + ((c4 ^ (lane & 7)) << 4));
            float4 q = *a4;
            q.x += v[4 * c4]; q.y += v[4 * c4 + 1]; q.z += v[4 * c4 + 2]; q.w += v[4 * c4 + 3];
            *a4 = q;
            if (p.ln_stats_out && on + 4 * c4 < n_out_total) {  // columns past N hold stale staging data
              ln_s1 += (q.x + q.y) + (q.z + q.w);
              ln_s2 += (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
            }
          }
        } else {
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 q = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            *reinterpret_cast<float4*>(slot + lane * 128 + ((c4 ^ (lane & 7)) << 4)) = q;
            if (p.ln_stats_out && on + 4 * c4 < n_out_total) {
              ln_s1 += (q.x + q.y) + (q.z + q.w);
              ln_s2 += (q.x * q.x + q.y * q.y) + (q.z * q.z + q.w * q.w);
            }
          }
        }
        __syncwarp();
        if (EARLY_LD && kc + 1 < nchunks_w) issue_ld(kc + 1);
        const int ocol = on + cc4 * 4;
        const bool col_ok = (cc + cc4 * 4 < ncols) && (ocol < n_out_total);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rl = i * 4 + crow;
          float4 x = *reinterpret_cast<const float4*>(slot + rl * 128 + ((cc4 ^ (rl & 7)) << 4));
          const long long m = mrow[i];
#ifdef IIR_GEMM_PROBE
          if (p.probe == 2) continue;  // ... and without the global stores?
#endif
          if (m >= 0 && col_ok) {
            if (p.residual && !pre) {
              float4 q = p.res_bf16
                  ? ld4(reinterpret_cast<const h16*>(p.residual) + m * p.ld_res + ocol)
                  : ld4(reinterpret_cast<const float*>(p.residual) + m * p.ld_res + ocol);
              x.x += q.x; x.y += q.y; x.z += q.z; x.w += q.w;
            }
            if (p.out_bf16) st4(reinterpret_cast<h16*>(p.out) + m * p.ld_out + ocol, x);
            else st4(reinterpret_cast<float*>(p.out) + m * p.ld_out + ocol, x);
            if (p.ln_out16) st4(reinterpret_cast<h16*>(p.ln_out16) + m * p.ld_ln16 + ocol, x);
          }
        }
        __syncwarp();
        if (pre && kc + R < nchunks_w) prefetch(kc + R);
      }
      if (p.ln_stats_out && valid) {
        // fixed-point accumulation: integer adds commute, so the row sums do not depend on the order in which
        // the N tiles (and the two warps of a lane quarter) arrive — results stay run-to-run bit-identical
        unsigned long long* acc = reinterpret_cast<unsigned long long*>(p.ln_stats_out) + m_own * 2;
        atomicAdd(acc, static_cast<unsigned long long>(__float2ll_rn(ln_s1 * LN_S1_SCALE)));
        atomicAdd(acc + 1, static_cast<unsigned long long>(__float2ll_rn(ln_s2 * LN_S2_SCALE)));
      }
      tc_fence_before();
      // the leader's MMA thread waits for both CTAs' epilogues before it reuses this accumulator buffer.  When no later
      // tile will reuse it nobody waits for the arrival, and the remote one is skipped: no cross-CTA traffic may still
      // be in flight towards the leader when the (execution-only) cluster rendezvous at the end lets it exit
      if (MMA2 && rank != 0) {
        if (tile + NBUF * tile_step < num_tiles) mbar_arrive_remote(&tempty_bar[buf], 0);
      } else {
        mbar_arrive(&tempty_bar[buf]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_relaxed();  // no CTA exits while a peer may still multicast into it / arrive on it
  if (warp == 1) {
    tc_fence_after();
    if (MMA2) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int pow2_floor(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}
int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r *= 2;
  return r;
}

}  // namespace

}  // namespace iir

extern "C" int iir_gemm_tc(const iir_gemm_args* a, void* stream) {
  using namespace iir;
  IIR_REQUIRE(a != nullptr, "iir_gemm_tc: null args");
  IIR_REQUIRE(dtype_ok(a->out_dtype) && (!a->residual || dtype_ok(a->res_dtype)) && (!a->aux || dtype_ok(a->aux_dtype)), "iir_gemm_tc: unsupported dtype for this library build (fp32 or %s only)", IIR_H16 == IIR_F16 ? "fp16" : "bf16");
  IIR_REQUIRE(a->a_dtype == IIR_H16 && a->w_dtype == IIR_H16,
              "iir_gemm_tc: operands must be bf16 (use iir_gemm_simt for the fp32 check mode)");
  IIR_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "iir_gemm_tc: empty problem M=%d N=%d K=%d", a->M,
              a->N, a->K);
  IIR_REQUIRE(a->K % 8 == 0, "iir_gemm_tc: K=%d must be a multiple of 8", a->K);
  IIR_REQUIRE(a->bn >= 32 && a->bn <= 256 && a->bn % 32 == 0, "iir_gemm_tc: bad bn=%d", a->bn);
  IIR_REQUIRE(a->pair == IIR_PAIR_NONE || (a->bn % 64 == 0 && a->N % a->bn == 0),
              "iir_gemm_tc: paired epilogue needs bn%%64==0 and N%%bn==0 (N=%d bn=%d)", a->N, a->bn);
  IIR_REQUIRE((reinterpret_cast<uintptr_t>(a->a) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->w) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(a->out) & 15) == 0,
              "iir_gemm_tc: pointers must be 16-byte aligned");
  const int n_out = a->pair ? a->N / 2 : a->N;
  IIR_REQUIRE(n_out % 8 == 0 && a->ld_out % 8 == 0, "iir_gemm_tc: N_out=%d / ld_out must be multiples of 8", n_out);
  IIR_REQUIRE(!a->residual || a->ld_res % 4 == 0, "iir_gemm_tc: ld_res must be a multiple of 4");
  IIR_REQUIRE(a->pair != IIR_PAIR_SFT || a->aux, "iir_gemm_tc: SFT epilogue needs aux (h)");
  IIR_REQUIRE(a->rows_per_sample > 0 || !a->rowvec, "iir_gemm_tc: rowvec needs rows_per_sample");
  IIR_REQUIRE(!a->rowvec || (a->ld_rowvec % 4 == 0 && (reinterpret_cast<uintptr_t>(a->rowvec) & 15) == 0), "iir_gemm_tc: rowvec / ld_rowvec must be 16-byte aligned");

  GemmTcParams p;
  memset(&p, 0, sizeof(p));
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.num_kb = (a->K + BK - 1) / BK;
  p.BN = a->bn;
  p.tiles_n = (a->N + a->bn - 1) / a->bn;
  p.conv = a->conv ? 1 : 0;
  int cl = 1;         // cluster size along M; decided below once tiles_m is known
  bool mma2 = false;  // cta_group::2 CTA pairs
  CUresult cr;
  if (a->conv) {
    IIR_REQUIRE(a->conv == 3 && a->stride == 1 && a->up2 == 0 && !a->conv_asym,
                "iir_gemm_tc: conv mode supports 3x3 stride 1 only");
    IIR_REQUIRE(a->Cin % BK == 0 && a->K == 9 * a->Cin, "iir_gemm_tc: conv needs Cin%%64==0, K==9*Cin");
    IIR_REQUIRE(a->M == a->n_img * a->H * a->W, "iir_gemm_tc: conv M mismatch");
    p.n_img = a->n_img; p.H = a->H; p.W = a->W; p.Cin = a->Cin;
    int Wt = 0;
    for (int cand = 128; cand >= 8; cand >>= 1)
      if (a->W % cand == 0) { Wt = cand; break; }
    if (Wt == 0) Wt = pow2_ceil(a->W) > 128 ? 128 : pow2_ceil(a->W);
    int Ht = 128 / Wt;
    if (Ht > pow2_ceil(a->H)) Ht = pow2_ceil(a->H);
    int Nt = 128 / (Wt * Ht);
    p.Wt = Wt; p.Ht = Ht; p.Nt = Nt;
    p.tiles_x = (a->W + Wt - 1) / Wt;
    p.tiles_y = (a->H + Ht - 1) / Ht;
    p.tiles_img = (a->n_img + Nt - 1) / Nt;
    p.tiles_m = p.tiles_x * p.tiles_y * p.tiles_img;
    uint64_t dims[4] = {(uint64_t)a->Cin, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
    uint64_t strides[3] = {(uint64_t)a->Cin * 2, (uint64_t)a->W * a->Cin * 2,
                           (uint64_t)a->H * a->W * a->Cin * 2};
    uint32_t box[4] = {BK, (uint32_t)Wt, (uint32_t)Ht, (uint32_t)Nt};
    cr = encode_tiled(&p.tmA, IIR_H16_TMA, 4, a->a, dims, strides, box,
                      CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    IIR_REQUIRE(a->lda % 8 == 0 && a->lda >= a->K, "iir_gemm_tc: lda=%lld invalid", (long long)a->lda);
    p.tiles_m = (a->M + BM - 1) / BM;
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->M};
    uint64_t strides[1] = {(uint64_t)a->lda * 2};
    uint32_t box[2] = {BK, BM};
    cr = encode_tiled(&p.tmA, IIR_H16_TMA, 2, a->a, dims, strides, box,
                      CU_TENSOR_MAP_SWIZZLE_128B);
  }
  if (cr != CUDA_SUCCESS) {
    set_error("iir_gemm_tc: cuTensorMapEncodeTiled(A) failed (%d)", (int)cr);
    return IIR_ERR_CUDA;
  }
  {
    static int cl_env = -1;
    if (cl_env < 0) {
      const char* e = getenv("IIR_GEMM_CLUSTER");
      cl_env = e ? atoi(e) : 0;  // 0 = automatic
    }
    // IIR_GEMM_CLUSTER: 0/unset = automatic (CTA pairs with cta_group::2 when there are >= 2 M tiles),
    // 1 = single CTA, 2/4 = cluster multicast with cta_group::1 MMAs, 22 = force CTA pairs
    if (cl_env == 1 || cl_env == 2 || cl_env == 4) cl = cl_env;
    else if (cl_env == 22 || a->cluster == 2) { cl = p.tiles_m >= 2 ? 2 : 1; mma2 = cl == 2; }
    else if (a->cluster == 1) cl = 1;
    else {
      // CTA pairs pay two cluster syncs (~1-2 us): worth it once the main loop is long enough to be
      // shared-memory-bandwidth bound (measured: wins for N >= 2560 with K >= 1024 and for long-K convs)
      const double flops = 2.0 * a->M * a->N * a->K;
      const bool big = (a->N >= 2560 && a->K >= 1024) || (a->conv && flops >= 6e10) || flops >= 2e11;
      cl = (big && p.tiles_m >= 2) ? 2 : 1;
      mma2 = cl == 2;
    }
    while (cl > 1 && (p.tiles_m < cl || (a->bn / cl) % 8 != 0)) cl >>= 1;
    if (cl != 2) mma2 = false;
    p.tiles_m_cl = (p.tiles_m + cl - 1) / cl;
    uint64_t dims[2] = {(uint64_t)a->K, (uint64_t)a->N};
    uint64_t strides[1] = {(uint64_t)a->K * 2};
    uint32_t box[2] = {BK, (uint32_t)(a->bn / cl)};
    cr = encode_tiled(&p.tmB, IIR_H16_TMA, 2, a->w, dims, strides, box,
                      CU_TENSOR_MAP_SWIZZLE_128B);
    if (cr != CUDA_SUCCESS) {
      set_error("iir_gemm_tc: cuTensorMapEncodeTiled(W) failed (%d)", (int)cr);
      return IIR_ERR_CUDA;
    }
  }
  p.bias = a->bias;
  p.rowvec = a->rowvec;
  p.ld_rowvec = a->ld_rowvec > 0 ? a->ld_rowvec : a->N;
  p.rows_per_sample = a->rows_per_sample > 0 ? a->rows_per_sample : a->M;
  p.residual = a->residual; p.res_bf16 = a->res_dtype == IIR_H16; p.ld_res = a->ld_res;
  p.aux = a->aux; p.aux_bf16 = a->aux_dtype == IIR_H16; p.ld_aux = a->ld_aux;
  p.out = a->out; p.out_bf16 = a->out_dtype == IIR_H16; p.ld_out = a->ld_out;
  p.act = a->act;
  {
    // a residual that is NOT prefetched through the staging ring (16-bit residual) keeps the transposed path: its loads
    // would be row-per-thread from global memory.  IIR_GEMM_DIRECT: 0 = transposed stores everywhere, 1 = direct for
    // 16-bit outputs only, 2 (default) = direct for fp32 outputs too (with 256-bit stores: 2048x1280x1280 + residual
    // 12.8 -> 11.7 us; with the 16-bit LayerNorm copy on top it is a wash, 13.1 vs 13.2 us)
    static int direct_env = -1;
    if (direct_env < 0) {
      const char* e = getenv("IIR_GEMM_DIRECT");
      direct_env = e ? atoi(e) : 2;
    }
    p.direct = direct_env && (a->out_dtype == IIR_H16 || direct_env == 2) && !(a->residual && a->res_dtype != IIR_F32);
    const long long esz = a->out_dtype == IIR_H16 ? 2 : 4;
    p.wide_out = (reinterpret_cast<uintptr_t>(a->out) % 32 == 0) && (a->ld_out * esz) % 32 == 0;
    p.wide_ln16 = a->ln_out16 && (reinterpret_cast<uintptr_t>(a->ln_out16) % 32 == 0) && (a->ld_ln_out16 * 2) % 32 == 0;
  }
#ifdef IIR_GEMM_PROBE
  {
    const char* e = getenv("IIR_GEMM_PROBE_MODE");  // read per call: the probe tool switches it between graphs
    p.probe = e ? atoi(e) : 0;
  }
#endif
  if (a->ln_stats_out) {
    IIR_REQUIRE(a->pair == IIR_PAIR_NONE && (!a->residual || a->res_dtype == IIR_F32) && !a->conv,
                "iir_gemm_tc: ln_stats_out needs a plain linear epilogue with no or an fp32 residual");
    IIR_REQUIRE((reinterpret_cast<uintptr_t>(a->ln_stats_out) & 15) == 0, "iir_gemm_tc: ln_stats_out must be 16-byte aligned");
    IIR_REQUIRE(!a->ln_out16 || (a->ld_ln_out16 % 8 == 0 && (reinterpret_cast<uintptr_t>(a->ln_out16) & 15) == 0),
                "iir_gemm_tc: ln_out16 / ld_ln_out16 must be 16-byte aligned");
    p.ln_stats_out = reinterpret_cast<long long*>(a->ln_stats_out); p.ln_out16 = a->ln_out16; p.ld_ln16 = a->ld_ln_out16;
  } else {
    IIR_REQUIRE(!a->ln_out16, "iir_gemm_tc: ln_out16 needs ln_stats_out");
  }
  if (a->ln_stats_in) {
    IIR_REQUIRE(a->ln_colsum && a->bias && !a->conv && a->pair != IIR_PAIR_SFT,
                "iir_gemm_tc: ln_stats_in needs ln_colsum, bias (= W beta + b) and a linear problem");
    IIR_REQUIRE((reinterpret_cast<uintptr_t>(a->ln_stats_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->ln_stats_zero) & 15) == 0,
                "iir_gemm_tc: ln_stats_in / ln_stats_zero must be 16-byte aligned");
    p.ln_stats_in = reinterpret_cast<const long long*>(a->ln_stats_in);
    p.ln_stats_zero = reinterpret_cast<long long*>(a->ln_stats_zero);
    p.ln_colsum = a->ln_colsum;
    p.ln_inv_k = 1.0f / static_cast<float>(a->K); p.ln_eps = a->ln_eps;
  } else {
    IIR_REQUIRE(!a->ln_stats_zero, "iir_gemm_tc: ln_stats_zero needs ln_stats_in");
  }

  if (a->gn_sums) {
    // GroupNorm statistics from this launch's epilogue (opt-in): see iir_gemm_args.gn_sums
    IIR_REQUIRE(a->pair == IIR_PAIR_NONE && !a->ln_stats_out && p.direct,
                "iir_gemm_tc: gn_sums needs a plain epilogue on the direct-store path (no pair, no LayerNorm producer, no 16-bit residual)");
    IIR_REQUIRE(a->gn_cpg >= 2 && a->gn_cpg % 2 == 0 && a->gn_groups > 0 && a->gn_groups * a->gn_cpg == a->N,
                "iir_gemm_tc: gn_sums needs N == gn_groups * gn_cpg with an even gn_cpg (N=%d groups=%d cpg=%d)", a->N, a->gn_groups, a->gn_cpg);
    IIR_REQUIRE((reinterpret_cast<uintptr_t>(a->gn_sums) & 7) == 0, "iir_gemm_tc: gn_sums must be 8-byte aligned");
    IIR_REQUIRE(cl == 1 || mma2, "iir_gemm_tc: gn_sums is built for single CTAs and cta_group::2 pairs only");
    // the 32 rows of an epilogue warp must belong to one sample
    if (a->conv) IIR_REQUIRE(p.Wt * p.Ht >= 32, "iir_gemm_tc: gn_sums needs >= 32 pixels of one image per tile (H=%d W=%d)", a->H, a->W);
    else IIR_REQUIRE(p.rows_per_sample % 32 == 0, "iir_gemm_tc: gn_sums needs rows_per_sample %% 32 == 0 (got %d)", p.rows_per_sample);
    p.gn_sums = reinterpret_cast<unsigned long long*>(a->gn_sums);
    p.gn_cpg = a->gn_cpg; p.gn_groups = a->gn_groups;
  }

  const int stage_bytes = A_STAGE_BYTES + (mma2 ? a->bn / 2 : a->bn) * BK * 2;
  // half-SM variant: only for launches that are a single wave of CTAs anyway (one accumulator buffer: no epilogue /
  // main-loop overlap between tiles of a CTA) and whose pipeline still gets >= 3 stages in 113 KB
  static int half_env = -1;
  if (half_env < 0) {
    const char* e = getenv("IIR_GEMM_HALF");
    half_env = e ? atoi(e) : 0;
  }
  int ew = 8;
  if (half_env == 1 && !a->gn_sums && p.tiles_m_cl * p.tiles_n * cl <= sm_count() &&
      (smem_budget(4) - 1024 - 256 - VEC_BYTES - 4 * EPI_STAGE_BYTES) / stage_bytes >= 3)
    ew = 4;
  const int CH_STRIDE = 32 * (ew / 4), SMEM_BUDGET = smem_budget(ew), EW = ew;
  const bool HALF_SM = ew == 4;
  // staging slots per epilogue warp: enough for this warp's chunks of the tile when an fp32 residual is
  // prefetched (at most 3), but never at the price of a shallow main-loop pipeline (measured: a CTA pair
  // wants >= 6 stages of 26 KB at K = 5120, one CTA >= 4 of 36 KB)
  int slots = 1;
  if (a->residual && a->res_dtype == IIR_F32) {
    const int ncols = a->pair ? a->bn / 2 : a->bn;
    slots = (ncols + CH_STRIDE - 1) / CH_STRIDE;
    if (slots > 3) slots = 3;
    int want = mma2 ? 6 : 4;
    if (want > p.num_kb + 1) want = p.num_kb + 1;
    if (HALF_SM) want = 3;
    while (slots > 1 && (SMEM_BUDGET - 1024 - 256 - VEC_BYTES - EW * slots * EPI_STAGE_BYTES) / stage_bytes < want) --slots;
  }
  p.epi_slots = slots;
  // the direct-store epilogue touches the staging tiles only to receive a prefetched fp32 residual: without one their
  // 32 KB go to the main-loop pipeline (one more stage for a 256-wide pair tile)
  const bool need_staging = !p.direct || (a->residual && a->res_dtype == IIR_F32);
  const int epi_bytes = need_staging ? EW * slots * EPI_STAGE_BYTES : 0;
  int stages = (SMEM_BUDGET - 1024 - 256 - VEC_BYTES - epi_bytes) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > p.num_kb + 1) stages = p.num_kb + 1 < 2 ? 2 : p.num_kb + 1;
  p.stages = stages;
  size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + VEC_BYTES + epi_bytes;
  if (!HALF_SM && smem < 120 * 1024) smem = 120 * 1024;  // force one CTA per SM (each allocates all of TMEM)
  IIR_REQUIRE(stages >= 2, "iir_gemm_tc: tile bn=%d does not fit the shared-memory budget of this build", a->bn);

  const int num_tiles = p.tiles_m_cl * p.tiles_n;  // super-tiles, one per cluster at a time
  int grid = (sm_count() / cl) * cl;
  if (grid > num_tiles * cl) grid = num_tiles * cl;

  cudaError_t e;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define LAUNCH3(PAIRV, CLV, M2, EWV)                                                                            \
  e = cudaFuncSetAttribute(gemm_tc_kernel<PAIRV, CLV, M2, EWV>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                           (int)smem);                                                                           \
  if (e == cudaSuccess)                                                                                          \
    e = launch_cluster_pdl(gemm_tc_kernel<PAIRV, CLV, M2, EWV>, dim3(grid), dim3(gemm_threads(EWV)), smem, st, CLV, p);
#define LAUNCH2(PAIRV, CLV, M2)               \
  if (ew == 4) { LAUNCH3(PAIRV, CLV, M2, 4) } \
  else { LAUNCH3(PAIRV, CLV, M2, 8) }
#define LAUNCH(PAIRV)                                \
  if (mma2) { LAUNCH2(PAIRV, 2, true) }              \
  else if (cl == 4) { LAUNCH2(PAIRV, 4, false) }     \
  else if (cl == 2) { LAUNCH2(PAIRV, 2, false) }     \
  else { LAUNCH2(PAIRV, 1, false) }
#define LAUNCH_GNS(CLV, M2)                                                                                       \
  e = cudaFuncSetAttribute(gemm_tc_kernel<0, CLV, M2, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                           (int)smem);                                                                            \
  if (e == cudaSuccess)                                                                                           \
    e = launch_cluster_pdl(gemm_tc_kernel<0, CLV, M2, 8, true>, dim3(grid), dim3(gemm_threads(8)), smem, st, CLV, p);
  if (a->gn_sums) {
    if (mma2) { LAUNCH_GNS(2, true) }
    else { LAUNCH_GNS(1, false) }
  }
  else if (a->pair == IIR_PAIR_NONE) { LAUNCH(0) }
  else if (a->pair == IIR_PAIR_GEGLU) { LAUNCH(1) }
  else if (a->pair == IIR_PAIR_SFT) { LAUNCH(2) }
  else { set_error("iir_gemm_tc: bad pair=%d", a->pair); return IIR_ERR_INVALID; }
#undef LAUNCH_GNS
#undef LAUNCH
#undef LAUNCH2
#undef LAUNCH3
  if (e != cudaSuccess) {
    set_error("iir_gemm_tc: %s", cudaGetErrorString(e));
    return IIR_ERR_CUDA;
  }
  count_launch();
  return check_launch("iir_gemm_tc");
}
