// gemm_tcgen05.cu — persistent, warp-specialised bf16 GEMM for sm_100a with fused epilogues.
//
//   C[M,N] = A[M,K] · W[N,K]^T   (A activations row-major, W = nn.Linear weight, both K-major)
//
// This one kernel replaces every cuBLASLt call + trailing elementwise kernel that the reference
// reaches through timm / nn.Sequential:
//   timm Attention.qkv / proj, Mlp.fc1 / fc2, LayerScale, residual add, PatchEmbed (as GEMM)
//   prismatic/util/nn_utils.py:42-48 (FusedMLPProjector Linear+GELU chain)
//
// Structure (one CTA, or a cta_group::2 CTA pair, per SM; static persistent tile schedule):
//   warps 0-7   epilogue       (tcgen05.ld → registers → fused bias/GELU/LayerScale/residual → global)
//   warp 8      TMA producer   (cp.async.bulk.tensor → 128B-swizzled smem ring, mbarrier tx-count)
//   warp 9      MMA issuer     (one elected lane issues tcgen05.mma, accumulators in TMEM)
//   warp 10     TMEM allocator
//   warp 11     idle
// The control warps carry the HIGHEST warp ids on purpose: the SM sub-partition arbiter prefers the highest
// eligible warp id, so a TMA / MMA issue never queues behind the (ALU-heavy) epilogue warps, and their loops are
// warp-uniform with a single elected lane so that descriptors stay in uniform registers.
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "gemm.h"
#include "ptx.cuh"

#ifndef BLB_GEMM_EPI_SLEEP_NS
#define BLB_GEMM_EPI_SLEEP_NS 0
#endif

namespace blb {

constexpr int BM = 128;          // rows per CTA (UMMA M = 128 per CTA; 256 for a CTA pair)
constexpr int BK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = 128 + NUM_EPI_WARPS * 32;

constexpr int RCHUNK_BYTES = 32 * 32 * 4;   // one epilogue warp's [32 rows x 32 cols] fp32 residual chunk

constexpr int OCHUNK_BYTES = 32 * 32 * 2;   // one epilogue warp's [32 rows x 32 cols] bf16 output staging chunk
constexpr int VEC_BYTES = 2 * 128 * 4;      // one epilogue warp's per-tile copy of two per-column fp32 vectors (<= 128 cols)

template <int BN, int CTAS, int RTMA = 0, bool OBUF = false, bool FBUF = false>
struct GemmCfg {
  static constexpr int B_ROWS = BN / CTAS;                 // rows of W this CTA stages per k-block
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // RTMA = 1 | 2: every epilogue warp owns that many TMA-fed residual chunk buffers (prefetched RTMA chunks ahead).
  //   2 for short-K GEMMs (out-proj: the epilogue is on the critical path), 1 for deep-K GEMMs (fc2: the tile takes
  //   4x longer than its epilogue, and the 32 KB saved buy a fifth operand stage for the HBM-streamed A operand)
  // OBUF: every epilogue warp owns one bf16 staging chunk used to transpose its row-per-lane results into
  //       row-contiguous (coalesced) global stores
  // every epilogue warp also keeps this tile's slice of bias and of gamma / LN column sums in smem (VEC_BYTES): one
  // coalesced load per tile, issued before the accumulator is ready, instead of 8-16 dependent LDGs per 32-column chunk
  // FBUF: one fp32 [32 x 32] staging chunk per epilogue warp (EPI_PATCH: transposed, coalesced fp32 row stores)
  static constexpr int STAGE_EPI_BYTES = RTMA ? NUM_EPI_WARPS * RTMA * RCHUNK_BYTES
                                              : (OBUF ? NUM_EPI_WARPS * OCHUNK_BYTES : (FBUF ? NUM_EPI_WARPS * RCHUNK_BYTES : 0));
  static constexpr int EPI_BYTES = STAGE_EPI_BYTES + NUM_EPI_WARPS * VEC_BYTES;
  // one CTA per SM: spend all 227 KB of shared memory on operand stages (the deeper ring lets the TMA producer run
  // further into the next tile while the current one drains)
  static constexpr int SMEM_MAX = 227 * 1024;
  static constexpr int STAGES_MAX = (SMEM_MAX - EPI_BYTES - 1024 - 512) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_MAX > 8 ? 8 : STAGES_MAX;
  static constexpr int TMEM_COLS = 512;                    // 2 accumulator stages of BN (<=256) columns
  static constexpr int ACC_STRIDE = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 512 /*barriers*/;
};

__device__ __forceinline__ bool map_row(const GemmEpilogue& e, int row, int& dst_row, int& tok) {
  if (e.tok_in <= 0) {
    dst_row = row;
    tok = 0;
    return true;
  }
  int b = row / e.tok_in;
  tok = row - b * e.tok_in;
  int t2 = tok + e.tok_shift;
  dst_row = b * e.tok_out + t2;
  return t2 >= 0 && t2 < e.tok_out;
}

template <int BN, int CTAS, int MODE, int RTMA>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_r, int M, int N, int K, GemmEpilogue epi) {
  constexpr bool OBUF = (MODE == EPI_BIAS || MODE == EPI_BIAS_GELU || MODE == EPI_BIAS_QGELU);
  constexpr bool FBUF = (MODE == EPI_PATCH);
  using Cfg = GemmCfg<BN, CTAS, RTMA, OBUF, FBUF>;
  constexpr int STAGES = Cfg::STAGES;
  static_assert(!RTMA || MODE == EPI_RESIDUAL, "TMA-staged residual only exists for EPI_RESIDUAL");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_r = smem + STAGES * Cfg::STAGE_BYTES;                 // [8 warps][2][4 KB] residual chunks (RTMA)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                      // [STAGES]   TMA → MMA
  uint64_t* empty_bar = bars + STAGES;            // [STAGES]   MMA → TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;        // [2]        MMA → epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;   // [2]        epilogue → MMA
  uint64_t* rld_bar = bars + 2 * STAGES + 4;      // [8][2]     residual chunk TMA → epilogue warp (RTMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + 2 * NUM_EPI_WARPS);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  constexpr int W_TMA = NUM_EPI_WARPS, W_MMA = NUM_EPI_WARPS + 1, W_ALLOC = NUM_EPI_WARPS + 2;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  const int tile_m = BM * CTAS;
  const int m_tiles = (M + tile_m - 1) / tile_m;
  const int n_tiles = N / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  const int first_tile = blockIdx.x / CTAS;
  const int tile_step = gridDim.x / CTAS;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (RTMA) tma_prefetch_desc(&tmap_r);
  }
  if (warp == W_MMA && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], CTAS * NUM_EPI_WARPS);
    }
    if (RTMA)
      for (int i = 0; i < 2 * NUM_EPI_WARPS; ++i) mbar_init(&rld_bar[i], 1);   // [8][2]; only [.][0] used when RTMA == 1
    fence_mbar_init();
  }
  if (warp == W_ALLOC) {
    tmem_alloc<CTAS>(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped with the previous kernel's tail; from here on we read its output
  pdl_launch_dependents();
  pdl_wait();

  if (warp == W_TMA) {
    // ===================================== TMA producer ==========================================
    // whole warp walks the (warp-uniform) loop; one elected lane issues
    int stage = 0;
    uint32_t phase = 0;
    for (int tile_i = first_tile; tile_i < num_tiles; tile_i += tile_step) {
      const int tile = epi.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int m_blk = tile / n_tiles;
      const int n_blk = tile - m_blk * n_tiles;
      const int row_a = m_blk * tile_m + static_cast<int>(cta_rank) * BM;
      const int row_b = n_blk * BN + static_cast<int>(cta_rank) * Cfg::B_ROWS;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          if constexpr (CTAS == 1) {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kb * BK, row_a);
            tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmap_b, &full_bar[stage], kb * BK, row_b);
          } else {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            tma_load_2d_2sm(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kb * BK, row_a);
            tma_load_2d_2sm(smem_b + stage * Cfg::B_BYTES, &tmap_b, &full_bar[stage], kb * BK, row_b);
          }
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == W_MMA) {
    // ====================================== MMA issuer ===========================================
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CTAS, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * Cfg::ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t desc_a = make_sw128_kmajor_desc(smem_u32(smem_a + stage * Cfg::A_BYTES));
            const uint64_t desc_b = make_sw128_kmajor_desc(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // +32 bytes per UMMA_K step inside the 128-byte swizzle row → +2 in the (addr>>4) field
              umma_bf16<CTAS>(tmem_d, desc_a + static_cast<uint64_t>(2 * k), desc_b + static_cast<uint64_t>(2 * k),
                              idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit<CTAS>(&empty_bar[stage]);   // smem slot reusable once these MMAs retire
            if (kb == num_kb - 1) umma_commit<CTAS>(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else if (warp < NUM_EPI_WARPS) {
    // ======================================= epilogue ============================================
    const int ew = warp;
    const int quarter = warp & 3;              // TMEM lane quarter this warp may touch (= warp_id % 4)
    const int half = ew >> 2;                  // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    constexpr int CHUNKS = COLS_PER_WARP / 32;
    static_assert(COLS_PER_WARP % 32 == 0, "BN/2 must be a multiple of 32");
    int acc = 0;
    uint32_t acc_phase = 0;
    // RTMA: flat sequence seq = local_tile * CHUNKS + chunk; residual chunk `seq` lands in buffer seq & 1 of this
    // warp, its TMA is issued RB sequence steps ahead (i.e. possibly already for the next tile).
    const uint32_t vec_s = smem_u32(smem_r) + Cfg::STAGE_EPI_BYTES + ew * VEC_BYTES;
    constexpr int RB = RTMA > 0 ? RTMA : 1;
    uint8_t* rbuf = smem_r + ew * RB * RCHUNK_BYTES;
    uint64_t* rbar = rld_bar + ew * 2;
    auto issue_resid = [&](int sq) {
      const int lt = sq / CHUNKS, c = sq - lt * CHUNKS;
      const int tl_i = first_tile + lt * tile_step;
      if (tl_i >= num_tiles) return;
      const int tl = epi.reverse ? num_tiles - 1 - tl_i : tl_i;
      const int mb = tl / n_tiles, nb = tl - mb * n_tiles;
      const int r0 = mb * tile_m + static_cast<int>(cta_rank) * BM + quarter * 32;
      const int c0 = nb * BN + half * COLS_PER_WARP + c * 32;
      mbar_expect_tx(&rbar[sq % RB], RCHUNK_BYTES);
      tma_load_2d(rbuf + (sq % RB) * RCHUNK_BYTES, &tmap_r, &rbar[sq % RB], c0, r0);
    };
    int seq = 0;
    if (RTMA && lane == 0) {
      issue_resid(0);
      if (RB == 2) issue_resid(1);
    }
    for (int tile_i = first_tile; tile_i < num_tiles; tile_i += tile_step) {
      const int tile = epi.reverse ? num_tiles - 1 - tile_i : tile_i;
      const int m_blk = tile / n_tiles;
      const int n_blk = tile - m_blk * n_tiles;
      const int row = m_blk * tile_m + static_cast<int>(cta_rank) * BM + quarter * 32 + lane;
      const bool row_ok = row < M;
      int dst_row = row, tok = 0;
      bool dst_ok = row_ok;
      dst_ok = map_row(epi, row, dst_row, tok) && row_ok;

      // coalesced-store mapping (bf16 modes): in store pass k this lane writes 16 B of row (k*8 + lane/4) of the
      // warp's 32-row slab; (RTMA fp32: row k*4 + lane/8).  Destination rows are computed once per tile.
      const int row0 = m_blk * tile_m + static_cast<int>(cta_rank) * BM + quarter * 32;
      long long t_dst[4];
      if constexpr (OBUF) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gr = row0 + k * 8 + (lane >> 2);
          int dr = gr, tk = 0;
          const bool ok = map_row(epi, gr, dr, tk) && gr < M;
          t_dst[k] = ok ? static_cast<long long>(dr) * epi.ld_out + epi.out_col_off : -1;
        }
      }
      const uint32_t obuf_s = smem_u32(smem_r) + ew * OCHUNK_BYTES;
      // EPI_PATCH: in store pass k this lane writes 16 B of row (k*4 + lane/8) of the slab (fp32, 128 B per row)
      long long p_dst[FBUF ? 8 : 1];
      int p_tok[FBUF ? 8 : 1];
      if constexpr (FBUF) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int gr = row0 + k * 4 + (lane >> 3);
          int dr = gr, tk = 0;
          const bool ok = map_row(epi, gr, dr, tk) && gr < M;
          p_dst[k] = ok ? static_cast<long long>(dr) * epi.ld_resid : -1;
          p_tok[k] = tk;
        }
      }
      const uint32_t fbuf_s = smem_u32(smem_r) + ew * RCHUNK_BYTES;

      // folded LayerNorm, consumer side: rebuild (mean, rstd) of this lane's row from the producer's partial sums
      float ln_mean = 0.f, ln_rstd = 1.f;
      if constexpr (OBUF) {
        if (epi.ln_stats != nullptr && row_ok) {
          // stats are [part][row]: a warp reads 32 consecutive pairs per part; four independent loads in flight so the
          // whole rebuild costs ~one L2 round trip (it sits in front of every tile's epilogue)
          const float2* sp = epi.ln_stats + row;
          float s1 = 0.f, s2 = 0.f;
          for (int p0 = 0; p0 < epi.ln_parts; p0 += 4) {
            float2 t[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              t[q] = p0 + q < epi.ln_parts ? __ldg(sp + static_cast<size_t>(p0 + q) * M) : make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              s1 += t[q].x;
              s2 += t[q].y;
            }
          }
          const float inv_k = 1.0f / static_cast<float>(K);
          ln_mean = s1 * inv_k;
          ln_rstd = rsqrtf(fmaxf(s2 * inv_k - ln_mean * ln_mean, 0.f) + epi.ln_eps);
          // rolling shift: the rows and their statistics are relative to shift_in; the absolute row mean = shift_in +
          // ln_mean becomes the shift of the NEXT producer (one store per row: the first column span's warp)
          if (epi.shift_out != nullptr && n_blk == 0 && half == 0)
            epi.shift_out[row] = (epi.shift_in != nullptr ? __ldg(epi.shift_in + row) : 0.f) + ln_mean;
        }
      }
      float part_sum = 0.f, part_sq = 0.f;   // producer side: this lane's row over this warp's column span
      // producer side, rolling shift: c = this row's mean one residual update ago, handed over by the folded consumer
      // that ran in between (it rebuilds the mean anyway); the bf16 copy and the statistics below are those of x_new − c
      float row_shift = 0.f;
      if constexpr (MODE == EPI_RESIDUAL) {
        if (epi.shift_in != nullptr && row_ok) row_shift = __ldg(epi.shift_in + row);
      }

      // this warp's COLS_PER_WARP-wide slices of bias (vec 0) and of gamma | LN column sums (vec 1) → smem
      {
        const int cw0 = n_blk * BN + half * COLS_PER_WARP;
        const float* v1 = MODE == EPI_RESIDUAL ? epi.gamma : (OBUF && epi.ln_stats != nullptr ? epi.ln_colsum : nullptr);
        const float fill1 = MODE == EPI_RESIDUAL ? 1.f : 0.f;
        if (lane < COLS_PER_WARP / 4) {
          const float4 a4 = epi.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(epi.bias + cw0) + lane)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 g4 = v1 != nullptr ? __ldg(reinterpret_cast<const float4*>(v1 + cw0) + lane)
                                          : make_float4(fill1, fill1, fill1, fill1);
          sts128(vec_s + lane * 16, make_uint4(__float_as_uint(a4.x), __float_as_uint(a4.y), __float_as_uint(a4.z),
                                               __float_as_uint(a4.w)));
          sts128(vec_s + 512 + lane * 16, make_uint4(__float_as_uint(g4.x), __float_as_uint(g4.y),
                                                     __float_as_uint(g4.z), __float_as_uint(g4.w)));
        }
        __syncwarp();
      }

#if BLB_GEMM_EPI_SLEEP_NS > 0
      // A/B switch: poll with test_wait + nanosleep instead of parking the 8 epilogue warps in try_wait
      mbar_wait_sleep<BLB_GEMM_EPI_SLEEP_NS>(&tfull_bar[acc], acc_phase);
#else
      mbar_wait(&tfull_bar[acc], acc_phase);
#endif
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < CHUNKS; ++c) {
        const int col_in_tile = half * COLS_PER_WARP + c * 32;
        const int col0 = n_blk * BN + col_in_tile;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                          static_cast<uint32_t>(acc * Cfg::ACC_STRIDE + col_in_tile),
                      r);
        float xres[RTMA ? 32 : 1];
        if constexpr (RTMA) {
          // this lane's row of the TMA-staged residual chunk (128-byte rows, 128B-swizzled → conflict-free)
          mbar_wait(&rbar[seq % RB], static_cast<uint32_t>((seq / RB) & 1));
          const uint32_t src = smem_u32(rbuf) + (seq % RB) * RCHUNK_BYTES + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 x4 = lds128(src + ((j ^ (lane & 7)) << 4));
            xres[4 * j] = __uint_as_float(x4.x); xres[4 * j + 1] = __uint_as_float(x4.y);
            xres[4 * j + 2] = __uint_as_float(x4.z); xres[4 * j + 3] = __uint_as_float(x4.w);
          }
        }
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if constexpr (OBUF) {
          if (epi.ln_stats != nullptr) {   // LN(x)·Wᵀ = rstd·(x·W'ᵀ − mean·colsum(W'))  (+ folded bias below)
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const uint4 c4 = lds128(vec_s + 512 + (c * 32 + j) * 4);
              v[j] = (v[j] - ln_mean * __uint_as_float(c4.x)) * ln_rstd;
              v[j + 1] = (v[j + 1] - ln_mean * __uint_as_float(c4.y)) * ln_rstd;
              v[j + 2] = (v[j + 2] - ln_mean * __uint_as_float(c4.z)) * ln_rstd;
              v[j + 3] = (v[j + 3] - ln_mean * __uint_as_float(c4.w)) * ln_rstd;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {   // + bias (zeros when there is none)
          const uint4 b4 = lds128(vec_s + (c * 32 + j) * 4);
          v[j] += __uint_as_float(b4.x); v[j + 1] += __uint_as_float(b4.y);
          v[j + 2] += __uint_as_float(b4.z); v[j + 3] += __uint_as_float(b4.w);
        }
        if constexpr (OBUF) {
          if constexpr (MODE == EPI_BIAS_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          }
          if constexpr (MODE == EPI_BIAS_QGELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_quick(v[j]);
          }
          // stage this lane's row (64 B) in the warp's smem chunk: 16-byte piece j sits at j ^ ((row>>1)&3) so that
          // both this row-per-lane write and the transposed read below are bank-conflict free
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
            pk.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            pk.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            pk.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            sts128(obuf_s + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pk);
          }
          __syncwarp();
          // transposed read → each store instruction now writes 8 rows x 64 contiguous bytes
          uint4 tv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = k * 8 + (lane >> 2), j = lane & 3;
            tv[k] = lds128(obuf_s + r * 64 + ((j ^ ((r >> 1) & 3)) << 4));
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (t_dst[k] >= 0) stg128(epi.out + t_dst[k] + col0 + (lane & 3) * 8, tv[k]);
          __syncwarp();
        } else if constexpr (MODE == EPI_PATCH) {
          // out_f32[dst_row, col] = acc + bias + pos_embed[token, col]   (timm PatchEmbed + _pos_embed): the lane's row
          // (acc + bias) goes through the warp's swizzled smem chunk so that both the pos_embed read and the fp32 store
          // are row-contiguous (4 rows x 128 B per instruction instead of 32 rows x 16 B)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(fbuf_s + lane * 128 + ((j ^ (lane & 7)) << 4),
                   make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                              __float_as_uint(v[4 * j + 3])));
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int r = k * 4 + (lane >> 3), j = lane & 7;
            const uint4 t4 = lds128(fbuf_s + r * 128 + ((j ^ (r & 7)) << 4));
            if (p_dst[k] >= 0) {
              float4 p4 = make_float4(0.f, 0.f, 0.f, 0.f);
              if (epi.pos != nullptr)
                p4 = __ldg(reinterpret_cast<const float4*>(epi.pos + static_cast<size_t>(p_tok[k]) * N + col0 + j * 4));
              const uint4 o4 = make_uint4(__float_as_uint(__uint_as_float(t4.x) + p4.x), __float_as_uint(__uint_as_float(t4.y) + p4.y),
                                          __float_as_uint(__uint_as_float(t4.z) + p4.z), __float_as_uint(__uint_as_float(t4.w) + p4.w));
              stg128(epi.resid + p_dst[k] + col0 + j * 4, o4);
            }
          }
          __syncwarp();
        } else {  // EPI_RESIDUAL:  x += gamma * (acc + bias)   (timm Block: x + ls(branch(x)))
          if (row_ok) {
            float* x = epi.resid + static_cast<size_t>(row) * epi.ld_resid + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {   // LayerScale gamma (ones when there is none)
              const uint4 g4 = lds128(vec_s + 512 + (c * 32 + j) * 4);
              v[j] *= __uint_as_float(g4.x); v[j + 1] *= __uint_as_float(g4.y);
              v[j + 2] *= __uint_as_float(g4.z); v[j + 3] *= __uint_as_float(g4.w);
            }
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 x4;
              if constexpr (RTMA) x4 = make_float4(xres[j], xres[j + 1], xres[j + 2], xres[j + 3]);
              else x4 = *reinterpret_cast<const float4*>(x + j);
              x4.x += v[j]; x4.y += v[j + 1]; x4.z += v[j + 2]; x4.w += v[j + 3];
              v[j] = x4.x; v[j + 1] = x4.y; v[j + 2] = x4.z; v[j + 3] = x4.w;
              if constexpr (!RTMA) *reinterpret_cast<float4*>(x + j) = x4;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = v[j] - row_shift;
              part_sum += a;
              part_sq = fmaf(a, a, part_sq);
            }
            if constexpr (!RTMA) {
              if (epi.xb_out != nullptr) {
                __nv_bfloat16* o = epi.xb_out + static_cast<size_t>(row) * epi.ld_xb + col0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                  uint4 pk;
                  pk.x = pack_bf16x2(v[j] - row_shift, v[j + 1] - row_shift);
                  pk.y = pack_bf16x2(v[j + 2] - row_shift, v[j + 3] - row_shift);
                  pk.z = pack_bf16x2(v[j + 4] - row_shift, v[j + 5] - row_shift);
                  pk.w = pack_bf16x2(v[j + 6] - row_shift, v[j + 7] - row_shift);
                  *reinterpret_cast<uint4*>(o + j) = pk;
                }
              }
            }
            if (epi.out != nullptr && dst_ok) {
              __nv_bfloat16* o = epi.out + static_cast<size_t>(dst_row) * epi.ld_out + epi.out_col_off + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                pk.x = pack_bf16x2(v[j], v[j + 1]);
                pk.y = pack_bf16x2(v[j + 2], v[j + 3]);
                pk.z = pack_bf16x2(v[j + 4], v[j + 5]);
                pk.w = pack_bf16x2(v[j + 6], v[j + 7]);
                *reinterpret_cast<uint4*>(o + j) = pk;
              }
            }
          }
          if constexpr (RTMA) {
            // write the updated row back into the (already consumed) residual chunk in place, then store the
            // chunk transposed: each store instruction writes 4 rows x 128 contiguous bytes
            const uint32_t cb = smem_u32(rbuf) + (seq % RB) * RCHUNK_BYTES;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(cb + lane * 128 + ((j ^ (lane & 7)) << 4),
                     make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                                __float_as_uint(v[4 * j + 3])));
            __syncwarp();
            uint4 tv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = k * 4 + (lane >> 3), j = lane & 7;
              tv[k] = lds128(cb + r * 128 + ((j ^ (r & 7)) << 4));
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int r = k * 4 + (lane >> 3), j = lane & 7;
              if (row0 + r < M)
                stg128(epi.resid + static_cast<size_t>(row0 + r) * epi.ld_resid + col0 + j * 4, tv[k]);
            }
            if (epi.xb_out != nullptr) {   // bf16 copy for the LN-folded consumer: 4 rows x 64 contiguous bytes per store
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int r = k * 4 + (lane >> 3), j = lane & 7;
                const float cr = __shfl_sync(0xffffffffu, row_shift, r);   // row r of the slab belongs to lane r
                if (row0 + r < M) {
                  uint2 pk;
                  pk.x = pack_bf16x2(__uint_as_float(tv[k].x) - cr, __uint_as_float(tv[k].y) - cr);
                  pk.y = pack_bf16x2(__uint_as_float(tv[k].z) - cr, __uint_as_float(tv[k].w) - cr);
                  *reinterpret_cast<uint2*>(epi.xb_out + static_cast<size_t>(row0 + r) * epi.ld_xb + col0 + j * 4) = pk;
                }
              }
            }
            // WAR across proxies: the refill is an async-proxy (TMA) write to a buffer just accessed through the
            // generic proxy → fence, converge, then lane 0 issues the TMA for the chunk two steps ahead.
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) issue_resid(seq + RB);
            ++seq;
          }
        }
      }
      if constexpr (MODE == EPI_RESIDUAL) {
        if (epi.stats_out != nullptr && row_ok)
          epi.stats_out[static_cast<size_t>(2 * n_blk + half) * M + row] = make_float2(part_sum, part_sq);
      }
      // all TMEM reads of this accumulator stage are complete → hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 1) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(&tempty_bar[acc], 0);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  if constexpr (CTAS == 2) cluster_sync(); else __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc<CTAS>(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows, cols] matrix (pitch ld elements of elem_bytes) → 2-D TMA map with a {box_cols, box_rows} box.
static int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* ptr, uint64_t rows,
                        uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return BLB_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * elem_bytes) % 16 != 0) return BLB_ERR_ALIGN;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * elem_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : BLB_ERR_DRIVER;
}

// bf16 [rows, cols] row-major (pitch ld elements) → TMA map with a {64, box_rows} box, 128B swizzle.
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, rows, cols, ld, static_cast<uint32_t>(BK),
                      box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

// Process-wide state.  The library may be driven from several host threads (one per stream / device): counters and
// switches are atomics, the per-launch timing records (a diagnostics facility, see blb_timing_*) sit behind a mutex.
static std::atomic<int> g_num_sms_dev[BLB_MAX_DEVICES];
static std::atomic<int> g_force_ctas{0};   // 0 = auto, 1 / 2 = forced (tests and A/B measurements)
static bool g_resid_direct = getenv("BLB_RESID_DIRECT") != nullptr;   // A/B switch: residual via plain loads
static std::atomic<long long> g_launches{0};

void gemm_set_cta_group(int ctas) { g_force_ctas.store(ctas, std::memory_order_relaxed); }
static int pick_bn(int N) { return N % 256 == 0 ? 256 : N % 192 == 0 ? 192 : N % 128 == 0 ? 128 : 0; }
int gemm_stats_parts(int N) {
  const int bn = pick_bn(N);
  return bn == 0 ? 0 : 2 * (N / bn);
}
bool pdl_enabled() {
  static const bool on = getenv("BLB_NO_PDL") == nullptr;
  return on;
}

// ---- optional per-launch timing ------------------------------------------------------------------
namespace {
struct TimingRec { int cat; long long tag; double work; cudaEvent_t e0, e1; };
std::atomic<bool> g_timing{false};
std::mutex g_timing_mu;                 // guards g_recs / g_event_pool
std::vector<TimingRec> g_recs;
std::vector<cudaEvent_t> g_event_pool;
thread_local cudaEvent_t g_pending_e0 = nullptr;   // begin event of the launch this thread is bracketing
cudaEvent_t get_event() {               // caller holds g_timing_mu
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e; cudaEventCreate(&e); return e;
}
}  // namespace
void timing_enable(int on) { g_timing.store(on != 0); }
bool timing_enabled() { return g_timing.load(std::memory_order_relaxed); }
void timing_begin(cudaStream_t s) {
  { std::lock_guard<std::mutex> lk(g_timing_mu); g_pending_e0 = get_event(); }
  cudaEventRecord(g_pending_e0, s);
}
void timing_end(int cat, double work, cudaStream_t s, long long tag) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  cudaEvent_t e1 = get_event();
  cudaEventRecord(e1, s);
  g_recs.push_back({cat, tag, work, g_pending_e0, e1});
  g_pending_e0 = nullptr;
}
void timing_reset() {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (auto& r : g_recs) { g_event_pool.push_back(r.e0); g_event_pool.push_back(r.e1); }
  g_recs.clear();
}
int timing_collect(int cat, double* ms, double* work, long long* launches) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  double t = 0, w = 0; long long n = 0;
  for (auto& r : g_recs) {
    if (r.cat != cat) continue;
    float f = 0.f;
    cudaError_t e = cudaEventElapsedTime(&f, r.e0, r.e1);
    if (e != cudaSuccess) return static_cast<int>(e);
    t += f; w += r.work; ++n;
  }
  if (ms) *ms = t;
  if (work) *work = w;
  if (launches) *launches = n;
  return 0;
}
int timing_records(int max_records, int* cat, long long* tag, double* ms, double* work) {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  int n = 0;
  for (auto& r : g_recs) {
    if (n >= max_records) break;
    float f = 0.f;
    cudaError_t e = cudaEventElapsedTime(&f, r.e0, r.e1);
    if (e != cudaSuccess) return -static_cast<int>(e);
    cat[n] = r.cat; tag[n] = r.tag; ms[n] = f; work[n] = r.work;
    ++n;
  }
  return n;
}
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= BLB_MAX_DEVICES) return 0;
  return dev;
}

int num_sms() {   // of the CURRENT device (cached per device)
  const int dev = current_device();
  int n = g_num_sms_dev[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    g_num_sms_dev[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}

template <int BN, int CTAS, int MODE, int RTMA>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, int M, int N, int K,
                  const GemmEpilogue& epi, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTAS, RTMA, (MODE == EPI_BIAS || MODE == EPI_BIAS_GELU || MODE == EPI_BIAS_QGELU),
                      (MODE == EPI_PATCH)>;
  auto kern = gemm_bf16_kernel<BN, CTAS, MODE, RTMA>;
  static std::atomic<bool> configured[BLB_MAX_DEVICES];   // the attribute is per device; setting it twice is harmless
  if (!configured[current_device()].load(std::memory_order_acquire)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[current_device()].store(true, std::memory_order_release);
  }
  const int tile_m = BM * CTAS;
  const int tiles = ((M + tile_m - 1) / tile_m) * (N / BN);
  int ctas = num_sms();
  if (CTAS == 2) ctas &= ~1;
  if (tiles * CTAS < ctas) ctas = tiles * CTAS;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  // tag: epilogue mode | LN-folded consumer | stats/xb producer | N | K  (blb_timing_records)
  const long long tag = (static_cast<long long>(MODE) << 44) | (static_cast<long long>(epi.ln_stats != nullptr) << 43) |
                        (static_cast<long long>(epi.xb_out != nullptr) << 42) | (static_cast<long long>(N) << 20) | K;
  TimingScope ts(TIME_GEMM, epi.alg_work > 0.0 ? epi.alg_work : 2.0 * M * N * K, stream, tag);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tr, M, N, K, epi);
  count_launch(1);
  return static_cast<int>(e);
}

template <int BN, int CTAS>
static int launch_mode(int mode, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K,
                       const GemmEpilogue& epi, cudaStream_t s) {
  switch (mode) {
    case EPI_BIAS: return launch<BN, CTAS, EPI_BIAS, 0>(ta, tb, ta, M, N, K, epi, s);
    case EPI_BIAS_GELU: return launch<BN, CTAS, EPI_BIAS_GELU, 0>(ta, tb, ta, M, N, K, epi, s);
    case EPI_BIAS_QGELU: return launch<BN, CTAS, EPI_BIAS_QGELU, 0>(ta, tb, ta, M, N, K, epi, s);
    case EPI_PATCH: return launch<BN, CTAS, EPI_PATCH, 0>(ta, tb, ta, M, N, K, epi, s);
    case EPI_RESIDUAL: {
      if (g_resid_direct) return launch<BN, CTAS, EPI_RESIDUAL, 0>(ta, tb, ta, M, N, K, epi, s);
      CUtensorMap tr;   // fp32 residual stream [M, N] → [32 x 32] boxes, 128B-swizzled rows
      int rc = make_tmap_2d(&tr, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, epi.resid, static_cast<uint64_t>(M),
                            static_cast<uint64_t>(N), static_cast<uint64_t>(epi.ld_resid), 32, 32,
                            CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc != 0) return rc;
      static const int deep_k = getenv("BLB_RESID_DEEPK") ? atoi(getenv("BLB_RESID_DEEPK")) : 2048;   // A/B switch
      if (K >= deep_k) return launch<BN, CTAS, EPI_RESIDUAL, 1>(ta, tb, tr, M, N, K, epi, s);
      return launch<BN, CTAS, EPI_RESIDUAL, 2>(ta, tb, tr, M, N, K, epi, s);
    }
  }
  return BLB_ERR_ARG;
}

int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, int mode,
              const GemmEpilogue& epi, cudaStream_t stream) {
  if (M <= 0 || N <= 0 || K <= 0 || A == nullptr || W == nullptr) return BLB_ERR_ARG;
  const int bn = pick_bn(N);
  if (bn == 0) return BLB_ERR_SHAPE;
  const bool bf16_out_mode = mode == EPI_BIAS || mode == EPI_BIAS_GELU || mode == EPI_BIAS_QGELU;
  if (epi.ln_stats != nullptr && (epi.ln_colsum == nullptr || epi.ln_parts <= 0 || !bf16_out_mode)) return BLB_ERR_ARG;
  if ((epi.stats_out != nullptr || epi.xb_out != nullptr) && mode != EPI_RESIDUAL) return BLB_ERR_ARG;
  if (epi.shift_out != nullptr && (epi.ln_stats == nullptr || epi.shift_in == epi.shift_out)) return BLB_ERR_ARG;   // ping-pong
  if (epi.shift_in != nullptr && mode == EPI_PATCH) return BLB_ERR_ARG;
  const int forced = g_force_ctas.load(std::memory_order_relaxed);
  const int ctas = forced != 0 ? forced : 2;
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), BM);
  if (rc != 0) return rc;
  rc = make_tmap_bf16_2d(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw),
                         static_cast<uint32_t>(bn / ctas));
  if (rc != 0) return rc;
  if (ctas == 1) {
    if (bn == 256) return launch_mode<256, 1>(mode, ta, tb, M, N, K, epi, stream);
    if (bn == 192) return launch_mode<192, 1>(mode, ta, tb, M, N, K, epi, stream);
    return launch_mode<128, 1>(mode, ta, tb, M, N, K, epi, stream);
  }
  if (bn == 256) return launch_mode<256, 2>(mode, ta, tb, M, N, K, epi, stream);
  if (bn == 192) return launch_mode<192, 2>(mode, ta, tb, M, N, K, epi, stream);
  return launch_mode<128, 2>(mode, ta, tb, M, N, K, epi, stream);
}

}  // namespace blb
