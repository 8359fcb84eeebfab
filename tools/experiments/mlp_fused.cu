// EXPERIMENT, NOT BUILT (round 1): fc1 + fc2 of one Block as one persistent kernel with tile-level dependencies.
// Result on B200 (same-box A/B, B = 256): numerically correct (matched the two-launch path and fp32 torch in
// tests), but 3.5 % SLOWER end to end (2465-2475 vs 2543-2562 images/s) although fc2 alone gains 8.7 % when its A
// operand is L2-resident.  Why: (1) a static round-robin schedule of mixed 1-unit (fc1) and 4-unit (fc2) tiles leaves
// CTA pairs unevenly loaded and makes dependency stalls cascade; (2) with two TMEM accumulator stages, the long fc2
// epilogue delays the short fc1 tiles that follow it on the same CTA; (3) fc1 loses one operand stage to the shared
// smem layout.  It also needs ALL CTAs co-resident (spin waits), so the two towers had to be serialised.  What a second
// attempt needs: dynamic tile claiming (atomic counter → balance + no residency assumption) and per-phase epilogue
// warps.  To build it again: add the file to bridgelang_b200/build.py SOURCES and restore the hooks described in the
// header comment below (mlp_fused_bf16 in gemm.h, the call in capi.cu::tower_blocks_ln_folded).
// mlp_fused.cu — timm Mlp (fc1 + GELU, fc2 + LayerScale + residual) of one Block as ONE persistent kernel.
//
// Why: the step is power-bound and an HBM byte costs ≈ 70-90 pJ.  As two launches, fc1 writes the
// [M, Hm] hidden tensor (547 MB at B = 256) to HBM and fc2 streams it back; with the A operand of fc2 L2-resident the
// fc2 GEMM runs 8.7 % faster under the power cap (measured).  Here both GEMMs share one static tile schedule in which
// the fc2 tiles of a group of row blocks come one round after the fc1 tiles that produce their A operand, so the
// hidden rows are read back out of the 126 MB L2 a few microseconds after they were written.
//
// Structure = gemm_tcgen05.cu (cta_group::2 pairs, 256 x BN tiles, TMA ring → tcgen05.mma → double-buffered TMEM
// accumulators → 8 epilogue warps), with
//   * a mixed schedule: round r = [fc1 tiles of row-block group r][fc2 tiles of group r-1], tiles dealt round-robin to
//     the CTA pairs; every role warp decodes (phase, m_blk, n_blk) from the flat tile index with a private cursor;
//   * tile-level dependencies: every fc1 epilogue warp publishes its stores (__threadfence + atomicAdd on
//     ready[m_blk][cta_rank]); the TMA producer of an fc2 tile polls that counter (acquire), issues a cross-proxy
//     fence, then loads the hidden rows.  Dependencies only point to smaller tile indices and every CTA walks its tiles
//     in increasing order, so the schedule cannot deadlock as long as all CTAs are resident (persistent grid, 1 CTA/SM);
//   * epilogues: fc1 = folded LayerNorm + bias + erf-GELU → bf16 (EPI_BIAS_GELU of gemm_tcgen05.cu);
//     fc2 = LayerScale + residual on the fp32 stream, TMA-staged (one chunk buffer), + partial row statistics and the
//     bf16 copy for the next block, or the concat write of the last block (EPI_RESIDUAL).
#include <cstdlib>

#include "gemm.h"
#include "ptx.cuh"

namespace blb {

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int THREADS = 128 + NUM_EPI_WARPS * 32;
constexpr int BN1 = 256;
constexpr int A_BYTES = BM * BK * 2;            // 16 KB
constexpr int B_SLOT = 128 * BK * 2;            // 16 KB: BN/2 rows of W per CTA, BN <= 256
constexpr int STAGE_BYTES = A_BYTES + B_SLOT;
constexpr int RCHUNK_BYTES = 32 * 32 * 4;       // fp32 residual chunk of one epilogue warp (also holds the 2 KB bf16 chunk)
constexpr int VEC_BYTES = 2 * 128 * 4;
constexpr int EPI_BYTES = NUM_EPI_WARPS * (RCHUNK_BYTES + VEC_BYTES);
constexpr int STAGES = (227 * 1024 - EPI_BYTES - 1024 - 512) / STAGE_BYTES;   // 5
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 512;
constexpr int ACC_STRIDE = 256;

struct MlpArgs {
  int M, N1, K1, N2, K2;
  int group;                 // row blocks (of 256 rows) per schedule group
  int reverse;               // walk the row blocks last-to-first (serpentine with the producer of xb)
  int* ready;                // [m_tiles][2] epilogue-warp completions of fc1 tiles, monotonic across launches
  int ready_target;          // value that means "all fc1 tiles of this row block are stored" for THIS launch
  GemmEpilogue e1, e2;
};

struct Tile {
  int phase, m_blk, n_blk;
};

// flat tile index → (phase, m_blk, n_blk); `r`/`base` are the caller's cursor (tile indices only ever increase)
struct Cursor {
  int r = 0, base = 0;
};
__device__ __forceinline__ Tile decode(const MlpArgs& a, int m_tiles, int n1, int n2, int n_groups, int t, Cursor& c) {
  auto gsz = [&](int g) { return min(a.group, m_tiles - g * a.group); };
  for (;;) {
    const int s1 = c.r < n_groups ? gsz(c.r) * n1 : 0;
    const int s2 = c.r >= 1 ? gsz(c.r - 1) * n2 : 0;
    if (t < c.base + s1 + s2) {
      int j = t - c.base;
      Tile tl;
      if (j < s1) {
        tl.phase = 0;
        tl.m_blk = c.r * a.group + j / n1;
        tl.n_blk = j % n1;
      } else {
        j -= s1;
        tl.phase = 1;
        tl.m_blk = (c.r - 1) * a.group + j / n2;
        tl.n_blk = j % n2;
      }
      if (a.reverse) tl.m_blk = m_tiles - 1 - tl.m_blk;
      return tl;
    }
    c.base += s1 + s2;
    ++c.r;
  }
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

template <int BN2>
__global__ void __launch_bounds__(THREADS, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_b1,
                 const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b2,
                 const __grid_constant__ CUtensorMap tmap_r, MlpArgs args) {
  extern __shared__ uint8_t smem_raw_mlp[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_mlp) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_e = smem + STAGES * STAGE_BYTES;                       // [8 warps][4 KB chunk] then [8 warps][1 KB vec]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;
  uint64_t* rld_bar = bars + 2 * STAGES + 4;                            // [8] residual chunk TMA → epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4 + NUM_EPI_WARPS);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  constexpr int W_TMA = NUM_EPI_WARPS, W_MMA = NUM_EPI_WARPS + 1, W_ALLOC = NUM_EPI_WARPS + 2;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  const int M = args.M;
  constexpr int tile_m = 2 * BM;
  const int m_tiles = (M + tile_m - 1) / tile_m;
  const int n1 = args.N1 / BN1, n2 = args.N2 / BN2;
  const int n_groups = (m_tiles + args.group - 1) / args.group;
  const int num_tiles = m_tiles * (n1 + n2);
  const int kb1 = (args.K1 + BK - 1) / BK, kb2 = (args.K2 + BK - 1) / BK;
  const int first_tile = blockIdx.x / 2;
  const int tile_step = gridDim.x / 2;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_a1); tma_prefetch_desc(&tmap_b1);
    tma_prefetch_desc(&tmap_a2); tma_prefetch_desc(&tmap_b2);
    tma_prefetch_desc(&tmap_r);
  }
  if (warp == W_MMA && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * NUM_EPI_WARPS);
    }
    for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(&rld_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == W_ALLOC) tmem_alloc<2>(tmem_slot, 512);
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == W_TMA) {
    // ===================================== TMA producer ==========================================
    int stage = 0;
    uint32_t phase = 0;
    Cursor cur;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      const Tile tl = decode(args, m_tiles, n1, n2, n_groups, t, cur);
      const bool p2 = tl.phase == 1;
      const int num_kb = p2 ? kb2 : kb1;
      const int b_rows = (p2 ? BN2 : BN1) / 2;
      const int row_a = tl.m_blk * tile_m + static_cast<int>(cta_rank) * BM;
      const int row_b = tl.n_blk * (p2 ? BN2 : BN1) + static_cast<int>(cta_rank) * b_rows;
      const CUtensorMap* ta = p2 ? &tmap_a2 : &tmap_a1;
      const CUtensorMap* tb = p2 ? &tmap_b2 : &tmap_b1;
      const uint32_t stage_tx = 2u * static_cast<uint32_t>(A_BYTES + b_rows * BK * 2);
      if (p2) {
        // the hidden rows of this CTA's half of the row block: all n1 fc1 tiles × 8 epilogue warps have published them
        const int* flag = args.ready + tl.m_blk * 2 + static_cast<int>(cta_rank);
        while (ld_acquire_gpu(flag) < args.ready_target) __nanosleep(40);
        fence_proxy_async_all();     // generic-proxy stores of other SMs → this thread's async-proxy (TMA) loads
        __syncwarp();
      }
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_2d_2sm(smem_a + stage * A_BYTES, ta, &full_bar[stage], kb * BK, row_a);
          tma_load_2d_2sm(smem_b + stage * B_SLOT, tb, &full_bar[stage], kb * BK, row_b);
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
      constexpr uint32_t idesc1 = make_idesc_bf16(2 * BM, BN1);
      constexpr uint32_t idesc2 = make_idesc_bf16(2 * BM, BN2);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      Cursor cur;
      for (int t = first_tile; t < num_tiles; t += tile_step) {
        const Tile tl = decode(args, m_tiles, n1, n2, n_groups, t, cur);
        const int num_kb = tl.phase == 1 ? kb2 : kb1;
        const uint32_t idesc = tl.phase == 1 ? idesc2 : idesc1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * ACC_STRIDE);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t desc_a = make_sw128_kmajor_desc(smem_u32(smem_a + stage * A_BYTES));
            const uint64_t desc_b = make_sw128_kmajor_desc(smem_u32(smem_b + stage * B_SLOT));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16<2>(tmem_d, desc_a + static_cast<uint64_t>(2 * k), desc_b + static_cast<uint64_t>(2 * k), idesc,
                           (kb | k) != 0 ? 1u : 0u);
            umma_commit<2>(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit<2>(&tfull_bar[acc]);
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
    const int quarter = warp & 3;
    const int half = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t rphase = 0;                                   // parity of this warp's residual-chunk barrier
    const uint32_t chunk_s = smem_u32(smem_e) + ew * RCHUNK_BYTES;
    uint8_t* chunk_p = smem_e + ew * RCHUNK_BYTES;
    const uint32_t vec_s = smem_u32(smem_e) + NUM_EPI_WARPS * RCHUNK_BYTES + ew * VEC_BYTES;
    uint64_t* rbar = rld_bar + ew;
    Cursor cur;
    for (int t = first_tile; t < num_tiles; t += tile_step) {
      const Tile tl = decode(args, m_tiles, n1, n2, n_groups, t, cur);
      const bool p2 = tl.phase == 1;
      const GemmEpilogue& epi = p2 ? args.e2 : args.e1;
      const int bn = p2 ? BN2 : BN1;
      const int cols_per_warp = bn / 2;
      const int chunks = cols_per_warp / 32;
      const int m_blk = tl.m_blk, n_blk = tl.n_blk;
      const int row0 = m_blk * tile_m + static_cast<int>(cta_rank) * BM + quarter * 32;
      const int row = row0 + lane;
      const bool row_ok = row < M;
      const int cw0 = n_blk * bn + half * cols_per_warp;

      // fc2: the first residual chunk of this tile can travel while the main loop still runs
      if (p2 && lane == 0) {
        mbar_expect_tx(rbar, RCHUNK_BYTES);
        tma_load_2d(chunk_p, &tmap_r, rbar, cw0, row0);
      }
      // per-tile slices of the per-column vectors: bias | (LN column sums or LayerScale gamma)
      {
        const float* v1 = p2 ? epi.gamma : epi.ln_colsum;
        const float fill1 = p2 ? 1.f : 0.f;
        if (lane < cols_per_warp / 4) {
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

      if (!p2) {
        // ------------------------- fc1: folded LayerNorm + bias + GELU → bf16 hidden -------------------------
        float ln_mean = 0.f, ln_rstd = 1.f;
        if (row_ok) {
          const float2* sp = epi.ln_stats + row;
          float s1 = 0.f, s2 = 0.f;
          for (int p0 = 0; p0 < epi.ln_parts; p0 += 4) {
            float2 tq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              tq[q] = p0 + q < epi.ln_parts ? __ldg(sp + static_cast<size_t>(p0 + q) * M) : make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              s1 += tq[q].x;
              s2 += tq[q].y;
            }
          }
          const float inv_k = 1.0f / static_cast<float>(args.K1);
          ln_mean = s1 * inv_k;
          ln_rstd = rsqrtf(fmaxf(s2 * inv_k - ln_mean * ln_mean, 0.f) + epi.ln_eps);
        }
        long long t_dst[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int gr = row0 + k * 8 + (lane >> 2);
          t_dst[k] = gr < M ? static_cast<long long>(gr) * epi.ld_out : -1;
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          const int col_in_tile = half * cols_per_warp + c * 32;
          const int col0 = n_blk * bn + col_in_tile;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                            static_cast<uint32_t>(acc * ACC_STRIDE + col_in_tile), r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const uint4 c4 = lds128(vec_s + 512 + (c * 32 + j) * 4);
            const uint4 b4 = lds128(vec_s + (c * 32 + j) * 4);
            v[j] = gelu_erf((__uint_as_float(r[j]) - ln_mean * __uint_as_float(c4.x)) * ln_rstd + __uint_as_float(b4.x));
            v[j + 1] = gelu_erf((__uint_as_float(r[j + 1]) - ln_mean * __uint_as_float(c4.y)) * ln_rstd + __uint_as_float(b4.y));
            v[j + 2] = gelu_erf((__uint_as_float(r[j + 2]) - ln_mean * __uint_as_float(c4.z)) * ln_rstd + __uint_as_float(b4.z));
            v[j + 3] = gelu_erf((__uint_as_float(r[j + 3]) - ln_mean * __uint_as_float(c4.w)) * ln_rstd + __uint_as_float(b4.w));
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 pk;
            pk.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
            pk.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            pk.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            pk.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            sts128(chunk_s + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pk);
          }
          __syncwarp();
          uint4 tv[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = k * 8 + (lane >> 2), j = lane & 3;
            tv[k] = lds128(chunk_s + rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4));
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (t_dst[k] >= 0) stg128(epi.out + t_dst[k] + col0 + (lane & 3) * 8, tv[k]);
          __syncwarp();
        }
        // publish this warp's part of the hidden tile: stores → gpu-scope fence → counter
        __threadfence();
        // the chunk buffer was accessed through the generic proxy; the next user may be a TMA write (fc2 tile)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) atomicAdd(args.ready + m_blk * 2 + static_cast<int>(cta_rank), 1);
      } else {
        // ------------------------- fc2: x += gamma·(acc + b) on the fp32 stream -------------------------------
        int dst_row = row, tok = 0;
        bool dst_ok = row_ok;
        if (epi.tok_in > 0) {
          const int b = row / epi.tok_in;
          tok = row - b * epi.tok_in;
          const int t2 = tok + epi.tok_shift;
          dst_row = b * epi.tok_out + t2;
          dst_ok = row_ok && t2 >= 0 && t2 < epi.tok_out;
        }
        float part_sum = 0.f, part_sq = 0.f;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          const int col_in_tile = half * cols_per_warp + c * 32;
          const int col0 = n_blk * bn + col_in_tile;
          uint32_t r[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                            static_cast<uint32_t>(acc * ACC_STRIDE + col_in_tile), r);
          float xres[32];
          mbar_wait(rbar, rphase);
          rphase ^= 1u;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 x4 = lds128(chunk_s + lane * 128 + ((j ^ (lane & 7)) << 4));
            xres[4 * j] = __uint_as_float(x4.x); xres[4 * j + 1] = __uint_as_float(x4.y);
            xres[4 * j + 2] = __uint_as_float(x4.z); xres[4 * j + 3] = __uint_as_float(x4.w);
          }
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const uint4 b4 = lds128(vec_s + (c * 32 + j) * 4);
            const uint4 g4 = lds128(vec_s + 512 + (c * 32 + j) * 4);
            v[j] = xres[j] + (__uint_as_float(r[j]) + __uint_as_float(b4.x)) * __uint_as_float(g4.x);
            v[j + 1] = xres[j + 1] + (__uint_as_float(r[j + 1]) + __uint_as_float(b4.y)) * __uint_as_float(g4.y);
            v[j + 2] = xres[j + 2] + (__uint_as_float(r[j + 2]) + __uint_as_float(b4.z)) * __uint_as_float(g4.z);
            v[j + 3] = xres[j + 3] + (__uint_as_float(r[j + 3]) + __uint_as_float(b4.w)) * __uint_as_float(g4.w);
          }
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              part_sum += v[j];
              part_sq = fmaf(v[j], v[j], part_sq);
            }
            if (epi.out != nullptr && dst_ok) {   // last block: concat write (prefix rows dropped)
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
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts128(chunk_s + lane * 128 + ((j ^ (lane & 7)) << 4),
                   make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                              __float_as_uint(v[4 * j + 3])));
          __syncwarp();
          uint4 tv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = k * 4 + (lane >> 3), j = lane & 7;
            tv[k] = lds128(chunk_s + rr * 128 + ((j ^ (rr & 7)) << 4));
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int rr = k * 4 + (lane >> 3), j = lane & 7;
            if (row0 + rr < M)
              stg128(epi.resid + static_cast<size_t>(row0 + rr) * epi.ld_resid + col0 + j * 4, tv[k]);
          }
          if (epi.xb_out != nullptr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int rr = k * 4 + (lane >> 3), j = lane & 7;
              if (row0 + rr < M) {
                uint2 pk;
                pk.x = pack_bf16x2(__uint_as_float(tv[k].x), __uint_as_float(tv[k].y));
                pk.y = pack_bf16x2(__uint_as_float(tv[k].z), __uint_as_float(tv[k].w));
                *reinterpret_cast<uint2*>(epi.xb_out + static_cast<size_t>(row0 + rr) * epi.ld_xb + col0 + j * 4) = pk;
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && c + 1 < chunks) {      // next chunk of this tile (the next tile's first chunk is issued above)
            mbar_expect_tx(rbar, RCHUNK_BYTES);
            tma_load_2d(chunk_p, &tmap_r, rbar, col0 + 32, row0);
          }
        }
        if (epi.stats_out != nullptr && row_ok)
          epi.stats_out[static_cast<size_t>(2 * n_blk + half) * M + row] = make_float2(part_sum, part_sq);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[acc], 0);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  cluster_sync();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_f32_map(CUtensorMap* map, const float* ptr, uint64_t rows, uint64_t cols, uint64_t ld) {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return BLB_ERR_DRIVER;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 4) % 16 != 0) return BLB_ERR_ALIGN;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : BLB_ERR_DRIVER;
}

template <int BN2>
int launch_fused(const CUtensorMap& a1, const CUtensorMap& b1, const CUtensorMap& a2, const CUtensorMap& b2,
                 const CUtensorMap& tr, const MlpArgs& args, cudaStream_t stream) {
  auto kern = mlp_fused_kernel<BN2>;
  static bool configured[BLB_MAX_DEVICES] = {};
  if (!configured[current_device()]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[current_device()] = true;
  }
  const int m_tiles = (args.M + 2 * BM - 1) / (2 * BM);
  const int tiles = m_tiles * (args.N1 / BN1 + args.N2 / BN2);
  int ctas = num_sms() & ~1;
  if (tiles * 2 < ctas) ctas = tiles * 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  // one record for both GEMMs: mode 1 (gelu) | folded LN | stats, N = N1, K = K1 marks the fused launch
  const long long tag = (1LL << 44) | (1LL << 43) | (1LL << 42) | (static_cast<long long>(args.N1) << 20) | args.K1;
  TimingScope ts(TIME_GEMM, 2.0 * args.M * (static_cast<double>(args.N1) * args.K1 + static_cast<double>(args.N2) * args.K2),
                 stream, tag);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a1, b1, a2, b2, tr, args);
  count_launch(1);
  return static_cast<int>(e);
}

}  // namespace

// hidden = gelu(LN-folded fc1(xb));  resid += gamma·(fc2(hidden) + b2)  [+ stats / bf16 copy / concat write], one launch.
// e1: bias, ln_stats, ln_colsum, ln_parts, ln_eps, out = hidden (ld_out = N1).   e2: an EPI_RESIDUAL epilogue.
// `ready`: device ints [ceil(M/256)][2], zeroed once by the caller; `ready_target` = launches so far (incl. this one) ×
// (N1/256) × 8.
int mlp_fused_bf16(const __nv_bfloat16* xb, int ldx, const __nv_bfloat16* W1, int ldw1, const __nv_bfloat16* W2,
                   int ldw2, int M, int N1, int K1, int N2, const GemmEpilogue& e1, const GemmEpilogue& e2, int* ready,
                   int ready_target, int reverse, cudaStream_t stream) {
  if (xb == nullptr || W1 == nullptr || W2 == nullptr || ready == nullptr || M <= 0) return BLB_ERR_ARG;
  if (e1.out == nullptr || e1.ln_stats == nullptr || e1.ln_colsum == nullptr || e1.ln_parts <= 0 || e2.resid == nullptr)
    return BLB_ERR_ARG;
  if (N1 % BN1 != 0 || e1.ld_out != N1) return BLB_ERR_SHAPE;
  const int bn2 = N2 % 256 == 0 ? 256 : N2 % 192 == 0 ? 192 : 0;
  if (bn2 == 0) return BLB_ERR_SHAPE;
  const int K2 = N1;
  CUtensorMap a1, b1, a2, b2, tr;
  int rc = make_tmap_bf16_2d(&a1, xb, static_cast<uint64_t>(M), static_cast<uint64_t>(K1), static_cast<uint64_t>(ldx), BM);
  if (rc == 0) rc = make_tmap_bf16_2d(&b1, W1, static_cast<uint64_t>(N1), static_cast<uint64_t>(K1), static_cast<uint64_t>(ldw1), BN1 / 2);
  if (rc == 0) rc = make_tmap_bf16_2d(&a2, e1.out, static_cast<uint64_t>(M), static_cast<uint64_t>(K2), static_cast<uint64_t>(N1), BM);
  if (rc == 0) rc = make_tmap_bf16_2d(&b2, W2, static_cast<uint64_t>(N2), static_cast<uint64_t>(K2), static_cast<uint64_t>(ldw2), static_cast<uint32_t>(bn2 / 2));
  if (rc == 0) rc = make_f32_map(&tr, e2.resid, static_cast<uint64_t>(M), static_cast<uint64_t>(N2), static_cast<uint64_t>(e2.ld_resid));
  if (rc != 0) return rc;
  MlpArgs args;
  args.M = M; args.N1 = N1; args.K1 = K1; args.N2 = N2; args.K2 = K2;
  static const int group = getenv("BLB_MLP_GROUP") ? atoi(getenv("BLB_MLP_GROUP")) : 6;
  args.group = group > 0 ? group : 6;
  args.reverse = reverse;
  args.ready = ready;
  args.ready_target = ready_target;
  args.e1 = e1;
  args.e2 = e2;
  if (bn2 == 256) return launch_fused<256>(a1, b1, a2, b2, tr, args, stream);
  return launch_fused<192>(a1, b1, a2, b2, tr, args, stream);
}

}  // namespace blb
