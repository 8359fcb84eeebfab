// attention_tc.cu — tcgen05/TMEM softmax attention for the two towers.
//
// timm Attention core: softmax(q·kᵀ·hd^-0.5)·v, no mask, 16 heads,
//   DINOv2-reg4  T = 261 tokens, head_dim 64   →  template <64, 16>  (256 keys + 16-key tail block, 5 of them real)
//   SigLIP       T = 256 tokens, head_dim 72   →  template <72, 0>   (d split 64 + 16; the 16-wide block is 32B-swizzled)
// One persistent CTA per SM walks (image, head) units; per unit it handles the first 256 query rows as two
// 128-row tiles (the T-256 remaining query rows of DINOv2 go to the mma.sync kernel in attention.cu).
//
//   warp 8     TMA producer: Q tiles, K, V straight out of the packed QKV GEMM output through a 4-D tensor map
//              {d, 3·H head slots, token, image}; rows >= T and d >= head_dim are zero-filled by TMA (OOB);
//              K and V are double-buffered across units
//   warp 9     MMA issuer:   S = Q·Kᵀ  (SS: M128 × N256[+16], fp32 in TMEM columns [0, 256+KX))
//                            O = P·V   (TS: A = bf16 P in TMEM, B = V as an MN-major smem operand)
//   warp 10    TMEM allocator   (control warps carry the highest ids: the sub-partition arbiter prefers the highest
//              eligible warp id, so TMA / MMA issue never queues behind the MUFU-bound softmax warps)
//   warps 0-7  softmax + epilogue: two groups of 4 warps split the S columns; a thread owns query row = TMEM lane.
//              It pulls its whole slice of the S row into registers with ONE pass of tcgen05.ld and releases S at
//              once (s_free) — the next tile's Q·Kᵀ runs on the tensor pipe underneath this tile's softmax — then
//              row max (partials combined through smem), exp2, row sum in fp32, P rounded to bf16 and written back
//              to TMEM with tcgen05.st, and finally O / rowsum → bf16 → global for the previous tile.
//   TMEM columns: S fp32 [0, 256+KX) | P bf16x2 [.., +(256+KX)/2) | O fp32 [.., +64/80)
#include <algorithm>
#include <cmath>

#include "gemm.h"
#include "ptx.cuh"

namespace blb {

namespace {

constexpr int QT = 128;        // query rows per tile (UMMA M)
constexpr int KMAIN = 256;     // keys in the main block (UMMA N of S)
constexpr int TC_THREADS = 384;   // 4 control warps + 8 softmax/epilogue warps

// ---- descriptors ------------------------------------------------------------------------------------
// layout_type: 2 = SWIZZLE_128B, 6 = SWIZZLE_32B.  K-major operands: SBO = 8 rows * row_bytes.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}
// instruction descriptor: bf16 x bf16 -> f32; A K-major; B K-major (b_mn = 0) or MN-major (b_mn = 1)
__host__ __device__ constexpr uint32_t idesc_bf16(int m, int n, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// D[tmem] (+)= A[tmem] · B[smem]: A = bf16 P, row i in TMEM lane i, two consecutive k per 32-bit column
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int HD, int KX>
struct AttnCfg {
  static constexpr bool SPLIT_D = HD > 64;                       // head_dim 72: d = 64 (SW128) + 16 (SW32, 8 real)
  static constexpr int HDP = SPLIT_D ? 80 : 64;                  // O columns
  static constexpr int Q_MAIN = QT * 128;                        // 16 KB
  static constexpr int Q_X = SPLIT_D ? QT * 32 : 0;              // 4 KB
  static constexpr int Q_BYTES = Q_MAIN + Q_X;
  static constexpr int KV_MAIN = KMAIN * 128;                    // 32 KB
  static constexpr int KV_TAIL = KX * 128;                       // 2 KB (16 keys)
  static constexpr int KV_X = SPLIT_D ? KMAIN * 32 : 0;          // 8 KB
  static constexpr int KV_BYTES = KV_MAIN + KV_TAIL + KV_X;
  static constexpr int OFF_Q0 = 0;
  static constexpr int OFF_Q1 = OFF_Q0 + Q_BYTES;
  static constexpr int OFF_K = OFF_Q1 + Q_BYTES;                 // [2 unit parities]
  static constexpr int OFF_V = OFF_K + 2 * KV_BYTES;             // [2 unit parities]
  static constexpr int OFF_XCHG = OFF_V + 2 * KV_BYTES;          // row max [2][128] + row sums [2][2][128] (fp32)
  static constexpr int OFF_OST = OFF_XCHG + 6 * QT * 4;          // per-warp [32 rows x 64 B] output staging chunks
  static constexpr int OFF_BAR = OFF_OST + 8 * 2048;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  // TMEM columns: S fp32 [0, 256+KX) | P bf16x2 [P_COL, P_COL + (256+KX)/2) | O fp32 [O_COL, O_COL + HDP)
  static constexpr int S_COLS = KMAIN + KX;
  static constexpr int P_COL = S_COLS;
  static constexpr int P_COLS = S_COLS / 2;
  static constexpr int O_COL = P_COL + P_COLS;
  static constexpr int NREG_S = 128 + KX / 2;                    // S columns one softmax thread keeps in registers
  static_assert(Q_BYTES % 1024 == 0 && KV_BYTES % 1024 == 0 && KV_MAIN % 1024 == 0, "1 KB aligned blocks");
  static_assert(O_COL + HDP <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "smem budget");
};

struct AttnMaps {
  CUtensorMap q_main, kv_main, kv_tail, q_x, kv_x;
};

__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x4(uint32_t taddr, const uint32_t (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int HD, int KX>
__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ AttnMaps maps, __nv_bfloat16* __restrict__ out, int B, int T, int H,
                    float scale_log2, long long* trace, int reverse) {
  // trace (debug only, normally nullptr): CTA 0 records clock64() at pipeline events of tiles [8, 16):
  // trace[(g-8)*16 + e]; e: 0 mma:s_free(g) seen, 1 mma:S(g+1) issued, 2 mma:p_full(g)+o_empty seen, 3 mma:PV(g) issued,
  //   8 sm:s_full seen, 9 sm:S in registers (s_free), 10 sm:max exchanged, 11 sm:p_empty seen, 12 sm:P stored/p_full,
  //   13 sm:epilogue o_full seen, 14 sm:epilogue done
#define BLB_TRACE(g_, e_)                                                                   \
  do {                                                                                      \
    if (trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g_) >= 8 && (g_) < 16)            \
      trace[((g_) - 8) * 16 + (e_)] = clock64();                                                                \
  } while (0)
  using Cfg = AttnCfg<HD, KX>;
  extern __shared__ uint8_t smem_raw_attn[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_attn) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* q_full = bars;         // [2 tiles of a unit]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* k_full = bars + 4;     // [2 unit parities]
  uint64_t* k_empty = bars + 6;    // [2]
  uint64_t* v_full = bars + 8;     // [2]
  uint64_t* v_empty = bars + 10;   // [2]
  uint64_t* s_full = bars + 12;    // MMA → softmax: S(g) is in TMEM
  uint64_t* s_free = bars + 13;    // softmax → MMA: S(g) is in registers, TMEM columns reusable
  uint64_t* p_full = bars + 14;    // softmax → MMA: P(g) is in TMEM
  uint64_t* p_empty = bars + 15;   // MMA → softmax: PV(g) has consumed P(g)
  uint64_t* o_full = bars + 16;    // MMA → epilogue
  uint64_t* o_empty = bars + 17;   // epilogue → MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int W_TMA = 8, W_MMA = 9, W_ALLOC = 10;
  const int n_units = B * H;
  const int my_units = (n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int G = 2 * my_units;      // tiles this CTA processes

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.q_main);
    tma_prefetch_desc(&maps.kv_main);
    if (KX > 0) tma_prefetch_desc(&maps.kv_tail);
    if (Cfg::SPLIT_D) { tma_prefetch_desc(&maps.q_x); tma_prefetch_desc(&maps.kv_x); }
  }
  if (warp == W_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, 8); mbar_init(p_full, 8); mbar_init(p_empty, 1);
    mbar_init(o_full, 1); mbar_init(o_empty, 8);
    fence_mbar_init();
  }
  if (warp == W_ALLOC) tmem_alloc<1>(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();   // PDL: the prologue above overlapped with the QKV GEMM's tail
  pdl_wait();

  uint8_t* sQ[2] = {smem + Cfg::OFF_Q0, smem + Cfg::OFF_Q1};
  uint8_t* sK[2] = {smem + Cfg::OFF_K, smem + Cfg::OFF_K + Cfg::KV_BYTES};
  uint8_t* sV[2] = {smem + Cfg::OFF_V, smem + Cfg::OFF_V + Cfg::KV_BYTES};

  if (warp >= 8) {
    setmaxnreg_dec<56>();          // hand registers to the softmax warps
    if (warp == W_TMA) {
      // ============================== TMA producer (warp-uniform loop, one elected lane issues) ===========
      {
        for (int i = 0; i < my_units; ++i) {
          const int u_i = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
          const int u = reverse ? n_units - 1 - u_i : u_i;
          const int b = u / H, h = u - b * H;
          const int kb = i & 1;                                          // K/V buffer of this unit
          const uint32_t ph = static_cast<uint32_t>(i & 1);              // Q barriers: one use per unit
          const uint32_t kph = static_cast<uint32_t>((i >> 1) & 1);      // K/V barriers: one use per two units
          mbar_wait(&k_empty[kb], kph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&k_full[kb], Cfg::KV_BYTES);
            tma_load_4d(sK[kb], &maps.kv_main, &k_full[kb], 0, H + h, 0, b);
            if (KX > 0) tma_load_4d(sK[kb] + Cfg::KV_MAIN, &maps.kv_tail, &k_full[kb], 0, H + h, KMAIN, b);
            if (Cfg::SPLIT_D)
              tma_load_4d(sK[kb] + Cfg::KV_MAIN + Cfg::KV_TAIL, &maps.kv_x, &k_full[kb], 64, H + h, 0, b);
          }
          __syncwarp();
          for (int t = 0; t < 2; ++t) {
            mbar_wait(&q_empty[t], ph ^ 1u);
            if (elect_one()) {
              mbar_expect_tx(&q_full[t], Cfg::Q_BYTES);
              tma_load_4d(sQ[t], &maps.q_main, &q_full[t], 0, h, t * QT, b);
              if (Cfg::SPLIT_D) tma_load_4d(sQ[t] + Cfg::Q_MAIN, &maps.q_x, &q_full[t], 64, h, t * QT, b);
            }
            __syncwarp();
          }
          mbar_wait(&v_empty[kb], kph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&v_full[kb], Cfg::KV_BYTES);
            tma_load_4d(sV[kb], &maps.kv_main, &v_full[kb], 0, 2 * H + h, 0, b);
            if (KX > 0) tma_load_4d(sV[kb] + Cfg::KV_MAIN, &maps.kv_tail, &v_full[kb], 0, 2 * H + h, KMAIN, b);
            if (Cfg::SPLIT_D)
              tma_load_4d(sV[kb] + Cfg::KV_MAIN + Cfg::KV_TAIL, &maps.kv_x, &v_full[kb], 64, 2 * H + h, 0, b);
          }
          __syncwarp();
        }
      }
    } else if (warp == W_MMA) {
      // =============================== MMA issuer (warp-uniform loop, one elected lane issues) ===========
      {
        constexpr uint32_t idesc_s_main = idesc_bf16(QT, KMAIN, 0);
        constexpr uint32_t idesc_s_tail = idesc_bf16(QT, 16, 0);
        constexpr uint32_t idesc_o_main = idesc_bf16(QT, 64, 1);
        constexpr uint32_t idesc_o_x = idesc_bf16(QT, 16, 1);
        // S(g) = Q·Kᵀ into TMEM columns [0, 256+KX)
        auto issue_s = [&](int g) {
          const int t = g & 1, i = g >> 1, kb = i & 1;
          if (t == 0) mbar_wait(&k_full[kb], static_cast<uint32_t>((i >> 1) & 1));
          mbar_wait(&q_full[t], static_cast<uint32_t>(i & 1));
          tc_fence_after();
          if (elect_one()) {
            const uint32_t qa = smem_u32(sQ[t]), ka = smem_u32(sK[kb]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16<1>(tmem_base, make_desc(qa, 1024, 2) + 2 * k, make_desc(ka, 1024, 2) + 2 * k, idesc_s_main,
                           k > 0 ? 1u : 0u);
            if (Cfg::SPLIT_D)
              umma_bf16<1>(tmem_base, make_desc(qa + Cfg::Q_MAIN, 256, 6),
                           make_desc(ka + Cfg::KV_MAIN + Cfg::KV_TAIL, 256, 6), idesc_s_main, 1u);
            if (KX > 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16<1>(tmem_base + KMAIN, make_desc(qa, 1024, 2) + 2 * k,
                             make_desc(ka + Cfg::KV_MAIN, 1024, 2) + 2 * k, idesc_s_tail, k > 0 ? 1u : 0u);
            }
            umma_commit<1>(s_full);
            umma_commit<1>(&q_empty[t]);
            if (t == 1) umma_commit<1>(&k_empty[kb]);
          }
          __syncwarp();
        };
        if (G > 0) issue_s(0);
        for (int g = 0; g < G; ++g) {
          if (g + 1 < G) {
            // softmax(g) holds S(g) in registers → the next Q·Kᵀ runs underneath this tile's softmax
            mbar_wait(s_free, static_cast<uint32_t>(g & 1));
            tc_fence_after();
            BLB_TRACE(g, 0);
            issue_s(g + 1);
            BLB_TRACE(g, 1);
          }
          // ---------------- O(g) = P(g) · V   (A = P in TMEM, B = V as an MN-major smem operand) ----------------
          const int t = g & 1, i = g >> 1, kb = i & 1;
          mbar_wait(p_full, static_cast<uint32_t>(g & 1));
          if (t == 0) mbar_wait(&v_full[kb], static_cast<uint32_t>((i >> 1) & 1));
          mbar_wait(o_empty, static_cast<uint32_t>(g & 1) ^ 1u);   // epilogue(g-1) drained the O accumulator
          tc_fence_after();
          BLB_TRACE(g, 2);
          if (elect_one()) {
            const uint32_t o_col = tmem_base + Cfg::O_COL;
            const uint32_t p_col = tmem_base + Cfg::P_COL;
            const uint32_t va = smem_u32(sV[kb]);
#pragma unroll
            for (int j = 0; j < 16; ++j) {   // 16 keys per step: 8 packed P columns, V rows 16j..16j+15
              umma_bf16_ts(o_col, p_col + j * 8, make_desc(va + j * 2048, 1024, 2), idesc_o_main, j > 0 ? 1u : 0u);
              if (Cfg::SPLIT_D)
                umma_bf16_ts(o_col + 64, p_col + j * 8,
                             make_desc(va + Cfg::KV_MAIN + Cfg::KV_TAIL + j * 512, 256, 6), idesc_o_x, j > 0 ? 1u : 0u);
            }
            if (KX > 0)
              umma_bf16_ts(o_col, p_col + 128, make_desc(va + Cfg::KV_MAIN, 1024, 2), idesc_o_main, 1u);
            umma_commit<1>(p_empty);
            umma_commit<1>(o_full);
            if (t == 1) umma_commit<1>(&v_empty[kb]);
          }
          __syncwarp();
          BLB_TRACE(g, 3);
        }
      }
    }
  } else {
    // ============================ softmax + epilogue ===============================================
    // 8 warps: group wg = 0/1 owns S columns [128·wg, 128·wg+128) plus 8 of the 16 tail columns; both groups see
    // all 128 rows (warps w and w+4 share TMEM lane quarter w%4), i.e. two softmax warps per SM sub-partition.
    setmaxnreg_inc<216>();
    const int q = warp & 3;
    const int wg = warp >> 2;
    const int row = q * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int D = H * HD;
    float* xmax = reinterpret_cast<float*>(smem + Cfg::OFF_XCHG);            // [2 groups][128 rows]
    float* xsum = xmax + 2 * QT;                                             // [2 tile parities][2 groups][128]
    float sum_prev = 0.f;
    const int col_base = wg * 128;
    const int tail_key0 = KMAIN + wg * 8;                // first key of this group's 8 tail columns

    auto epilogue = [&](int gp, float own_sum) {
      const int t = gp & 1, i = gp >> 1;
      const int u_i = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int u = reverse ? n_units - 1 - u_i : u_i;
      const int b = u / H, h = u - b * H;
      const float inv = 1.0f / (own_sum + xsum[(gp & 1) * 2 * QT + (wg ^ 1) * QT + row]);
      mbar_wait(o_full, static_cast<uint32_t>(gp & 1));
      tc_fence_after();
      if (warp == 0 && lane == 0) BLB_TRACE(gp + 1, 13);
      const uint32_t o_addr = lane_addr + Cfg::O_COL;
      __nv_bfloat16* dst = out + (static_cast<size_t>(b) * T + t * QT + row) * D + h * HD;
      uint32_t r[32];
      uint32_t rx[16];
      tmem_ld_32x32(o_addr + wg * 32, r);               // this group's 32 of the 64 main O columns
      if (Cfg::SPLIT_D && wg == 1) tmem_ld_32x16(o_addr + 64, rx);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);              // O is in registers: PV(g) may overwrite the accumulator
      // stage the lane's 64-byte row slice in smem (16-byte piece j at j ^ ((row>>1)&3): conflict-free both ways),
      // then store transposed so that one instruction writes 8 rows x 64 contiguous bytes instead of 32 x 16
      const uint32_t ob = smem_u32(smem + Cfg::OFF_OST) + warp * 2048;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(r[8 * j]) * inv, __uint_as_float(r[8 * j + 1]) * inv);
        pk.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]) * inv, __uint_as_float(r[8 * j + 3]) * inv);
        pk.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]) * inv, __uint_as_float(r[8 * j + 5]) * inv);
        pk.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]) * inv, __uint_as_float(r[8 * j + 7]) * inv);
        sts128(ob + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pk);
      }
      __syncwarp();
      {
        __nv_bfloat16* slab = out + (static_cast<size_t>(b) * T + t * QT + q * 32) * D + h * HD + wg * 32;
        uint4 tv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rr = k * 8 + (lane >> 2), j = lane & 3;
          tv[k] = lds128(ob + rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int rr = k * 8 + (lane >> 2), j = lane & 3;
          stg128(slab + static_cast<size_t>(rr) * D + j * 8, tv[k]);
        }
      }
      __syncwarp();
      if (Cfg::SPLIT_D && wg == 1) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(rx[0]) * inv, __uint_as_float(rx[1]) * inv);
        pk.y = pack_bf16x2(__uint_as_float(rx[2]) * inv, __uint_as_float(rx[3]) * inv);
        pk.z = pack_bf16x2(__uint_as_float(rx[4]) * inv, __uint_as_float(rx[5]) * inv);
        pk.w = pack_bf16x2(__uint_as_float(rx[6]) * inv, __uint_as_float(rx[7]) * inv);
        *reinterpret_cast<uint4*>(dst + 64) = pk;        // d 64..71 (72..79 are padding)
      }
      if (warp == 0 && lane == 0) BLB_TRACE(gp + 1, 14);
    };

    for (int g = 0; g < G; ++g) {
      mbar_wait(s_full, static_cast<uint32_t>(g & 1));
      tc_fence_after();
      if (warp == 0 && lane == 0) BLB_TRACE(g, 8);
      // ---- this thread's slice of the S row → registers in one pass, then release S ----
      uint32_t sv[Cfg::NREG_S];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tmem_ld_32x32(lane_addr + col_base + c * 32, reinterpret_cast<uint32_t(&)[32]>(sv[c * 32]));
      if (KX > 0) tmem_ld_32x8(lane_addr + tail_key0, reinterpret_cast<uint32_t(&)[8]>(sv[128]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      if (warp == 0 && lane == 0) BLB_TRACE(g, 9);
      // ---- row max ----
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 128; j += 4) {
        m0 = fmaxf(m0, __uint_as_float(sv[j]));
        m1 = fmaxf(m1, __uint_as_float(sv[j + 1]));
        m2 = fmaxf(m2, __uint_as_float(sv[j + 2]));
        m3 = fmaxf(m3, __uint_as_float(sv[j + 3]));
      }
      if (KX > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (tail_key0 + j < T) m0 = fmaxf(m0, __uint_as_float(sv[128 + j]));
      }
      float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      xmax[wg * QT + row] = m;
      asm volatile("bar.sync 1, 256;" ::: "memory");    // the two groups exchange their partial maxima
      m = fmaxf(m, xmax[(wg ^ 1) * QT + row]);
      const float ms = m * scale_log2;
      if (warp == 0 && lane == 0) BLB_TRACE(g, 10);
      // The two groups are deliberately out of phase: group 1 drains the previous tile's O (epilogue: TMEM load,
      // scale, coalesced stores — no MUFU) while group 0 already runs its exp2 stream, and group 0 does its half of
      // that epilogue after its exp2s while group 1 is still in MUFU.  The MUFU pipe (the bottleneck of this kernel,
      // 16 exp2/clk/SM) therefore sees a continuous stream instead of two warps per sub-partition stalling together.
      if (wg == 1 && g > 0) epilogue(g - 1, sum_prev);
      // P(g-1) must have been consumed by PV(g-1) before it is overwritten
      if (g > 0) mbar_wait(p_empty, static_cast<uint32_t>((g - 1) & 1));
      if (warp == 0 && lane == 0) BLB_TRACE(g, 11);
      // ---- p = 2^(s*c - m*c), partial row sum, bf16 P → TMEM (16 packed columns per 32 keys) ----
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c * 32 + j]), scale_log2, -ms));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sv[c * 32 + j + 1]), scale_log2, -ms));
          const float p2 = ex2_approx(fmaf(__uint_as_float(sv[c * 32 + j + 2]), scale_log2, -ms));
          const float p3 = ex2_approx(fmaf(__uint_as_float(sv[c * 32 + j + 3]), scale_log2, -ms));
          s0 += p0 + p1;
          s1 += p2 + p3;
          pk[j / 2] = pack_bf16x2(p0, p1);
          pk[j / 2 + 1] = pack_bf16x2(p2, p3);
        }
        tmem_st_32x16(lane_addr + Cfg::P_COL + (col_base + c * 32) / 2, pk);
      }
      if (KX > 0) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float p0 = (tail_key0 + j < T) ? ex2_approx(fmaf(__uint_as_float(sv[128 + j]), scale_log2, -ms)) : 0.f;
          const float p1 =
              (tail_key0 + j + 1 < T) ? ex2_approx(fmaf(__uint_as_float(sv[128 + j + 1]), scale_log2, -ms)) : 0.f;
          s0 += p0 + p1;
          pk[j / 2] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x4(lane_addr + Cfg::P_COL + tail_key0 / 2, pk);
      }
      const float sum = s0 + s1;
      xsum[(g & 1) * 2 * QT + wg * QT + row] = sum;   // read by the other group in this tile's epilogue
      tmem_st_wait();               // P is in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (warp == 0 && lane == 0) BLB_TRACE(g, 12);
      // ---- epilogue of the previous tile (its PV ran underneath this tile's softmax; the other group's partial
      //      sum of tile g-1 was published before this tile's bar.sync) ----
      if (wg == 0 && g > 0) epilogue(g - 1, sum_prev);
      sum_prev = sum;
    }
    if (G > 0) {
      asm volatile("bar.sync 1, 256;" ::: "memory");    // partial sums of the last tile are visible
      epilogue(G - 1, sum_prev);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
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

// 4-D view of the packed qkv tensor: {d (hd), head slot (3H), token (T), image (B)}
int make_qkv_map(CUtensorMap* map, const void* qkv, int B, int T, int H, int hd, int box_d, int box_rows,
                 CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return BLB_ERR_DRIVER;
  const uint64_t D3 = static_cast<uint64_t>(3) * H * hd;
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(hd), static_cast<cuuint64_t>(3 * H), static_cast<cuuint64_t>(T),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(hd) * 2, D3 * 2, D3 * 2 * static_cast<cuuint64_t>(T)};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_d), 1, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(qkv), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : BLB_ERR_DRIVER;
}

long long* g_attn_trace = nullptr;

template <int HD, int KX>
int launch_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, cudaStream_t stream, int reverse) {
  using Cfg = AttnCfg<HD, KX>;
  AttnMaps maps;
  int rc = make_qkv_map(&maps.q_main, qkv, B, T, H, HD, 64, QT, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == 0) rc = make_qkv_map(&maps.kv_main, qkv, B, T, H, HD, 64, KMAIN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == 0) rc = make_qkv_map(&maps.kv_tail, qkv, B, T, H, HD, 64, 16, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == 0) rc = make_qkv_map(&maps.q_x, qkv, B, T, H, HD, 16, QT, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc == 0) rc = make_qkv_map(&maps.kv_x, qkv, B, T, H, HD, 16, KMAIN, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc != 0) return rc;
  auto kern = attention_tc_kernel<HD, KX>;
  static bool configured[BLB_MAX_DEVICES] = {};   // the attribute is per device
  if (!configured[current_device()]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[current_device()] = true;
  }
  const int grid = std::min(num_sms(), B * H);
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  TimingScope ts(TIME_ATTENTION, 4.0 * B * H * 256.0 * T * HD, stream);
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, stream, maps, out, B, T, H,
                              scale_log2, g_attn_trace, reverse);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

void attention_set_trace(long long* device_buffer) { g_attn_trace = device_buffer; }

// Handles query rows [0, 256) of every (image, head); returns BLB_ERR_SHAPE when (T, hd) is not one of the two
// tower configurations this kernel is built for (the caller then uses the mma.sync kernel for everything).
int attention_tc_first256(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd,
                          cudaStream_t stream, int reverse) {
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) != 0) return BLB_ERR_ALIGN;
  if (hd == 64 && T == 256) return launch_tc<64, 0>(qkv, out, B, T, H, stream, reverse);
  if (hd == 64 && T > 256 && T <= 272) return launch_tc<64, 16>(qkv, out, B, T, H, stream, reverse);
  if (hd == 72 && T == 256) return launch_tc<72, 0>(qkv, out, B, T, H, stream, reverse);
  return BLB_ERR_SHAPE;
}

}  // namespace blb
