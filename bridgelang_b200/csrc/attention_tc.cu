// attention_tc.cu — tcgen05/TMEM softmax attention for the two towers.
//
// timm Attention core: softmax(q·kᵀ·hd^-0.5)·v, no mask, 16 heads,
//   DINOv2-reg4  T = 261 tokens, head_dim 64   →  template <64, 16>  (256 keys + 16-key tail block, 5 of them real)
//   SigLIP       T = 256 tokens, head_dim 72   →  template <72, 0>   (d split 64 + 16; the 16-wide block is 32B-swizzled)
// One persistent CTA per SM walks (image, head) units; per unit it handles the first 256 query rows as two
// 128-row tiles and — DINOv2 only — the T-256 = 5 remaining query rows as a third "tail" tile against the K/V that
// are already in shared memory (round 1 sent them to a separate mma.sync kernel that re-read all K/V: 274 MB per
// launch).  The tail tile's Q rows are REPLICATED into all four TMEM lane quarters (four 8-row TMA boxes), so that all
// eight softmax warps share its 261 score columns (34 each instead of 136 on two warps); every warp zero-fills the P
// columns it does not own, the four quarters' partial O / row sums are added through shared memory.
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

#include <atomic>

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
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
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
  static constexpr int OFF_XCHG = OFF_V + 2 * KV_BYTES;          // partial row maxima [2 tile parities][2 groups][128] (fp32)
  static constexpr int OFF_OST = OFF_XCHG + 4 * QT * 4;          // per-warp [32 rows x 64 B] output staging chunks
  static constexpr int OFF_ONES = OFF_OST + 8 * 2048;            // 512 B of bf16 1.0: the B operand of the row-sum MMAs
  // tail tile (KX > 0): partial O [4 quarters][8 rows][64] + partial sums [4][8] + partial maxima [8 warps][32] (fp32)
  static constexpr int OFF_TAIL = OFF_ONES + 512;
  static constexpr int TAIL_BYTES = KX > 0 ? (4 * 8 * 64 + 4 * 8 + 8 * 32) * 4 : 0;
  static constexpr int OFF_BAR = OFF_TAIL + TAIL_BYTES;
  static constexpr int NT = KX > 0 ? 3 : 2;                      // tiles per (image, head) unit
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  // TMEM columns: S fp32 [0, 256+KX) | P bf16x2 [P_COL, P_COL + (256+KX)/2) | O fp32 [O_COL, O_COL + HDP) |
  //               row sums [SUM_COL, SUM_COL + 16)  (P · ones: every column holds Σ_k P[row, k])
  static constexpr int S_COLS = KMAIN + KX;
  static constexpr int P_COL = S_COLS;
  static constexpr int P_COLS = S_COLS / 2;
  static constexpr int O_COL = P_COL + P_COLS;
  static constexpr int SUM_COL = O_COL + HDP;
  static constexpr int NREG_S = 128 + KX / 2;                    // S columns one softmax thread keeps in registers
  static_assert(Q_BYTES % 1024 == 0 && KV_BYTES % 1024 == 0 && KV_MAIN % 1024 == 0, "1 KB aligned blocks");
  static_assert(SUM_COL + 16 <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "smem budget");
};

struct AttnMaps {
  CUtensorMap q_main, kv_main, kv_tail, q_x, kv_x, q_tail;
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

// Compiler-level fence for registers filled by an asynchronous tcgen05.ld: placed right after tcgen05.wait::ld, it makes
// every later use of the values depend on a statement that cannot move above the wait (emits no instruction).
__device__ __forceinline__ void reg_fence16(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ void reg_fence8(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}

// 3-input max (FMNMX3 on sm_100a): halves the instruction count of the row-max pass
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// named barrier over the 64 threads of one row quarter (softmax warps q and q+4): id 1 + q
__device__ __forceinline__ void pair_barrier(int q) { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); }

#ifndef BLB_ATTN_PCH
#define BLB_ATTN_PCH 2
#endif
constexpr int PCH = BLB_ATTN_PCH;   // P / PV chunks per tile: 128/PCH keys of each group per chunk (4 or 2)
static_assert(PCH == 4 || PCH == 2, "PCH");

// 2^t for t <= 0 on the FMA pipe (the MUFU pipe, 16 results/clk/SM, is this kernel's ceiling): t = n + f with
// n = round(t) taken from the low mantissa bits of t + 1.5·2^23 and f in [-0.5, 0.5]; degree-3 minimax fit of 2^f
// (relative error 7.5e-5 = 1/26 of the bf16 half-ulp that P is rounded to), exponent added with one integer op.
#ifndef BLB_ATTN_POLY_MASK
#define BLB_ATTN_POLY_MASK 0   // bit i set: element i of every group of four uses exp2_poly instead of MUFU.EX2
                               // (measured round 2: 8 → -3 %; the softmax warps are short of issue slots, not of MUFU)
#endif
__device__ __forceinline__ float exp2_poly(float t) {
  t = fmaxf(t, -126.0f);
  const float r = t + 12582912.0f;
  const float f = t - (r - 12582912.0f);
  float p = 0.0551716685f;
  p = fmaf(p, f, 0.242611125f);
  p = fmaf(p, f, 0.693260968f);
  p = fmaf(p, f, 0.999928057f);
  return __uint_as_float(__float_as_uint(p) + (__float_as_uint(r) << 23));
}

__device__ __forceinline__ void tmem_ld_32x2(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

template <int HD, int KX, bool TRACE>
__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ AttnMaps maps, __nv_bfloat16* __restrict__ out, int B, int T, int H,
                    float scale_log2, long long* trace, int reverse) {
  // trace (debug only, TRACE instantiation): every warp of CTA 0 records clock64() at pipeline events of tiles [8, 16):
  // trace[((g-8)*12 + warp)*16 + e]; MMA warp e: 0 s_free(g) seen, 1 S(g+1) issued, 2 o_empty seen, 3 PV(g) issued,
  //   4+c p_full[c](g) seen;  softmax warps e: 8 s_full seen, 9 S in registers (s_free), 10 max exchanged,
  //   4+c P chunk c stored, 12 last P chunk published, 13 O(g-1) in registers, 14 epilogue(g-1) stores issued
#define BLB_TRACE(g_, e_)                                                                   \
  do {                                                                                      \
    if (TRACE && trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g_) >= 8 && (g_) < 16)   \
      trace[(((g_) - 8) * 12 + (threadIdx.x >> 5)) * 16 + (e_)] = clock64();                                  \
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
  uint64_t* o_full = bars + 14;    // MMA → epilogue: every PV chunk of tile g has retired
  uint64_t* o_empty = bars + 15;   // epilogue → MMA: O(g) is in registers
  uint64_t* p_full = bars + 16;    // [PCH] softmax → MMA: P chunk c of tile g is in TMEM
  uint64_t* p_empty = bars + 16 + PCH;   // [PCH] MMA → softmax: PV chunk c of tile g has consumed its P columns
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16 + 2 * PCH);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int W_TMA = 8, W_MMA = 9, W_ALLOC = 10;
  const int n_units = B * H;
  const int my_units = (n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  constexpr int NT = Cfg::NT;
  const int G = NT * my_units;     // tiles this CTA processes

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.q_main);
    tma_prefetch_desc(&maps.kv_main);
    if (KX > 0) { tma_prefetch_desc(&maps.kv_tail); tma_prefetch_desc(&maps.q_tail); }
    if (Cfg::SPLIT_D) { tma_prefetch_desc(&maps.q_x); tma_prefetch_desc(&maps.kv_x); }
  }
  if (warp == W_MMA && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1); mbar_init(s_free, 8);
    mbar_init(o_full, 1); mbar_init(o_empty, 8);
    for (int c = 0; c < PCH; ++c) { mbar_init(&p_full[c], 8); mbar_init(&p_empty[c], 1); }
    fence_mbar_init();
  }
  if (warp == W_ALLOC) tmem_alloc<1>(tmem_slot, 512);
  if (warp == 0) {           // 512 B of bf16 1.0 (layout-invariant: every element is one) → read by the async proxy (UMMA)
    reinterpret_cast<uint4*>(smem + Cfg::OFF_ONES)[lane] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    fence_proxy_async_smem();
  }
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
          for (int t = 0; t < NT; ++t) {
            const int g = i * NT + t, qb = g & 1;                        // Q buffers alternate per tile
            mbar_wait(&q_empty[qb], static_cast<uint32_t>((g >> 1) & 1) ^ 1u);
            if (elect_one()) {
              if (KX > 0 && t == 2) {
                // tail tile: query rows 256..263 (rows >= T are zero-filled) into rows 0..7 of every 32-row quarter
                mbar_expect_tx(&q_full[qb], 4 * 1024);
#pragma unroll
                for (int r = 0; r < 4; ++r) tma_load_4d(sQ[qb] + r * 4096, &maps.q_tail, &q_full[qb], 0, h, 2 * QT, b);
              } else {
                mbar_expect_tx(&q_full[qb], Cfg::Q_BYTES);
                tma_load_4d(sQ[qb], &maps.q_main, &q_full[qb], 0, h, t * QT, b);
                if (Cfg::SPLIT_D) tma_load_4d(sQ[qb] + Cfg::Q_MAIN, &maps.q_x, &q_full[qb], 64, h, t * QT, b);
              }
            }
            __syncwarp();
            if (t == 0) {
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
          const int t = g % NT, i = g / NT, kb = i & 1, qb = g & 1;
          if (t == 0) mbar_wait(&k_full[kb], static_cast<uint32_t>((i >> 1) & 1));
          BLB_TRACE(g - 1, 8);
          mbar_wait(&q_full[qb], static_cast<uint32_t>((g >> 1) & 1));
          tc_fence_after();
          BLB_TRACE(g - 1, 9);
          if (elect_one()) {
            const uint32_t qa = smem_u32(sQ[qb]), ka = smem_u32(sK[kb]);
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
            umma_commit<1>(&q_empty[qb]);
            if (t == NT - 1) umma_commit<1>(&k_empty[kb]);
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
          // issued chunk by chunk as the softmax warps publish P: chunk c = keys [32c, 32c+32) of group 0 and
          // [128+32c, 128+32c+32) of group 1 (+ the 16-key tail block with the last chunk), so that only the last
          // chunk's MMAs are left when the tile's exp2 stream ends
          const int t = g % NT, i = g / NT, kb = i & 1;
          const uint32_t ph = static_cast<uint32_t>(g & 1);
          if (t == 0) mbar_wait(&v_full[kb], static_cast<uint32_t>((i >> 1) & 1));
          mbar_wait(o_empty, ph ^ 1u);   // epilogue(g-1) holds O(g-1) in registers
          BLB_TRACE(g, 2);
          const uint32_t o_col = tmem_base + Cfg::O_COL;
          const uint32_t p_col = tmem_base + Cfg::P_COL;
          const uint32_t sum_col = tmem_base + Cfg::SUM_COL;
          const uint32_t va = smem_u32(sV[kb]);
          const uint64_t ones = make_desc(smem_u32(smem + Cfg::OFF_ONES), 256, 6);   // [16 keys x 16] of 1.0
#pragma unroll 1
          for (int c = 0; c < PCH; ++c) {
            mbar_wait(&p_full[c], ph);
            tc_fence_after();
            BLB_TRACE(g, 4 + c);
            if (elect_one()) {
              constexpr int MPG = 8 / PCH;         // 16-key MMAs per group and chunk
#pragma unroll
              for (int jj = 0; jj < 2 * MPG; ++jj) {   // 16 keys per MMA: 8 packed P columns, V rows 16j..16j+15
                const int j = (jj / MPG) * 8 + MPG * c + (jj % MPG);
                umma_bf16_ts(o_col, p_col + j * 8, make_desc(va + j * 2048, 1024, 2), idesc_o_main,
                             (c | jj) != 0 ? 1u : 0u);
                if (Cfg::SPLIT_D)
                  umma_bf16_ts(o_col + 64, p_col + j * 8,
                               make_desc(va + Cfg::KV_MAIN + Cfg::KV_TAIL + j * 512, 256, 6), idesc_o_x,
                               (c | jj) != 0 ? 1u : 0u);
                // row sums on the tensor pipe: Σ_k P[row, k] · 1 (of the bf16-rounded P, i.e. exactly what PV uses)
                umma_bf16_ts(sum_col, p_col + j * 8, ones, idesc_o_x, (c | jj) != 0 ? 1u : 0u);
              }
              if (KX > 0 && c == PCH - 1) {
                umma_bf16_ts(o_col, p_col + 128, make_desc(va + Cfg::KV_MAIN, 1024, 2), idesc_o_main, 1u);
                umma_bf16_ts(sum_col, p_col + 128, ones, idesc_o_x, 1u);
              }
              umma_commit<1>(&p_empty[c]);
              if (c == PCH - 1) {
                umma_commit<1>(o_full);
                if (t == NT - 1) umma_commit<1>(&v_empty[kb]);
              }
            }
            __syncwarp();
          }
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
    float* xmax = reinterpret_cast<float*>(smem + Cfg::OFF_XCHG);            // [2 tile parities][2 groups][128 rows]
    const int col_base = wg * 128;
    const int tail_key0 = KMAIN + wg * 8;                // first key of this group's 8 tail columns
    const uint32_t ob = smem_u32(smem + Cfg::OFF_OST) + warp * 2048;         // this warp's output staging chunk

    // The epilogue of tile g-1 is split into three pieces that are threaded through the exp2 stream of tile g (the
    // MUFU pipe is this kernel's bottleneck; everything else should issue while it is busy):
    //   epi_load   O(g-1) → registers (the accumulator is then free for PV(g));  epi_stage  scale by 1/rowsum, round,
    //   stage in smem;  epi_store  transposed read + coalesced global stores (8 rows x 64 contiguous bytes each)
    uint32_t o_r[32];
    uint32_t o_rx[Cfg::SPLIT_D ? 16 : 1];
    uint32_t o_sum = 0;
    auto epi_load = [&](int gp) {
      mbar_wait(o_full, static_cast<uint32_t>(gp & 1));
      tc_fence_after();
      const uint32_t o_addr = lane_addr + Cfg::O_COL;
      tmem_ld_32x32(o_addr + wg * 32, o_r);               // this group's 32 of the 64 main O columns
      if (Cfg::SPLIT_D && wg == 1) tmem_ld_32x16(o_addr + 64, reinterpret_cast<uint32_t(&)[16]>(o_rx[0]));
      tmem_ld_32x1(lane_addr + Cfg::SUM_COL, o_sum);      // Σ_k P[row, k] from the ones-MMA
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);              // O is in registers: PV(g) may overwrite the accumulator
      BLB_TRACE(gp + 1, 13);
    };
    float* tail_o = reinterpret_cast<float*>(smem + Cfg::OFF_TAIL);          // [4 quarters][8 rows][64]
    float* tail_sum = tail_o + 4 * 8 * 64;                                   // [4][8]
    float* tail_max = tail_sum + 4 * 8;                                      // [8 warps][32 lanes]
    auto epi_stage = [&](int gp) {
      if (KX > 0 && gp % NT == 2) {
        // tail tile: this warp holds, for the 8 replicated rows of its quarter, the partial O over the quarter's key
        // columns (its 32-column half) and the partial row sum → shared memory, combined in epi_store
        if (lane < 8) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(tail_o + (q * 8 + lane) * 64 + wg * 32 + j) =
                make_float4(__uint_as_float(o_r[j]), __uint_as_float(o_r[j + 1]), __uint_as_float(o_r[j + 2]),
                            __uint_as_float(o_r[j + 3]));
          if (wg == 0) tail_sum[q * 8 + lane] = __uint_as_float(o_sum);
        }
        return;
      }
      const float inv = 1.0f / __uint_as_float(o_sum);
      // the lane's 64-byte row slice → smem (16-byte piece j at j ^ ((row>>1)&3): conflict-free both ways)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(o_r[8 * j]) * inv, __uint_as_float(o_r[8 * j + 1]) * inv);
        pk.y = pack_bf16x2(__uint_as_float(o_r[8 * j + 2]) * inv, __uint_as_float(o_r[8 * j + 3]) * inv);
        pk.z = pack_bf16x2(__uint_as_float(o_r[8 * j + 4]) * inv, __uint_as_float(o_r[8 * j + 5]) * inv);
        pk.w = pack_bf16x2(__uint_as_float(o_r[8 * j + 6]) * inv, __uint_as_float(o_r[8 * j + 7]) * inv);
        sts128(ob + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pk);
      }
      if (Cfg::SPLIT_D && wg == 1) {
        const int t = gp % NT, i = gp / NT;
        const int u_i = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
        const int u = reverse ? n_units - 1 - u_i : u_i;
        const int b = u / H, h = u - b * H;
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(o_rx[0]) * inv, __uint_as_float(o_rx[1]) * inv);
        pk.y = pack_bf16x2(__uint_as_float(o_rx[2]) * inv, __uint_as_float(o_rx[3]) * inv);
        pk.z = pack_bf16x2(__uint_as_float(o_rx[4]) * inv, __uint_as_float(o_rx[5]) * inv);
        pk.w = pack_bf16x2(__uint_as_float(o_rx[6]) * inv, __uint_as_float(o_rx[7]) * inv);
        __nv_bfloat16* dst = out + (static_cast<size_t>(b) * T + t * QT + row) * D + h * HD;
        *reinterpret_cast<uint4*>(dst + 64) = pk;        // d 64..71 (72..79 are padding)
      }
      __syncwarp();
    };
    auto epi_store = [&](int gp) {
      const int t = gp % NT, i = gp / NT;
      const int u_i = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int u = reverse ? n_units - 1 - u_i : u_i;
      const int b = u / H, h = u - b * H;
      if (KX > 0 && t == 2) {
        asm volatile("bar.sync 6, 256;" ::: "memory");    // every quarter's partials are in shared memory
        if (q == 0 && lane < T - 2 * QT && lane < 8) {     // warps 0 and 4: one lane per tail row, 32 columns each
          float tot = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) tot += tail_sum[qq * 8 + lane];
          const float inv = 1.0f / tot;
          __nv_bfloat16* dst = out + (static_cast<size_t>(b) * T + 2 * QT + lane) * D + h * HD + wg * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              acc[e] = 0.f;
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) acc[e] += tail_o[(qq * 8 + lane) * 64 + wg * 32 + j + e];
            }
            uint4 pk;
            pk.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
            pk.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
            pk.z = pack_bf16x2(acc[4] * inv, acc[5] * inv);
            pk.w = pack_bf16x2(acc[6] * inv, acc[7] * inv);
            *reinterpret_cast<uint4*>(dst + j) = pk;
          }
        }
        return;
      }
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
      __syncwarp();
      BLB_TRACE(gp + 1, 14);
    };

    // ---- tail tile (KX > 0): 8 replicated query rows per lane quarter, 272 score columns shared by ALL eight warps ----
    auto tail_tile = [&](int g) {
      const int wid = wg * 4 + q;                          // which eighth of the keys
      const int mcol = wg * 128 + q * 32;                  // 32 main score columns of this warp
      const int tkey0 = KMAIN + wid * 2;                   // + 2 of the 16 tail columns
      uint32_t tv[32], tx[2];
      tmem_ld_32x32(lane_addr + mcol, tv);
      tmem_ld_32x2(lane_addr + tkey0, tx);
      tmem_ld_wait();
      reg_fence16(&tv[0]);
      reg_fence16(&tv[16]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        m0 = fmax3(m0, __uint_as_float(tv[j]), __uint_as_float(tv[j + 1]));
        m1 = fmax3(m1, __uint_as_float(tv[j + 2]), __uint_as_float(tv[j + 3]));
      }
      if (tkey0 < T) m0 = fmaxf(m0, __uint_as_float(tx[0]));
      if (tkey0 + 1 < T) m1 = fmaxf(m1, __uint_as_float(tx[1]));
      tail_max[wid * 32 + lane] = fmaxf(m0, m1);
      asm volatile("bar.sync 5, 256;" ::: "memory");      // all eight warps hold columns of the same rows
      float m = -INFINITY;
#pragma unroll
      for (int w = 0; w < 8; ++w) m = fmaxf(m, tail_max[w * 32 + lane]);
      const float ms = m * scale_log2;
      if (g > 0) {
#pragma unroll
        for (int c = 0; c < PCH; ++c) mbar_wait(&p_empty[c], static_cast<uint32_t>((g - 1) & 1));
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2)
        pk[j / 2] = pack_bf16x2(ex2_approx(fmaf(__uint_as_float(tv[j]), scale_log2, -ms)),
                                ex2_approx(fmaf(__uint_as_float(tv[j + 1]), scale_log2, -ms)));
      const float p0 = tkey0 < T ? ex2_approx(fmaf(__uint_as_float(tx[0]), scale_log2, -ms)) : 0.f;
      const float p1 = tkey0 + 1 < T ? ex2_approx(fmaf(__uint_as_float(tx[1]), scale_log2, -ms)) : 0.f;
      // P of this quarter's lanes: own 16 packed columns, zeros in the other 48 of this group's half (the partner warp
      // of the quarter fills the other half), own tail column, zeros in the group's other three
      const uint32_t zero16[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        if (blk == q) tmem_st_32x16(lane_addr + Cfg::P_COL + wg * 64 + blk * 16, pk);
        else tmem_st_32x16(lane_addr + Cfg::P_COL + wg * 64 + blk * 16, zero16);
      }
      uint32_t pt[4];
      const uint32_t ptail = pack_bf16x2(p0, p1);
#pragma unroll
      for (int j = 0; j < 4; ++j) pt[j] = j == q ? ptail : 0u;
      tmem_st_32x4(lane_addr + Cfg::P_COL + 128 + wg * 4, pt);
      if (g > 0) {                                         // the whole epilogue of tile g-1 (a regular tile)
        epi_load(g - 1);
        epi_stage(g - 1);
        epi_store(g - 1);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < PCH; ++c) mbar_arrive(&p_full[c]);
      }
    };

    for (int g = 0; g < G; ++g) {
      mbar_wait(s_full, static_cast<uint32_t>(g & 1));
      tc_fence_after();
      BLB_TRACE(g, 8);
      if (KX > 0 && g % NT == 2) {
        tail_tile(g);
        continue;
      }
      // ---- this thread's slice of the S row → registers in one pass, then release S ----
      uint32_t sv[Cfg::NREG_S];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        tmem_ld_32x32(lane_addr + col_base + c * 32, reinterpret_cast<uint32_t(&)[32]>(sv[c * 32]));
      if (KX > 0) tmem_ld_32x8(lane_addr + tail_key0, reinterpret_cast<uint32_t(&)[8]>(sv[128]));
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < Cfg::NREG_S / 8; ++c) reg_fence8(&sv[c * 8]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      BLB_TRACE(g, 9);
      // ---- row max (3-input max, four independent chains) ----
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 128; j += 8) {
        m0 = fmax3(m0, __uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
        m1 = fmax3(m1, __uint_as_float(sv[j + 2]), __uint_as_float(sv[j + 3]));
        m2 = fmax3(m2, __uint_as_float(sv[j + 4]), __uint_as_float(sv[j + 5]));
        m3 = fmax3(m3, __uint_as_float(sv[j + 6]), __uint_as_float(sv[j + 7]));
      }
      if (KX > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (tail_key0 + j < T) m0 = fmaxf(m0, __uint_as_float(sv[128 + j]));
      }
      float m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      xmax[(g & 1) * 2 * QT + wg * QT + row] = m;
      pair_barrier(q);              // the two warps of this row quarter exchange their partial maxima
      m = fmaxf(m, xmax[(g & 1) * 2 * QT + (wg ^ 1) * QT + row]);
      const float ms = m * scale_log2;
      BLB_TRACE(g, 10);
      // ---- p = 2^(s*c - m*c) → bf16 P → TMEM in PCH chunks (CK keys = CK/2 packed columns each); the row sum is an
      // MMA.  Chunk c is published (p_full[c]) one chunk late, after its tcgen05.st had a whole chunk of exp2s to
      // complete; the epilogue of tile g-1 is threaded through the chunks (group 0 one chunk ahead of group 1 where
      // there is room) so that the two warps of a sub-partition are not in their non-MUFU pieces at the same time.
      constexpr int CK = 128 / PCH;
#pragma unroll
      for (int c = 0; c < PCH; ++c) {
        if (g > 0) mbar_wait(&p_empty[c], static_cast<uint32_t>((g - 1) & 1));   // PV(g-1) consumed these P columns
        uint32_t pk[CK / 2];
#pragma unroll
        for (int j = 0; j < CK; j += 4) {
          const float t0 = fmaf(__uint_as_float(sv[c * CK + j]), scale_log2, -ms);
          const float t1 = fmaf(__uint_as_float(sv[c * CK + j + 1]), scale_log2, -ms);
          const float t2 = fmaf(__uint_as_float(sv[c * CK + j + 2]), scale_log2, -ms);
          const float t3 = fmaf(__uint_as_float(sv[c * CK + j + 3]), scale_log2, -ms);
          const float p0 = (BLB_ATTN_POLY_MASK & 1) ? exp2_poly(t0) : ex2_approx(t0);
          const float p1 = (BLB_ATTN_POLY_MASK & 2) ? exp2_poly(t1) : ex2_approx(t1);
          const float p2 = (BLB_ATTN_POLY_MASK & 4) ? exp2_poly(t2) : ex2_approx(t2);
          const float p3 = (BLB_ATTN_POLY_MASK & 8) ? exp2_poly(t3) : ex2_approx(t3);
          pk[j / 2] = pack_bf16x2(p0, p1);
          pk[j / 2 + 1] = pack_bf16x2(p2, p3);
        }
        if (c > 0) {                  // publish the previous chunk: its store was issued a whole chunk ago
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[c - 1]);
        }
        if constexpr (CK == 32) tmem_st_32x16(lane_addr + Cfg::P_COL + (col_base + c * CK) / 2, pk);
        else tmem_st_32x32(lane_addr + Cfg::P_COL + (col_base + c * CK) / 2, pk);
        BLB_TRACE(g, 4 + c);
        if (KX > 0 && c == PCH - 1) {
          uint32_t pt[4];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const float p0 = (tail_key0 + j < T) ? ex2_approx(fmaf(__uint_as_float(sv[128 + j]), scale_log2, -ms)) : 0.f;
            const float p1 =
                (tail_key0 + j + 1 < T) ? ex2_approx(fmaf(__uint_as_float(sv[128 + j + 1]), scale_log2, -ms)) : 0.f;
            pt[j / 2] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x4(lane_addr + Cfg::P_COL + tail_key0 / 2, pt);
        }
        if (g > 0) {
          if (c == 0) epi_load(g - 1);
          if (PCH >= 4) {
            if (c == wg) epi_stage(g - 1);
            if (c == wg + 1) epi_store(g - 1);
          } else {
            if (c == 0 && wg == 0) epi_stage(g - 1);
            if (c == PCH - 1 && wg == 1) epi_stage(g - 1);
            if (c == PCH - 1) epi_store(g - 1);
          }
        }
      }
      tmem_st_wait();               // the last chunk of P is in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[PCH - 1]);
      BLB_TRACE(g, 12);
    }
    if (G > 0) {
      epi_load(G - 1);
      epi_stage(G - 1);
      epi_store(G - 1);
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

template <int HD, int KX, bool TRACE>
int launch_tc_impl(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, cudaStream_t stream, int reverse) {
  using Cfg = AttnCfg<HD, KX>;
  AttnMaps maps;
  int rc = make_qkv_map(&maps.q_main, qkv, B, T, H, HD, 64, QT, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == 0) rc = make_qkv_map(&maps.kv_main, qkv, B, T, H, HD, 64, KMAIN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == 0) rc = make_qkv_map(&maps.kv_tail, qkv, B, T, H, HD, 64, 16, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc == 0) rc = make_qkv_map(&maps.q_x, qkv, B, T, H, HD, 16, QT, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc == 0) rc = make_qkv_map(&maps.kv_x, qkv, B, T, H, HD, 16, KMAIN, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc == 0) rc = make_qkv_map(&maps.q_tail, qkv, B, T, H, HD, 64, 8, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != 0) return rc;
  auto kern = attention_tc_kernel<HD, KX, TRACE>;
  static std::atomic<bool> configured[BLB_MAX_DEVICES];   // the attribute is per device
  if (!configured[current_device()]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[current_device()] = true;
  }
  const int grid = std::min(num_sms(), B * H);
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  TimingScope ts(TIME_ATTENTION, 4.0 * B * H * static_cast<double>(T) * T * HD, stream);
  cudaError_t le = launch_pdl(kern, dim3(grid), dim3(TC_THREADS), Cfg::SMEM_BYTES, stream, maps, out, B, T, H,
                              scale_log2, g_attn_trace, reverse);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

template <int HD, int KX>
int launch_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, cudaStream_t stream, int reverse) {
  // the pipeline-trace instrumentation is a separate instantiation: the production kernel carries none of it
  if (g_attn_trace != nullptr) return launch_tc_impl<HD, KX, true>(qkv, out, B, T, H, stream, reverse);
  return launch_tc_impl<HD, KX, false>(qkv, out, B, T, H, stream, reverse);
}

}  // namespace

void attention_set_trace(long long* device_buffer) { g_attn_trace = device_buffer; }

// Handles every query row of the two tower configurations this kernel is built for — T = 256 (hd 64 / 72) and
// 256 < T <= 264 (hd 64: the 261 tokens of DINOv2-reg4, tail tile of up to 8 rows) — and returns BLB_ERR_SHAPE for
// anything else (the caller then uses the mma.sync kernels).
int attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd, cudaStream_t stream,
                 int reverse) {
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) != 0) return BLB_ERR_ALIGN;
  if (hd == 64 && T == 256) return launch_tc<64, 0>(qkv, out, B, T, H, stream, reverse);
  if (hd == 64 && T > 256 && T <= 264) return launch_tc<64, 16>(qkv, out, B, T, H, stream, reverse);
  if (hd == 72 && T == 256) return launch_tc<72, 0>(qkv, out, B, T, H, stream, reverse);
  return BLB_ERR_SHAPE;
}

}  // namespace blb
