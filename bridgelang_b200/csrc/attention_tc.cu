// attention_tc.cu — tcgen05/TMEM softmax attention for the two towers.
//
// timm Attention core: softmax(q·kᵀ·hd^-0.5)·v, no mask, 16 heads,
//   DINOv2-reg4  T = 261 tokens, head_dim 64   →  template <64, 16>  (256 keys + 16-key tail block, 5 of them real)
//   SigLIP       T = 256 tokens, head_dim 72   →  template <72, 0>   (d split 64 + 16; the 16-wide block is 32B-swizzled)
// One persistent CTA per SM walks (image, head) units; per unit it handles the first 256 query rows as two
// 128-row tiles and — DINOv2 only — the T-256 = 5 remaining query rows as a third "tail" tile against the K/V that
// are already in shared memory.  The tail tile's Q rows are REPLICATED into all four TMEM lane quarters (four 8-row
// TMA boxes), so that all softmax warps share its 261 score columns (34 each for eight warps); every warp zero-fills the P
// columns it does not own, the four quarters' partial O / row sums are added through shared memory.
//
// Two floors bound the kernel, within 20 % of each other: the MUFU pipe (one exp2 per score, 16/clk/SM) and HBM (the
// packed qkv tensor read once, the output written once: 547 / 604 MB per launch at B = 256).  The warps that feed the
// MUFU pipe do nothing else (warp ids for the default NSG = 2 softmax warpgroups):
//
//   warps 0-7    softmax: NSG groups of 4 warps split the S columns; a thread owns query row = TMEM lane.  It pulls
//                its slice of the S row into registers with ONE pass of tcgen05.ld and releases S at once (s_free) —
//                the next tile's Q·Kᵀ runs on the tensor pipe underneath this tile's softmax — reads the row max that
//                the helper warps prepared, then exp2 (+ the fp32 row sum, packed f32x2 adds) → bf16 P → tcgen05.st,
//                published chunk by chunk.
//   warp 8       TMA producer: Q tiles, K, V straight out of the packed QKV GEMM output through a 4-D tensor map
//                {d, 3·H head slots, token, image}; rows >= T and d >= head_dim are zero-filled by TMA (OOB);
//                K and V are double-buffered across units
//   warp 9       MMA issuer:   S = Q·Kᵀ  (SS: M128 × N256[+16], fp32 in TMEM columns [0, 256+KX))
//                              O = P·V   (TS: A = bf16 P in TMEM, B = V as an MN-major smem operand), issued chunk
//                              by chunk under the exp2 stream
//   warp 10      TMEM allocator   (warp 11 idles; setmaxnreg is warpgroup-granular)
//   warps 12-15  helpers (one per TMEM lane quarter), everything that is neither MUFU nor MMA:
//                  row max of S(g+2) over all key columns → shared memory (two tiles ahead of the softmax warps:
//                  S(g+1) is in TMEM long before softmax(g) ends), and the epilogue of tile g: O / rowsum → bf16 →
//                  swizzled smem staging → coalesced 128-byte row stores.
//                Round 1 / early round 2 had the softmax warps do both (max exchange through a 64-thread barrier, the
//                epilogue threaded through the exp2 stream): the softmax warps of a sub-partition run in lock step,
//                so every non-MUFU phase left the MUFU pipe idle for all of them (tile ≈ 4500 cycles against a MUFU
//                floor of 2176; ≈ 3000 now).
//   TMEM columns: S fp32 [0, 256+KX) | P bf16x2 [.., +(256+KX)/2) | O fp32 [.., +64/80)   (40 / 48 columns spare)
// DESIGN.md §4.3 has the measurements behind each of these choices and the list of what was tried and dropped.
#include <algorithm>
#include <cmath>

#include <atomic>

#include "gemm.h"
#include "ptx.cuh"

namespace blb {

namespace {

constexpr int QT = 128;        // query rows per tile (UMMA M)
constexpr int KMAIN = 256;     // keys in the main block (UMMA N of S)
#ifndef BLB_ATTN_NSG
#define BLB_ATTN_NSG 2
#endif
constexpr int NSG = BLB_ATTN_NSG;   // softmax warpgroups: each owns 256/NSG of the main key columns (NSG warps per sub-partition)
static_assert(NSG == 2 || NSG == 4, "NSG");
constexpr int CG = KMAIN / NSG;     // main S columns per softmax group
constexpr int NSW = 4 * NSG;        // softmax warps
constexpr int TC_THREADS = 128 * (NSG + 2);   // softmax warpgroups + 4 control warps + 4 helper (row max / epilogue) warps
// register budgets per thread (setmaxnreg), summed over the warpgroups they may not exceed what the CTA was launched with
constexpr int REG_SOFTMAX = NSG == 2 ? 176 : 88, REG_CONTROL = NSG == 2 ? 56 : 40, REG_HELPER = NSG == 2 ? 104 : 88;
constexpr int REG_LAUNCH = (65536 / TC_THREADS) / 8 * 8;   // what every thread owns when the kernel starts
// setmaxnreg moves registers inside the CTA's own allocation (threads x REG_LAUNCH), not the SM's file: an increase
// beyond it blocks forever
static_assert(NSG * REG_SOFTMAX + REG_CONTROL + REG_HELPER <= (NSG + 2) * REG_LAUNCH, "register budget");

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
[[maybe_unused]] __device__ __forceinline__ void tmem_st_32x2(uint32_t taddr, const uint32_t (&r)[2]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
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
  static constexpr int OFF_XCHG = OFF_V + 2 * KV_BYTES;          // scaled row maxima [2 tile parities][128 rows] (fp32), helper → softmax
  static constexpr int OFF_OST = OFF_XCHG + 2 * QT * 4;          // per-helper-warp [32 rows x 128 B] output staging chunks
  static constexpr int OFF_SUMS = OFF_OST + 4 * 4096;            // partial row sums [2 tile parities][NSG groups][128 rows] (fp32), softmax → helper
  // tail tile (KX > 0): partial O [4 quarters][8 rows][64] + partial sums [4][8] (fp32)
  static constexpr int OFF_TAIL = OFF_SUMS + 2 * NSG * QT * 4;
  static constexpr int TAIL_BYTES = KX > 0 ? (4 * 8 * 64 + 4 * 8) * 4 : 0;
  static constexpr int OFF_BAR = OFF_TAIL + TAIL_BYTES;
  static constexpr int NT = KX > 0 ? 3 : 2;                      // tiles per (image, head) unit
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
  // TMEM columns: S fp32 [0, 256+KX) | P bf16x2 [P_COL, P_COL + (256+KX)/2) | O fp32 [O_COL, O_COL + HDP)
  static constexpr int S_COLS = KMAIN + KX;
  static constexpr int P_COL = S_COLS;
  static constexpr int P_COLS = S_COLS / 2;
  static constexpr int O_COL = P_COL + P_COLS;
  static constexpr int NREG_S = CG + KX / 2;                     // S columns one softmax thread keeps in registers (the last group: + keys 256..263)
  static_assert(Q_BYTES % 1024 == 0 && KV_BYTES % 1024 == 0 && KV_MAIN % 1024 == 0, "1 KB aligned blocks");
  static_assert(O_COL + HDP <= 512, "TMEM budget");
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
template <int N, int FROM>
__device__ __forceinline__ void setmaxnreg_to() {
  if constexpr (N > FROM) setmaxnreg_inc<N>();
  else if constexpr (N < FROM) setmaxnreg_dec<N>();
}

// Compiler-level fence for registers filled by an asynchronous tcgen05.ld: placed right after tcgen05.wait::ld, it makes
// every later use of the values depend on a statement that cannot move above the wait (emits no instruction).
__device__ __forceinline__ void reg_fence16(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                    "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
__device__ __forceinline__ void reg_fence8(uint32_t* r) {
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}

// packed f32x2 arithmetic (sm_100a): one FMA-pipe instruction per two scores
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void ffma2(float& x0, float& x1, uint64_t b, uint64_t c) {   // (x0, x1) = (x0, x1)·b + c
  asm("{\n\t.reg .b64 a;\n\tmov.b64 a, {%0, %1};\n\tfma.rn.f32x2 a, a, %2, %3;\n\tmov.b64 {%0, %1}, a;\n\t}"
      : "+f"(x0), "+f"(x1)
      : "l"(b), "l"(c));
}
__device__ __forceinline__ void fadd2(uint64_t& acc, float p0, float p1) {               // acc += (p0, p1)
  asm("{\n\t.reg .b64 a;\n\tmov.b64 a, {%1, %2};\n\tadd.rn.f32x2 %0, %0, a;\n\t}" : "+l"(acc) : "f"(p0), "f"(p1));
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float hsum_f32x2(uint64_t v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

// 3-input max (FMNMX3 on sm_100a): halves the instruction count of the row-max pass
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
#ifndef BLB_ATTN_PCH
#define BLB_ATTN_PCH 2
#endif
constexpr int PCH = BLB_ATTN_PCH;   // P / PV chunks per tile: 128/PCH keys of each group per chunk (4 or 2)
static_assert(PCH == 4 || PCH == 2, "PCH");

#ifndef BLB_ATTN_SLEEP_NS
#define BLB_ATTN_SLEEP_NS 32   // poll interval of the helper / TMA warps' barrier waits (0: park in try_wait like the others)
#endif
__device__ __forceinline__ void slack_wait(uint64_t* bar, uint32_t parity) {
  if constexpr (BLB_ATTN_SLEEP_NS > 0) mbar_wait_sleep<BLB_ATTN_SLEEP_NS>(bar, parity);
  else mbar_wait(bar, parity);
}

#ifndef BLB_ATTN_DIAG
#define BLB_ATTN_DIAG 0      // diagnostic builds only: bit 0 drops the exp2s, bit 1 all but one PV MMA per chunk
#endif
#ifndef BLB_ATTN_STAGGER
#define BLB_ATTN_STAGGER 0
#endif
__device__ __forceinline__ void tmem_ld_32x2(uint32_t taddr, uint32_t (&r)[2]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void tmem_ld_32x1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
}

template <int HD, int KX, bool TRACE>
__global__ void __launch_bounds__(TC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ AttnMaps maps, __nv_bfloat16* __restrict__ out, int B, int T, int H,
                    float scale_log2, long long* trace, int reverse) {
  // trace (debug only, TRACE instantiation): every warp of CTA 0 records clock64() at pipeline events of tiles [8, 16):
  // trace[((g-8)*32 + warp)*16 + e]; MMA warp e: 0 s_free(g) seen, 1 S(g+1) issued, 2 o_empty seen, 3 PV(g) issued,
  //   4+c p_full[c](g) seen, 8 k_full seen, 9 q_full seen;  softmax warps e: 8 s_full seen, 9 S in registers (s_free),
  //   10 row max read, 4+c P chunk c stored, 12 last P chunk published;  helper warps e: 11 row max of tile g published,
  //   13 O(g) in registers (o_empty), 14 epilogue(g) stores issued
#define BLB_TRACE(g_, e_)                                                                   \
  do {                                                                                      \
    if (TRACE && trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (g_) >= 8 && (g_) < 16)   \
      trace[(((g_) - 8) * 32 + (threadIdx.x >> 5)) * 16 + (e_)] = clock64();                                  \
  } while (0)
  using Cfg = AttnCfg<HD, KX>;
  extern __shared__ uint8_t smem_raw_attn[];
  // Values every role needs — the 1 KB-aligned base of the dynamic shared memory, the TMEM base, the CTA's unit count —
  // live in static shared memory and are READ BACK INSIDE EVERY ROLE BRANCH (volatile loads, so nothing is hoisted):
  // a value that is live across the setmaxnreg boundaries is allocated against the smallest budget (the control
  // warps'), gets spilled, and the hot loops then wait for an LDL (an L1-miss round trip: the 227 KB shared-memory
  // carve-out leaves almost no L1) — ncu attributed ≈ 215 cycles per tile to one such reload.
  __shared__ uint32_t s_params[4];   // [0] dynamic-smem base (shared-window address), [1] TMEM base, [2] my_units
  constexpr int NT = Cfg::NT;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  constexpr int W_TMA = NSW, W_MMA = NSW + 1, W_ALLOC = NSW + 2, W_HELP = NSW + 4;
  const int n_units = B * H;

  // everything a role derives from the shared-memory base (expanded inside each role branch)
#define BLB_ATTN_ROLE_PROLOGUE                                                                                         \
  uint8_t* smem;                                                                                                       \
  uint32_t tmem_base;                                                                                                  \
  int my_units;                                                                                                        \
  {                                                                                                                    \
    uint32_t sb_, mu_;                                                                                                 \
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(sb_) : "r"(smem_u32(&s_params[0])));                                 \
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(smem_u32(&s_params[1])));                           \
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(mu_) : "r"(smem_u32(&s_params[2])));                                 \
    smem = reinterpret_cast<uint8_t*>(__cvta_shared_to_generic(sb_));                                                  \
    my_units = static_cast<int>(mu_);                                                                                  \
  }                                                                                                                    \
  const int G = NT * my_units;     /* tiles this CTA processes */                                                      \
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);                                             \
  uint64_t* const q_full = bars;         /* [2 tiles of a unit] */                                                     \
  uint64_t* const q_empty = bars + 2;    /* [2] */                                                                     \
  uint64_t* const k_full = bars + 4;     /* [2 unit parities] */                                                       \
  uint64_t* const k_empty = bars + 6;    /* [2] */                                                                     \
  uint64_t* const v_full = bars + 8;     /* [2] */                                                                     \
  uint64_t* const v_empty = bars + 10;   /* [2] */                                                                     \
  uint64_t* const s_full = bars + 12;    /* MMA → softmax + helpers: S(g) is in TMEM */                                \
  uint64_t* const s_free = bars + 13;    /* softmax + helpers → MMA: every reader of S(g) is done with its columns */  \
  uint64_t* const o_full = bars + 14;    /* MMA → helpers: every PV chunk of tile g has retired */                     \
  uint64_t* const o_empty = bars + 15;   /* helpers → MMA: O(g) is in registers */                                     \
  uint64_t* const p_full = bars + 16;    /* [PCH] softmax → MMA: P chunk c of tile g is in TMEM */                     \
  uint64_t* const p_empty = bars + 16 + PCH;       /* [PCH] MMA → softmax: PV chunk c consumed its P columns */        \
  uint64_t* const max_full = bars + 16 + 2 * PCH;  /* [2 tile parities] helpers → softmax: scaled row maxima in smem */\
  uint64_t* const sum_full = bars + 18 + 2 * PCH;  /* [2 tile parities] softmax → helpers: partial row sums in smem */ \
  /* buffer addresses by arithmetic (a dynamically indexed pointer array would live in local memory) */               \
  auto sQ = [&](int i) { return smem + Cfg::OFF_Q0 + i * Cfg::Q_BYTES; };                                              \
  auto sK = [&](int i) { return smem + Cfg::OFF_K + i * Cfg::KV_BYTES; };                                              \
  auto sV = [&](int i) { return smem + Cfg::OFF_V + i * Cfg::KV_BYTES; };                                              \
  const uint32_t xmax = smem_u32(smem + Cfg::OFF_XCHG);   /* float [2 tile parities][128 rows]: max·scale·log2e */     \
  const uint32_t psum = smem_u32(smem + Cfg::OFF_SUMS);   /* float [2 parities][NSG groups][128 rows]: Σ_k p */        \
  (void)G; (void)q_full; (void)q_empty; (void)k_full; (void)k_empty; (void)v_full; (void)v_empty; (void)s_full;        \
  (void)s_free; (void)o_full; (void)o_empty; (void)p_full; (void)p_empty; (void)max_full; (void)sum_full; (void)sQ;    \
  (void)sK; (void)sV; (void)xmax; (void)psum

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&maps.q_main);
    tma_prefetch_desc(&maps.kv_main);
    if (KX > 0) { tma_prefetch_desc(&maps.kv_tail); tma_prefetch_desc(&maps.q_tail); }
    if (Cfg::SPLIT_D) { tma_prefetch_desc(&maps.q_x); tma_prefetch_desc(&maps.kv_x); }
  }
  if (warp == W_MMA && lane == 0) {
    uint8_t* smem0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_attn) + 1023) & ~uintptr_t(1023));
    s_params[0] = smem_u32(smem0);
    s_params[2] = static_cast<uint32_t>((n_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                        static_cast<int>(gridDim.x));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem0 + Cfg::OFF_BAR);
    for (int i = 0; i < 12; ++i) mbar_init(&bars[i], 1);                 // q/k/v full + empty
    mbar_init(&bars[12], 1); mbar_init(&bars[13], NSW + 4);              // s_full, s_free
    mbar_init(&bars[14], 1); mbar_init(&bars[15], 4);                    // o_full, o_empty
    for (int c = 0; c < PCH; ++c) { mbar_init(&bars[16 + c], NSW); mbar_init(&bars[16 + PCH + c], 1); }   // p_full, p_empty
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[16 + 2 * PCH + i], 4); mbar_init(&bars[18 + 2 * PCH + i], NSW); }   // max_full, sum_full
    fence_mbar_init();
  }
  if (warp == W_ALLOC) tmem_alloc<1>(&s_params[1], 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();   // PDL: the prologue above overlapped with the QKV GEMM's tail
  pdl_wait();

  auto lds_f32 = [](uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; };
  auto sts_f32 = [](uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); };

  if (warp >= W_HELP) {
    // ============================ helpers: row max (two tiles ahead) + epilogue =========================
    setmaxnreg_to<REG_HELPER, REG_LAUNCH>();
    BLB_ATTN_ROLE_PROLOGUE;
    const int q = warp & 3;                              // TMEM lane quarter of this warp
    const int row = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int D = H * HD;
    const uint32_t ob = smem_u32(smem + Cfg::OFF_OST) + q * 4096;            // this warp's output staging chunk
    float* tail_o = reinterpret_cast<float*>(smem + Cfg::OFF_TAIL);          // [4 quarters][8 rows][64]
    float* tail_sum = tail_o + 4 * 8 * 64;                                   // [4][8]

    // row max of S(g) over every real key column, scaled, → xmax[g & 1]; the same code serves the tail tile (its
    // rows are replicated per quarter; lanes 8-31 hold stale rows whose maxima nobody reads)
    auto row_max = [&](int g) {
      slack_wait(s_full, static_cast<uint32_t>(g & 1));
      tc_fence_after();
      uint32_t ta[16], tb[16], tx[KX > 0 ? 8 : 1];
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
      auto fold = [&](const uint32_t (&t)[16]) {
#pragma unroll
        for (int j = 0; j < 16; j += 8) {
          m0 = fmax3(m0, __uint_as_float(t[j]), __uint_as_float(t[j + 1]));
          m1 = fmax3(m1, __uint_as_float(t[j + 2]), __uint_as_float(t[j + 3]));
          m2 = fmax3(m2, __uint_as_float(t[j + 4]), __uint_as_float(t[j + 5]));
          m3 = fmax3(m3, __uint_as_float(t[j + 6]), __uint_as_float(t[j + 7]));
        }
      };
      tmem_ld_32x16(lane_addr, ta);
      if (KX > 0) tmem_ld_32x8(lane_addr + KMAIN, reinterpret_cast<uint32_t(&)[8]>(tx[0]));
      tmem_ld_wait();
      reg_fence16(ta);
      if (KX > 0) reg_fence8(&tx[0]);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {                    // 16 chunks of 16 columns, one load in flight under each fold
        tmem_ld_32x16(lane_addr + (2 * kk + 1) * 16, tb);
        fold(ta);
        tmem_ld_wait();
        reg_fence16(tb);
        if (kk < 7) tmem_ld_32x16(lane_addr + (2 * kk + 2) * 16, ta);
        fold(tb);
        if (kk < 7) {
          tmem_ld_wait();
          reg_fence16(ta);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      if (KX > 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (KMAIN + j < T) m0 = fmaxf(m0, __uint_as_float(tx[j]));
      }
      sts_f32(xmax + ((g & 1) * QT + row) * 4, fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * scale_log2);
      __syncwarp();
      if (lane == 0) mbar_arrive(&max_full[g & 1]);       // release: the stores above are visible to the waiters
      BLB_TRACE(g, 11);
    };

    auto epilogue = [&](int g) {
      const int t = g % NT, i = g / NT;
      const int u_i = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
      const int u = reverse ? n_units - 1 - u_i : u_i;
      const int b = u / H, h = u - b * H;
      slack_wait(o_full, static_cast<uint32_t>(g & 1));
      tc_fence_after();
      const uint32_t o_addr = lane_addr + Cfg::O_COL;
      // O in 16-column pieces, two loads in flight (the helpers' register budget is small: the softmax warps need it)
      uint32_t oa[16], ob2[16];
      tmem_ld_32x16(o_addr, oa);
      tmem_ld_32x16(o_addr + 16, ob2);
      // Σ_k p[row, k]: the softmax groups' partial sums (fp32, of the unrounded p)
      slack_wait(&sum_full[g & 1], static_cast<uint32_t>((g >> 1) & 1));
      float o_sum = 0.f;
#pragma unroll
      for (int sg = 0; sg < NSG; ++sg) o_sum += lds_f32(psum + (((g & 1) * NSG + sg) * QT + row) * 4);
      tmem_ld_wait();
      reg_fence16(oa); reg_fence16(ob2);
      if (KX > 0 && t == 2) {
        // tail tile: this quarter's lanes 0-7 hold, for the 8 replicated rows, the partial O / partial row sum over
        // the key columns that the softmax warps of quarter q own → shared memory, combined below by all four helpers
        auto put16 = [&](const uint32_t (&v)[16], int c0) {
          if (lane < 8) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(tail_o + (q * 8 + lane) * 64 + c0 + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          }
        };
        put16(oa, 0);
        put16(ob2, 16);
        if (lane < 8) tail_sum[q * 8 + lane] = o_sum;
        tmem_ld_32x16(o_addr + 32, oa);
        tmem_ld_32x16(o_addr + 48, ob2);
        tmem_ld_wait();
        reg_fence16(oa); reg_fence16(ob2);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty);              // O is in registers: PV(g+1) may overwrite the accumulator
        BLB_TRACE(g, 13);
        put16(oa, 32);
        put16(ob2, 48);
        asm volatile("bar.sync 1, 128;" ::: "memory");    // the four quarters' partials are in shared memory
        const int tid = q * 32 + lane, r = tid >> 4, c4 = (tid & 15) * 4;   // 8 rows x 16 groups of 4 columns
        if (r < T - 2 * QT) {
          float tot = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            tot += tail_sum[qq * 8 + r];
            const float4 v = *reinterpret_cast<const float4*>(tail_o + (qq * 8 + r) * 64 + c4);
            acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
          }
          const float inv = 1.0f / tot;
          uint2 pk;
          pk.x = pack_bf16x2(acc[0] * inv, acc[1] * inv);
          pk.y = pack_bf16x2(acc[2] * inv, acc[3] * inv);
          *reinterpret_cast<uint2*>(out + (static_cast<size_t>(b) * T + 2 * QT + r) * D + h * HD + c4) = pk;
        }
        BLB_TRACE(g, 14);
        return;
      }
      const float inv = 1.0f / o_sum;
      // the lane's 128-byte row → smem staging (16-byte piece j at j ^ (row & 7): conflict-free both ways)
      auto stage16 = [&](const uint32_t (&v)[16], int piece0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          uint4 pk;
          pk.x = pack_bf16x2(__uint_as_float(v[8 * j]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
          pk.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
          pk.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
          pk.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
          sts128(ob + lane * 128 + (((piece0 + j) ^ (lane & 7)) << 4), pk);
        }
      };
      stage16(oa, 0);
      stage16(ob2, 2);
      uint32_t ox[Cfg::SPLIT_D ? 8 : 1];
      tmem_ld_32x16(o_addr + 32, oa);
      tmem_ld_32x16(o_addr + 48, ob2);
      if (Cfg::SPLIT_D) tmem_ld_32x8(o_addr + 64, reinterpret_cast<uint32_t(&)[8]>(ox[0]));   // d 64..71 (72..79: padding)
      tmem_ld_wait();
      reg_fence16(oa); reg_fence16(ob2);
      if (Cfg::SPLIT_D) reg_fence8(&ox[0]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty);                // O is in registers: PV(g+1) may overwrite the accumulator
      BLB_TRACE(g, 13);
      stage16(oa, 4);
      stage16(ob2, 6);
      __nv_bfloat16* slab = out + (static_cast<size_t>(b) * T + t * QT + q * 32) * D + h * HD;
      if (Cfg::SPLIT_D) {
        uint4 pk;
        pk.x = pack_bf16x2(__uint_as_float(ox[0]) * inv, __uint_as_float(ox[1]) * inv);
        pk.y = pack_bf16x2(__uint_as_float(ox[2]) * inv, __uint_as_float(ox[3]) * inv);
        pk.z = pack_bf16x2(__uint_as_float(ox[4]) * inv, __uint_as_float(ox[5]) * inv);
        pk.w = pack_bf16x2(__uint_as_float(ox[6]) * inv, __uint_as_float(ox[7]) * inv);
        stg128(slab + static_cast<size_t>(lane) * D + 64, pk);
      }
      __syncwarp();
      // transposed read: one instruction stores 4 rows x 128 contiguous bytes
      uint4 tv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = k * 4 + (lane >> 3), j = lane & 7;
        tv[k] = lds128(ob + rr * 128 + ((j ^ (rr & 7)) << 4));
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int rr = k * 4 + (lane >> 3), j = lane & 7;
        stg128(slab + static_cast<size_t>(rr) * D + j * 8, tv[k]);
      }
      __syncwarp();
      BLB_TRACE(g, 14);
    };

    if (G > 0) row_max(0);
    if (G > 1) row_max(1);
#pragma unroll 1
    for (int g = 0; g < G; ++g) {
      epilogue(g);
      if (g + 2 < G) row_max(g + 2);
    }
  } else if (warp >= NSW) {
    setmaxnreg_to<REG_CONTROL, REG_LAUNCH>();          // hand registers to the softmax warps
    if (warp == W_TMA) {
      // ============================== TMA producer (warp-uniform loop, one elected lane issues) ===========
      {
        BLB_ATTN_ROLE_PROLOGUE;
        (void)tmem_base;
        for (int i = 0; i < my_units; ++i) {
          const int u_i = static_cast<int>(blockIdx.x) + i * static_cast<int>(gridDim.x);
          const int u = reverse ? n_units - 1 - u_i : u_i;
          const int b = u / H, h = u - b * H;
          const int kb = i & 1;                                          // K/V buffer of this unit
          const uint32_t kph = static_cast<uint32_t>((i >> 1) & 1);      // K/V barriers: one use per two units
          slack_wait(&k_empty[kb], kph ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(&k_full[kb], Cfg::KV_BYTES);
            tma_load_4d(sK(kb), &maps.kv_main, &k_full[kb], 0, H + h, 0, b);
            if (KX > 0) tma_load_4d(sK(kb) + Cfg::KV_MAIN, &maps.kv_tail, &k_full[kb], 0, H + h, KMAIN, b);
            if (Cfg::SPLIT_D)
              tma_load_4d(sK(kb) + Cfg::KV_MAIN + Cfg::KV_TAIL, &maps.kv_x, &k_full[kb], 64, H + h, 0, b);
          }
          __syncwarp();
          for (int t = 0; t < NT; ++t) {
            const int g = i * NT + t, qb = g & 1;                        // Q buffers alternate per tile
            slack_wait(&q_empty[qb], static_cast<uint32_t>((g >> 1) & 1) ^ 1u);
            if (elect_one()) {
              if (KX > 0 && t == 2) {
                // tail tile: query rows 256..263 (rows >= T are zero-filled) into rows 0..7 of every 32-row quarter
                mbar_expect_tx(&q_full[qb], 4 * 1024);
#pragma unroll
                for (int r = 0; r < 4; ++r) tma_load_4d(sQ(qb) + r * 4096, &maps.q_tail, &q_full[qb], 0, h, 2 * QT, b);
              } else {
                mbar_expect_tx(&q_full[qb], Cfg::Q_BYTES);
                tma_load_4d(sQ(qb), &maps.q_main, &q_full[qb], 0, h, t * QT, b);
                if (Cfg::SPLIT_D) tma_load_4d(sQ(qb) + Cfg::Q_MAIN, &maps.q_x, &q_full[qb], 64, h, t * QT, b);
              }
            }
            __syncwarp();
            if (t == 0) {
              slack_wait(&v_empty[kb], kph ^ 1u);
              if (elect_one()) {
                mbar_expect_tx(&v_full[kb], Cfg::KV_BYTES);
                tma_load_4d(sV(kb), &maps.kv_main, &v_full[kb], 0, 2 * H + h, 0, b);
                if (KX > 0) tma_load_4d(sV(kb) + Cfg::KV_MAIN, &maps.kv_tail, &v_full[kb], 0, 2 * H + h, KMAIN, b);
                if (Cfg::SPLIT_D)
                  tma_load_4d(sV(kb) + Cfg::KV_MAIN + Cfg::KV_TAIL, &maps.kv_x, &v_full[kb], 64, 2 * H + h, 0, b);
              }
              __syncwarp();
            }
          }
        }
      }
    } else if (warp == W_MMA) {
      // =============================== MMA issuer (warp-uniform loop, one elected lane issues) ===========
      {
        BLB_ATTN_ROLE_PROLOGUE;
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
            const uint32_t qa = smem_u32(sQ(qb)), ka = smem_u32(sK(kb));
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
            // every reader of S(g) holds what it needs in registers → the next Q·Kᵀ runs underneath this tile's softmax
            mbar_wait(s_free, static_cast<uint32_t>(g & 1));
            tc_fence_after();
            BLB_TRACE(g, 0);
            issue_s(g + 1);
            BLB_TRACE(g, 1);
          }
          // ---------------- O(g) = P(g) · V   (A = P in TMEM, B = V as an MN-major smem operand) ----------------
          // issued chunk by chunk as the softmax warps publish P: chunk c = keys [CG/PCH·c, CG/PCH·(c+1)) of every
          // group's CG columns (+ the 16-key tail block with the last chunk), so that only the last chunk's MMAs are
          // left when the tile's exp2 stream ends
          const int t = g % NT, i = g / NT, kb = i & 1;
          const uint32_t ph = static_cast<uint32_t>(g & 1);
          if (t == 0) mbar_wait(&v_full[kb], static_cast<uint32_t>((i >> 1) & 1));
          mbar_wait(o_empty, ph ^ 1u);   // the helpers hold O(g-1) in registers
          BLB_TRACE(g, 2);
          const uint32_t o_col = tmem_base + Cfg::O_COL;
          const uint32_t p_col = tmem_base + Cfg::P_COL;
          // V descriptors: built once per tile; key block j is the base plus an immediate (2048 B = 128 in the
          // descriptor's 16-byte units; the extension block 512 B = 32) — the single issuing thread is itself a
          // bottleneck (≈ 35-50 MMAs per tile), so its per-MMA address arithmetic is one 64-bit add
          const uint64_t vdesc = make_desc(smem_u32(sV(kb)), 1024, 2);
          const uint64_t vxdesc = make_desc(smem_u32(sV(kb)) + Cfg::KV_MAIN + Cfg::KV_TAIL, 256, 6);
#pragma unroll
          for (int c = 0; c < PCH; ++c) {
            mbar_wait(&p_full[c], ph);
            tc_fence_after();
            BLB_TRACE(g, 4 + c);
            if (elect_one()) {
              constexpr int KBG = CG / 16;         // 16-key blocks per softmax group
              constexpr int MPG = KBG / PCH;       // 16-key MMAs per group and chunk
#pragma unroll
              for (int jj = 0; jj < ((BLB_ATTN_DIAG & 2) ? 1 : NSG * MPG); ++jj) {   // 16 keys per MMA: 8 packed P columns, V rows 16j..16j+15
                const int j = (jj / MPG) * KBG + MPG * c + (jj % MPG);
                umma_bf16_ts(o_col, p_col + j * 8, vdesc + static_cast<uint64_t>(j * 128), idesc_o_main,
                             (c | jj) != 0 ? 1u : 0u);
                if (Cfg::SPLIT_D)
                  umma_bf16_ts(o_col + 64, p_col + j * 8, vxdesc + static_cast<uint64_t>(j * 32), idesc_o_x,
                               (c | jj) != 0 ? 1u : 0u);
              }
              if (KX > 0 && c == PCH - 1)
                umma_bf16_ts(o_col, p_col + 128, vdesc + static_cast<uint64_t>(Cfg::KV_MAIN >> 4), idesc_o_main, 1u);
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
    // ============================ softmax: pull S, exp2, pack, publish P — nothing else ==================
    // 4·NSG warps: group wg owns S columns [CG·wg, CG·wg+CG) (the last group also keys 256..263 of the tail block);
    // every group sees all 128 rows (warps w, w+4, ... share TMEM lane quarter w%4), i.e. NSG softmax warps per SM
    // sub-partition: while one of them is in a MUFU-free piece (S pull, barrier waits, the last store's completion)
    // the others keep the MUFU pipe busy.
    setmaxnreg_to<REG_SOFTMAX, REG_LAUNCH>();
    BLB_ATTN_ROLE_PROLOGUE;
    const int q = warp & 3;
    const int wg = warp >> 2;
    const int row = q * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int col_base = wg * CG;
    const bool has_tail = KX > 0 && wg == NSG - 1;       // this group also owns keys 256..263 (264..271 are always padding)

    // ---- tail tile (KX > 0): 8 replicated query rows per lane quarter, 272 score columns shared by ALL softmax warps ----
    auto tail_tile = [&](int g) {
      constexpr int MC = KMAIN / NSW;                      // main score columns per warp: 32 | 16
      constexpr int TK = 16 / NSW;                         // tail-block keys per warp: 2 | 1
      const int wid = wg * 4 + q;                          // which slice of the keys
      const int mcol = wg * CG + q * MC;
      const int tkey0 = KMAIN + wid * TK;
      uint32_t tv[MC], tx[2];
      if constexpr (MC == 32) tmem_ld_32x32(lane_addr + mcol, reinterpret_cast<uint32_t(&)[32]>(tv[0]));
      else tmem_ld_32x16(lane_addr + mcol, reinterpret_cast<uint32_t(&)[16]>(tv[0]));
      if constexpr (TK == 2) tmem_ld_32x2(lane_addr + tkey0, tx);
      else { tmem_ld_32x1(lane_addr + tkey0, tx[0]); tx[1] = 0u; }
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < MC / 8; ++c) reg_fence8(&tv[c * 8]);
      asm volatile("" : "+r"(tx[0]), "+r"(tx[1]));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      mbar_wait(&max_full[g & 1], static_cast<uint32_t>((g >> 1) & 1));
      const float ms = lds_f32(xmax + ((g & 1) * QT + row) * 4);
      uint32_t pk[MC / 2];
      const float p0 = tkey0 < T ? ex2_approx(fmaf(__uint_as_float(tx[0]), scale_log2, -ms)) : 0.f;
      const float p1 = (TK == 2 && tkey0 + 1 < T) ? ex2_approx(fmaf(__uint_as_float(tx[1]), scale_log2, -ms)) : 0.f;
      float l0 = p0, l1 = p1;
#pragma unroll
      for (int j = 0; j < MC; j += 2) {
        const float e0 = ex2_approx(fmaf(__uint_as_float(tv[j]), scale_log2, -ms));
        const float e1 = ex2_approx(fmaf(__uint_as_float(tv[j + 1]), scale_log2, -ms));
        pk[j / 2] = pack_bf16x2(e0, e1);
        l0 += e0; l1 += e1;
      }
      if (g > 0) {                                         // PV(g-1) must have consumed the P columns
#pragma unroll
        for (int c = 0; c < PCH; ++c) mbar_wait(&p_empty[c], static_cast<uint32_t>((g - 1) & 1));
        tc_fence_after();
      }
      // the partial sum over this warp's key columns; the helper of quarter q adds the groups' partials, exactly
      // the keys its quarter's partial O covers.  (Written after the p_empty waits: PV(g-1) issued ⇒ the helpers have
      // read the sums of tile g-2, which used this slot.)
      sts_f32(psum + (((g & 1) * NSG + wg) * QT + row) * 4, l0 + l1);
      // P of this quarter's lanes: own MC/2 packed columns, zeros in the other three blocks of this group's CG/2 (the
      // partner warps of the quarter fill the other groups' columns), own tail key, zeros in the rest of the group's
      // share of the 8 packed tail columns
      const uint32_t zero16[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {
        const uint32_t pa = lane_addr + Cfg::P_COL + wg * (CG / 2) + blk * (MC / 2);
        if constexpr (MC == 32) {
          if (blk == q) tmem_st_32x16(pa, reinterpret_cast<const uint32_t(&)[16]>(pk[0]));
          else tmem_st_32x16(pa, zero16);
        } else {
          if (blk == q) tmem_st_32x8(pa, reinterpret_cast<const uint32_t(&)[8]>(pk[0]));
          else tmem_st_32x8(pa, reinterpret_cast<const uint32_t(&)[8]>(zero16[0]));
        }
      }
      if constexpr (TK == 2) {
        uint32_t pt[4];
        const uint32_t ptail = pack_bf16x2(p0, p1);
#pragma unroll
        for (int j = 0; j < 4; ++j) pt[j] = j == q ? ptail : 0u;
        tmem_st_32x4(lane_addr + Cfg::P_COL + 128 + wg * 4, pt);
      } else {
        // key 256 + 4·wg + q lives in packed column 2·wg + (q >> 1), low or high half by q & 1
        uint32_t pt[2];
        const uint32_t ptail = (q & 1) ? pack_bf16x2(0.f, p0) : pack_bf16x2(p0, 0.f);
        pt[0] = (q >> 1) == 0 ? ptail : 0u;
        pt[1] = (q >> 1) == 1 ? ptail : 0u;
        tmem_st_32x2(lane_addr + Cfg::P_COL + 128 + wg * 2, pt);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
#pragma unroll
        for (int c = 0; c < PCH; ++c) mbar_arrive(&p_full[c]);
        mbar_arrive(&sum_full[g & 1]);
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
      for (int c = 0; c < CG / 32; ++c)
        tmem_ld_32x32(lane_addr + col_base + c * 32, reinterpret_cast<uint32_t(&)[32]>(sv[c * 32]));
      if (has_tail) tmem_ld_32x8(lane_addr + KMAIN, reinterpret_cast<uint32_t(&)[8]>(sv[KX > 0 ? CG : 0]));
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < Cfg::NREG_S / 8; ++c) reg_fence8(&sv[c * 8]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      BLB_TRACE(g, 9);
#if BLB_ATTN_STAGGER > 0
      // One-time phase shift of the odd groups against the even ones (cycles); nothing re-synchronises the groups
      // afterwards: the shared barriers only bound the lead of one group over another (to roughly a tile).
      if (g == 0 && (wg & 1)) {
        const long long t_st = clock64();
        while (clock64() - t_st < BLB_ATTN_STAGGER) {
        }
      }
#endif
      // ---- the row max (times scale·log2e) comes from the helper warps, computed while the previous tile ran ----
      mbar_wait(&max_full[g & 1], static_cast<uint32_t>((g >> 1) & 1));
      const float ms = lds_f32(xmax + ((g & 1) * QT + row) * 4);
      BLB_TRACE(g, 10);
      // ---- p = 2^(s*c - m*c) → bf16 P → TMEM in PCH chunks (CK keys = CK/2 packed columns each).  A chunk is
      // published (p_full[c]) one 16-key sub-block into the next chunk: its tcgen05.st has completed by then.
      constexpr int CK = CG / PCH;
      static_assert(CK == 64 || CK == 32 || CK == 16, "chunk size");
      float l0 = 0.f, l1 = 0.f;                         // this thread's partial row sum: tail keys ...
      uint64_t l01 = 0ull, l23 = 0ull;                   // ... and two packed f32x2 chains over the main columns
      const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nms2 = pack_f32x2(-ms, -ms);
      uint32_t pt[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // packed tail columns (keys 256..271), last group only
      if (has_tail) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float p0 = (KMAIN + j < T) ? ex2_approx(fmaf(__uint_as_float(sv[(KX > 0 ? CG : 0) + j]), scale_log2, -ms)) : 0.f;
          const float p1 =
              (KMAIN + j + 1 < T) ? ex2_approx(fmaf(__uint_as_float(sv[(KX > 0 ? CG : 0) + j + 1]), scale_log2, -ms)) : 0.f;
          pt[j / 2] = pack_bf16x2(p0, p1);
          l0 += p0; l1 += p1;
        }
      }
#pragma unroll
      for (int c = 0; c < PCH; ++c) {
        uint32_t pk[CK / 2];
        // packed f32x2 arithmetic (sm_100: FFMA2 / FADD2): the scale-and-shift and the row sum cost one instruction per
        // TWO scores.  A warp issues in order and ptxas groups a chunk into [FFMAs][MUFUs][packs, adds] (it does not
        // keep an interleaved asm order), so every FMA-pipe instruction of the softmax warps is a cycle in which none of
        // the sub-partition's softmax warps feeds the MUFU pipe: fewer of them is the lever.
#pragma unroll
        for (int sb = 0; sb < CK / 16; ++sb) {            // sub-blocks of 16 keys
#pragma unroll
          for (int j = sb * 16; j < sb * 16 + 16; j += 4) {
            float t0 = __uint_as_float(sv[c * CK + j]), t1 = __uint_as_float(sv[c * CK + j + 1]);
            float t2 = __uint_as_float(sv[c * CK + j + 2]), t3 = __uint_as_float(sv[c * CK + j + 3]);
            ffma2(t0, t1, sc2, nms2);
            ffma2(t2, t3, sc2, nms2);
#if BLB_ATTN_DIAG & 1   // diagnostic build (wrong numerics): no MUFU in the main loop — is the tile MUFU-bound at all?
            const float p0 = t0 * 0.5f, p1 = t1 * 0.5f, p2 = t2 * 0.5f, p3 = t3 * 0.5f;
#else
            const float p0 = ex2_approx(t0), p1 = ex2_approx(t1), p2 = ex2_approx(t2), p3 = ex2_approx(t3);
#endif
            pk[j / 2] = pack_bf16x2(p0, p1);
            pk[j / 2 + 1] = pack_bf16x2(p2, p3);
            fadd2(l01, p0, p1);
            fadd2(l23, p2, p3);
          }
          if (c > 0 && sb == 0) {     // publish the previous chunk: its store was issued a sub-block of exp2s ago
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[c - 1]);
          }
        }
        if (g > 0) {                  // PV(g-1) must have consumed these P columns — checked after the exp2s, not before
          mbar_wait(&p_empty[c], static_cast<uint32_t>((g - 1) & 1));
          tc_fence_after();
        }
        const uint32_t pa = lane_addr + Cfg::P_COL + (col_base + c * CK) / 2;
        if constexpr (CK == 64) tmem_st_32x32(pa, reinterpret_cast<const uint32_t(&)[32]>(pk[0]));
        else if constexpr (CK == 32) tmem_st_32x16(pa, reinterpret_cast<const uint32_t(&)[16]>(pk[0]));
        else tmem_st_32x8(pa, reinterpret_cast<const uint32_t(&)[8]>(pk[0]));
        BLB_TRACE(g, 4 + c);
        if (has_tail && c == PCH - 1) tmem_st_32x8(lane_addr + Cfg::P_COL + KMAIN / 2, pt);
      }
      sts_f32(psum + (((g & 1) * NSG + wg) * QT + row) * 4, (l0 + l1) + hsum_f32x2(fadd2(l01, l23)));
      tmem_st_wait();               // the last chunk of P is in TMEM
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&p_full[PCH - 1]);
        mbar_arrive(&sum_full[g & 1]);            // release: the warp's partial sums are visible to the helpers
      }
      BLB_TRACE(g, 12);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == W_ALLOC) {
    tc_fence_after();
    uint32_t tb;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tb) : "r"(smem_u32(&s_params[1])));
    tmem_dealloc<1>(tb, 512);
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
