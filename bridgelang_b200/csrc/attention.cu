// attention.cu — fused softmax attention for the two ViT towers (whole K/V of one head in shared memory).
//
// Replaces timm Attention.forward's reshape/permute + F.scaled_dot_product_attention + transpose
// (SURVEY.md §2 K5-K7, §8 a8): no mask, scale = head_dim^-0.5, fp32 softmax, 16 heads,
//   DINOv2-reg4: T = 261 tokens, head_dim 64;   SigLIP: T = 256 tokens, head_dim 72.
// Reads q/k/v in place from the packed QKV GEMM output [B*T, 3*H*hd] (timm's [B,N,3,H,hd] view) and
// writes token-major [B*T, H*hd], so none of the reference's layout copies exist.
// head_dim 72 is zero-padded to 80 inside shared memory only (the fork needed
// PT_SDPA_ENABLE_HEAD_DIM_PADDING for it, run_openvla.sh:14).
//
// These are the mma.sync kernels for shapes the tcgen05 kernel (attention_tc.cu: the two 224 px tower shapes, every
// query row) does not take: attention_kernel (short sequences, whole K/V of a head in shared memory: one CTA per
// (image, head), 4 warps x 16 query rows per pass) and attention_stream_kernel (336 / 384 px towers, K/V streamed).
// Both: online softmax over 64-key blocks, bf16 mma.sync m16n8k16 with fp32 accumulation.
#include <cstdlib>

#include <atomic>

#include "gemm.h"
#include "ptx.cuh"

namespace blb {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int HD, int HDP>
__global__ void __launch_bounds__(128, 2)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int H,
                 float scale_log2, int q_begin, int reverse) {
  constexpr int PITCH = HDP + 8;        // +16 B per row keeps ldmatrix conflict-free
  constexpr int KSTEPS = HDP / 16;
  constexpr int NT_O = HDP / 8;
  constexpr int CHUNKS = HDP / 8;       // 16-byte chunks per padded row
  constexpr int CHUNKS_REAL = HD / 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int TKP = (T + 63) & ~63;
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sV = sK + TKP * PITCH;
  __nv_bfloat16* sQ = sV + TKP * PITCH;

  const int unit = reverse ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int b = unit / H, h = unit % H;
  const int D = H * HD;
  const size_t row_pitch = static_cast<size_t>(3) * D;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * row_pitch + h * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  pdl_launch_dependents();
  pdl_wait();

  // ---- stage K and V of this (image, head) once -------------------------------------------------
  for (int i = threadIdx.x; i < TKP * CHUNKS; i += blockDim.x) {
    const int r = i / CHUNKS, c = i - r * CHUNKS;
    __nv_bfloat16* dk = sK + r * PITCH + c * 8;
    __nv_bfloat16* dv = sV + r * PITCH + c * 8;
    if (r < T && c < CHUNKS_REAL) {
      const __nv_bfloat16* g = base + static_cast<size_t>(r) * row_pitch + c * 8;
      cp_async16(dk, g + D);
      cp_async16(dv, g + 2 * D);
    } else {
      *reinterpret_cast<uint4*>(dk) = zero4;
      *reinterpret_cast<uint4*>(dv) = zero4;
    }
  }

  const int n_kblocks = TKP / 64;
  const int g = lane >> 2, t4 = lane & 3;

  for (int q0 = q_begin; q0 < T; q0 += 64) {
    __syncthreads();   // previous pass finished reading sQ
    for (int i = threadIdx.x; i < 64 * CHUNKS; i += blockDim.x) {
      const int r = i / CHUNKS, c = i - r * CHUNKS;
      __nv_bfloat16* dq = sQ + r * PITCH + c * 8;
      if (q0 + r < T && c < CHUNKS_REAL) cp_async16(dq, base + static_cast<size_t>(q0 + r) * row_pitch + c * 8);
      else *reinterpret_cast<uint4*>(dq) = zero4;
    }
    cp_async_wait_all();
    __syncthreads();
    if (q0 + warp * 16 >= T) continue;   // this warp's 16 query rows are all padding (warp-uniform; the block-wide
                                         // barrier at the top of the next pass is still reached by every warp)

    uint32_t qf[KSTEPS][4];
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
      const uint32_t addr = smem_u32(sQ + (warp * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8);
      ldsm_x4(addr, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }

    float o[NT_O][4];
#pragma unroll
    for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kb = 0; kb < n_kblocks; ++kb) {
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int key = kb * 64 + p * 16 + (lane & 7) + ((lane >> 4) << 3);
          const uint32_t addr = smem_u32(sK + key * PITCH + ks * 16 + ((lane >> 3) & 1) * 8);
          uint32_t b0, b1, b2, b3;
          ldsm_x4(addr, b0, b1, b2, b3);
          mma_bf16_16816(s[2 * p], qf[ks], b0, b1);
          mma_bf16_16816(s[2 * p + 1], qf[ks], b2, b3);
        }
      }
      if (kb == n_kblocks - 1 && TKP != T) {   // mask the zero-padded keys of the last block
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int key = kb * 64 + nt * 8 + 2 * t4;
          if (key >= T) s[nt][0] = s[nt][2] = -INFINITY;
          if (key + 1 >= T) s[nt][1] = s[nt][3] = -INFINITY;
        }
      }
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
        bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);   // finite: every block holds >= 1 real key
      const float c0 = exp2f((m0 - mn0) * scale_log2), c1 = exp2f((m1 - mn1) * scale_log2);
      m0 = mn0; m1 = mn1;
      const float ms0 = mn0 * scale_log2, ms1 = mn1 * scale_log2;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        s[nt][0] = exp2f(s[nt][0] * scale_log2 - ms0);
        s[nt][1] = exp2f(s[nt][1] * scale_log2 - ms0);
        s[nt][2] = exp2f(s[nt][2] * scale_log2 - ms1);
        s[nt][3] = exp2f(s[nt][3] * scale_log2 - ms1);
        rs0 += s[nt][0] + s[nt][1];
        rs1 += s[nt][2] + s[nt][3];
      }
      l0 = l0 * c0 + rs0;
      l1 = l1 * c1 + rs1;
#pragma unroll
      for (int i = 0; i < NT_O; ++i) {
        o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // 16 keys per PV k-step
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
        pa[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
        pa[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
        pa[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
        const int key = kb * 64 + j * 16 + (lane & 15);
#pragma unroll
        for (int p = 0; p < NT_O / 2; ++p) {
          const uint32_t addr = smem_u32(sV + key * PITCH + p * 16 + (lane >> 4) * 8);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(addr, b0, b1, b2, b3);
          mma_bf16_16816(o[2 * p], pa, b0, b1);
          mma_bf16_16816(o[2 * p + 1], pa, b2, b3);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __nv_bfloat16* ob = out + static_cast<size_t>(b) * T * D + h * HD;
#pragma unroll
    for (int nt = 0; nt < NT_O; ++nt) {
      const int col = nt * 8 + 2 * t4;
      if (col < HD) {
        if (r0 < T)
          *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * D + col) =
              pack_bf16x2(o[nt][0] * i0, o[nt][1] * i0);
        if (r1 < T)
          *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * D + col) =
              pack_bf16x2(o[nt][2] * i1, o[nt][3] * i1);
      }
    }
  }
}

template <int HD, int HDP>
static int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int q_begin,
                            cudaStream_t stream, int reverse) {
  constexpr int PITCH = HDP + 8;
  const int TKP = (T + 63) & ~63;
  const int smem = (2 * TKP + 64) * PITCH * 2;
  if (smem > 227 * 1024) return BLB_ERR_SHAPE;
  auto kern = attention_kernel<HD, HDP>;
  static std::atomic<int> configured_smem[BLB_MAX_DEVICES];   // the attribute is per device
  if (smem > configured_smem[current_device()]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured_smem[current_device()] = smem;
  }
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  TimingScope ts(TIME_ATTENTION, 4.0 * B * H * static_cast<double>(T - q_begin) * T * HD, stream);
  cudaError_t le = launch_pdl(kern, dim3(B * H), dim3(128), smem, stream, qkv, out, T, H, scale_log2, q_begin, reverse);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

// ---- streaming variant for long sequences (577 tokens of CLIP ViT-L/14-336, 729 / 734 tokens of the
// 384 px SigLIP / DINOv2 towers) -----------------------------------------------------------------------------------
// K/V of one head no longer fit next to each other in shared memory (729 keys x 80 padded dims x 2 tensors = 233 KB),
// so a CTA of NW warps takes 16·NW query rows per pass and streams K/V through shared memory in chunks of KC keys,
// carrying the online-softmax state (m, l, O) of its rows in registers across the chunks.  Same arithmetic as
// attention_kernel: bf16 mma.sync m16n8k16, fp32 softmax, exp2 with the scale folded in.
template <int HD, int HDP, int NW, int KC>
__global__ void __launch_bounds__(NW * 32, 2)
attention_stream_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int H,
                        float scale_log2, int reverse) {
  constexpr int PITCH = HDP + 8;
  constexpr int KSTEPS = HDP / 16;
  constexpr int NT_O = HDP / 8;
  constexpr int CHUNKS = HDP / 8;
  constexpr int CHUNKS_REAL = HD / 8;
  constexpr int QROWS = 16 * NW;
  extern __shared__ __align__(16) uint8_t smem_attn_s[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn_s);
  __nv_bfloat16* sV = sK + KC * PITCH;
  __nv_bfloat16* sQ = sV + KC * PITCH;

  const int q_passes = (T + QROWS - 1) / QROWS;
  const int wi = reverse ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int unit = wi / q_passes, q0 = (wi - unit * q_passes) * QROWS;
  const int b = unit / H, h = unit % H;
  const int D = H * HD;
  const size_t row_pitch = static_cast<size_t>(3) * D;
  const __nv_bfloat16* base = qkv + static_cast<size_t>(b) * T * row_pitch + h * HD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint4 zero4 = make_uint4(0, 0, 0, 0);
  const int g = lane >> 2, t4 = lane & 3;
  pdl_launch_dependents();
  pdl_wait();

  for (int i = threadIdx.x; i < QROWS * CHUNKS; i += blockDim.x) {
    const int r = i / CHUNKS, c = i - r * CHUNKS;
    __nv_bfloat16* dq = sQ + r * PITCH + c * 8;
    if (q0 + r < T && c < CHUNKS_REAL) cp_async16(dq, base + static_cast<size_t>(q0 + r) * row_pitch + c * 8);
    else *reinterpret_cast<uint4*>(dq) = zero4;
  }
  const bool live = q0 + warp * 16 < T;   // warp-uniform: all 16 rows of a dead warp are padding

  uint32_t qf[KSTEPS][4];
  float o[NT_O][4];
#pragma unroll
  for (int i = 0; i < NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int kc0 = 0; kc0 < T; kc0 += KC) {
    __syncthreads();   // every warp is done with the previous chunk
    for (int i = threadIdx.x; i < KC * CHUNKS; i += blockDim.x) {
      const int r = i / CHUNKS, c = i - r * CHUNKS;
      __nv_bfloat16* dk = sK + r * PITCH + c * 8;
      __nv_bfloat16* dv = sV + r * PITCH + c * 8;
      if (kc0 + r < T && c < CHUNKS_REAL) {
        const __nv_bfloat16* gp = base + static_cast<size_t>(kc0 + r) * row_pitch + c * 8;
        cp_async16(dk, gp + D);
        cp_async16(dv, gp + 2 * D);
      } else {
        *reinterpret_cast<uint4*>(dk) = zero4;
        *reinterpret_cast<uint4*>(dv) = zero4;
      }
    }
    cp_async_wait_all();
    __syncthreads();
    if (!live) continue;
    if (kc0 == 0) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint32_t addr = smem_u32(sQ + (warp * 16 + (lane & 15)) * PITCH + ks * 16 + (lane >> 4) * 8);
        ldsm_x4(addr, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    for (int kb = 0; kb < KC / 64 && kc0 + kb * 64 < T; ++kb) {   // blocks without a real key are skipped
      float sc[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int key = kb * 64 + p * 16 + (lane & 7) + ((lane >> 4) << 3);
          const uint32_t addr = smem_u32(sK + key * PITCH + ks * 16 + ((lane >> 3) & 1) * 8);
          uint32_t b0, b1, b2, b3;
          ldsm_x4(addr, b0, b1, b2, b3);
          mma_bf16_16816(sc[2 * p], qf[ks], b0, b1);
          mma_bf16_16816(sc[2 * p + 1], qf[ks], b2, b3);
        }
      }
      if (kc0 + kb * 64 + 64 > T) {   // mask the zero-padded keys of the last block
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const int key = kc0 + kb * 64 + nt * 8 + 2 * t4;
          if (key >= T) sc[nt][0] = sc[nt][2] = -INFINITY;
          if (key + 1 >= T) sc[nt][1] = sc[nt][3] = -INFINITY;
        }
      }
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        bm0 = fmaxf(bm0, fmaxf(sc[nt][0], sc[nt][1]));
        bm1 = fmaxf(bm1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);   // finite: every processed block holds >= 1 real key
      const float c0 = exp2f((m0 - mn0) * scale_log2), c1 = exp2f((m1 - mn1) * scale_log2);
      m0 = mn0; m1 = mn1;
      const float ms0 = mn0 * scale_log2, ms1 = mn1 * scale_log2;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        sc[nt][0] = exp2f(sc[nt][0] * scale_log2 - ms0);
        sc[nt][1] = exp2f(sc[nt][1] * scale_log2 - ms0);
        sc[nt][2] = exp2f(sc[nt][2] * scale_log2 - ms1);
        sc[nt][3] = exp2f(sc[nt][3] * scale_log2 - ms1);
        rs0 += sc[nt][0] + sc[nt][1];
        rs1 += sc[nt][2] + sc[nt][3];
      }
      l0 = l0 * c0 + rs0;
      l1 = l1 * c1 + rs1;
#pragma unroll
      for (int i = 0; i < NT_O; ++i) {
        o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // 16 keys per PV k-step
        uint32_t pa[4];
        pa[0] = pack_bf16x2(sc[2 * j][0], sc[2 * j][1]);
        pa[1] = pack_bf16x2(sc[2 * j][2], sc[2 * j][3]);
        pa[2] = pack_bf16x2(sc[2 * j + 1][0], sc[2 * j + 1][1]);
        pa[3] = pack_bf16x2(sc[2 * j + 1][2], sc[2 * j + 1][3]);
        const int key = kb * 64 + j * 16 + (lane & 15);
#pragma unroll
        for (int p = 0; p < NT_O / 2; ++p) {
          const uint32_t addr = smem_u32(sV + key * PITCH + p * 16 + (lane >> 4) * 8);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(addr, b0, b1, b2, b3);
          mma_bf16_16816(o[2 * p], pa, b0, b1);
          mma_bf16_16816(o[2 * p + 1], pa, b2, b3);
        }
      }
    }
  }
  if (!live) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  __nv_bfloat16* ob = out + static_cast<size_t>(b) * T * D + h * HD;
#pragma unroll
  for (int nt = 0; nt < NT_O; ++nt) {
    const int col = nt * 8 + 2 * t4;
    if (col < HD) {
      if (r0 < T)
        *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r0) * D + col) = pack_bf16x2(o[nt][0] * i0, o[nt][1] * i0);
      if (r1 < T)
        *reinterpret_cast<uint32_t*>(ob + static_cast<size_t>(r1) * D + col) = pack_bf16x2(o[nt][2] * i1, o[nt][3] * i1);
    }
  }
}

template <int HD, int HDP>
static int launch_attention_stream(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H,
                                   cudaStream_t stream, int reverse) {
  constexpr int NW = 8, KC = 256, PITCH = HDP + 8;
  constexpr int smem = (2 * KC + 16 * NW) * PITCH * 2;
  auto kern = attention_stream_kernel<HD, HDP, NW, KC>;
  static std::atomic<bool> configured[BLB_MAX_DEVICES];
  if (!configured[current_device()]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured[current_device()] = true;
  }
  const int q_passes = (T + 16 * NW - 1) / (16 * NW);
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  TimingScope ts(TIME_ATTENTION, 4.0 * B * H * static_cast<double>(T) * T * HD, stream);
  cudaError_t le = launch_pdl(kern, dim3(B * H * q_passes), dim3(NW * 32), smem, stream, qkv, out, T, H, scale_log2,
                              reverse);
  if (le != cudaSuccess) return static_cast<int>(le);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

int attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd, cudaStream_t stream,
                 int reverse);   // attention_tc.cu

static bool g_attn_v1_only = getenv("BLB_ATTN_V1") != nullptr;   // A/B switch: mma.sync kernels for everything

int attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd, cudaStream_t stream,
                   int reverse) {
  if (qkv == nullptr || out == nullptr || B <= 0 || T <= 0 || H <= 0) return BLB_ERR_ARG;
  if (hd != 64 && hd != 72) return BLB_ERR_SHAPE;
  if (!g_attn_v1_only) {
    // tower shapes (224 px: 256 / 261 tokens): the tcgen05 kernel takes every query row (DINOv2's 5 rows beyond the two
    // 128-row tiles are a third tile inside it)
    const int rc = attention_tc(qkv, out, B, T, H, hd, stream, reverse);
    if (rc != BLB_ERR_SHAPE) return rc;
  }
  if (T > 320) {   // long sequences (336 / 384 px towers): K/V streamed through shared memory
    if (hd == 64) return launch_attention_stream<64, 64>(qkv, out, B, T, H, stream, reverse);
    return launch_attention_stream<72, 80>(qkv, out, B, T, H, stream, reverse);
  }
  if (hd == 64) return launch_attention<64, 64>(qkv, out, B, T, H, 0, stream, reverse);
  return launch_attention<72, 80>(qkv, out, B, T, H, 0, stream, reverse);
}

}  // namespace blb
