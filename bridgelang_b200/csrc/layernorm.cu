// layernorm.cu — fused LayerNorm over the fp32 residual stream, bf16 output (the GEMM A operand).
//
// Replaces timm Block.norm1 / norm2 = nn.LayerNorm(D, eps=1e-6): biased variance,
// fp32 statistics, affine.  One warp per token row; the row lives in registers between the
// mean pass, the variance pass (about the mean, like ATen) and the normalise/store pass, so HBM
// sees exactly one fp32 read and one bf16 write per element (6 B/element — the roofline for this op).
#include "gemm.h"
#include "ptx.cuh"

namespace blb {

// F32OUT: fp32 output (may alias the input: the row lives in registers) — timm `norm_pre` of the CLIP towers, which
// rewrites the residual stream itself before block 0
template <int VEC, bool F32OUT = false>  // D = VEC * 128
__global__ void __launch_bounds__(256) layernorm_kernel(const float* x, int ldx,
                                                        const float* __restrict__ w, const float* __restrict__ b,
                                                        void* y, int ldy, int rows, float eps, int reverse) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk = reverse ? static_cast<int>(gridDim.x - 1 - blockIdx.x) : static_cast<int>(blockIdx.x);
  const int row = blk * 8 + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= rows) return;
  constexpr int D = VEC * 128;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
  float4 v[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) v[j] = xr[lane + 32 * j];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
    q += (a * a + c * c) + (d * d + e * e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.0f / D) + eps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  uint2* yr = reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(y) + static_cast<size_t>(row) * ldy);
  float4* yf = reinterpret_cast<float4*>(static_cast<float*>(y) + static_cast<size_t>(row) * ldy);
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float4 ww = __ldg(w4 + lane + 32 * j);
    const float4 bb = __ldg(b4 + lane + 32 * j);
    const float o0 = (v[j].x - mean) * rstd * ww.x + bb.x, o1 = (v[j].y - mean) * rstd * ww.y + bb.y;
    const float o2 = (v[j].z - mean) * rstd * ww.z + bb.z, o3 = (v[j].w - mean) * rstd * ww.w + bb.w;
    if constexpr (F32OUT) {
      yf[lane + 32 * j] = make_float4(o0, o1, o2, o3);
    } else {
      uint2 o;
      o.x = pack_bf16x2(o0, o1);
      o.y = pack_bf16x2(o2, o3);
      yr[lane + 32 * j] = o;
    }
  }
}

int layernorm_f32_bf16(const float* x, int ldx, const float* w, const float* b, __nv_bfloat16* y, int ldy, int rows,
                       int D, float eps, cudaStream_t stream, int reverse) {
  if (x == nullptr || w == nullptr || b == nullptr || y == nullptr || rows <= 0) return BLB_ERR_ARG;
  if (D % 128 != 0 || D > 2048 || ldx % 4 != 0 || ldy % 4 != 0) return BLB_ERR_SHAPE;
  const dim3 grid((rows + 7) / 8), block(256);
  cudaError_t le = cudaSuccess;
  TimingScope ts(TIME_LAYERNORM, 6.0 * rows * D, stream);   // bytes: fp32 read + bf16 write
  switch (D / 128) {
#define BLB_LN_CASE(V) \
  case V: le = launch_pdl(layernorm_kernel<V, false>, grid, block, 0, stream, x, ldx, w, b, static_cast<void*>(y), ldy, rows, eps, reverse); break;
    BLB_LN_CASE(1) BLB_LN_CASE(2) BLB_LN_CASE(3) BLB_LN_CASE(4) BLB_LN_CASE(5) BLB_LN_CASE(6) BLB_LN_CASE(7)
    BLB_LN_CASE(8) BLB_LN_CASE(9) BLB_LN_CASE(10) BLB_LN_CASE(11) BLB_LN_CASE(12) BLB_LN_CASE(13)
    BLB_LN_CASE(14) BLB_LN_CASE(15) BLB_LN_CASE(16)
#undef BLB_LN_CASE
    default: return BLB_ERR_SHAPE;
  }
  count_launch(1);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

// timm norm_pre (pre_norm=True, the CLIP towers): fp32 rows → fp32 rows, in place allowed
int layernorm_f32_f32(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int rows, int D,
                      float eps, cudaStream_t stream) {
  if (x == nullptr || w == nullptr || b == nullptr || y == nullptr || rows <= 0) return BLB_ERR_ARG;
  if (D % 128 != 0 || D > 2048 || ldx % 4 != 0 || ldy % 4 != 0) return BLB_ERR_SHAPE;
  const dim3 grid((rows + 7) / 8), block(256);
  cudaError_t le = cudaSuccess;
  TimingScope ts(TIME_LAYERNORM, 8.0 * rows * D, stream);
  switch (D / 128) {
#define BLB_LNF_CASE(V) \
  case V: le = launch_pdl(layernorm_kernel<V, true>, grid, block, 0, stream, x, ldx, w, b, static_cast<void*>(y), ldy, rows, eps, 0); break;
    BLB_LNF_CASE(6) BLB_LNF_CASE(8) BLB_LNF_CASE(9) BLB_LNF_CASE(10) BLB_LNF_CASE(12)
#undef BLB_LNF_CASE
    default: return BLB_ERR_SHAPE;
  }
  count_launch(1);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

// Primer of the LN-folded chain: per row the exact mean c (→ shift), the bf16 copy of x − c and the
// (sum, sum of squares) of x − c.  Runs once per tower, after the patch-embed stage; every later block gets all three
// from the EPI_RESIDUAL epilogues (which roll the shift forward, see gemm_tcgen05.cu).  shift == nullptr: c = 0.
template <int VEC>
__global__ void __launch_bounds__(256) rowstats_cast_kernel(const float* __restrict__ x, int ldx,
                                                            __nv_bfloat16* __restrict__ y, int ldy,
                                                            float2* __restrict__ stats, int parts, int rows,
                                                            float* __restrict__ shift) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = static_cast<int>(blockIdx.x) * 8 + warp;
  pdl_launch_dependents();
  pdl_wait();
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * ldx);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * ldy);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    v[j] = xr[lane + 32 * j];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float c = shift != nullptr ? s / static_cast<float>(VEC * 128) : 0.f;
  float s1 = 0.f, q = 0.f;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const float a0 = v[j].x - c, a1 = v[j].y - c, a2 = v[j].z - c, a3 = v[j].w - c;
    s1 += (a0 + a1) + (a2 + a3);
    q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    uint2 o;
    o.x = pack_bf16x2(a0, a1);
    o.y = pack_bf16x2(a2, a3);
    yr[lane + 32 * j] = o;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  for (int p = lane; p < parts; p += 32)   // stats are [part][row]
    stats[static_cast<size_t>(p) * rows + row] = p == 0 ? make_float2(s1, q) : make_float2(0.f, 0.f);
  if (shift != nullptr && lane == 0) shift[row] = c;
}

int rowstats_cast_f32_bf16(const float* x, int ldx, __nv_bfloat16* y, int ldy, float2* stats, int parts, int rows,
                           int D, cudaStream_t stream, float* shift) {
  if (x == nullptr || y == nullptr || stats == nullptr || rows <= 0 || parts <= 0) return BLB_ERR_ARG;
  if (D % 128 != 0 || D > 2048 || ldx % 4 != 0 || ldy % 4 != 0) return BLB_ERR_SHAPE;
  const dim3 grid((rows + 7) / 8), block(256);
  cudaError_t le = cudaSuccess;
  TimingScope ts(TIME_LAYERNORM, 6.0 * rows * D, stream);
  switch (D / 128) {
#define BLB_RS_CASE(V) \
  case V: le = launch_pdl(rowstats_cast_kernel<V>, grid, block, 0, stream, x, ldx, y, ldy, stats, parts, rows, shift); break;
    BLB_RS_CASE(1) BLB_RS_CASE(2) BLB_RS_CASE(3) BLB_RS_CASE(4) BLB_RS_CASE(5) BLB_RS_CASE(6) BLB_RS_CASE(7)
    BLB_RS_CASE(8) BLB_RS_CASE(9) BLB_RS_CASE(10) BLB_RS_CASE(11) BLB_RS_CASE(12) BLB_RS_CASE(13)
    BLB_RS_CASE(14) BLB_RS_CASE(15) BLB_RS_CASE(16)
#undef BLB_RS_CASE
    default: return BLB_ERR_SHAPE;
  }
  count_launch(1);
  if (le != cudaSuccess) return static_cast<int>(le);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace blb
