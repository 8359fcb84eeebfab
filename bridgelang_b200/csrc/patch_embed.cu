// patch_embed.cu — staging for timm PatchEmbed = Conv2d(3, D, kernel 14, stride 14, bias) run as a GEMM
// (SURVEY.md §2 K1/K2, §8 a4/a5).  Images are S x S with S = 14·grid (224 px: grid 16, 336: 24, 378/384: 27).
//
// im2col_patch14: pixels [B,3,S,S] bf16 → cols [B*grid², ldk] bf16 with column index k = c*196 + kh*14 + kw,
// i.e. exactly the flattening order of the conv weight [D,3,14,14] → [D,588]; patch rows are ordered h then w
// (flatten(2).transpose(1,2)).  Columns 588..ldk-1 are zero so the row pitch is 16-byte aligned for TMA.
// The conv bias and the position embedding are added by the GEMM epilogue (EPI_PATCH).
//
// u8_to_patches (the north star's "TMA-staged im2col"): because stride == kernel, im2col is a pure
// PERMUTATION of the frame.  The uint8 HWC frame is therefore emitted once, directly in patch-major order
// (k = kh*42 + kw*3 + c, the order in which a patch row lies in HWC memory), as exact bf16 integers 0..255; ToTensor and
// each tower's Normalize are folded into that tower's patch-embed weights / bias at pack time (vision.py), so this ONE
// matrix is the TMA-loaded A operand of BOTH towers' patch-embed GEMMs: no normalized frames, no second im2col pass.
//
// write_prefix_tokens: DINOv2-reg4 prepends cls + 4 register tokens *after* the pos-embed add
// (no_embed_class=True), so those 5 rows of the residual stream are input independent.
#include "gemm.h"
#include "ptx.cuh"

namespace blb {

constexpr int IMG = 224, PATCH = 14, KREAL = 3 * PATCH * PATCH;  // 588  (IMG: the LUT path below is 224 px only)

__global__ void __launch_bounds__(256) im2col_kernel(const __nv_bfloat16* __restrict__ px,
                                                     __nv_bfloat16* __restrict__ cols, int B, int ldk, int grid,
                                                     int S) {
  // one thread per (patch row, c*14+kh): 14 contiguous pixels = 28 bytes = 7 x 4-byte words
  // S = image side in pixels (>= 14·grid: a 384 px frame has grid 27 and 6 unused trailing rows / columns)
  const int P = grid * grid;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * P * 42;
  if (idx >= total) return;
  const int seg = static_cast<int>(idx % 42);
  const long long row = idx / 42;
  const int p = static_cast<int>(row % P);
  const int b = static_cast<int>(row / P);
  const int c = seg / 14, kh = seg % 14;
  const int ph = p / grid, pw = p % grid;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(
      px + ((static_cast<size_t>(b) * 3 + c) * S + ph * PATCH + kh) * S + pw * PATCH);
  uint32_t* dst = reinterpret_cast<uint32_t*>(cols + static_cast<size_t>(row) * ldk + seg * PATCH);
#pragma unroll
  for (int i = 0; i < 7; ++i) dst[i] = __ldg(src + i);
  if (seg == 41) {
    __nv_bfloat16* tail = cols + static_cast<size_t>(row) * ldk;
    for (int k = KREAL; k < ldk; ++k) tail[k] = __float2bfloat16(0.f);
  }
}

int im2col_patch14(const __nv_bfloat16* pixels, __nv_bfloat16* cols, int B, int ldk, cudaStream_t stream, int grid,
                   int img) {
  if (pixels == nullptr || cols == nullptr || B <= 0 || grid <= 0) return BLB_ERR_ARG;
  if (img <= 0) img = grid * PATCH;
  if (ldk < KREAL || ldk % 8 != 0 || img % 2 != 0 || img < grid * PATCH) return BLB_ERR_SHAPE;   // 4-byte aligned 14-pixel runs
  const long long total = static_cast<long long>(B) * grid * grid * 42;
  TimingScope ts(TIME_OTHER, 4.0 * B * grid * grid * 588, stream);
  im2col_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(pixels, cols, B, ldk, grid, img);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

// uint8 HWC frame [B,S,S,3] → bf16 [B*grid², ldk], k = kh*42 + kw*3 + c (a patch row = 42 contiguous source bytes);
// one thread per (patch, kh): 42 bytes in (as 21 aligned 2-byte loads: S*3 and 42 are even), 84 bytes out.
__global__ void __launch_bounds__(256) u8_to_patches_kernel(const uint8_t* __restrict__ frames,
                                                            __nv_bfloat16* __restrict__ cols, int B, int ldk,
                                                            int grid, int S) {
  const int P = grid * grid;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * P * PATCH;
  if (idx >= total) return;
  const int kh = static_cast<int>(idx % PATCH);
  const long long row = idx / PATCH;
  const int p = static_cast<int>(row % P);
  const int b = static_cast<int>(row / P);
  const int ph = p / grid, pw = p % grid;
  const uint16_t* src = reinterpret_cast<const uint16_t*>(
      frames + ((static_cast<size_t>(b) * S + ph * PATCH + kh) * S + pw * PATCH) * 3);
  uint32_t* dst = reinterpret_cast<uint32_t*>(cols + static_cast<size_t>(row) * ldk + kh * 42);
#pragma unroll
  for (int i = 0; i < 21; ++i) {
    const uint32_t w = __ldg(src + i);
    // bf16 of an integer 0..255 is exact: build the bit pattern through the fp32 conversion
    const uint32_t lo = __float_as_uint(static_cast<float>(w & 0xFFu)) >> 16;
    const uint32_t hi = __float_as_uint(static_cast<float>(w >> 8)) >> 16;
    dst[i] = lo | (hi << 16);
  }
  if (kh == PATCH - 1) {
    __nv_bfloat16* tail = cols + static_cast<size_t>(row) * ldk;
    for (int k = KREAL; k < ldk; ++k) tail[k] = __float2bfloat16(0.f);
  }
}

int u8_to_patches(const uint8_t* frames, __nv_bfloat16* cols, int B, int ldk, int grid, cudaStream_t stream, int img) {
  if (frames == nullptr || cols == nullptr || B <= 0 || grid <= 0) return BLB_ERR_ARG;
  if (img <= 0) img = grid * PATCH;
  if (ldk < KREAL || ldk % 8 != 0 || img % 2 != 0 || img < grid * PATCH) return BLB_ERR_SHAPE;
  if ((reinterpret_cast<uintptr_t>(frames) & 1) != 0) return BLB_ERR_ALIGN;
  const long long total = static_cast<long long>(B) * grid * grid * PATCH;
  TimingScope ts(TIME_OTHER, 3.0 * B * grid * grid * 588, stream);   // bytes: 1 in + 2 out per element
  u8_to_patches_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(frames, cols, B, ldk, grid, img);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

// ---- device-side image preprocessing -------------------------------------------------------------
// One uint8 HWC frame feeds BOTH towers: ToTensor (x/255) + per-tower Normalize ((t-mean)/std) + the bf16 cast of
// `.to(device, dtype=bf16)` (processing_prismatic.py:128-145, dinosiglip_vit.py:33-40, run_openvla_demo.py:38-41)
// collapse to a 256-entry table per (tower, channel) — built on the host with the reference's own torch ops, so the
// result is bit-identical by construction — and the kernel is a pure gather: 3 B read, 12 B written per pixel.
// lut: bf16 [2 towers][3 channels][256].   frames: uint8 [B,224,224,3]   →   out_*: bf16 [B,3,224,224]
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ frames,
                                                            const __nv_bfloat16* __restrict__ lut,
                                                            __nv_bfloat16* __restrict__ out_dino,
                                                            __nv_bfloat16* __restrict__ out_siglip, int B) {
  __shared__ __nv_bfloat16 s_lut[2 * 3 * 256];
  for (int i = threadIdx.x; i < 2 * 3 * 256; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  // one thread per 4 consecutive pixels of a row: 12 contiguous input bytes, 3 channels x 4 bf16 (8 B) out per tower
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * IMG * (IMG / 4);
  if (idx >= total) return;
  const int x4 = static_cast<int>(idx % (IMG / 4));
  const long long rowi = idx / (IMG / 4);
  const int y = static_cast<int>(rowi % IMG);
  const int b = static_cast<int>(rowi / IMG);
  const uint32_t* src = reinterpret_cast<const uint32_t*>(frames + ((static_cast<size_t>(b) * IMG + y) * IMG + x4 * 4) * 3);
  const uint32_t w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
  uint8_t px[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    px[i] = (w0 >> (8 * i)) & 0xFF;
    px[4 + i] = (w1 >> (8 * i)) & 0xFF;
    px[8 + i] = (w2 >> (8 * i)) & 0xFF;
  }
#pragma unroll
  for (int tower = 0; tower < 2; ++tower) {
    __nv_bfloat16* out = tower == 0 ? out_dino : out_siglip;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const __nv_bfloat16* t = s_lut + (tower * 3 + c) * 256;
      __nv_bfloat16 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = t[px[3 * i + c]];
      uint2 pk;
      pk.x = static_cast<uint32_t>(__bfloat16_as_ushort(v[0])) | (static_cast<uint32_t>(__bfloat16_as_ushort(v[1])) << 16);
      pk.y = static_cast<uint32_t>(__bfloat16_as_ushort(v[2])) | (static_cast<uint32_t>(__bfloat16_as_ushort(v[3])) << 16);
      *reinterpret_cast<uint2*>(out + ((static_cast<size_t>(b) * 3 + c) * IMG + y) * IMG + x4 * 4) = pk;
    }
  }
}

int preprocess_u8(const uint8_t* frames, int B, const __nv_bfloat16* lut, __nv_bfloat16* out_dino,
                  __nv_bfloat16* out_siglip, cudaStream_t stream) {
  if (frames == nullptr || lut == nullptr || out_dino == nullptr || out_siglip == nullptr || B <= 0) return BLB_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(frames) & 3) != 0 || (reinterpret_cast<uintptr_t>(out_dino) & 7) != 0 ||
      (reinterpret_cast<uintptr_t>(out_siglip) & 7) != 0)
    return BLB_ERR_ALIGN;
  const long long total = static_cast<long long>(B) * IMG * (IMG / 4);
  TimingScope ts(TIME_OTHER, 15.0 * B * IMG * IMG, stream);
  preprocess_u8_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(frames, lut, out_dino, out_siglip, B);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

__global__ void prefix_kernel(const float* __restrict__ prefix, float* __restrict__ resid, int T, int n_prefix,
                              int D) {
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < n_prefix * D; i += blockDim.x)
    resid[static_cast<size_t>(b) * T * D + i] = prefix[i];
}

int write_prefix_tokens(const float* prefix, float* resid, int B, int T, int n_prefix, int D, cudaStream_t stream) {
  if (n_prefix == 0) return 0;
  if (prefix == nullptr || resid == nullptr || B <= 0 || n_prefix > T) return BLB_ERR_ARG;
  prefix_kernel<<<B, 256, 0, stream>>>(prefix, resid, T, n_prefix, D);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace blb
