// resize.cu — antialiased bicubic resize of uint8 HWC frames on the device, bit-exact with Pillow.
//
// Replaces the host-side `Resize(..., interpolation=BICUBIC)` of the image transform (prismatic
// dinosiglip_vit.py:91-111 "resize-naive", processing_prismatic.py:128-145 → torchvision → PIL.Image.resize), the last
// CPU stage in front of the visual-prefix path.  Pillow's ImagingResample for 8-bit images is a
// deterministic fixed-point algorithm: per output coordinate a window [xmin, xmin+xmax) of integer coefficients
// (the normalised filter weights scaled by 2^22 and rounded), a horizontal pass into an 8-bit intermediate and a
// vertical pass, each `clip8((2^21 + Σ pixel·coeff) >> 22)`.  The coefficient tables are built on the host with the
// very same double-precision arithmetic (bridgelang_b200/resize.py) — so the two integer passes below reproduce
// PIL's bytes exactly (tests compare against PIL.Image.resize itself).
#include "gemm.h"

namespace blb {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;   // Pillow Resample.c

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PRECISION_BITS;   // arithmetic shift, as in Pillow's clip8 lookup
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [B,H,Ws,3] → dst [B,H,Wd,3]
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       long long total, int Ws, int Wd, const int* __restrict__ kk,
                                                       const int* __restrict__ bounds, int ksize) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int xo = static_cast<int>(idx % Wd);
  const long long row = idx / Wd;   // b*H + y
  const int xmin = __ldg(bounds + 2 * xo), xmax = __ldg(bounds + 2 * xo + 1);
  const int* k = kk + static_cast<size_t>(xo) * ksize;
  const uint8_t* s = src + (static_cast<size_t>(row) * Ws + xmin) * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int i = 0; i < xmax; ++i) {
    const int w = __ldg(k + i);
    s0 += static_cast<int>(s[3 * i]) * w;
    s1 += static_cast<int>(s[3 * i + 1]) * w;
    s2 += static_cast<int>(s[3 * i + 2]) * w;
  }
  uint8_t* d = dst + static_cast<size_t>(idx) * 3;
  d[0] = clip8(s0);
  d[1] = clip8(s1);
  d[2] = clip8(s2);
}

// vertical pass: src [B,Hs,W,3] → dst [B,Hd,W,3]
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                       long long total, int Hs, int Hd, int W,
                                                       const int* __restrict__ kk, const int* __restrict__ bounds,
                                                       int ksize) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = static_cast<int>(idx % W);
  const long long t = idx / W;
  const int yo = static_cast<int>(t % Hd);
  const long long b = t / Hd;
  const int ymin = __ldg(bounds + 2 * yo), ymax = __ldg(bounds + 2 * yo + 1);
  const int* k = kk + static_cast<size_t>(yo) * ksize;
  const uint8_t* s = src + ((static_cast<size_t>(b) * Hs + ymin) * W + x) * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int i = 0; i < ymax; ++i) {
    const int w = __ldg(k + i);
    const uint8_t* p = s + static_cast<size_t>(i) * W * 3;
    s0 += static_cast<int>(p[0]) * w;
    s1 += static_cast<int>(p[1]) * w;
    s2 += static_cast<int>(p[2]) * w;
  }
  uint8_t* d = dst + static_cast<size_t>(idx) * 3;
  d[0] = clip8(s0);
  d[1] = clip8(s1);
  d[2] = clip8(s2);
}

}  // namespace

// kx/bx: horizontal coefficient table [Wd, ksx] and bounds [Wd, 2] (xmin, count); ky/by likewise for the rows.
// tmp: [B, Hs, Wd, 3] bytes.  A pass whose size does not change is skipped like Pillow does (then tmp may be NULL
// if neither... the horizontal pass writes tmp only when both passes run).
int resize_u8(const uint8_t* src, int B, int Hs, int Ws, uint8_t* dst, int Hd, int Wd, const int* kx, const int* bx,
              int ksx, const int* ky, const int* by, int ksy, uint8_t* tmp, cudaStream_t stream) {
  if (src == nullptr || dst == nullptr || B <= 0 || Hs <= 0 || Ws <= 0 || Hd <= 0 || Wd <= 0) return BLB_ERR_ARG;
  const bool need_h = Wd != Ws, need_v = Hd != Hs;
  if ((need_h && (kx == nullptr || bx == nullptr || ksx <= 0)) || (need_v && (ky == nullptr || by == nullptr || ksy <= 0)))
    return BLB_ERR_ARG;
  if (need_h && need_v && tmp == nullptr) return BLB_ERR_ARG;
  TimingScope ts(TIME_OTHER, 3.0 * B * (static_cast<double>(Hs) * Ws + static_cast<double>(Hd) * Wd), stream);
  if (!need_h && !need_v) {
    cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(B) * Hs * Ws * 3, cudaMemcpyDeviceToDevice, stream);
    return static_cast<int>(e);
  }
  const uint8_t* vin = src;
  if (need_h) {
    uint8_t* hout = need_v ? tmp : dst;
    const long long total = static_cast<long long>(B) * Hs * Wd;
    resize_h_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src, hout, total, Ws, Wd, kx, bx, ksx);
    count_launch(1);
    vin = hout;
  }
  if (need_v) {
    const long long total = static_cast<long long>(B) * Hd * Wd;
    resize_v_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(vin, dst, total, Hs, Hd, Wd, ky, by, ksy);
    count_launch(1);
  }
  return static_cast<int>(cudaGetLastError());
}

}  // namespace blb
