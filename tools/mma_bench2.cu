// mma_bench2.cu — does SIMT-side background work slow tcgen05.mma down?  One control warp issues 32x16 unrolled
// MMAs (M128 N64) either SS (A,B smem; B MN-major) or TS (A in TMEM); 8 other warps run a background loop:
//   bg 0: idle   1: tcgen05.ld loop   2: MUFU.EX2 loop   3: st.shared loop   4: FFMA loop
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/mma_bench2 tools/mma_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../bridgelang_b200/csrc/ptx.cuh"
using namespace blb;

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc(int m, int n, int bmn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)bmn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <int TS, int CTRL_WARP>
__global__ void __launch_bounds__(384, 1) k(int bg, int trips, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int done;
  uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (warp == 10) tmem_alloc<1>(&slot, 512);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t a = smem_u32(base), b = smem_u32(base + 64 * 1024);
  constexpr uint32_t id = idesc(128, 64, 1);
  if (warp == CTRL_WARP) {
    long long t0 = clock64();
    for (int r = 0; r < trips; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (TS) mma_ts(tm + 400, tm + 256 + j * 8, mk_desc(b + j * 2048, 1024, 2), id, (r | j) != 0);
          else umma_bf16<1>(tm + 400, mk_desc(a + (j >> 2) * 16384, 1024, 2) + 2 * (j & 3), mk_desc(b + j * 2048, 1024, 2), id, (r | j) != 0);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit<1>(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[0] = t1 - t0; done = 1; }
  } else if (warp < 8 && bg != 0) {
    float acc = 0.f;
    const uint32_t lane_addr = tm + ((uint32_t)((warp & 3) * 32) << 16);
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
    while (!done) {
      if (bg == 1) {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) { tmem_ld_32x32(lane_addr + c * 32 + (warp >> 2) * 128, r); tmem_ld_wait(); acc += __uint_as_float(r[c]); }
      } else if (bg == 2) {
#pragma unroll
        for (int it = 0; it < 8; ++it)
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = ex2_approx(x[j]) - 1.0f;
      } else if (bg == 3) {
#pragma unroll
        for (int it = 0; it < 16; ++it)
          *reinterpret_cast<uint4*>(base + 128 * 1024 + ((threadIdx.x * 16 + it * 4096) & 0x7fff)) = make_uint4(it, 0, 0, 0);
      } else {
#pragma unroll
        for (int it = 0; it < 16; ++it)
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = fmaf(x[j], 1.0001f, 0.5f);
      }
    }
    for (int j = 0; j < 8; ++j) acc += x[j];
    if (acc == 123.456f) sink[0] = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<1>(tm, 512);
}

template <int TS, int CW>
void run(long long* d, float* sink) {
  const char* bgn[] = {"idle", "tcgen05.ld", "MUFU.EX2", "st.shared", "FFMA"};
  auto kern = k<TS, CW>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int bg = 0; bg < 5; ++bg) {
    kern<<<1, 384, 200 * 1024>>>(bg, 64, d, sink);
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%s ctrl-warp=%d background=%-11s: %.1f cyc/mma\n", TS ? "TS" : "SS", CW, bgn[bg], (double)h / (64 * 16));
  }
}

int main() {
  long long* d; float* sink; cudaMalloc(&d, 16); cudaMalloc(&sink, 4);
  run<0, 9>(d, sink); run<1, 9>(d, sink); run<0, 8>(d, sink);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
