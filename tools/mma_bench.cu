// mma_bench.cu — cycles per tcgen05.mma (cta_group::1, kind::f16, M=128) vs N, operand source and ISSUE STYLE.
//   style 0: divergent single thread (`if (threadIdx.x == 32)`), 16 MMAs unrolled per loop trip
//   style 1: warp-uniform loop, one elect_one() region around the 16 unrolled MMAs
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/mma_bench tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../bridgelang_b200/csrc/ptx.cuh"
using namespace blb;

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t sbo, uint32_t lbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
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

// MODE: 0 = SS B K-major SW128 | 1 = SS B MN-major SW128 | 2 = TS B MN-major SW128
template <int N, int MODE>
__device__ __forceinline__ void issue16(uint32_t tm, uint32_t a, uint32_t b, uint32_t first_acc) {
  constexpr uint32_t id = idesc(128, N, MODE == 0 ? 0 : 1);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t acc = j > 0 ? 1u : first_acc;
    if (MODE == 0) umma_bf16<1>(tm, mk_desc(a, 1024, 16, 2) + 2 * (j & 3), mk_desc(b, 1024, 16, 2) + 2 * (j & 3), id, acc);
    else if (MODE == 1) umma_bf16<1>(tm, mk_desc(a + (j >> 2) * 16384, 1024, 16, 2) + 2 * (j & 3), mk_desc(b + j * 2048, 1024, 16, 2), id, acc);
    else mma_ts(tm, tm + 300 + j * 8, mk_desc(b + j * 2048, 1024, 16, 2), id, acc);
  }
}

template <int N, int MODE, int STYLE>
__global__ void __launch_bounds__(128, 1) mma_kernel(int trips, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc<1>(&slot, 512);
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t a = smem_u32(base), b = smem_u32(base + 64 * 1024);
  if (STYLE == 0) {
    if (threadIdx.x == 32) {
      long long t0 = clock64();
      for (int r = 0; r < trips; ++r) issue16<N, MODE>(tm, a, b, r > 0);
      umma_commit<1>(&bar);
      long long t1 = clock64();
      mbar_wait(&bar, 0);
      out[0] = t1 - t0;
      out[1] = clock64() - t0;
    }
  } else {
    if (warp == 1) {
      long long t0 = clock64();
      for (int r = 0; r < trips; ++r) {
        if (elect_one()) issue16<N, MODE>(tm, a, b, r > 0);
        __syncwarp();
      }
      if (elect_one()) umma_commit<1>(&bar);
      __syncwarp();
      long long t1 = clock64();
      mbar_wait(&bar, 0);
      if (threadIdx.x == 32) {
        out[0] = t1 - t0;
        out[1] = clock64() - t0;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(tm, 512);
}

template <int N, int MODE, int STYLE>
void run(long long* d) {
  const char* names[] = {"SS B=K-major", "SS B=MN-major", "TS B=MN-major"};
  const int trips = 32;
  auto k = mma_kernel<N, MODE, STYLE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<<<1, 128, 200 * 1024>>>(trips, d);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-14s N=%3d style=%d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", names[MODE], N, STYLE,
         (double)h[0] / (trips * 16), (double)h[1] / (trips * 16));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<16, 0, 0>(d); run<16, 0, 1>(d);
  run<64, 0, 0>(d); run<64, 0, 1>(d);
  run<128, 0, 0>(d); run<128, 0, 1>(d);
  run<144, 0, 0>(d); run<144, 0, 1>(d);
  run<192, 0, 0>(d); run<192, 0, 1>(d);
  run<256, 0, 0>(d); run<256, 0, 1>(d);
  run<64, 1, 0>(d); run<64, 1, 1>(d);
  run<80, 1, 0>(d); run<80, 1, 1>(d);
  run<64, 2, 0>(d); run<64, 2, 1>(d);
  run<80, 2, 0>(d); run<80, 2, 1>(d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
