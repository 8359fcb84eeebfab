// mma_bench3.cu — why does O = P·V (17 chained M128 x N64 x K16 MMAs into ONE accumulator) take ~110 cycles per MMA
// inside attention_tc?  Hypothesis: back-to-back MMAs that accumulate into the same TMEM columns are serialised by
// the accumulate latency, which a 32-cycle MMA cannot hide.  One control warp issues `trips` x 16 MMAs, rotating over
// NACC accumulators; TS (A in TMEM) or SS; N = 64 / 128 / 256.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/mma_bench3 tools/mma_bench3.cu
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

template <int TS, int N, int NACC, int BMN>
__global__ void __launch_bounds__(384, 1) k(int trips, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  uint8_t* base = (uint8_t*)(((uintptr_t)sm + 1023) & ~(uintptr_t)1023);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (warp == 10) tmem_alloc<1>(&slot, 512);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  const uint32_t a = smem_u32(base), b = smem_u32(base + 64 * 1024);
  constexpr uint32_t id = idesc(128, N, BMN);
  // TMEM: A (P) at columns [0, 128) for TS; accumulators from column 128 on (NACC * N <= 384)
  if (warp == 9) {
    long long t0 = clock64();
    for (int r = 0; r < trips; ++r) {
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t d = tm + 128 + (j % NACC) * N;
          const uint32_t accum = (r != 0 || j >= NACC) ? 1u : 0u;
          const uint64_t bd = BMN ? mk_desc(b + j * 2048, 1024, 2) : mk_desc(b + (j >> 2) * 32768, 1024, 2) + 2 * (j & 3);
          if (TS) mma_ts(d, tm + j * 8, bd, id, accum);
          else umma_bf16<1>(d, mk_desc(a + (j >> 2) * 16384, 1024, 2) + 2 * (j & 3), bd, id, accum);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit<1>(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<1>(tm, 512);
}

template <int TS, int N, int NACC, int BMN>
void run(long long* d) {
  auto kern = k<TS, N, NACC, BMN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  kern<<<1, 384, 200 * 1024>>>(64, d);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s N=%3d B=%s accumulators=%d: %.1f cyc/mma  (ideal %d)  %s\n", TS ? "TS" : "SS", N, BMN ? "MN-major" : "K-major",
         NACC, (double)h / (64 * 16), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<1, 64, 1, 1>(d); run<1, 64, 2, 1>(d); run<1, 64, 4, 1>(d);
  run<0, 64, 1, 1>(d); run<0, 64, 2, 1>(d); run<0, 64, 4, 1>(d);
  run<0, 64, 1, 0>(d); run<0, 64, 4, 0>(d);
  run<1, 128, 1, 1>(d); run<1, 128, 2, 1>(d);
  run<0, 128, 1, 0>(d); run<0, 256, 1, 0>(d);
  run<1, 16, 1, 1>(d); run<1, 16, 4, 1>(d);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
