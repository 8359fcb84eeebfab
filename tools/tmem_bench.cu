// tmem_bench.cu — microbenchmarks that size the attention / epilogue designs on B200:
//   (1) tcgen05.ld 32x32b.x32 throughput per SM with 4 and 8 reader warps
//   (2) MUFU.EX2 throughput per SM
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../bridgelang_b200/csrc/ptx.cuh"
using namespace blb;

__global__ void __launch_bounds__(256, 1) tmem_read_kernel(int iters, int nwarps, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<1>(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(base + lane_base + c * 32 + (warp >> 2) * 256, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 8) acc += __uint_as_float(r[j]);
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(base, 512);
}

__global__ void __launch_bounds__(256, 1) tmem_read_nowait_kernel(int iters, int nwarps, long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<1>(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  float acc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    for (int it = 0; it < iters; ++it) {
      uint32_t r0[32], r1[32], r2[32], r3[32];
      const uint32_t a = base + lane_base + (warp >> 2) * 256;
      tmem_ld_32x32(a, r0); tmem_ld_32x32(a + 32, r1); tmem_ld_32x32(a + 64, r2); tmem_ld_32x32(a + 96, r3);
      tmem_ld_wait();
      acc += __uint_as_float(r0[0]) + __uint_as_float(r1[1]) + __uint_as_float(r2[2]) + __uint_as_float(r3[3]);
      tmem_ld_32x32(a + 128, r0); tmem_ld_32x32(a + 160, r1); tmem_ld_32x32(a + 192, r2); tmem_ld_32x32(a + 224, r3);
      tmem_ld_wait();
      acc += __uint_as_float(r0[0]) + __uint_as_float(r1[1]) + __uint_as_float(r2[2]) + __uint_as_float(r3[3]);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<1>(base, 512);
}

__global__ void __launch_bounds__(256, 1) mufu_kernel(int iters, int nwarps, long long* cycles, float* sink) {
  const int warp = threadIdx.x >> 5;
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  long long t0 = clock64();
  if (warp < nwarps) {
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = ex2_approx(x[j]) - 1.0f;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  if (s == 123.456f) sink[0] = s;
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * sizeof(long long)); cudaMalloc(&sink, 4);
  long long h[148];
  const int iters = 2000;
  for (int nw : {4, 8}) {
    tmem_read_kernel<<<148, 256>>>(iters, nw, cyc, sink);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = double(iters) * 8 * 32 * 32 * 4 * nw;
    printf("tmem ld (wait each) warps=%d: %lld cycles, %.1f B/cycle/SM\n", nw, h[0], bytes / h[0]);
    tmem_read_nowait_kernel<<<148, 256>>>(iters, nw, cyc, sink);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("tmem ld (4 in flight) warps=%d: %lld cycles, %.1f B/cycle/SM\n", nw, h[0], bytes / h[0]);
  }
  for (int nw : {4, 8}) {
    mufu_kernel<<<148, 256>>>(iters, nw, cyc, sink);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double n = double(iters) * 8 * 32 * nw;
    printf("ex2 warps=%d: %lld cycles, %.2f ex2/cycle/SM\n", nw, h[0], n / h[0]);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
