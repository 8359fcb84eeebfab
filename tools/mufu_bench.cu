// mufu_bench.cu — per-SM throughput of MUFU.EX2 (f32 / packed bf16 / packed f16) and of the exp2 inner step variants
// used by the attention softmax, 8 warps per SM (2 per sub-partition), 1 CTA per SM on every SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/mufu_bench tools/mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters, float sc, float ms) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = -0.01f * (threadIdx.x + j);
  uint32_t acc = 0; float facc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      if (MODE == 0) {          // FFMA + MUFU.EX2 f32 + FADD, pack (today's inner step)
        float p0, p1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(x[j], sc, -ms)));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(x[j + 1], sc, -ms)));
        facc += p0 + p1;
        uint32_t pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(p1), "f"(p0));
        acc ^= pk;
      } else if (MODE == 1) {   // bare MUFU.EX2 f32
        float p0, p1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(x[j]));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(x[j + 1]));
        facc += p0; acc ^= __float_as_uint(p1);
      } else if (MODE == 2) {   // FFMA x2 + pack + ex2.bf16x2 (no fp32 sum)
        uint32_t pk, r;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(fmaf(x[j + 1], sc, -ms)), "f"(fmaf(x[j], sc, -ms)));
        asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(pk));
        acc ^= r;
      } else if (MODE == 3) {   // f16x2
        uint32_t pk, r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(fmaf(x[j + 1], sc, -ms)), "f"(fmaf(x[j], sc, -ms)));
        asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(pk));
        acc ^= r;
      } else if (MODE == 4) {   // FFMA + MUFU f32 + pack, no FADD (row sum comes from the MMA)
        float p0, p1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(fmaf(x[j], sc, -ms)));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(fmaf(x[j + 1], sc, -ms)));
        uint32_t pk; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(p1), "f"(p0));
        acc ^= pk;
      } else if (MODE == 5) {   // 3-input max (FMNMX3?) throughput
        float m;
        asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(x[j]), "f"(x[j + 1]), "f"(facc));
        facc = m;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] += 1e-6f;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (facc == 123.f || acc == 77u) out[0] = facc + acc;
}

template <int MODE>
void run(const char* name, float* out, long long* cyc, int threads = 256) {
  const int iters = 2000;
  k<MODE><<<148, threads>>>(out, cyc, iters, 0.18f, 0.3f);
  long long h[148];
  cudaError_t e = cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  // elements per SM = 256 threads * iters * 16
  printf("%-48s %4d threads/SM  %.2f elements/clk/SM   (%s)\n", name, threads, (double)threads * iters * 16 / avg, cudaGetErrorString(e));
}

int main() {
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
  run<1>("bare ex2.approx.ftz.f32", out, cyc);
  run<0>("FFMA + ex2.f32 + FADD + cvt.bf16x2 (current)", out, cyc);
  run<4>("FFMA + ex2.f32 + cvt.bf16x2 (sum via MMA)", out, cyc);
  run<2>("FFMA + cvt.bf16x2 + ex2.bf16x2", out, cyc);
  run<3>("FFMA + cvt.f16x2 + ex2.f16x2", out, cyc);
  run<5>("max.f32 3-input", out, cyc);
  run<1>("bare ex2.approx.ftz.f32", out, cyc, 128);
  run<4>("FFMA + ex2.f32 + cvt.bf16x2 (sum via MMA)", out, cyc, 128);
  run<4>("FFMA + ex2.f32 + cvt.bf16x2 (sum via MMA)", out, cyc, 512);
  run<4>("FFMA + ex2.f32 + cvt.bf16x2 (sum via MMA)", out, cyc, 1024);
  return 0;
}
