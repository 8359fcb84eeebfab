// decode_tail.cu — predict_action's decode tail on the device.
//
//   argmax        HF GenerationMixin greedy step, torch.argmax(logits[:, -1, :]) over the FULL vocab row
//                 (32064 rows for Llama-2 + pad), first max index wins, NaN counts as maximal
//                 (called from prismatic/models/vlas/openvla.py:81-86, extern/hf/modeling_prismatic.py:518)
//   de-tokenize   ActionTokenizer.decode_token_ids_to_actions, prismatic/vla/action_tokenizer.py:65-68:
//                 d = vocab_size - id;  idx = clip(d - 1, 0, len(bin_centers) - 1);  bin_centers[idx]
//   un-normalize  openvla.py:94-101 / modeling_prismatic.py:527-534:
//                 where(mask, 0.5*(n+1)*(q99-q01)+q01, n)  — float64, every op rounded separately like NumPy
//                 (__dadd_rn/__dmul_rn forbid FMA contraction), so results are bit-identical.
#include <cuda_fp16.h>

#include "gemm.h"
#include "ptx.cuh"

namespace blb {

struct Best {
  float v;
  int i;
};

// true if a beats b under torch.argmax semantics (NaN is maximal, ties → lowest index)
__device__ __forceinline__ bool beats(const Best& a, const Best& b) {
  if (b.i < 0) return a.i >= 0;
  if (a.i < 0) return false;
  const bool an = a.v != a.v, bn = b.v != b.v;
  if (an || bn) return (an && !bn) || (an && bn && a.i < b.i);
  return a.v > b.v || (a.v == b.v && a.i < b.i);
}

template <typename T>
__device__ __forceinline__ float to_f32(T x);
template <>
__device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }

template <typename T>
__device__ int block_argmax(const T* __restrict__ row, int vocab) {
  Best best{0.f, -1};
  for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
    Best c{to_f32<T>(row[i]), i};
    if (beats(c, best)) best = c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Best other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
    if (beats(other, best)) best = other;
  }
  __shared__ Best wbest[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) wbest[warp] = best;
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    best = lane < nw ? wbest[lane] : Best{0.f, -1};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Best other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      if (beats(other, best)) best = other;
    }
  }
  return best.i;   // valid in warp 0
}

__device__ __forceinline__ void detok_one(long long id, int j, int vocab_size, const double* bin_centers,
                                          int n_centers, int action_dim, const double* q01, const double* q99,
                                          const uint8_t* mask, double* norm_out, double* act_out) {
  const int sj = action_dim > 0 ? j % action_dim : j;   // which action dimension's statistics apply
  long long d = static_cast<long long>(vocab_size) - id - 1;
  if (d < 0) d = 0;
  if (d > n_centers - 1) d = n_centers - 1;
  const double n = bin_centers[d];
  if (norm_out != nullptr) norm_out[j] = n;
  if (act_out != nullptr) {
    double a = n;
    const bool m = mask == nullptr ? true : (mask[sj] != 0);
    if (m && q01 != nullptr && q99 != nullptr) {
      const double lo = q01[sj], hi = q99[sj];
      const double t0 = __dmul_rn(0.5, __dadd_rn(n, 1.0));
      const double t1 = __dmul_rn(t0, __dsub_rn(hi, lo));
      a = __dadd_rn(t1, lo);
    }
    act_out[j] = a;
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) argmax_kernel(const T* __restrict__ logits, int vocab, long long ld,
                                                      int64_t* __restrict__ ids, int vocab_size,
                                                      const double* bin_centers, int n_centers, int action_dim,
                                                      const double* q01, const double* q99, const uint8_t* mask,
                                                      double* norm_out, double* act_out) {
  const int r = blockIdx.x;
  const int best = block_argmax<T>(logits + static_cast<size_t>(r) * ld, vocab);
  if (threadIdx.x == 0) {
    ids[r] = best;
    if (bin_centers != nullptr)
      detok_one(best, r, vocab_size, bin_centers, n_centers, action_dim, q01, q99, mask, norm_out, act_out);
  }
}

__global__ void detok_kernel(const int64_t* __restrict__ ids, int n, int vocab_size, const double* bin_centers,
                             int n_centers, int action_dim, const double* q01, const double* q99,
                             const uint8_t* mask, double* norm_out, double* act_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) detok_one(ids[j], j, vocab_size, bin_centers, n_centers, action_dim, q01, q99, mask, norm_out, act_out);
}

int argmax_detokenize_unnormalize(const void* logits, int dtype, int rows, int vocab, int64_t ld, int vocab_size,
                                  const double* bin_centers, int n_centers, int action_dim, const double* q01,
                                  const double* q99,
                                  const uint8_t* mask, int64_t* ids, double* norm_out, double* act_out,
                                  cudaStream_t stream) {
  if (logits == nullptr || ids == nullptr || rows <= 0 || vocab <= 0 || ld < vocab) return BLB_ERR_ARG;
  if (bin_centers != nullptr && n_centers <= 0) return BLB_ERR_ARG;
  const int threads = 1024;
  switch (dtype) {
    case 0:
      argmax_kernel<float><<<rows, threads, 0, stream>>>(static_cast<const float*>(logits), vocab, ld, ids,
                                                         vocab_size, bin_centers, n_centers, action_dim, q01, q99,
                                                         mask, norm_out, act_out);
      break;
    case 1:
      argmax_kernel<__nv_bfloat16><<<rows, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), vocab,
                                                                 ld, ids, vocab_size, bin_centers, n_centers,
                                                                 action_dim, q01, q99, mask, norm_out, act_out);
      break;
    case 2:
      argmax_kernel<__half><<<rows, threads, 0, stream>>>(static_cast<const __half*>(logits), vocab, ld, ids,
                                                          vocab_size, bin_centers, n_centers, action_dim, q01, q99,
                                                          mask, norm_out, act_out);
      break;
    default:
      return BLB_ERR_ARG;
  }
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

int argmax_rows(const void* logits, int dtype, int rows, int vocab, int64_t ld, int64_t* ids, cudaStream_t stream) {
  return argmax_detokenize_unnormalize(logits, dtype, rows, vocab, ld, 0, nullptr, 0, 0, nullptr, nullptr, nullptr, ids,
                                       nullptr, nullptr, stream);
}

int detokenize_unnormalize(const int64_t* ids, int n, int vocab_size, const double* bin_centers, int n_centers,
                           int action_dim, const double* q01, const double* q99, const uint8_t* mask, double* norm_out,
                           double* act_out, cudaStream_t stream) {
  if (ids == nullptr || bin_centers == nullptr || n <= 0 || n_centers <= 0) return BLB_ERR_ARG;
  detok_kernel<<<(n + 127) / 128, 128, 0, stream>>>(ids, n, vocab_size, bin_centers, n_centers, action_dim, q01, q99, mask,
                                                    norm_out, act_out);
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

}  // namespace blb
