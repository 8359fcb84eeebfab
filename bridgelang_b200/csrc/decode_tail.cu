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
                                                      int win_begin, int64_t* __restrict__ ids, int vocab_size,
                                                      const double* bin_centers, int n_centers, int action_dim,
                                                      const double* q01, const double* q99, const uint8_t* mask,
                                                      double* norm_out, double* act_out) {
  const int r = blockIdx.x;
  // `vocab` columns starting at win_begin (full-row mode: win_begin = 0, vocab = row length)
  const int best = block_argmax<T>(logits + static_cast<size_t>(r) * ld + win_begin, vocab) + win_begin;
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

// ---- ActionTokenizer.__call__, id part (action_tokenizer.py:38-47): clip → np.digitize(action, bins) → vocab_size − idx.
// np.digitize with increasing bins returns #{bins <= x} (right=False); NaN sorts after every bin (→ n_bins), as np.clip
// propagates NaN and NumPy orders NaN last.  float32 inputs are widened exactly, as NumPy's comparison does.
template <typename T>
__global__ void encode_kernel(const T* __restrict__ actions, int n, const double* __restrict__ bins, int n_bins,
                              double lo, double hi, int vocab_size, int64_t* __restrict__ ids) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  T a = actions[j];
  if (a < static_cast<T>(lo)) a = static_cast<T>(lo);   // np.clip in the array's own dtype
  if (a > static_cast<T>(hi)) a = static_cast<T>(hi);
  const double x = static_cast<double>(a);
  int idx;
  if (x != x) {
    idx = n_bins;
  } else {
    int l = 0, r = n_bins;                              // first index with bins[i] > x
    while (l < r) {
      const int mid = (l + r) >> 1;
      if (bins[mid] <= x) l = mid + 1; else r = mid;
    }
    idx = l;
  }
  ids[j] = static_cast<int64_t>(vocab_size) - idx;
}

int encode_actions(const void* actions, int dtype, int n, const double* bins, int n_bins, double lo, double hi,
                   int vocab_size, int64_t* ids, cudaStream_t stream) {
  if (actions == nullptr || bins == nullptr || ids == nullptr || n <= 0 || n_bins <= 0) return BLB_ERR_ARG;
  const int blocks = (n + 127) / 128;
  if (dtype == 0) encode_kernel<float><<<blocks, 128, 0, stream>>>(static_cast<const float*>(actions), n, bins, n_bins, lo, hi, vocab_size, ids);
  else if (dtype == 3) encode_kernel<double><<<blocks, 128, 0, stream>>>(static_cast<const double*>(actions), n, bins, n_bins, lo, hi, vocab_size, ids);
  else return BLB_ERR_ARG;
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

// ---- training-side action metrics (base_strategy.py:314-329, finetune.py:270-286) ---------------------------------
//   preds = logits[:, P:-1].argmax(2);  gt = labels[:, 1:];  mask = gt > action_token_begin_idx
//   accuracy = sum((preds == gt) & mask) / sum(mask);  l1 = mean |decode(preds[mask]) − decode(gt[mask])|
// One block per (sample, position).  Positions whose label is not an action token never enter either metric, so
// their 32064-wide argmax is skipped (pred = −1): with 7 action tokens per sample that is most of the rows.  The
// per-position results are reduced by one block in a fixed order (deterministic; float64 sum).
template <typename T>
__global__ void __launch_bounds__(1024) action_metrics_rows_kernel(const T* __restrict__ logits, int vocab, long long ld_row,
                                                                   long long ld_batch, int first_pos, int n_pos,
                                                                   const int64_t* __restrict__ labels, long long ld_labels,
                                                                   int begin_idx, int vocab_size,
                                                                   const double* __restrict__ bin_centers, int n_centers,
                                                                   int64_t* __restrict__ preds, double* __restrict__ absdiff) {
  const int b = blockIdx.x / n_pos, p = blockIdx.x - b * n_pos;
  const long long gt = labels[static_cast<size_t>(b) * ld_labels + 1 + p];
  if (gt <= begin_idx) {                       // block-uniform
    if (threadIdx.x == 0) { preds[blockIdx.x] = -1; absdiff[blockIdx.x] = 0.0; }
    return;
  }
  const int best = block_argmax<T>(logits + static_cast<size_t>(b) * ld_batch + static_cast<size_t>(first_pos + p) * ld_row, vocab);
  if (threadIdx.x == 0) {
    preds[blockIdx.x] = best;
    double a = 0.0, g = 0.0;
    detok_one(best, 0, vocab_size, bin_centers, n_centers, 0, nullptr, nullptr, nullptr, &a, nullptr);
    detok_one(gt, 0, vocab_size, bin_centers, n_centers, 0, nullptr, nullptr, nullptr, &g, nullptr);
    absdiff[blockIdx.x] = fabs(__dsub_rn(a, g));
  }
}

__global__ void __launch_bounds__(1024) action_metrics_reduce_kernel(const int64_t* __restrict__ preds,
                                                                     const double* __restrict__ absdiff,
                                                                     const int64_t* __restrict__ labels, long long ld_labels,
                                                                     int batch, int n_pos, int begin_idx,
                                                                     int64_t* __restrict__ counts, double* __restrict__ l1_sum) {
  __shared__ long long s_c[1024], s_m[1024];
  __shared__ double s_l[1024];
  long long c = 0, m = 0;
  double l = 0.0;
  const int total = batch * n_pos;
  for (int i = threadIdx.x; i < total; i += 1024) {      // fixed assignment + fixed tree → deterministic
    const int b = i / n_pos, p = i - b * n_pos;
    const long long gt = labels[static_cast<size_t>(b) * ld_labels + 1 + p];
    if (gt > begin_idx) {
      ++m;
      c += preds[i] == gt;
      l += absdiff[i];
    }
  }
  s_c[threadIdx.x] = c; s_m[threadIdx.x] = m; s_l[threadIdx.x] = l;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_c[threadIdx.x] += s_c[threadIdx.x + o];
      s_m[threadIdx.x] += s_m[threadIdx.x + o];
      s_l[threadIdx.x] += s_l[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { counts[0] = s_c[0]; counts[1] = s_m[0]; *l1_sum = s_l[0]; }
}

int action_token_metrics(const void* logits, int dtype, int batch, int seq, int vocab, int64_t ld_row, int64_t ld_batch,
                         int num_patches, const int64_t* labels, int64_t ld_labels, int begin_idx, int vocab_size,
                         const double* bin_centers, int n_centers, int64_t* preds, double* absdiff, int64_t* counts,
                         double* l1_sum, cudaStream_t stream) {
  const int n_pos = seq - 1 - num_patches;     // logits[:, P:-1]
  if (logits == nullptr || labels == nullptr || bin_centers == nullptr || preds == nullptr || absdiff == nullptr ||
      counts == nullptr || l1_sum == nullptr || batch <= 0 || vocab <= 0 || n_pos <= 0 || n_centers <= 0)
    return BLB_ERR_ARG;
  const int blocks = batch * n_pos;
#define BLB_AM(T) action_metrics_rows_kernel<T><<<blocks, 1024, 0, stream>>>(static_cast<const T*>(logits), vocab, ld_row, \
      ld_batch, num_patches, n_pos, labels, ld_labels, begin_idx, vocab_size, bin_centers, n_centers, preds, absdiff)
  switch (dtype) {
    case 0: BLB_AM(float); break;
    case 1: BLB_AM(__nv_bfloat16); break;
    case 2: BLB_AM(__half); break;
    default: return BLB_ERR_ARG;
  }
#undef BLB_AM
  action_metrics_reduce_kernel<<<1, 1024, 0, stream>>>(preds, absdiff, labels, ld_labels, batch, n_pos, begin_idx, counts,
                                                       l1_sum);
  count_launch(2);
  return static_cast<int>(cudaGetLastError());
}

// full-row mode: [win_begin, win_end) = [0, vocab).  A window of <= 1024 columns is reduced by ONE warp with shuffles
// (256 action bins: 8 elements per lane), wider ranges by a 1024-thread block (per-warp shuffles + one smem stage).
int argmax_window_detokenize_unnormalize(const void* logits, int dtype, int rows, int vocab, int64_t ld, int win_begin,
                                         int win_end, int vocab_size, const double* bin_centers, int n_centers,
                                         int action_dim, const double* q01, const double* q99, const uint8_t* mask,
                                         int64_t* ids, double* norm_out, double* act_out, cudaStream_t stream) {
  if (logits == nullptr || ids == nullptr || rows <= 0 || vocab <= 0 || ld < vocab) return BLB_ERR_ARG;
  if (win_begin < 0 || win_end > vocab || win_begin >= win_end) return BLB_ERR_ARG;
  if (bin_centers != nullptr && n_centers <= 0) return BLB_ERR_ARG;
  const int width = win_end - win_begin;
  const int threads = width <= 1024 ? 32 : 1024;
  switch (dtype) {
    case 0:
      argmax_kernel<float><<<rows, threads, 0, stream>>>(static_cast<const float*>(logits), width, ld, win_begin, ids,
                                                         vocab_size, bin_centers, n_centers, action_dim, q01, q99,
                                                         mask, norm_out, act_out);
      break;
    case 1:
      argmax_kernel<__nv_bfloat16><<<rows, threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), width,
                                                                 ld, win_begin, ids, vocab_size, bin_centers, n_centers,
                                                                 action_dim, q01, q99, mask, norm_out, act_out);
      break;
    case 2:
      argmax_kernel<__half><<<rows, threads, 0, stream>>>(static_cast<const __half*>(logits), width, ld, win_begin, ids,
                                                          vocab_size, bin_centers, n_centers, action_dim, q01, q99,
                                                          mask, norm_out, act_out);
      break;
    default:
      return BLB_ERR_ARG;
  }
  count_launch(1);
  return static_cast<int>(cudaGetLastError());
}

int argmax_detokenize_unnormalize(const void* logits, int dtype, int rows, int vocab, int64_t ld, int vocab_size,
                                  const double* bin_centers, int n_centers, int action_dim, const double* q01,
                                  const double* q99,
                                  const uint8_t* mask, int64_t* ids, double* norm_out, double* act_out,
                                  cudaStream_t stream) {
  return argmax_window_detokenize_unnormalize(logits, dtype, rows, vocab, ld, 0, vocab, vocab_size, bin_centers,
                                              n_centers, action_dim, q01, q99, mask, ids, norm_out, act_out, stream);
}

int argmax_rows_window(const void* logits, int dtype, int rows, int vocab, int64_t ld, int win_begin, int win_end,
                       int64_t* ids, cudaStream_t stream) {
  return argmax_window_detokenize_unnormalize(logits, dtype, rows, vocab, ld, win_begin, win_end, 0, nullptr, 0, 0,
                                              nullptr, nullptr, nullptr, ids, nullptr, nullptr, stream);
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
