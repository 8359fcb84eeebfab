// gemm.h — internal C++ interface shared by the kernels in csrc/ (not part of the C ABI).
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/bridgelang_b200.h"

namespace blb {

// status codes BLB_OK / BLB_ERR_* come from the public header

enum GemmEpilogueMode : int {
  EPI_BIAS = 0,       // out_bf16 = acc + bias
  EPI_BIAS_GELU = 1,  // out_bf16 = gelu_erf(acc + bias)
  EPI_RESIDUAL = 2,   // resid_f32 += gamma * (acc + bias); optional bf16 copy of the new residual (row-remapped)
  EPI_PATCH = 3,      // resid_f32[remap(row)] = acc + bias + pos[token]
  EPI_BIAS_QGELU = 4  // out_bf16 = quick_gelu(acc + bias) = x·σ(1.702 x)   (OpenAI CLIP towers, clip_vit.py:15-27)
};

struct GemmEpilogue {
  const float* bias = nullptr;    // [N]
  const float* gamma = nullptr;   // [N] LayerScale (nullptr → 1)
  float* resid = nullptr;         // fp32 residual stream
  int ld_resid = 0;
  __nv_bfloat16* out = nullptr;   // bf16 destination
  int ld_out = 0;
  int out_col_off = 0;
  const float* pos = nullptr;     // [tok_in, N] position embedding (EPI_PATCH)
  // row remap image-wise: src row = b*tok_in + t  →  dst row = b*tok_out + t + tok_shift (dropped if out of range)
  int tok_in = 0;                 // 0 → identity
  int tok_out = 0;
  int tok_shift = 0;
  // walk the tiles last-to-first: a kernel that starts where its producer finished finds that data still in the
  // 126 MB L2 (the schedule in capi.cu alternates directions along the producer → consumer chain)
  int reverse = 0;
  // ---- LayerNorm folded into the neighbouring GEMMs -----------------------------------------
  // consumer (EPI_BIAS / EPI_BIAS_GELU): A is the bf16 copy of the UN-normalised fp32 rows, W is the LN-weight-folded
  // bf16 matrix W' = W·diag(ln_w), bias is b + W·ln_b, and  out = rstd_m·(acc − mean_m·colsum_n) + bias_n
  // with (mean, rstd) of row m rebuilt from `ln_parts` partial (sum, sum of squares) pairs written by the producer.
  const float2* ln_stats = nullptr;   // [ln_parts, M]
  const float* ln_colsum = nullptr;   // [N]  Σ_k W'[n,k] (of the bf16-rounded values, in fp32)
  int ln_parts = 0;
  float ln_eps = 0.f;
  // producer (EPI_RESIDUAL): per-row partial statistics of the new fp32 rows, one pair per epilogue warp column
  // span → [gemm_stats_parts(N), M], and the bf16 copy of the new rows that the consumer uses as its A operand
  float2* stats_out = nullptr;
  __nv_bfloat16* xb_out = nullptr;
  int ld_xb = 0;
  // rolling per-row shift (makes the fold independent of the row mean): xb, the statistics and the
  // consumer's algebra live in "x − c_m" coordinates, c_m = the row mean one residual update ago.
  //   producer (EPI_RESIDUAL): c_m = shift_in[m]; stores xb = bf16(x_new − c_m) and the statistics of x_new − c_m
  //   consumer (folded BIAS / BIAS_GELU): unchanged algebra ((x − c) − mean(x − c) = x − mean(x)); it rebuilds
  //     mean(x − c) anyway, so it also hands the next producer its shift: shift_out[m] = shift_in[m] + mean(x − c)
  // nullptr: c = 0.  shift_in and shift_out must be different buffers (other CTAs still read shift_in).
  const float* shift_in = nullptr;
  float* shift_out = nullptr;
  // algorithmic FLOPs of this launch for the timing records when N or K carry zero padding (fc1 / fc2 of SigLIP:
  // Hm 4304 → 4352, patch embed: K 588 → 592); 0 → 2·M·N·K
  double alg_work = 0.0;
};

int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K, int mode,
              const GemmEpilogue& epi, cudaStream_t stream);
void gemm_set_cta_group(int ctas);
int gemm_stats_parts(int N);   // number of (sum, sumsq) pairs per row an EPI_RESIDUAL launch with this N writes
int make_tmap_bf16_2d(CUtensorMap* map, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int num_sms();
int current_device();   // cudaGetDevice, clamped to [0, BLB_MAX_DEVICES): index for per-device one-time setup flags
constexpr int BLB_MAX_DEVICES = 64;
bool pdl_enabled();   // launch kernels with cudaLaunchAttributeProgrammaticStreamSerialization (default on)

// <<<>>> replacement that adds the PDL attribute; the kernel must call pdl_wait() before touching global data
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Optional per-launch device timing (bench.py's roofline): when enabled, every kernel launch is bracketed by a
// pair of CUDA events on its own stream; timing_collect() sums them per category after the caller synchronised.
enum TimingCategory : int { TIME_GEMM = 0, TIME_ATTENTION = 1, TIME_LAYERNORM = 2, TIME_OTHER = 3, TIME_NCAT = 4 };
void timing_enable(int on);
bool timing_enabled();
void timing_begin(cudaStream_t s);
void timing_end(int cat, double work, cudaStream_t s, long long tag);
int timing_records(int max_records, int* cat, long long* tag, double* ms, double* work);
int timing_collect(int cat, double* ms, double* work, long long* launches);
void timing_reset();
struct TimingScope {
  int cat; double work; cudaStream_t s; bool on; long long tag;
  TimingScope(int c, double w, cudaStream_t st, long long tg = 0)
      : cat(c), work(w), s(st), on(timing_enabled()), tag(tg) { if (on) timing_begin(s); }
  ~TimingScope() { if (on) timing_end(cat, work, s, tag); }
};
long long launch_count();
void count_launch(int n);

// layernorm.cu — timm LayerNorm(eps) over the last dim of an fp32 [rows, D] stream → bf16
int layernorm_f32_bf16(const float* x, int ldx, const float* w, const float* b, __nv_bfloat16* y, int ldy, int rows,
                       int D, float eps, cudaStream_t stream, int reverse = 0);

int layernorm_f32_f32(const float* x, int ldx, const float* w, const float* b, float* y, int ldy, int rows, int D,
                      float eps, cudaStream_t stream);

// fp32 rows → bf16 copy + full-row (sum, sumsq) in part 0 of `parts` (others zero): primes the LN-folded chain after
// the patch-embed stage (the later blocks get both from the EPI_RESIDUAL epilogues)
int rowstats_cast_f32_bf16(const float* x, int ldx, __nv_bfloat16* y, int ldy, float2* stats, int parts, int rows,
                           int D, cudaStream_t stream, float* shift = nullptr);

// attention.cu — softmax(q kᵀ · hd^-0.5) v over packed qkv [B*T, 3*H*hd] → out [B*T, H*hd]
int attention_bf16(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int T, int H, int hd, cudaStream_t stream,
                   int reverse = 0);

// attention_tc.cu debug hook: CTA 0 writes clock64() stamps of pipeline events into this device buffer (or nullptr)
void attention_set_trace(long long* device_buffer);

// patch_embed.cu — im2col for Conv2d(3, D, 14, stride 14) and the cls/reg prefix rows
int im2col_patch14(const __nv_bfloat16* pixels, __nv_bfloat16* cols, int B, int ldk, cudaStream_t stream, int grid = 16,
                   int img = 0);
int u8_to_patches(const uint8_t* frames, __nv_bfloat16* cols, int B, int ldk, int grid, cudaStream_t stream, int img = 0);
// resize.cu — PIL-exact antialiased bicubic resize of uint8 HWC frames (two separable fixed-point passes)
int resize_u8(const uint8_t* src, int B, int Hs, int Ws, uint8_t* dst, int Hd, int Wd, const int* kx, const int* bx,
              int ksx, const int* ky, const int* by, int ksy, uint8_t* tmp, cudaStream_t stream);
int preprocess_u8(const uint8_t* frames, int B, const __nv_bfloat16* lut, __nv_bfloat16* out_dino,
                  __nv_bfloat16* out_siglip, cudaStream_t stream);
int write_prefix_tokens(const float* prefix, float* resid, int B, int T, int n_prefix, int D, cudaStream_t stream);

// decode_tail.cu
int argmax_rows(const void* logits, int dtype, int rows, int vocab, int64_t ld, int64_t* ids, cudaStream_t stream);
int argmax_rows_window(const void* logits, int dtype, int rows, int vocab, int64_t ld, int win_begin, int win_end,
                       int64_t* ids, cudaStream_t stream);
int detokenize_unnormalize(const int64_t* ids, int n, int vocab_size, const double* bin_centers, int n_centers,
                           int action_dim, const double* q01, const double* q99, const uint8_t* mask, double* norm_out,
                           double* act_out, cudaStream_t stream);
int argmax_detokenize_unnormalize(const void* logits, int dtype, int rows, int vocab, int64_t ld, int vocab_size,
                                  const double* bin_centers, int n_centers, int action_dim, const double* q01,
                                  const double* q99,
                                  const uint8_t* mask, int64_t* ids, double* norm_out, double* act_out,
                                  cudaStream_t stream);

int argmax_window_detokenize_unnormalize(const void* logits, int dtype, int rows, int vocab, int64_t ld, int win_begin,
                                         int win_end, int vocab_size, const double* bin_centers, int n_centers,
                                         int action_dim, const double* q01, const double* q99, const uint8_t* mask,
                                         int64_t* ids, double* norm_out, double* act_out, cudaStream_t stream);

int encode_actions(const void* actions, int dtype, int n, const double* bins, int n_bins, double lo, double hi,
                   int vocab_size, int64_t* ids, cudaStream_t stream);
int action_token_metrics(const void* logits, int dtype, int batch, int seq, int vocab, int64_t ld_row, int64_t ld_batch,
                         int num_patches, const int64_t* labels, int64_t ld_labels, int begin_idx, int vocab_size,
                         const double* bin_centers, int n_centers, int64_t* preds, double* absdiff, int64_t* counts,
                         double* l1_sum, cudaStream_t stream);

}  // namespace blb
