/* bridgelang_b200.h — C ABI of libbridgelang_b200.so: the B200-native (sm_100a) visual-prefix hot path of
 * OpenVLA / prismatic as found in CliffKai/BridgeLang.
 *
 * The reference has NO native layer: its boundary for this path is a set of Python classes
 * (SURVEY.md §8b).  Each entry point below names the reference interface whose device work it replaces;
 * the Python mirrors of those classes live in bridgelang_b200/*.py and bind these symbols with ctypes
 * (INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); nothing synchronises;
 *   - return value: 0 = ok, < 0 = bad argument (BLB_ERR_*), > 0 = cudaError_t of the failed runtime call;
 *   - never throws; callers own all memory (PyTorch tensors on the reference side);
 *   - bf16 activations/weights, fp32 biases / LayerNorm / LayerScale parameters and fp32 residual stream.
 */
#ifndef BRIDGELANG_B200_H_
#define BRIDGELANG_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLB_OK 0
#define BLB_ERR_ARG (-1)       /* null pointer / non-positive size */
#define BLB_ERR_SHAPE (-2)     /* shape unsupported by the tiling (N %% 128, head_dim not in {64,72}, ...) */
#define BLB_ERR_ALIGN (-3)     /* pointer or row pitch not 16-byte aligned */
#define BLB_ERR_DRIVER (-4)    /* cuTensorMapEncodeTiled unavailable / failed */
#define BLB_ERR_WORKSPACE (-5) /* workspace too small */

/* GEMM epilogue modes */
#define BLB_EPI_BIAS 0      /* out_bf16 = acc + bias                                   (Attention.qkv, projector fc3) */
#define BLB_EPI_BIAS_GELU 1 /* out_bf16 = gelu_erf(acc + bias)                         (Mlp.fc1+GELU, projector fc1/fc2) */
#define BLB_EPI_RESIDUAL 2  /* resid_f32 += gamma*(acc+bias) [+ bf16 copy, row-remapped] (Attention.proj / Mlp.fc2 + LayerScale + add) */
#define BLB_EPI_PATCH 3     /* resid_f32[remap(row)] = acc + bias + pos[token]          (PatchEmbed + _pos_embed) */
#define BLB_EPI_BIAS_QGELU 4 /* out_bf16 = quick_gelu(acc + bias)                       (Mlp.fc1 of the OpenAI CLIP towers) */

/* logits dtypes for the decode tail */
#define BLB_DTYPE_F32 0
#define BLB_DTYPE_BF16 1
#define BLB_DTYPE_F16 2

typedef struct blb_epilogue {
  const float* bias;   /* [N] or NULL */
  const float* gamma;  /* [N] LayerScale or NULL (= 1) */
  float* resid;        /* fp32 residual stream [M, ld_resid] */
  int32_t ld_resid;
  void* out;           /* bf16 destination */
  int32_t ld_out;
  int32_t out_col_off;
  const float* pos;    /* [tok_in, N] position embedding (BLB_EPI_PATCH) */
  int32_t tok_in;      /* row remap per image: src row b*tok_in+t -> dst row b*tok_out+t+tok_shift; 0 = identity */
  int32_t tok_out;
  int32_t tok_shift;
  /* LayerNorm folded into the neighbouring GEMMs (timm Block.norm1 -> attn.qkv, Block.norm2 -> mlp.fc1):
   * consumer (BIAS / BIAS_GELU), ln_stats != NULL: A is the bf16 copy of the UN-normalised fp32 rows, W holds
   * W*diag(ln_w), bias holds b + W*ln_b, and out = rstd_m*(acc - mean_m*ln_colsum_n) + bias_n, where mean/rstd of
   * row m come from ln_parts partial (sum, sum-of-squares) float pairs. */
  const float* ln_stats;   /* [ln_parts, M, 2] or NULL */
  const float* ln_colsum;  /* [N] sum_k W'[n,k] over the bf16-rounded folded weight */
  int32_t ln_parts;
  float ln_eps;
  /* producer (RESIDUAL): partial statistics [blb_gemm_stats_parts(N), M, 2] and bf16 copy [M, ld_xb] of the new rows */
  float* stats_out;        /* or NULL */
  void* xb_out;            /* or NULL */
  int32_t ld_xb;
  /* rolling per-row shift (ABI v3): xb, the statistics and the consumer's algebra live in "x - c_m" coordinates,
   * c_m = the row mean one residual update ago, so the fold's rounding error no longer grows with |row mean| / row std.
   *   producer (RESIDUAL): c_m = shift_in[m]; xb_out = bf16(x_new - c_m), stats_out = statistics of x_new - c_m;
   *   folded consumer: algebra unchanged ((x-c) - mean(x-c) = x - mean(x)); it also writes the next producer's shift,
   *   shift_out[m] = shift_in[m] + mean_m(x - c).  NULL: c = 0.  shift_in != shift_out (ping-pong buffers). */
  const float* shift_in;   /* [M] or NULL */
  float* shift_out;        /* [M] or NULL (folded consumer only) */
} blb_epilogue;

/* One timm `Block` (vision_transformer.Block as instantiated at dinosiglip_vit.py:50-58). */
typedef struct blb_block_weights {
  const float* ln1_w; const float* ln1_b;       /* norm1 [D] */
  const void* qkv_w;  const float* qkv_b;       /* attn.qkv  bf16 [3D, D], fp32 [3D] */
  const void* proj_w; const float* proj_b;      /* attn.proj bf16 [D, D],  fp32 [D] */
  const float* ls1;                             /* ls1.gamma / scale_factor [D] or NULL (SigLIP) */
  const float* ln2_w; const float* ln2_b;       /* norm2 [D] */
  const void* fc1_w;  const float* fc1_b;       /* mlp.fc1 bf16 [Hm_pad, D], fp32 [Hm_pad] (rows >= Hm zero) */
  const void* fc2_w;  const float* fc2_b;       /* mlp.fc2 bf16 [D, Hm_pad] (cols >= Hm zero), fp32 [D] */
  const float* ls2;                             /* ls2.gamma or NULL */
  /* only when blb_vit_weights.ln_folded: qkv_w = W*diag(ln1_w), qkv_b = b + W*ln1_b (same for fc1 with ln2), and
   * these hold sum_k of the bf16-rounded folded rows: fp32 [3D] / [Hm_pad].  ln1_ / ln2_ pointers are then unused. */
  const float* qkv_colsum;
  const float* fc1_colsum;
} blb_block_weights;

/* One timm VisionTransformer restricted to what get_intermediate_layers(n={depth-2}) needs. */
typedef struct blb_vit_weights {
  int32_t dim;        /* D: 1024 (DINOv2 ViT-L/14-reg4) / 1152 (SigLIP SO400M/14) */
  int32_t heads;      /* 16 */
  int32_t head_dim;   /* 64 / 72 */
  int32_t hidden_pad; /* MLP hidden padded to a multiple of 256: 4096 / 4352 (from 4304) */
  int32_t n_prefix;   /* 5 (cls + 4 reg) / 0 */
  int32_t n_blocks;   /* blocks to run = depth - 1 (23 / 26): output of block depth-2, final norm NOT applied */
  int32_t patch_ldk;  /* row pitch of patch_w: 592 (588 padded to 16 B) */
  float ln_eps;       /* 1e-6 */
  const void* patch_w;       /* patch_embed.proj.weight as bf16 [D, patch_ldk], k = c*196 + kh*14 + kw */
  const float* patch_b;      /* [D] or NULL (CLIP: the pre_norm conv has no bias) */
  const float* pos_embed;    /* [grid*grid, D] position embedding of the PATCH tokens (a checkpoint whose pos_embed also
                              * covers the class token, timm no_embed_class=False, has that row folded into `prefix`) */
  const float* prefix;       /* [n_prefix, D]: cls_token then reg_token rows, or NULL */
  const blb_block_weights* blocks_host; /* HOST array of n_blocks entries (device pointers inside) */
  int32_t ln_folded;  /* 1: no LayerNorm kernel - norm1/norm2 are folded into the qkv / fc1 GEMMs and their statistics come
                       * from the producing GEMM's epilogue; 0: explicit LayerNorm kernel before qkv / fc1 */
  /* ---- ABI v3 ---- */
  int32_t hidden;     /* unpadded MLP hidden (4096 / 4304): algorithmic FLOP accounting of the timing records; 0 = hidden_pad */
  int32_t grid;       /* patches per side: 16 (224 px, the default when 0), 24 (336 px), 27 (378 px of a 384 px frame);
                       * pixels are [batch,3,14*grid,14*grid], tokens = grid*grid + n_prefix, pos_embed is [grid*grid, D] */
  int32_t img_size;   /* image side in pixels (0 = 14*grid): 384 for the 384 px checkpoints, whose conv ignores the last 6 rows/cols */
  int32_t act;        /* MLP activation: 0 exact-erf GELU (nn.GELU), 1 quick-GELU x*sigmoid(1.702x) (clip_vit.py:15-27) */
  const float* norm_pre_w;   /* timm pre_norm=True (OpenAI CLIP): LayerNorm over the embedded tokens (prefix rows included) */
  const float* norm_pre_b;   /* before block 0; both NULL = identity */
  /* uint8 entry: patch_embed weight with ToTensor + this tower's Normalize folded in,
   * W'[n,k] = W[n,c,kh,kw] / (255*std_c) at k = kh*42 + kw*3 + c (HWC patch order), bf16 [D, patch_ldk], and
   * b'[n] = b[n] - sum_k W[n,c,kh,kw]*mean_c/std_c, fp32 [D].  NULL: the tower has no uint8 entry. */
  const void* patch_w_u8;
  const float* patch_b_u8;
} blb_vit_weights;

/* prismatic/util/nn_utils.py:37-53 FusedMLPProjector == extern/hf/modeling_prismatic.py:146-158 fc1/fc2/fc3 */
typedef struct blb_projector_weights {
  int32_t in_dim;      /* 2176 */
  int32_t hidden_dim;  /* 8704 = 4 * in_dim */
  int32_t out_dim;     /* 4096 (llm_dim) */
  const void* fc1_w; const float* fc1_b;  /* bf16 [hidden, in] */
  const void* fc2_w; const float* fc2_b;  /* bf16 [out, hidden] */
  const void* fc3_w; const float* fc3_b;  /* bf16 [out, out] */
} blb_projector_weights;

/* ---- library info ------------------------------------------------------------------------------------- */
int blb_abi_version(void);
const char* blb_status_string(int status);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
long long blb_launch_count(void);
/* 0 = auto, 1 = cta_group::1 tiles (128xBN), 2 = cta_group::2 CTA pairs (256xBN) */
void blb_set_gemm_cta_group(int ctas);

/* Debug hook: device buffer of >= 128 int64 into which CTA 0 of the tcgen05 attention kernel records clock64()
 * stamps of its pipeline events (tiles 8..15); NULL (default) disables it. */
void blb_debug_attention_trace(void* device_buffer);

/* Per-launch device timing for the roofline report (off by default).  When enabled every kernel launch is
 * bracketed by CUDA events on its stream; after the caller has synchronised, blb_timing_collect sums the elapsed
 * milliseconds, the algorithmic work (FLOPs for categories 0/1, bytes for 2/3) and the launch count of one
 * category: 0 = tcgen05 GEMM, 1 = attention, 2 = LayerNorm, 3 = other.  blb_timing_reset drops the records. */
void blb_timing_enable(int on);
void blb_timing_reset(void);
int blb_timing_collect(int category, double* ms, double* work, long long* launches);
/* every record since the last reset, in launch order (call after synchronising): returns the number written (<= max_records)
 * or a negated cudaError_t.  GEMM tag = mode << 44 | ln-folded consumer << 43 | stats/xb producer << 42 | N << 20 | K. */
int blb_timing_records(int max_records, int* category, long long* tag, double* ms, double* work);

/* ---- primitive operators (each replaces one library call of the reference; used by the parity tests) --- */
/* C = A[M,K] (bf16, pitch lda) x W[N,K]^T (bf16, pitch ldw) with a fused epilogue.
 * Replaces nn.Linear (+GELU / +LayerScale+residual) reached via timm Attention/Mlp and nn_utils.py:42-48. */
int blb_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, int mode,
                  const blb_epilogue* epi, void* stream);
/* partial-statistics pairs per row that a BLB_EPI_RESIDUAL launch with this N writes to stats_out (0: N unsupported) */
int blb_gemm_stats_parts(int N);
/* primes the folded chain: per row c = mean(x) -> shift[row] (shift == NULL: c = 0), bf16 copy of x - c, and the
 * (sum, sum of squares) of x - c in pair 0 of `parts` pairs per row (others zero) */
int blb_rowstats_cast(const float* x, int ldx, void* y_bf16, int ldy, float* stats, int parts, int rows, int D,
                      float* shift, void* stream);
/* timm LayerNorm(D, eps) on fp32 rows -> bf16 rows (Block.norm1 / norm2). */
int blb_layernorm(const float* x, int ldx, const float* w, const float* b, void* y_bf16, int ldy, int rows, int D,
                  float eps, void* stream);
/* timm Attention core: softmax(q k^T hd^-0.5) v on packed qkv bf16 [B*T, 3*H*hd] -> bf16 [B*T, H*hd]. */
int blb_attention(const void* qkv_bf16, void* out_bf16, int B, int T, int H, int head_dim, void* stream);
/* timm PatchEmbed staging: pixels bf16 [B,3,224,224] -> bf16 [B*256, ldk], k = c*196+kh*14+kw, zero padded. */
int blb_im2col_patch14(const void* pixels_bf16, void* cols_bf16, int B, int ldk, void* stream);
/* same for uint8 HWC frames [B,14*grid,14*grid,3]: -> bf16 [B*grid*grid, ldk] of exact integers 0..255, k = kh*42+kw*3+c
 * (stride == kernel: im2col is a permutation of the frame; pair with blb_vit_weights.patch_w_u8 / patch_b_u8). */
int blb_u8_to_patches(const uint8_t* frames_hwc, void* cols_bf16, int B, int ldk, int grid, int img_size, void* stream);
/* Antialiased resize of uint8 HWC frames [B,Hs,Ws,3] -> [B,Hd,Wd,3], bit-exact with PIL.Image.resize (Pillow's
 * fixed-point ImagingResample): the Resize of the reference's image transform (dinosiglip_vit.py:91-111,
 * processing_prismatic.py:128-145 via torchvision).  kx [Wd,ksx] / bx [Wd,2] and ky [Hd,ksy] / by [Hd,2]: int32 device
 * tables from bridgelang_b200/resize.py (Pillow's precompute_coeffs + normalize_coeffs_8bpc); tmp: Hs*Wd*3*B bytes. */
int blb_resize_u8(const uint8_t* src, int B, int Hs, int Ws, uint8_t* dst, int Hd, int Wd, const int32_t* kx,
                  const int32_t* bx, int ksx, const int32_t* ky, const int32_t* by, int ksy, uint8_t* tmp, void* stream);

/* ---- towers, projector, fused path ---------------------------------------------------------------------- */
size_t blb_vit_workspace_bytes(const blb_vit_weights* w, int batch);
/* timm VisionTransformer.get_intermediate_layers(n={depth-2}) + unpack_tuple (base_vision.py:27-32):
 * writes patch tokens (prefix dropped) as bf16 into out[b*256+p, out_col_off : out_col_off+D], pitch ld_out. */
int blb_vit_tower_forward(const blb_vit_weights* w, const void* pixels_bf16, int batch, void* out_bf16, int ld_out,
                          int out_col_off, void* workspace, size_t workspace_bytes, void* stream);

/* the same from a uint8 HWC frame [batch,14*grid,14*grid,3] (needs patch_w_u8 / patch_b_u8): the patch matrix is built in
 * the workspace (blb_vit_workspace_bytes covers it) by blb_u8_to_patches and TMA-loaded by the patch-embed GEMM. */
int blb_vit_tower_forward_u8(const blb_vit_weights* w, const uint8_t* frames_hwc, int batch, void* out_bf16, int ld_out,
                             int out_col_off, void* workspace, size_t workspace_bytes, void* stream);

size_t blb_projector_workspace_bytes(const blb_projector_weights* w, int rows);
/* FusedMLPProjector.forward (nn_utils.py:52-53): x bf16 [rows, in_dim] -> out bf16.
 * tok_in/tok_out/tok_shift (0,0,0 = plain [rows, out_dim]) let fc3 store straight into an
 * inputs_embeds buffer [B, tok_out, out_dim] at token offset tok_shift (prismatic.py:389-396 splice). */
int blb_projector_forward(const blb_projector_weights* w, const void* x_bf16, int ldx, int rows, void* out_bf16,
                          int ld_out, int tok_in, int tok_out, int tok_shift, void* workspace,
                          size_t workspace_bytes, void* stream);

size_t blb_fused_workspace_bytes(const blb_vit_weights* dino, const blb_vit_weights* siglip,
                                 const blb_projector_weights* proj, int batch);
/* DinoSigLIPViTBackbone.forward (dinosiglip_vit.py:142-147) [+ projector when proj != NULL].
 * pixels_*: bf16 [batch,3,224,224] per tower (the dict of dinosiglip_vit.py:39-40; HF packs them as
 * [batch,6,224,224], modeling_prismatic.py:120).  features_bf16 [batch*256, 2176] is always written;
 * projected_bf16 [batch*256, out_dim] when proj != NULL. */
int blb_fused_featurize_project_forward(const blb_vit_weights* dino, const blb_vit_weights* siglip,
                                        const blb_projector_weights* proj, const void* pixels_dino,
                                        const void* pixels_siglip, int batch, void* features_bf16,
                                        void* projected_bf16, void* workspace, size_t workspace_bytes,
                                        void* stream);

/* uint8 entry of the fused path: ONE frame [batch,224,224,3] feeds both towers (their Normalize is folded into
 * patch_w_u8 / patch_b_u8); workspace: blb_fused_workspace_bytes(...) + blb_patch_matrix_bytes(dino, batch). */
size_t blb_patch_matrix_bytes(const blb_vit_weights* w, int batch);
int blb_fused_featurize_project_forward_u8(const blb_vit_weights* dino, const blb_vit_weights* siglip,
                                           const blb_projector_weights* proj, const uint8_t* frames_hwc, int batch,
                                           void* features_bf16, void* projected_bf16, void* workspace,
                                           size_t workspace_bytes, void* stream);

/* ---- rows SURVEY.md section 8f marks "next": the data formats either side of the path ------------------------------ */
/* Image preprocessing on the device: one uint8 HWC frame [batch,224,224,3] -> both towers' normalized bf16
 * [batch,3,224,224] tensors.  Replaces ToTensor + Normalize of DinoSigLIPImageTransform / PrismaticImageProcessor
 * (dinosiglip_vit.py:33-40, processing_prismatic.py:128-145) and the .to(bf16) of the callers; lut_bf16 is
 * [2 towers (dino, siglip)][3 channels][256] = the reference transform evaluated on every uint8 value (host-built,
 * hence bit-identical).  Resizing stays on the host (PIL bicubic, upstream of the path). */
int blb_preprocess_u8(const uint8_t* frames_hwc, int batch, const void* lut_bf16, void* out_dino_bf16,
                      void* out_siglip_bf16, void* stream);
/* ActionTokenizer.__call__ up to the token ids (action_tokenizer.py:38-47): clip, np.digitize(action, bins),
 * vocab_size - index.  actions: BLB_DTYPE_F32 or BLB_DTYPE_F64 [n]; bins: float64 [n_bins] increasing. */
#define BLB_DTYPE_F64 3
int blb_encode_actions(const void* actions, int dtype, int n, const double* bins, int n_bins, double min_action,
                       double max_action, int vocab_size, int64_t* ids, void* stream);
/* Training-side action metrics (training/strategies/base_strategy.py:314-329, vla-scripts/finetune.py:270-286):
 * preds = logits[:, num_patches:-1].argmax(2); gt = labels[:, 1:]; mask = gt > action_token_begin_idx.
 * logits [batch, seq, vocab] with strides (ld_batch, ld_row, 1); labels int64 [batch, >= seq - num_patches], pitch ld_labels.
 * Outputs: preds int64 [batch*(seq-1-num_patches)] (-1 where mask is false: those rows are not reduced at all),
 * absdiff float64 same shape, counts[0] = #correct&mask, counts[1] = #mask, l1_sum = sum |decode(pred) - decode(gt)|. */
int blb_action_token_metrics(const void* logits, int dtype, int batch, int seq, int vocab, int64_t ld_row,
                             int64_t ld_batch, int num_patches, const int64_t* labels, int64_t ld_labels,
                             int action_token_begin_idx, int vocab_size, const double* bin_centers, int n_centers,
                             int64_t* preds, double* absdiff, int64_t* counts, double* l1_sum, void* stream);

/* ---- decode tail ---------------------------------------------------------------------------------------- */
/* torch.argmax over each full logits row (first max wins; NaN maximal) -> int64 ids. */
int blb_argmax(const void* logits, int dtype, int rows, int vocab, int64_t ld, int64_t* ids, void* stream);
/* Windowed mode: torch.argmax(logits[:, win_begin:win_end], -1) + win_begin, e.g. [31744, 32000) = the 256 action bins of
 * the 32000-entry tokenizer vocabulary (NOT the last 256 logits rows: 32000..32063 are <PAD> + padding, llama2.py:74-76).
 * This is NOT what the reference's greedy step computes (it arg-maxes the full row, openvla.py:81-86) and equals
 * blb_argmax only when the full-row winner lies inside the window - an explicit, opt-in mode for constrained decoding. */
int blb_argmax_window(const void* logits, int dtype, int rows, int vocab, int64_t ld, int win_begin, int win_end,
                      int64_t* ids, void* stream);
/* the windowed argmax feeding the bin-centre lookup and the q01/q99 un-normalize in the same launch (one warp per row,
 * shuffle reduction over the <= 1024 window columns). */
int blb_argmax_window_detokenize_unnormalize(const void* logits, int dtype, int rows, int vocab, int64_t ld,
                                             int win_begin, int win_end, int vocab_size, const double* bin_centers,
                                             int n_centers, int action_dim, const double* q01, const double* q99,
                                             const uint8_t* mask, int64_t* ids, double* normalized_out,
                                             double* actions_out, void* stream);
/* ActionTokenizer.decode_token_ids_to_actions (action_tokenizer.py:65-68) + un-normalize (openvla.py:94-101).
 * stats index = j %% action_dim; q01/q99/mask may be NULL (then actions == normalized). float64 bit-exact. */
int blb_detokenize_unnormalize(const int64_t* ids, int n, int vocab_size, const double* bin_centers, int n_centers,
                               int action_dim, const double* q01, const double* q99, const uint8_t* mask,
                               double* normalized_out, double* actions_out, void* stream);
/* both of the above in one launch: row r of logits is decode step r. */
int blb_argmax_detokenize_unnormalize(const void* logits, int dtype, int rows, int vocab, int64_t ld, int vocab_size,
                                      const double* bin_centers, int n_centers, int action_dim, const double* q01,
                                      const double* q99, const uint8_t* mask, int64_t* ids, double* normalized_out,
                                      double* actions_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BRIDGELANG_B200_H_ */
