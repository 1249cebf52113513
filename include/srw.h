/* semireward_b200 — C ABI of the B200-native SemiReward train-step hot path.
 *
 * One shared library (libsrw_b200.so, built by __graft_entry__.build()) loaded by the Python plugin through ctypes.
 * The reference (Westlake-AI/SemiReward) is pure Python/PyTorch and has no FFI of its own ("no native
 * code"), so every entry point below replaces a *PyTorch library call site* of the reference; each declaration cites
 * the reference file:line whose arithmetic it takes over.  Conventions:
 *   - extern "C", plain pointers and sizes, no torch types;
 *   - device pointers are BORROWED from the caller (PyTorch tensors); nothing is allocated inside except through the
 *     caller-provided workspaces;
 *   - every call is asynchronous on the cudaStream_t passed as void* (0 = legacy default stream), no hidden syncs;
 *   - returns 0 on success, a negative SRW_ERR_* code otherwise; srw_last_error() gives the message; never throws;
 *   - one process per GPU, single caller thread per device.
 *
 * "planes" = split-bf16 operand format: an fp32 matrix X[rows, ld] stored as two bf16 planes (hi, lo),
 * hi = bf16(X), lo = bf16(X - hi), lo plane `plane_stride` elements after the hi plane (see csrc/srw_common.cuh).
 */
#ifndef SRW_H_
#define SRW_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRW_OK 0
#define SRW_ERR_CUDA (-1)
#define SRW_ERR_ARG (-2)
#define SRW_ERR_UNSUPPORTED (-3)
#define SRW_ERR_DRIVER (-4)

/* ---- library ---------------------------------------------------------------------------------------------- */
int srw_version(void);                 /* ABI version, bumped on any struct change */
const char* srw_last_error(void);      /* message of the last failing call on this thread */
int srw_device_check(int* sm_major, int* sm_minor, int* sm_count); /* fails unless the current device is sm_100 */
int64_t srw_kernel_launches(void);     /* number of kernels this library has launched so far (bench.py gpu_launches) */

/* Per-kernel-class device timing for bench.py's roofline block: when enabled, every launch of the classes below is
 * bracketed by CUDA events on its own stream; srw_profile_collect synchronises, sums and resets.  Off by default
 * (the bench's headline timing runs with it off). */
enum srw_profile_class { SRW_PROF_GEMM = 0, SRW_PROF_ATTN_FWD = 1, SRW_PROF_ATTN_BWD = 2, SRW_PROF_ADAMW = 3, SRW_PROF_NUM = 4 };
typedef struct { int64_t launches; double total_ms; double flops; double bytes; } srw_profile_stats;
int srw_profile_enable(int on);
int srw_profile_collect(srw_profile_stats* out /* [SRW_PROF_NUM] */);

/* ---- counter-based dropout ----------------------------------------------------------------------------------- */
/* nn.Dropout of the text / audio encoders (HF BertModel: embeddings, attention probabilities, attention output, FFN output;
 * bert.py:14 pooled features).  torch draws its masks from a Philox stream no other implementation can reproduce, so the
 * native path (and the oracle) define the mask as a pure function: element idx of site `site` of a sequence is kept iff
 *   (lowbias32(idx + lowbias32(seq_key + site * 0x9E3779B9)) >> 8) < (1 - p) * 2^24        (csrc/srw_common.cuh)
 * kept values are scaled by 1 / (1 - p).  seq_key[s] identifies the sequence's stream (one key per backbone call of the
 * reference: the three calls of `use_cat: False`, the K sampling passes), seq_row[s] is the sequence's row inside that call
 * (element indices are those of the call's own tensor, e.g. ((row * H + h) * L + q) * L + k for attention probabilities,
 * (row * L + l) * D + d for hidden states).  p == 0 or seq_key == NULL: no dropout. */
typedef struct {
  const uint32_t* seq_key;   /* device [num sequences] */
  const int32_t* seq_row;    /* device [num sequences] */
  uint32_t site;
  double p;
} srw_dropout;

/* ---- split-plane conversion -------------------------------------------------------------------------------- */
/* planes[r, c] = split(x[r, c] * (row_scale ? row_scale[r / rows_per_scale] : 1)).  Used for weights once per optimizer
 * step and for gradient tensors entering a GEMM.  transposed != 0 additionally writes planes_t[c, r] (ld = rows). */
typedef struct {
  const float* x; int64_t ldx;
  int rows, cols;
  const float* row_scale; int rows_per_scale;
  void* planes; int64_t ldp; int64_t plane_stride;
  void* planes_t; int64_t ldpt; int64_t plane_stride_t;
  float* colsum_out; int colsum_accumulate;   /* optional: out[c] (+)= sum_r scaled x[r, c] in the same pass (bias gradient) */
  float* colsum_workspace;                    /* >= 64 * cols floats when colsum_out != NULL */
} srw_split_args;
int srw_split_planes(const srw_split_args* a, void* stream);

/* ---- linear layers: torch.nn.functional.linear and its two backward GEMMs ----------------------------------- */
/* D[M,N] = A[M,K] * B[N,K]^T on tcgen05 tensor cores (bf16x3, fp32 accumulate in TMEM), operands staged by TMA.
 * Replaces F.linear in vit.py:93 (qkv), :105 (proj), :70/:73 (fc1/fc2) and autograd's dgrad / wgrad for them
 * (param_update.py:33 loss.backward()).
 *   a_mn_major = 0: A is stored [M, lda] (K contiguous);  1: A is stored [K, lda] (M contiguous)  (wgrad: dOut^T)
 *   b_mn_major = 0: B is stored [N, ldb] (K contiguous);  1: B is stored [K, ldb] (N contiguous)  (wgrad: activations)
 */
enum srw_epilogue {
  SRW_EPI_F32 = 0,          /* out_f32 = acc (+ bias) */
  SRW_EPI_PLANES = 1,       /* out_planes = split(acc + bias) */
  SRW_EPI_GELU = 2,         /* out_f32 = z = acc + bias ; out_planes = split(gelu(z))   (vit.py:70-71) */
  SRW_EPI_RESID = 3,        /* out_f32 = resid + row_scale[row / rows_per_scale] * (acc + bias)  (vit.py:164-165 + DropPath) */
  SRW_EPI_DGELU = 4,        /* out_planes = split(acc * gelu'(aux))  (backward of vit.py:71) */
  SRW_EPI_SPLITK = 5        /* workspace[split, M, N] = partial acc (reduced by srw_splitk_reduce) */
};
enum srw_gemm_impl { SRW_GEMM_TCGEN05 = 0 /* 2-CTA pairs where they fit, else 1-CTA */, SRW_GEMM_SIMT = 1, SRW_GEMM_TCGEN05_1CTA = 2 };

typedef struct {
  int M, N, K;
  const void* a; int64_t lda; int64_t a_plane_stride; int a_mn_major;
  const void* b; int64_t ldb; int64_t b_plane_stride; int b_mn_major;
  int epilogue;
  const float* bias;                          /* [N] or NULL */
  const float* resid; int64_t ldr;            /* SRW_EPI_RESID */
  const float* row_scale; int rows_per_scale; /* SRW_EPI_RESID; NULL = 1.0 */
  const float* aux; int64_t ldaux;            /* SRW_EPI_DGELU: pre-activation z */
  float* out_f32; int64_t ldo;
  void* out_planes; int64_t ldp; int64_t out_plane_stride;
  int split_k; float* workspace;              /* SRW_EPI_SPLITK: split_k >= 1 slices of K, workspace [split_k, M, N] */
  int impl;                                   /* srw_gemm_impl; SIMT is the on-device verification twin */
  int max_ctas;                               /* 0 = the whole GPU; > 0: persistent grid of at most this many CTAs (the GEMM
                                                 shares the GPU with kernels on other streams, e.g. the four wgrads of a block) */
  /* SRW_EPI_RESID only: out = resid + dropout(acc + bias) (BertSelfOutput / BertOutput: dropout between the dense layer and the
   * residual add, HF modeling_bert.py:287-298, 330-356).  Rows are tokens: sequence = row / drop_rows_per_seq, element index
   * (seq_row * drop_rows_per_seq + row % drop_rows_per_seq) * N + col. */
  srw_dropout drop; int drop_rows_per_seq;
  /* K segments (K-major A only; 0 = off): the reduction dimension is nseg segments of a_seg_k valid elements each, padded to
   * whole 64-element blocks (K = nseg * ceil(a_seg_k / 64) * 64; B holds zeros in the padding); segment s of output row m reads
   * A row m + s * a_seg_rows.  With lda < a_seg_k (overlapping rows) this is a 3x3 convolution over a zero-bordered NHWC tensor
   * as ONE GEMM: row = pixel, segment = kernel row, a_seg_k = 3 * C_in contiguous elements, a_seg_rows = padded image width
   * (wrn.py:35,39 nn.Conv2d(kernel_size=3, padding=1)). */
  int a_seg_k; int64_t a_seg_rows;
} srw_gemm_args;
int srw_gemm(const srw_gemm_args* a, void* stream);

/* out[M,N] (+)= sum_s workspace[s,M,N]; optionally also out_planes = split(result) */
typedef struct {
  const float* workspace; int split_k; int M, N;
  float* out; int64_t ldo; int accumulate;
} srw_splitk_reduce_args;
int srw_splitk_reduce(const srw_splitk_reduce_args* a, void* stream);

/* out[c] (+)= sum_r x[r, c]  (bias gradients).  Input either fp32 (x) or planes. */
typedef struct {
  const float* x; int64_t ldx;
  const void* planes; int64_t ldp; int64_t plane_stride;
  const float* row_scale; int rows_per_scale;
  int rows, cols;
  float* out; int accumulate;
  float* workspace;  /* >= 256 * cols floats */
} srw_colsum_args;
int srw_colsum(const srw_colsum_args* a, void* stream);   /* out == NULL: leave the partial sums in `workspace` ([srw_colsum_nparts(rows)][cols]) for srw_grad_fold */
int srw_colsum_nparts(int rows);
int srw_layernorm_bwd_nparts(int rows);                    /* partial sets srw_layernorm_bwd leaves in its workspace ([nparts][3][cols]: dgamma, dbeta, column sums) */

/* One launch that folds everything a transformer block's backward leaves behind: up to 4 split-K workspaces (the weight
 * gradients) and up to 8 sets of partial column sums (bias / LayerNorm parameter gradients; srw_colsum with out == NULL,
 * srw_layernorm_bwd with dgamma == NULL).  Replaces 8 launches of 3-6 us each per block; results are bit-identical to the
 * separate srw_splitk_reduce / reduce launches (same summation order). */
typedef struct {
  const float* partial; int nparts; int64_t stride_p;   /* partial[p * stride_p + c] */
  int cols; float* out; int accumulate;
} srw_fold_colsum;
typedef struct {
  int n_splitk; srw_splitk_reduce_args splitk[4];
  int n_colsum; srw_fold_colsum colsum[8];
} srw_grad_fold_args;
int srw_grad_fold(const srw_grad_fold_args* a, void* stream);

/* ---- LayerNorm (vit.py:164-165, 282: nn.LayerNorm eps 1e-6) -------------------------------------------------- */
typedef struct {
  const float* x; int64_t ldx; int rows, cols; float eps;
  const float* gamma; const float* beta;
  float* mean; float* rstd;                  /* [rows], saved for backward */
  void* y_planes; int64_t ldp; int64_t plane_stride;   /* may be NULL */
  float* y_f32; int64_t ldy;                 /* may be NULL */
} srw_layernorm_fwd_args;
int srw_layernorm_fwd(const srw_layernorm_fwd_args* a, void* stream);

typedef struct {
  const float* dy; int64_t lddy;             /* upstream grad of the LN output */
  const float* x; int64_t ldx; int rows, cols;
  const float* gamma; const float* mean; const float* rstd;
  float* dx; int64_t lddx; int accumulate_dx;  /* dx (+)= LN backward */
  float* dgamma; float* dbeta; int accumulate_dparams;
  float* workspace;                          /* >= 3 * 256 * cols floats */
  /* optional fused hand-over to the next backward GEMM: dx_planes = split(row_scale[row / rows_per_scale] * dx_new)
   * (dx_new = the value written to dx) and colsum_out[c] (+)= its column sums (the bias gradient of the layer whose
   * output gradient dx is).  Saves the separate srw_split_planes pass over dx. */
  void* dx_planes; int64_t ldp; int64_t plane_stride;
  const float* row_scale; int rows_per_scale;
  float* colsum_out; int colsum_accumulate;
  /* post-LN encoders: the handed-over operand is the gradient entering the dropout that precedes the residual add,
   * dx_planes = split(dropout_mask * dx_new / keep) (same mask as the forward's SRW_EPI_RESID dropout); rows per sequence as there */
  srw_dropout drop; int drop_rows_per_seq;
} srw_layernorm_bwd_args;
int srw_layernorm_bwd(const srw_layernorm_bwd_args* a, void* stream);

/* ---- attention (vit.py:100-104: softmax(q k^T * scale) v, per image and head) --------------------------------- */
/* qkv planes are token-major [B*N, 3*D] exactly as the qkv Linear produces them (vit.py:93-98 does the head split
 * with reshape/permute; here the kernel indexes heads in place).  o planes [B*N, D].  lse [B, H, N]. */
typedef struct {
  int B, N, H, head_dim; float scale;
  const void* qkv; int64_t ld_qkv; int64_t qkv_plane_stride;
  void* o; int64_t ld_o; int64_t o_plane_stride;
  float* lse;
  /* Extensions for the text / audio encoders (NULL / 0 = plain ViT attention).  With any of them set, or N > 272, the
   * key-streaming kernels run (N <= 512):
   *   key_bias [B, ld_bias] fp32: additive key-padding bias, 0 for real tokens and -inf for padding (HF create_bidirectional_mask
   *     adds finfo.min, whose exp is exactly 0 as well); kv_len[b] = number of leading keys that can be non-masked (key chunks past
   *     it are skipped).  Build both from an attention_mask with srw_attn_mask_prepare.
   *   drop: dropout on the attention probabilities (BertSelfAttention, p = attention_probs_dropout_prob). */
  const float* key_bias; int64_t ld_bias; const int32_t* kv_len;
  srw_dropout drop;
} srw_attn_fwd_args;
int srw_attn_fwd(const srw_attn_fwd_args* a, void* stream);

typedef struct {
  int B, N, H, head_dim; float scale;
  const void* qkv; int64_t ld_qkv; int64_t qkv_plane_stride;
  const void* o; int64_t ld_o; int64_t o_plane_stride;
  const void* d_o; int64_t ld_do; int64_t do_plane_stride;
  const float* lse;
  float* delta;                               /* scratch [B, H, N] */
  void* dqkv; int64_t ld_dqkv; int64_t dqkv_plane_stride;
  const float* key_bias; int64_t ld_bias; const int32_t* kv_len;   /* as in srw_attn_fwd_args */
  srw_dropout drop;
} srw_attn_bwd_args;
int srw_attn_bwd(const srw_attn_bwd_args* a, void* stream);

/* attention_mask int64 [B, L] (1 = token, 0 = padding; NULL = all ones) -> key_bias fp32 [B, ld_bias] (0 / -inf; columns in
 * [L, ld_bias) are -inf) and kv_len int32 [B] = 1 + index of the last non-zero mask entry (>= 1).  ld_bias >= L rounded up to 64. */
int srw_attn_mask_prepare(const int64_t* attention_mask, int B, int L, float* key_bias, int64_t ld_bias, int32_t* kv_len, void* stream);

/* ---- ViT engine: the whole backbone forward / backward as one native call ------------------------------------ */
/* Mirrors VisionTransformer.forward (vit.py:277-306).  Parameters stay PyTorch-owned fp32 tensors in nn.Linear layout;
 * `params` / `grads` are arrays of device pointers in state_dict order (cls_token, pos_embed, patch_embed.proj.weight,
 * patch_embed.proj.bias, then per block norm1.w, norm1.b, qkv.w, qkv.b, proj.w, proj.b, norm2.w, norm2.b, fc1.w,
 * fc1.b, fc2.w, fc2.b, then norm.w, norm.b, head.w, head.b  = 4 + 12*depth + 4 pointers). */
typedef struct {
  int img_size, patch_size, in_chans, embed_dim, depth, num_heads, hidden_dim, num_classes;
  float ln_eps;
} srw_vit_config;

int64_t srw_vit_weight_planes_bytes(const srw_vit_config* c);           /* split (+transposed) weight cache */
int64_t srw_vit_workspace_bytes(const srw_vit_config* c, int batch, int grad_batch);

/* Where parameter `param_index` (state_dict order) lives inside the weight-plane cache: byte offset of its hi plane,
 * logical columns, leading dimension and plane stride (elements).  Returns SRW_ERR_ARG for parameters that have no
 * planes (biases, norms, cls/pos, head).  Lets the fused optimizer rewrite the cache in place. */
int srw_vit_weight_plane_slot(const srw_vit_config* c, int param_index, int64_t* byte_offset, int* cols, int* ldp, int64_t* plane_stride);

/* Refresh the split-plane weight cache from the fp32 parameters (call after every optimizer step). */
int srw_vit_prepare_weights(const srw_vit_config* c, const float* const* params, void* weight_planes, void* stream);

typedef struct {
  const srw_vit_config* cfg;
  const float* const* params;
  const void* weight_planes;
  const float* x; int batch;                  /* [batch, C, H, W] fp32 NCHW */
  int grad_batch;                             /* the first grad_batch images will be back-propagated (activations kept) */
  const float* drop_scale;                    /* [depth, 2, batch] DropPath multipliers mask/keep, or NULL */
  float* logits; float* feat;                 /* [batch, num_classes], [batch, embed_dim] */
  void* workspace; int64_t workspace_bytes;
  int gemm_impl;
  float* tokens_out;                          /* optional [batch, N, embed_dim]: the final LayerNorm of ALL tokens = VisionTransformer.extract (vit.py:277-283) */
} srw_vit_fwd_args;
int srw_vit_forward(const srw_vit_fwd_args* a, void* stream);

typedef struct {
  const srw_vit_config* cfg;
  const float* const* params;
  const void* weight_planes;
  const float* x; int batch; int grad_batch;
  const float* drop_scale;
  const float* dlogits; const float* dfeat;   /* [grad_batch, C] and [grad_batch, D] (dfeat may be NULL) */
  float* const* grads;                        /* same order as params */
  int accumulate_grads;                       /* 0: grads are overwritten, 1: grads += */
  void* workspace; int64_t workspace_bytes;
  int gemm_impl;
  /* block range of this call: blocks block_hi down to block_lo (block_hi < 0 = the whole backward).  The head / final-norm
   * stage runs iff block_hi is the last block, the embedding stage iff block_lo == 0.  Successive calls over descending
   * ranges on the same workspace equal one full call; data parallelism uses it to all-reduce the gradients of the
   * finished blocks (the tail of the flat gradient buffer) while the remaining blocks run (DDP bucket overlap). */
  int block_lo, block_hi;
} srw_vit_bwd_args;
int srw_vit_backward(const srw_vit_bwd_args* a, void* stream);

/* srw_vit_forward / srw_vit_backward capture their launches into a CUDA graph the second time they see an identical
 * argument set (same pointers and sizes) and replay it afterwards: a training loop that keeps its buffers stable pays
 * one graph launch instead of ~90 / ~340 kernel launches of host time per call.  on = 0 disables replay and drops the
 * cached graphs (also: environment SRW_GRAPHS=0).  Replay is bypassed while srw_profile_enable(1) is active. */
int srw_set_graph_mode(int on);

/* The engine's kernels are launched with programmatic dependent launch (each kernel's launch latency and prologue overlap
 * its predecessor's tail; griddepcontrol.wait orders the memory accesses).  on = 0 launches them plainly (also SRW_PDL=0).
 * Changing the mode does not affect CUDA graphs that were already captured. */
int srw_set_pdl_mode(int on);

/* ---- BERT engine: ClassificationBert.forward (bert.py:22-48) around HF BertModel, as one native call ------------------------ */
/* Arithmetic restated from transformers 5.5.0 modeling_bert.py (not under /root/reference):
 *   e = word[ids] + type[0] + pos[:L] -> LayerNorm(eps) -> dropout(p_hidden)                                  (:72-112)
 *   12 post-LN layers: a = dropout_attn(softmax(q k^T / 8 + key_bias)) v ;  x = LN(dropout(a Wo + bo) + x)     (:287-298)
 *                      x = LN(dropout(gelu(x W1 + b1) W2 + b2) + x)                                           (:330-356)
 *   feat = mean over ALL L positions of dropout(x) (padding included, bert.py:36-37); logits = gelu(feat Wc1 + bc1) Wc2 + bc2.
 * `params` / `grads`: device pointers in ClassificationBert.state_dict() order = embeddings {word, position, token_type,
 * LayerNorm.w, LayerNorm.b}, per layer {query.w, query.b, key.w, key.b, value.w, value.b, attention.output.dense.w, .b,
 * attention.output.LayerNorm.w, .b, intermediate.dense.w, .b, output.dense.w, .b, output.LayerNorm.w, .b}, pooler.dense.w, .b
 * (built by BertModel, executed by the reference but unused by bert.py:35: never read here, no gradient, may be NULL),
 * classifier.0.w, .b, classifier.2.w, .b  = 5 + 16 * layers + 6 pointers.
 * Gradient layout requirement: the gradients of query / key / value weights must be contiguous in that order (one [3 hidden,
 * hidden] matrix), and so must their biases: the three projections run as one GEMM.  The host's flat gradient buffer is laid out
 * accordingly.  Dropout: counter-based (srw_dropout); sites 0 = embeddings, 1 + 3 l = attention probabilities of layer l,
 * 2 + 3 l = attention output, 3 + 3 l = FFN output, 1 + 3 layers = pooled features. */
typedef struct {
  int vocab_size, max_position, type_vocab, hidden, layers, heads, intermediate, num_classes;
  float ln_eps;
  double p_hidden, p_attn, p_pooled;
} srw_bert_config;

int64_t srw_bert_weight_planes_bytes(const srw_bert_config* c);
int64_t srw_bert_workspace_bytes(const srw_bert_config* c, int batch, int seq_len, int grad_batch);
int srw_bert_weight_plane_slot(const srw_bert_config* c, int param_index, int64_t* byte_offset, int* cols, int* ldp, int64_t* plane_stride);
int srw_bert_prepare_weights(const srw_bert_config* c, const float* const* params, void* weight_planes, void* stream);

typedef struct {
  const srw_bert_config* cfg;
  const float* const* params;
  const void* weight_planes;
  const int64_t* input_ids; const int64_t* attention_mask;   /* [batch, seq_len] int64; attention_mask may be NULL (all ones) */
  int batch, seq_len;
  int grad_batch;                             /* the first grad_batch sequences will be back-propagated */
  const uint32_t* drop_seq_key; const int32_t* drop_seq_row;   /* [batch] each, or NULL: no dropout (eval mode / deterministic parity mode) */
  const int32_t* pool_len;                    /* [batch] or NULL (= seq_len): positions the mean pool runs over.  The reference pools over the
                                                 padded length of EACH of its three calls (bert.py:36-37); when the host pads calls of different
                                                 length to one seq_len for the concatenated launch, pool_len keeps each call's own length */
  float* logits; float* feat;                 /* [batch, num_classes], [batch, hidden] */
  void* workspace; int64_t workspace_bytes;
  int gemm_impl;
} srw_bert_fwd_args;
int srw_bert_forward(const srw_bert_fwd_args* a, void* stream);

typedef struct {
  const srw_bert_config* cfg;
  const float* const* params;
  const void* weight_planes;
  const int64_t* input_ids; const int64_t* attention_mask;
  int batch, seq_len, grad_batch;
  const uint32_t* drop_seq_key; const int32_t* drop_seq_row;
  const int32_t* pool_len;
  const float* dlogits; const float* dfeat;   /* [grad_batch, C] and [grad_batch, hidden] (dfeat may be NULL) */
  float* const* grads;                        /* same order as params; pooler entries are ignored */
  int accumulate_grads;
  void* workspace; int64_t workspace_bytes;
  int gemm_impl;
  int layer_lo, layer_hi;                     /* layer range, like block_lo / block_hi of the ViT backward; layer_hi < 0 = everything */
} srw_bert_bwd_args;
int srw_bert_backward(const srw_bert_bwd_args* a, void* stream);

/* ---- HuBERT engine: ClassificationHubert.forward / backward as two native calls -------------------------------- */
/* Replaces `self.model(x, ...)` + dropout + mean + classifier of semilearn/nets/hubert/hubert.py:24-50 and autograd's backward of it.
 * Arithmetic restated from transformers 5.5.0 modeling_hubert.py (not under /root/reference), architecture of
 * facebook/hubert-base-ls960 (`feat_extract_norm='group'`, `do_stable_layer_norm=False`):
 *   conv stem: 7 Conv1d without bias on the raw waveform, 512 channels, kernels (10,3,3,3,3,2,2), strides (5,2,2,2,2,2,2), GELU after each,
 *              GroupNorm(512 groups = per channel over time) after the first                    -> F frames (199 for 64 000 samples)
 *   feature projection: LayerNorm(512) -> Linear(512, hidden) -> dropout(p_feat_proj)
 *   SpecAugment: frames flagged in `mask_time` are replaced by `masked_spec_embed`  (the host draws the spans like _compute_mask_indices)
 *   positional conv: Conv1d(hidden, hidden, k = 128, pad 64, groups 16), weight-normalised per tap, last frame dropped, GELU;
 *              x = dropout(LayerNorm(h + pos), p_hidden)
 *   `layers` post-LN encoder layers (attention without any mask: hubert.py:45 passes none), each skipped per model call by LayerDrop
 *   feat = mean over the F frames of dropout(x, p_pooled);  logits = gelu(feat Wc1 + bc1) Wc2 + bc2          (hubert.py:17-21,44-49)
 * Layout in HBM: every conv-stem activation is time-major [clip, frame, channel] split planes, clips stacked with one padding frame each so
 * that a strided Conv1d is ONE GEMM over an overlapping row view (row t of the im2col matrix = frames s t .. s t + k - 1, contiguous: lda =
 * s * 512 < K = k * 512); the grouped positional conv is 16 GEMMs (one per group) over a group-major zero-padded copy of h (row stride 48).
 * `params` / `grads`: device pointers in ClassificationHubert.state_dict() order = masked_spec_embed, conv0.weight, conv0.layer_norm.w, .b,
 * conv{1..6}.weight, feature_projection {layer_norm.w, .b, projection.w, .b}, pos_conv_embed {bias, weight_g (original0), weight_v
 * (original1)}, encoder.layer_norm {w, b}, per layer {k_proj.w, .b, v_proj.w, .b, q_proj.w, .b, out_proj.w, .b, layer_norm.w, .b,
 * intermediate_dense.w, .b, output_dense.w, .b, final_layer_norm.w, .b}, classifier.0.w, .b, classifier.2.w, .b = 19 + 16 * layers + 4.
 * Gradient layout requirement: the gradients of q_proj / k_proj / v_proj weights must be contiguous in THAT order (and their biases).
 * Dropout sites (srw_dropout): 0 feature projection, 1 encoder input, 2 + 4 l attention probabilities, 3 + 4 l attention output,
 * 4 + 4 l FFN activation, 5 + 4 l FFN output, 2 + 4 layers pooled features. */
#define SRW_HUBERT_MAX_CONV 8
typedef struct {
  int hidden, layers, heads, intermediate, num_classes;
  int conv_dim;                                    /* 512, all conv layers */
  int num_conv; int conv_kernel[SRW_HUBERT_MAX_CONV]; int conv_stride[SRW_HUBERT_MAX_CONV];
  int pos_kernel, pos_groups;                      /* 128, 16 */
  float ln_eps;                                    /* 1e-5 (GroupNorm / LayerNorms) */
  double p_feat_proj, p_hidden, p_attn, p_act, p_pooled;
} srw_hubert_config;

int srw_hubert_frames(const srw_hubert_config* c, int samples);                 /* F for clips of `samples` samples (< 0: invalid) */
int64_t srw_hubert_weight_planes_bytes(const srw_hubert_config* c);
int64_t srw_hubert_workspace_bytes(const srw_hubert_config* c, int batch, int samples, int grad_batch);
/* rebuilds the whole plane cache (re-laid-out conv weights, weight-normalised positional taps, packed q|k|v) */
int srw_hubert_prepare_weights(const srw_hubert_config* c, const float* const* params, void* weight_planes, void* stream);
/* The encoder matrices (q | k | v, out_proj, intermediate_dense, output_dense: 85 of the 94 M parameters) have per-parameter plane slots
 * the fused AdamW kernel rewrites in the same pass as the parameter (like the ViT / BERT caches); after such a step only the front end
 * (conv relayouts, projection, weight-normalised positional taps, packed biases) needs rebuilding: */
int srw_hubert_prepare_front(const srw_hubert_config* c, const float* const* params, void* weight_planes, void* stream);
int srw_hubert_weight_plane_slot(const srw_hubert_config* c, int param_index, int64_t* byte_offset, int* cols, int* ldp, int64_t* plane_stride);

typedef struct {
  const srw_hubert_config* cfg;
  const float* const* params;
  const void* weight_planes;
  const float* wav; int64_t ld_wav;             /* [batch, samples] fp32 */
  int batch, samples;
  int grad_batch;                               /* the first grad_batch clips will be back-propagated */
  const uint8_t* mask_time;                     /* device [batch, F] (1 = frame replaced by masked_spec_embed) or NULL */
  const uint32_t* drop_seq_key; const int32_t* drop_seq_row;   /* [batch] each, or NULL: no dropout */
  /* LayerDrop: the launch is `num_segments` consecutive clip ranges (the model calls it stands for); segment_start is a HOST array
   * [num_segments + 1] (first clip of each, last = batch), layer_skip a HOST array [num_segments, layers] (1 = that call skips the
   * layer).  NULL / 0 = nothing skipped.  A boundary must not fall inside the gradient rows' range end: grad_batch is a boundary. */
  int num_segments; const int32_t* segment_start; const uint8_t* layer_skip;
  float* logits; float* feat;                   /* [batch, num_classes], [batch, hidden] */
  void* workspace; int64_t workspace_bytes;
  int gemm_impl;
} srw_hubert_fwd_args;
int srw_hubert_forward(const srw_hubert_fwd_args* a, void* stream);

typedef struct {
  const srw_hubert_config* cfg;
  const float* const* params;
  const void* weight_planes;
  const float* wav; int64_t ld_wav;
  int batch, samples, grad_batch;
  const uint8_t* mask_time;
  const uint32_t* drop_seq_key; const int32_t* drop_seq_row;
  int num_segments; const int32_t* segment_start; const uint8_t* layer_skip;
  const float* dlogits; const float* dfeat;     /* [grad_batch, C] and [grad_batch, hidden] (dfeat may be NULL) */
  float* const* grads;                          /* same order as params */
  int accumulate_grads;
  void* workspace; int64_t workspace_bytes;
  int gemm_impl;
} srw_hubert_bwd_args;
int srw_hubert_backward(const srw_hubert_bwd_args* a, void* stream);

/* ---- WideResNet engine: WideResNet.forward / backward as two native calls ----------------------------------------- */
/* Replaces `self.model(inputs)` of the `use_cat: True` step (srflexmatch.py:111-117) for net wrn_28_2 / wrn_28_8
 * (semilearn/nets/wrn/wrn.py:73-146; BasicBlock :30-54) and autograd's backward of it.  NHWC with a zero border as operand planes,
 * 3x3 convolutions as ONE srw_gemm with K segments over an overlapping row view (srw_gemm_args.a_seg_k), train-mode BatchNorm
 * (batch statistics over ALL rows of the launch, running statistics advanced with momentum bn_momentum and the unbiased variance,
 * wrn.py:33,37,97; final BatchNorm eps 1e-3, others 1e-5) + LeakyReLU(slope) fused into the operand writers.
 * `params` / `grads`: device pointers in WideResNet.named_parameters() order = conv1.weight, conv1.bias, per BasicBlock {bn1.weight,
 * bn1.bias, conv1.weight, bn2.weight, bn2.bias, conv2.weight[, convShortcut.weight]}, bn1.weight, bn1.bias, classifier.weight,
 * classifier.bias (81 tensors for WRN-28-2).  The bn1 of the first layer of block2 / block3 normalises a tensor the block then
 * drops (wrn.py:46-51): its running statistics advance, its parameters get NO gradient (grads[] entries may be NULL, never written).
 * BatchNorm buffers: bn_running_mean / bn_running_var / bn_num_batches_tracked [2 * blocks + 1] pointers, per block bn1 then bn2,
 * last the final bn1 (num_batches_tracked: int64 scalars, may be NULL). */
typedef struct {
  int num_classes, depth, widen, img_size;
  float bn_momentum, slope;                    /* 0.001, 0.1 */
} srw_wrn_config;

int srw_wrn_num_params(const srw_wrn_config* c);
int64_t srw_wrn_weight_planes_bytes(const srw_wrn_config* c);
int64_t srw_wrn_workspace_bytes(const srw_wrn_config* c, int batch);
int srw_wrn_prepare_weights(const srw_wrn_config* c, const float* const* params, void* weight_planes, void* stream);

/* Data parallel = SyncBatchNorm (core/utils/misc.py:54 converts every BatchNorm2d under DDP): the batch statistics of a layer are sums over
 * the rows of ALL ranks.  The engine folds its per-CTA partials into `sync_buf` (device, >= 2 * 64 * widen floats: per-channel sum and sum of
 * squares; in the backward sum(du) and sum(du * xhat)), calls `sync_fn(sync_ctx, pointer into sync_buf, count)` on the host thread — the
 * caller all-reduces (SUM) that range on the same stream (torch.distributed / NCCL) — and continues with the global sums and
 * world_size * rows as the count.  dgamma / dbeta stay LOCAL sums (averaged later with the other gradients, as DDP does).  With
 * sync_fn == NULL nothing changes.  Calls with a sync_fn run eagerly (no CUDA-graph replay: the collective is host-enqueued). */
typedef int (*srw_allreduce_sum_fn)(void* ctx, float* device_buf, int count);

typedef struct {
  const srw_wrn_config* cfg;
  const float* const* params;
  float* const* bn_running_mean; float* const* bn_running_var; int64_t* const* bn_num_batches_tracked;
  const void* weight_planes;
  const float* x;                              /* [batch, 3, img, img] fp32 NCHW */
  int batch;
  int training;                                /* 1: batch statistics + running-statistics update; 0: running statistics (eval) */
  int stat_repeats;                            /* training: advance the running statistics this many EXTRA times with the same batch
                                                  statistics — the K sampling passes of stage 2 re-run the identical deterministic
                                                  forward (srflexmatch.py:72-104), whose only effect is on these buffers */
  float* logits; float* feat;                  /* [batch, num_classes], [batch, 64 * widen] */
  void* workspace; int64_t workspace_bytes;
  int gemm_impl;
  srw_allreduce_sum_fn sync_fn; void* sync_ctx; float* sync_buf; int world_size;   /* SyncBatchNorm (see above); NULL / 0 = single rank */
} srw_wrn_fwd_args;
int srw_wrn_forward(const srw_wrn_fwd_args* a, void* stream);

typedef struct {
  const srw_wrn_config* cfg;
  const float* const* params;
  const void* weight_planes;
  int batch;
  int grad_rows;                               /* dlogits / dfeat cover the first grad_rows rows; the others (weak rows) have zero
                                                  logit gradient but still take part in every BatchNorm backward */
  const float* dlogits; const float* dfeat;    /* [grad_rows, C], [grad_rows, 64 * widen] (dfeat may be NULL) */
  float* const* grads;
  int accumulate_grads;
  void* workspace; int64_t workspace_bytes;    /* the forward's workspace */
  int gemm_impl;
  srw_allreduce_sum_fn sync_fn; void* sync_ctx; float* sync_buf; int world_size;
} srw_wrn_bwd_args;
int srw_wrn_backward(const srw_wrn_bwd_args* a, void* stream);


/* x[i] *= *scale for i < n, where scale is a DEVICE scalar; returns without touching memory when *scale == 1.  Used to
 * apply autograd's upstream gradient of the loss (normally exactly 1, param_update.py:33) to gradients that were
 * computed ahead of loss.backward(). */
int srw_scale_inplace(float* x, int64_t n, const float* scale, void* stream);

/* ---- fused SSL epilogue (the a6-a12 rows of SURVEY.md §8a) ---------------------------------------------------- */
/* Rewarder.forward (semireward.py:52-72): reward[b] in (0,1), one CTA, softmax over the 2B rows done on chip.
 * rp = 17 device pointers in Rewarder.state_dict order.
 * Deliberate divergence: a label outside [0, label_rows) is CLAMPED into range in every Rewarder kernel (forward, training step,
 * fused SSL loss), where the reference's nn.Embedding / F.one_hot would raise a device-side assert (semireward.py:56,136); the
 * Generator's relu(...).long() output is unbounded in principle (0 at initialisation, never trained, semireward.py:21-24). */
typedef struct {
  int B, feature_dim, label_rows;
  const float* const* rp;
  const float* feats; int64_t ld_feats;
  const int64_t* labels;
  float* reward;                               /* [B] */
  float* workspace;                            /* >= srw_rewarder_workspace_floats(B) */
} srw_rewarder_fwd_args;
int64_t srw_rewarder_workspace_floats(int B, int feature_dim);  /* also covers srw_rewarder_train */
int srw_rewarder_fwd(const srw_rewarder_fwd_args* a, void* stream);

/* Generator.forward (semireward.py:21-24) followed by .long() (srflexmatch.py:157-158): integer fake labels. */
typedef struct {
  int B, feature_dim;
  const float* const* gp;                      /* 8 pointers, Generator.state_dict order */
  const float* feats; int64_t ld_feats;
  int64_t* labels;                             /* [B] */
  float* workspace;                            /* >= B * 448 floats */
} srw_generator_fwd_args;
int srw_generator_fwd(const srw_generator_fwd_args* a, void* stream);

/* Rewarder training step (srflexmatch.py:173-208): reward = R(feats, gen_label); target = (gen==true ? 1 : .5)
 * (== cosine_similarity_n of the two one-hots, semireward.py:130-139); grads of MSE(reward,1) + MSE(reward,target);
 * one torch.optim.Adam step (lr, betas .9/.999, eps 1e-8) on the 17 Rewarder tensors.  m/v = Adam moments, same order. */
typedef struct {
  int B, feature_dim, label_rows, num_classes;
  float* const* rp; float* const* g; float* const* m; float* const* v;   /* params, grad scratch, Adam moments */
  const float* feats; int64_t ld_feats;
  const int64_t* gen_labels; const int64_t* true_labels;
  float lr; int step;                          /* step = t (1-based) for bias correction */
  int phase;                                   /* 0: forward+backward+Adam fused; 1: forward+backward only (g written);
                                                  2: Adam only from g (lets the caller all-reduce g between 1 and 2: DDP, C3) */
  float* losses;                               /* [2] generator_loss, rewarder_loss */
  float* workspace;                            /* >= srw_rewarder_workspace_floats(B, feature_dim) */
  /* Data parallel, matching torch DDP on the reference's sequence `generator_loss.backward(); rewarder_loss.backward()` after ONE
   * forward of the DDP-wrapped Rewarder (srflexmatch.py:204-205): DDP's reducer synchronises only the FIRST backward, so the
   * optimizer sees mean_ranks(g_generator_loss) + LOCAL g_rewarder_loss (measured with the reference's own module under gloo DDP,
   * scripts/c3_ddp_probe.py).  loss_select: 0 = both losses (single process), 1 = gradient of generator_loss only, 2 = of
   * rewarder_loss only (phases 0/1).  g_add (phase 2, may be NULL): a second gradient set added to g before the Adam step. */
  int loss_select; float* const* g_add;
} srw_rewarder_train_args;
int srw_rewarder_train(const srw_rewarder_train_args* a, void* stream);

/* FlexMatchThresholdingHook.masking + update (srflexmatch/utils.py:23-63) fused with compute_prob (algorithmbase.py:332-333)
 * and PseudoLabelingHook.gen_ulb_targets hard labels (hooks/pseudo_label.py:40).  One CTA.  The host Counter over all
 * ulb_dest_len entries (utils.py:25-29) is replaced by a device-resident histogram kept incrementally:
 * hist[c+1] = #entries of selected_label equal to c, hist[0] = #entries still -1. */
typedef struct {
  int B, num_classes, ulb_dest_len;
  const float* logits_w; int64_t ld_logits;
  const int64_t* idx_ulb;
  float p_cutoff; int thresh_warmup;
  int64_t* selected_label;  /* [ulb_dest_len] hook state; NULL = FixedThresholdingHook (hooks/masking.py:42-57): stateless
                               mask = max_p >= p_cutoff, idx_ulb / hist / classwise_acc unused (SRFixMatch) */
  int32_t* hist;            /* [num_classes + 1] hook state */
  float* classwise_acc;     /* [num_classes] hook state */
  float* probs_w;           /* [B, C] out, row stride num_classes (may be NULL) */
  int64_t* pseudo;          /* [B] out: argmax of probs_w */
  float* mask;              /* [B] out */
  float* max_probs;         /* [B] out (may be NULL) */
} srw_flexmatch_mask_args;
int srw_flexmatch_mask(const srw_flexmatch_mask_args* a, void* stream);

/* Losses of the SSL step and their gradient w.r.t. the logits (srflexmatch.py:132,100-102,152,210; cross_entropy.py:11-31;
 * consistency.py:13-45):  sup = mean CE(logits_lb, y_lb);  mask2 = reward >= mean(reward) (when reward != NULL);
 * unsup = mean_B(CE(logits_s, pseudo) * mask * mask2);  total = sup + lambda_u * unsup;  util = mean(mask).  One CTA. */
typedef struct {
  int B_lb, B_ulb, num_classes;
  const float* logits_lb; const float* logits_s; int64_t ld_logits;
  const int64_t* y_lb; const int64_t* pseudo;
  const float* mask;        /* [B_ulb] per-sample weight (0/1 for FlexMatch/FreeMatch, soft for SoftMatch) */
  const float* reward;      /* [B_ulb] or NULL */
  float lambda_u;
  float* mask2;             /* [B_ulb] out (ones when reward == NULL; may be NULL) */
  float* losses;            /* [>= 4] out: sup, unsup, total, util_ratio ([4] = FreeMatch entropy term, written by srw_freematch_entropy) */
  float* dlogits_lb; float* dlogits_s; int64_t ld_dlogits;   /* d total / d logits (may be NULL) */
  /* Fused stage-2 epilogue (north_star: "Rewarder MLP forward, reward score, mean-threshold mask and masked cross-entropy consistency
   * loss collapse into one fused epilogue kernel"): when rp != NULL the kernel first evaluates Rewarder.forward(feats, pseudo)
   * (semireward.py:52-72, srflexmatch.py:99) itself — `reward` is then ignored — and continues with mask2 / losses / dlogits in the
   * same launch.  rp = 17 device pointers in Rewarder.state_dict order; rew_workspace >= srw_rewarder_workspace_floats(B_ulb, .)
   * floats (only touched when the intermediates do not fit in shared memory); reward_out [B_ulb] optional. */
  const float* const* rp; const float* feats; int64_t ld_feats; int feature_dim, label_rows;
  float* rew_workspace; float* reward_out;
} srw_ssl_loss_args;
int srw_ssl_loss(const srw_ssl_loss_args* a, void* stream);

/* FreeMatchThresholdingHook.masking + update (semilearn/algorithms/freematch/utils.py:23-66) fused with compute_prob
 * (algorithmbase.py:332-333) and PseudoLabelingHook.gen_ulb_targets hard labels (hooks/pseudo_label.py:40).  One CTA.
 * State (device): time_p [1], p_model [C], label_hist [C]; every EMA in the reference's fp32 evaluation order.
 * pseudo_from_probs: 0 = argmax of the logits (train_step, srfreematch.py:146-150), 1 = argmax of the probabilities
 * (data_generator, :93-97). */
typedef struct {
  int B, num_classes;
  const float* logits_w; int64_t ld_logits;
  double momentum; int use_quantile; int clip_thresh;
  float* time_p; float* p_model; float* label_hist;
  float* probs_w;           /* [B, C] out: softmax(logits_w), row stride C */
  int64_t* pseudo; int pseudo_from_probs;
  float* mask;              /* [B] out */
  float* max_probs;         /* [B] out (may be NULL) */
  /* data parallel (C4, freematch/utils.py:25-26): update() sees the probabilities of ALL ranks, masking() the local ones.
   * phase 0 = everything in one launch (world size 1, probs_all = NULL); phase 1 = softmax + pseudo-labels only; then the
   * caller all-gathers probs_w into probs_all [B_all, C]; phase 2 = update() from probs_all + mask of the local rows. */
  int phase; const float* probs_all; int B_all;
} srw_freematch_mask_args;
int srw_freematch_mask(const srw_freematch_mask_args* a, void* stream);

/* entropy_loss of SRFreeMatch (srfreematch.py:12-44) over the rows with mask != 0, and `total += lambda_e * ent`
 * (srfreematch.py:214-219): losses[4] = ent, losses[2] += lambda_e * ent; dlogits_s (+)= lambda_e * d ent / d logits_s.
 * ent = 0 and no gradient when no row is selected. */
typedef struct {
  int B, num_classes;
  const float* mask; const float* logits_s; int64_t ld_logits;
  const float* p_model; const float* label_hist;
  float lambda_e;
  float* losses;            /* [5]: see srw_ssl_loss for [0..3] */
  float* dlogits_s; int64_t ld_dlogits; int accumulate;   /* may be NULL; accumulate = 0 overwrites (zeros for unselected rows) */
} srw_freematch_entropy_args;
int srw_freematch_entropy(const srw_freematch_entropy_args* a, void* stream);

/* SoftMatch: softmax, optional DistAlignEMAHook.dist_align (hooks/dist_align.py:25-55, uniform target), then
 * SoftMatchWeightingHook.update + masking with per_class = False (srsoftmatch/utils.py:31-77): EMA of the mean and the
 * unbiased variance of the (aligned) max probabilities, weight = exp(-clamp(max_p - mu, max=0)^2 / (2 var / n_sigma^2)).
 * The reference's two .item() host syncs per call are gone; the EMA keeps its mixed fp32 / double arithmetic.  One CTA. */
typedef struct {
  int B, num_classes;
  const float* logits_w; int64_t ld_logits;
  double momentum; int n_sigma;
  int dist_align;                      /* 1: train_step (srsoftmatch.py:134-141); 0: data_generator passes (:84-90) */
  float* da_p_model; const float* da_p_target; int32_t* da_initialized;   /* DistAlign state ([C], [C], [1]); p_model is set on first use */
  float* prob_max_mu_t; float* prob_max_var_t;                             /* [1], [1] state */
  float* probs_w;                      /* [B, C] out: softmax(logits_w) (un-aligned) */
  float* probs_aligned;                /* [B, C] out (may be NULL) */
  int64_t* pseudo; int pseudo_from_probs;
  float* mask;                         /* [B] out: the soft weights */
  float* max_probs;                    /* [B] out (may be NULL): the max of the probabilities the weights were formed from */
  /* data parallel (C4: dist_align.py:40-42, srsoftmatch/utils.py:33-34): DistAlign averages every rank's probabilities,
   * the weighting statistics run over every rank's (aligned) max probabilities.  phase 0 = one launch (world size 1);
   * phase 1 = softmax + pseudo-labels (max_probs = un-aligned row max); all-gather probs_w -> probs_all [B_all, C];
   * phase 2 (dist_align only) = DistAlign update + aligned row max -> max_probs; all-gather max_probs -> maxp_all [n_all];
   * phase 3 = EMA of mean / variance over maxp_all + weights of the local rows (from max_probs). */
  int phase; const float* probs_all; int B_all; const float* maxp_all; int n_all;
} srw_softmatch_mask_args;
int srw_softmatch_mask(const srw_softmatch_mask_args* a, void* stream);

/* ---- optimizer: torch.optim.AdamW / Adam over many tensors in one launch --------------------------------------------- */
/* Replaces optimizer.step() of param_update.py:36 for the AdamW built by get_optimizer with the reference's layer-decay
 * param groups (build.py:193-224, nets/utils.py:143-204), and model.zero_grad() is unnecessary because the backward
 * overwrites gradients.  Same arithmetic as torch's single-tensor Adam: p *= 1 - lr*wd (decoupled) ; m.lerp_(g, 1-b1) ;
 * v = v*b2 + (1-b2) g g ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps).  Optionally rewrites the split-bf16 planes of the
 * updated parameter in the same pass (the ViT engine's weight cache), saving srw_vit_prepare_weights.
 * The table (one row per tensor) lives in DEVICE memory; the caller fills it with a plain copy of this struct array. */
#define SRW_ADAMW_BLOCK_ELEMS 4096
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  void* planes;             /* NULL, or planes of this parameter: element i -> planes[(i / cols) * ldp + i % cols] */
  int64_t numel;
  int64_t plane_stride;
  int32_t cols, ldp;
  double lr;                /* param-group base lr (layer-decay scale included), before the schedule factor */
  double weight_decay;
  int64_t first_block;      /* exclusive prefix sum of ceil(numel / SRW_ADAMW_BLOCK_ELEMS) */
} srw_adamw_row;

typedef struct {
  int num_tensors; int64_t total_blocks;
  const srw_adamw_row* table;   /* device pointer */
  double lr_factor;             /* schedule multiplier of this step (LambdaLR) */
  double beta1, beta2, eps;
  int step;                     /* 1-based */
  int decoupled;                /* 1 = AdamW, 0 = Adam with L2 (grad += wd * p) */
} srw_adamw_args;
int srw_adamw_step(const srw_adamw_args* a, void* stream);

/* torch.optim.SGD(momentum, nesterov, weight_decay) over all tensors in one launch (core/utils/build.py:219-220: the optimizer of
 * config/classic_cv).  Rows are srw_adamw_row with exp_avg = momentum buffer (exp_avg_sq / planes unused).
 *   g' = g + wd p;  buf = first_step ? g' : momentum buf + g';  p -= lr (nesterov ? g' + momentum buf : buf) */
typedef struct {
  int num_tensors; int64_t total_blocks;
  const void* table;                           /* device array of srw_adamw_row */
  double lr_factor, momentum;
  int nesterov, first_step;
} srw_sgd_args;
int srw_sgd_step(const srw_sgd_args* a, void* stream);

/* ---- EMA of the parameters: EMAHook.after_train_step / EMA.update (core/hooks/ema.py:20-24, core/utils/misc.py:152-155) ---- */
/* shadow = (1 - decay) * param + decay * shadow over many tensors in one launch (bit-exact with the reference's fp32 tensor
 * expression).  The table lives in device memory; first_block as in srw_adamw_row.  HBM-bound: 12 B per parameter. */
typedef struct { const float* param; float* shadow; int64_t numel; int64_t first_block; } srw_ema_row;
typedef struct { int num_tensors; int64_t total_blocks; const srw_ema_row* table; double decay; } srw_ema_args;
int srw_ema_step(const srw_ema_args* a, void* stream);

/* ---- image input pipeline: transform_weak / transform_strong on the device ------------------------- */
/* Replaces, per sample, BasicDataset.__getitem__'s PIL pipelines (semilearn/datasets/cv_datasets/datasetbase.py:85-115) built by
 * get_cifar (semilearn/datasets/cv_datasets/cifar.py:34-49): Resize (identity at the source size) -> RandomCrop(size, padding,
 * 'reflect') -> RandomHorizontalFlip -> [RandAugment(3, 5): up to 3 ops of augment_list() + Cutout,
 * semilearn/datasets/augmentation/randaugment.py:16-206] -> ToTensor -> Normalize.  The uint8 dataset stays resident in HBM; the
 * host sends the random DECISIONS (drawn from the generators the reference draws from, semireward_b200/datasets/gpu_augment.py);
 * one CTA per sample writes the normalised fp32 CHW image.  Results are bit-identical to Pillow 12.2 / torchvision 0.26.
 * Op ids are the positions in augment_list() (randaugment.py:157-174). */
enum srw_aug_op {
  SRW_AUG_AUTOCONTRAST = 0, SRW_AUG_BRIGHTNESS = 1, SRW_AUG_COLOR = 2, SRW_AUG_CONTRAST = 3, SRW_AUG_EQUALIZE = 4, SRW_AUG_IDENTITY = 5,
  SRW_AUG_POSTERIZE = 6, SRW_AUG_ROTATE = 7, SRW_AUG_SHARPNESS = 8, SRW_AUG_SHEAR_X = 9, SRW_AUG_SHEAR_Y = 10, SRW_AUG_SOLARIZE = 11,
  SRW_AUG_TRANSLATE_X = 12, SRW_AUG_TRANSLATE_Y = 13
};
typedef struct {
  int32_t op;         /* enum srw_aug_op */
  int32_t ival;       /* Posterize: bits kept (1..8); Solarize: first inverted level = ceil(threshold) */
  float alpha;        /* Brightness / Color / Contrast / Sharpness: Image.blend factor as a C float */
  int32_t identity;   /* affine ops: 1 = Image.rotate's `angle % 360 == 0` copy, nothing to do */
  double a[6];        /* Rotate / Shear / Translate: inverse-map coefficients exactly as Pillow's Python layer computes them */
} srw_aug_op_desc;
typedef struct {
  int64_t src_index;               /* image of the resident dataset */
  int32_t crop_top, crop_left;     /* RandomCrop.get_params inside the padded image: 0 .. 2 * padding */
  int32_t flip;
  int32_t n_ops;                   /* 0 = transform_weak */
  srw_aug_op_desc ops[3];
  int32_t cut_x0, cut_y0, cut_x1, cut_y1;   /* Cutout rectangle after ImageDraw's int truncation, corners inclusive; x1 < x0 = none */
} srw_aug_sample;
typedef struct {
  const uint8_t* src;              /* device [n_src, img_size, img_size, 3] uint8 HWC */
  int64_t n_src;
  int img_size, padding;
  const srw_aug_sample* samples;   /* device [n] */
  int n;
  float mean[3], std[3];           /* Normalize constants rounded to fp32 (torch.as_tensor(mean, dtype=float32)) */
  float* out;                      /* device [n, 3, img_size, img_size] fp32 */
  uint8_t* out_u8;                 /* optional device [n, img_size, img_size, 3]: the image handed to ToTensor */
} srw_augment_args;
int srw_augment_batch(const srw_augment_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SRW_H_ */
