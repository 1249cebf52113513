// srw_bert — ClassificationBert.forward (semilearn/nets/bert/bert.py:22-48) and its backward as two native calls.
//
// The encoder arithmetic is Hugging Face transformers' BertModel (un-vendored dependency of the reference, `transformers>=4.30.0`,
// 5.5.0 in this image; modeling_bert.py:72-112 embeddings, :287-298 / :330-356 post-LN layers) restated on this library's kernels:
//   forward, per layer:   [q|k|v] GEMM (one packed [3D, D] weight) -> planes | key-streaming attention (key-padding bias, dropout on
//                         the probabilities) | Wo GEMM + dropout + residual -> u1 | LN -> x_mid (fp32 + planes) | W1 GEMM + GELU |
//                         W2 GEMM + dropout + residual -> u2 | LN -> x_out (fp32 + planes)
//   backward, per layer:  mirrored.  Post-LN means every LayerNorm backward produces du, which goes BOTH to the residual path (it is
//                         the `resid` operand of the branch's last dgrad GEMM) and, through the dropout mask, into the branch
//                         (the fused plane hand-over of srw_layernorm_bwd).  wgrad GEMMs are split-K into workspaces, folded once
//                         per layer by srw_grad_fold together with all bias / LayerNorm-parameter column sums.
// Dropout is counter-based (include/srw.h: srw_dropout): the backward regenerates every mask, nothing is stored.
// `use_cat: False` (srsoftmatch.py:119-130): the reference calls the model three times; rows do not interact in a LayerNorm
// network, so the host runs ONE concatenated call whose sequences carry their own dropout stream key and row index, and only
// the first grad_batch sequences (labelled + strong) are back-propagated (the weak call runs under no_grad there).
#include <cuda.h>

#include <vector>

#include "../../include/srw.h"
#include "srw_common.cuh"
#include "srw_encoder.cuh"

namespace srw {

// ------------------------------------------------------------------------------------------------
// small kernels: embeddings, pooling, classifier head
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum256(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// e = word[id] + type[0] + pos[l]; x0 = dropout(LN(e)).  One warp per token, the row lives in registers (D = 128 * J4).
template <int J4>
__global__ void __launch_bounds__(256) bert_embed_fwd_kernel(const int64_t* __restrict__ ids, int T, int Lq, int vocab, const float* __restrict__ word,
                                                             const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, float* __restrict__ e_out, float* __restrict__ mean_out,
                                                             float* __restrict__ rstd_out, float* __restrict__ xf, __nv_bfloat16* __restrict__ xp, int64_t ps,
                                                             const DropParams dr) {
  constexpr int D = J4 * 128;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= T) return;
  const int l = t % Lq, sq = t / Lq;
  int64_t id = ids[t];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);   // nn.Embedding raises on an out-of-range id; the clamp only keeps the kernel in bounds
  float4 v[J4];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    const int c = j * 128 + lane * 4;
    const float4 w = *reinterpret_cast<const float4*>(word + id * D + c);
    const float4 ty = *reinterpret_cast<const float4*>(type0 + c);
    const float4 po = *reinterpret_cast<const float4*>(pos + (int64_t)l * D + c);
    v[j] = make_float4((w.x + ty.x) + po.x, (w.y + ty.y) + po.y, (w.z + ty.z) + po.z, (w.w + ty.w) + po.w);   // modeling_bert.py: inputs + token_type, then + position
    *reinterpret_cast<float4*>(e_out + (int64_t)t * D + c) = v[j];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) { mean_out[t] = mean; rstd_out[t] = rstd; }
  uint32_t key = 0, base = 0;
  if (dr.on) {
    key = drop_site_key(dr.seq_key[sq], dr.site);
    base = ((uint32_t)dr.seq_row[sq] * (uint32_t)Lq + (uint32_t)l) * (uint32_t)D;
  }
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    const int c = j * 128 + lane * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 y = make_float4(v[j].x * rstd * g.x + b.x, v[j].y * rstd * g.y + b.y, v[j].z * rstd * g.z + b.z, v[j].w * rstd * g.w + b.w);
    if (dr.on) {
      y.x = drop_kept(key, base + c, dr.thr24) ? y.x * dr.inv_keep : 0.f;
      y.y = drop_kept(key, base + c + 1, dr.thr24) ? y.y * dr.inv_keep : 0.f;
      y.z = drop_kept(key, base + c + 2, dr.thr24) ? y.z * dr.inv_keep : 0.f;
      y.w = drop_kept(key, base + c + 3, dr.thr24) ? y.w * dr.inv_keep : 0.f;
    }
    *reinterpret_cast<float4*>(xf + (int64_t)t * D + c) = y;
    uint32_t h0, l0, h1, l1;
    split2(y.x, y.y, h0, l0);
    split2(y.z, y.w, h1, l1);
    __nv_bfloat16* hp = xp + (int64_t)t * D + c;
    *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(hp + ps) = make_uint2(l0, l1);
  }
}

// word-embedding gradient: dword[id[t], :] += de[t, :] for id != padding_idx (0): nn.Embedding(padding_idx=0) never updates that row.
// Repeated tokens collide, hence atomics (fp32 adds in arrival order: the one non-deterministic summation of the path, ~1 ulp).
__global__ void bert_word_grad_kernel(const float* __restrict__ de, const int64_t* __restrict__ ids, int Tg, int D, int vocab, float* __restrict__ dword) {
  const int64_t total = (int64_t)Tg * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / D), c = (int)(i % D);
    const int64_t id = ids[t];
    if (id <= 0 || id >= vocab) continue;
    atomicAdd(dword + id * D + c, de[i]);
  }
}

// possum[l, :] = sum over sequences of de[(s, l), :];  dpos[l, :] (+)= possum[l, :] (rows l >= L of the table get no gradient)
__global__ void bert_pos_grad_kernel(const float* __restrict__ de, int Sg, int Lq, int D, int max_pos, float* __restrict__ possum, float* __restrict__ dpos,
                                     int accumulate) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= (int64_t)max_pos * D) return;
  const int l = (int)(i / D);
  float acc = 0.f;
  if (l < Lq) {
    for (int s = 0; s < Sg; ++s) acc += de[(int64_t)s * Lq * D + i];
    possum[i] = acc;
  }
  dpos[i] = accumulate ? dpos[i] + acc : acc;
}
// token_type_ids are all 0 (bert.py:34 passes none): dtype[0, :] (+)= sum over positions of possum; the other rows get zero
__global__ void bert_type_grad_kernel(const float* __restrict__ possum, int Lq, int D, int type_vocab, float* __restrict__ dtype, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float acc = 0.f;
  for (int l = 0; l < Lq; ++l) acc += possum[(int64_t)l * D + c];
  dtype[c] = accumulate ? dtype[c] + acc : acc;
  if (!accumulate)
    for (int r = 1; r < type_vocab; ++r) dtype[(int64_t)r * D + c] = 0.f;
}

// packed [3D] bias of the fused q|k|v projection, for every layer, from the three separate bias tensors
struct QkvBiasPtrs { const float* p[3 * 48]; };
__global__ void bert_pack_qkv_bias_kernel(const __grid_constant__ QkvBiasPtrs P, int layers, int D, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= layers * 3 * D) return;
  const int which = i / D, d = i % D;
  out[i] = P.p[which][d];
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
struct BDims {
  int S, Sg, Lq, D, H, F, NL, C, V;
  int64_t T, Tg, ldb;
};

static int make_bdims(const srw_bert_config* c, int batch, int seq_len, int grad_batch, BDims& d) {
  SRW_REQUIRE(c && batch > 0 && grad_batch >= 0 && grad_batch <= batch, "srw_bert: bad batch (%d, grad %d)", batch, grad_batch);
  SRW_REQUIRE(c->hidden == 384 || c->hidden == 768 || c->hidden == 1024, "srw_bert: hidden must be 384, 768 or 1024 (got %d)", c->hidden);
  SRW_REQUIRE(c->heads * 64 == c->hidden, "srw_bert: heads * 64 must equal hidden");
  SRW_REQUIRE(c->intermediate % 64 == 0 && c->layers > 0 && c->layers <= 48 && c->num_classes > 0, "srw_bert: bad intermediate / layers / num_classes");
  SRW_REQUIRE(seq_len >= 16 && seq_len <= 512 && seq_len <= c->max_position, "srw_bert: 16 <= seq_len <= min(512, max_position) (got %d)", seq_len);
  SRW_REQUIRE((int64_t)batch * c->heads * seq_len * (int64_t)seq_len < (int64_t)1 << 32, "srw_bert: batch too large for the 32-bit dropout counters");
  d.S = batch; d.Sg = grad_batch; d.Lq = seq_len; d.D = c->hidden; d.H = c->heads; d.F = c->intermediate; d.NL = c->layers; d.C = c->num_classes;
  d.V = c->vocab_size;
  d.T = (int64_t)batch * seq_len; d.Tg = (int64_t)grad_batch * seq_len;
  d.ldb = (seq_len + 63) / 64 * 64;
  return SRW_OK;
}

struct BLayout {
  int64_t e, mean0, rstd0, key_bias, kv_len, qkv_bias;
  std::vector<int64_t> xf, xp;     // [NL + 1]: layer inputs / outputs, fp32 and planes
  std::vector<EncLayerBufs> lay;
  int64_t feat, z1, a1;
  // backward scratch
  int64_t dx, dfeat, dz1;
  EncBwdScratch bw;
  int64_t total;
};

static EncDims enc_dims(const BDims& d) {
  EncDims e;
  e.S = d.S; e.Sg = d.Sg; e.Lq = d.Lq; e.D = d.D; e.H = d.H; e.F = d.F; e.T = d.T; e.Tg = d.Tg; e.ld_bias = d.ldb;
  return e;
}

static BLayout make_blayout(const BDims& d) {
  BLayout L;
  Carver c;
  const EncDims ed = enc_dims(d);
  const int64_t T = d.T, D = d.D;
  L.e = c.take(T * D * 4); L.mean0 = c.take(T * 4); L.rstd0 = c.take(T * 4);
  L.key_bias = c.take((int64_t)d.S * d.ldb * 4); L.kv_len = c.take((int64_t)d.S * 4);
  L.qkv_bias = c.take((int64_t)d.NL * 3 * D * 4);
  L.xf.resize(d.NL + 1); L.xp.resize(d.NL + 1); L.lay.resize(d.NL);
  for (int l = 0; l <= d.NL; ++l) { L.xf[l] = c.take(T * D * 4); L.xp[l] = c.take(T * D * 4); }
  for (int l = 0; l < d.NL; ++l) enc_take_layer(c, ed, L.lay[l]);
  L.feat = c.take((int64_t)d.S * D * 4); L.z1 = c.take((int64_t)d.S * D * 4); L.a1 = c.take((int64_t)d.S * D * 4);
  const int64_t Tg = std::max<int64_t>(d.Tg, 1), Sg = std::max(d.Sg, 1);
  L.dx = c.take(Tg * D * 4); L.dfeat = c.take(Sg * D * 4); L.dz1 = c.take(Sg * D * 4);
  enc_take_bwd_scratch(c, ed, L.bw);
  L.total = c.off;
  return L;
}

// parameter indices (ClassificationBert.state_dict() order, include/srw.h)
enum { E_WORD = 0, E_POS, E_TYPE, E_LNW, E_LNB };
enum { Y_QW = 0, Y_QB, Y_KW, Y_KB, Y_VW, Y_VB, Y_OW, Y_OB, Y_LN1W, Y_LN1B, Y_F1W, Y_F1B, Y_F2W, Y_F2B, Y_LN2W, Y_LN2B };
static inline int play(int l, int which) { return 5 + 16 * l + which; }
static inline int pcls(int NL, int which) { return 5 + 16 * NL + which; }   // 0 pooler.w, 1 pooler.b, 2 classifier.0.w, 3 .b, 4 classifier.2.w, 5 .b
static int bert_num_params(const srw_bert_config* c) { return 5 + 16 * c->layers + 6; }

struct BWOff { int64_t qkv, o, f1, f2; };
static int64_t bert_weight_layout(const BDims& d, std::vector<BWOff>& w) {
  Carver c;
  w.resize(d.NL);
  for (int l = 0; l < d.NL; ++l) {
    w[l].qkv = c.take((int64_t)3 * d.D * d.D * 4);
    w[l].o = c.take((int64_t)d.D * d.D * 4);
    w[l].f1 = c.take((int64_t)d.F * d.D * 4);
    w[l].f2 = c.take((int64_t)d.D * d.F * 4);
  }
  return c.off;
}

static DropParams site_drop(const uint32_t* key, const int32_t* row, uint32_t site, double p) {
  srw_dropout d;
  d.seq_key = key; d.seq_row = row; d.site = site; d.p = p;
  return make_drop(d);
}
static EncLayerW layer_weights(const float* const* P, const uint8_t* wp, const std::vector<BWOff>& w, const float* qkv_bias, int l, int D) {
  EncLayerW lw;
  lw.qkv = wp + w[l].qkv; lw.o = wp + w[l].o; lw.f1 = wp + w[l].f1; lw.f2 = wp + w[l].f2;
  lw.qkv_bias = qkv_bias + (int64_t)l * 3 * D;
  lw.ob = P[play(l, Y_OB)]; lw.ln1w = P[play(l, Y_LN1W)]; lw.ln1b = P[play(l, Y_LN1B)]; lw.f1b = P[play(l, Y_F1B)]; lw.f2b = P[play(l, Y_F2B)];
  lw.ln2w = P[play(l, Y_LN2W)]; lw.ln2b = P[play(l, Y_LN2B)];
  return lw;
}
// dropout sites: 0 embeddings, 1 + 3l attention probabilities, 2 + 3l attention output, 3 + 3l FFN output, 1 + 3 NL pooled features
static EncDropSites layer_sites(const uint32_t* key, const int32_t* row, int l, const srw_bert_config& cf) {
  EncDropSites ds;
  ds.key = key; ds.row = row; ds.attn = 1 + 3 * l; ds.o = 2 + 3 * l; ds.f2 = 3 + 3 * l; ds.act = 0;
  ds.p_attn = cf.p_attn; ds.p_hidden = cf.p_hidden; ds.p_act = 0.0;
  return ds;
}

}  // namespace srw

using namespace srw;

extern "C" int64_t srw_bert_weight_planes_bytes(const srw_bert_config* c) {
  BDims d;
  if (make_bdims(c, 1, 16, 0, d)) return -1;
  std::vector<BWOff> w;
  return bert_weight_layout(d, w);
}

extern "C" int64_t srw_bert_workspace_bytes(const srw_bert_config* c, int batch, int seq_len, int grad_batch) {
  BDims d;
  if (make_bdims(c, batch, seq_len, grad_batch, d)) return -1;
  return make_blayout(d).total;
}

extern "C" int srw_bert_weight_plane_slot(const srw_bert_config* c, int param_index, int64_t* byte_offset, int* cols, int* ldp, int64_t* plane_stride) {
  BDims d;
  SRW_TRY(make_bdims(c, 1, 16, 0, d));
  SRW_REQUIRE(byte_offset && cols && ldp && plane_stride, "srw_bert_weight_plane_slot: null pointer");
  std::vector<BWOff> w;
  bert_weight_layout(d, w);
  const int rel = param_index - 5;
  if (rel >= 0 && rel < 16 * d.NL) {
    const int l = rel / 16, which = rel % 16;
    const int64_t D = d.D, F = d.F;
    int64_t off = -1, cc = 0, ps = 0;
    if (which == Y_QW) { off = w[l].qkv; cc = D; ps = 3 * D * D; }
    else if (which == Y_KW) { off = w[l].qkv + D * D * 2; cc = D; ps = 3 * D * D; }        // rows [D, 2D) of the packed [3D, D] hi plane (bf16)
    else if (which == Y_VW) { off = w[l].qkv + 2 * D * D * 2; cc = D; ps = 3 * D * D; }
    else if (which == Y_OW) { off = w[l].o; cc = D; ps = D * D; }
    else if (which == Y_F1W) { off = w[l].f1; cc = D; ps = F * D; }
    else if (which == Y_F2W) { off = w[l].f2; cc = F; ps = D * F; }
    if (off >= 0) {
      *byte_offset = off; *cols = (int)cc; *ldp = (int)cc; *plane_stride = ps;
      return SRW_OK;
    }
  }
  set_last_error("srw_bert_weight_plane_slot: parameter %d has no planes", param_index);
  return SRW_ERR_ARG;
}

extern "C" int srw_bert_prepare_weights(const srw_bert_config* c, const float* const* params, void* weight_planes, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  BDims d;
  SRW_TRY(make_bdims(c, 1, 16, 0, d));
  SRW_REQUIRE(params && weight_planes, "srw_bert_prepare_weights: null pointer");
  std::vector<BWOff> w;
  bert_weight_layout(d, w);
  uint8_t* base = reinterpret_cast<uint8_t*>(weight_planes);
  const int D = d.D, F = d.F;
  for (int l = 0; l < d.NL; ++l) {
    for (int i = 0; i < 3; ++i) {   // q, k, v -> rows [i D, (i + 1) D) of one [3D, D] plane pair
      srw_split_args a = {};
      a.x = params[play(l, Y_QW + 2 * i)]; a.ldx = D; a.rows = D; a.cols = D; a.rows_per_scale = 1;
      a.planes = base + w[l].qkv + (int64_t)i * D * D * 2; a.ldp = D; a.plane_stride = (int64_t)3 * D * D;
      SRW_TRY(srw_split_planes(&a, s));
    }
    SRW_TRY(split_to(params[play(l, Y_OW)], D, D, D, base + w[l].o, D, nullptr, 1, s));
    SRW_TRY(split_to(params[play(l, Y_F1W)], D, F, D, base + w[l].f1, D, nullptr, 1, s));
    SRW_TRY(split_to(params[play(l, Y_F2W)], F, D, F, base + w[l].f2, F, nullptr, 1, s));
  }
  return SRW_OK;
}

static int bert_forward_body(const srw_bert_fwd_args* a, cudaStream_t s) {
  SRW_REQUIRE(a && a->cfg && a->params && a->weight_planes && a->input_ids && a->logits && a->feat && a->workspace, "srw_bert_forward: null pointer");
  SRW_REQUIRE((a->drop_seq_key == nullptr) == (a->drop_seq_row == nullptr), "srw_bert_forward: drop_seq_key and drop_seq_row go together");
  BDims d;
  SRW_TRY(make_bdims(a->cfg, a->batch, a->seq_len, a->grad_batch, d));
  const BLayout L = make_blayout(d);
  SRW_REQUIRE(a->workspace_bytes >= L.total, "srw_bert_forward: workspace too small (%lld < %lld)", (long long)a->workspace_bytes, (long long)L.total);
  SRW_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 1023) == 0, "srw_bert_forward: workspace must be 1024-byte aligned");
  std::vector<BWOff> w;
  bert_weight_layout(d, w);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(a->weight_planes);
  const float* const* P = a->params;
  auto F32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const int T = (int)d.T, D = d.D, impl = a->gemm_impl, Lq = d.Lq;
  const float eps = a->cfg->ln_eps;
  const uint32_t* dk = a->drop_seq_key;
  const int32_t* drw = a->drop_seq_row;
  const srw_bert_config& cf = *a->cfg;

  // ---- key-padding bias, packed q|k|v biases ----
  SRW_TRY(srw_attn_mask_prepare(a->attention_mask, d.S, Lq, F32(L.key_bias), d.ldb, reinterpret_cast<int32_t*>(ws + L.kv_len), s));
  {
    QkvBiasPtrs bp = {};
    for (int l = 0; l < d.NL; ++l)
      for (int i = 0; i < 3; ++i) bp.p[3 * l + i] = P[play(l, Y_QB + 2 * i)];
    const int n = d.NL * 3 * D;
    bert_pack_qkv_bias_kernel<<<cdiv(n, 256), 256, 0, s>>>(bp, d.NL, D, F32(L.qkv_bias));
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // ---- embeddings ----
  {
    const DropParams dr = site_drop(dk, drw, 0, cf.p_hidden);
    __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(ws + L.xp[0]);
#define SRW_EMB(J) bert_embed_fwd_kernel<J><<<cdiv(T, 8), 256, 0, s>>>(a->input_ids, T, Lq, d.V, P[E_WORD], P[E_POS], P[E_TYPE], P[E_LNW], P[E_LNB], eps, \
                                                                      F32(L.e), F32(L.mean0), F32(L.rstd0), F32(L.xf[0]), xp, (int64_t)T * D, dr)
    if (D == 384) SRW_EMB(3); else if (D == 768) SRW_EMB(6); else SRW_EMB(8);
#undef SRW_EMB
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // ---- layers ----
  const EncDims ed = enc_dims(d);
  for (int l = 0; l < d.NL; ++l) {
    const EncLayerW lw = layer_weights(P, wp, w, F32(L.qkv_bias), l, D);
    const EncDropSites ds = layer_sites(dk, drw, l, cf);
    SRW_TRY(enc_layer_forward(ws, ed, 0, d.S, L.lay[l], lw, L.xf[l], L.xp[l], L.xf[l + 1], L.xp[l + 1], F32(L.key_bias),
                              reinterpret_cast<const int32_t*>(ws + L.kv_len), eps, 0.125f, ds, impl, s));
  }
  // ---- pooled features + classifier ----
  {
    const DropParams dr = site_drop(dk, drw, 1 + 3 * d.NL, cf.p_pooled);
    const float* cls[4] = {P[pcls(d.NL, 2)], P[pcls(d.NL, 3)], P[pcls(d.NL, 4)], P[pcls(d.NL, 5)]};
    SRW_TRY(enc_head_forward(F32(L.xf[d.NL]), d.S, Lq, D, d.C, dr, a->pool_len, cls, F32(L.feat), F32(L.z1), F32(L.a1), a->logits, a->feat, s));
  }
  return SRW_OK;
}

static int bert_backward_body(const srw_bert_bwd_args* a, cudaStream_t s) {
  SRW_REQUIRE(a && a->cfg && a->params && a->weight_planes && a->input_ids && a->dlogits && a->grads && a->workspace, "srw_bert_backward: null pointer");
  BDims d;
  SRW_TRY(make_bdims(a->cfg, a->batch, a->seq_len, a->grad_batch, d));
  SRW_REQUIRE(d.Sg > 0, "srw_bert_backward: grad_batch must be > 0");
  const BLayout L = make_blayout(d);
  SRW_REQUIRE(a->workspace_bytes >= L.total, "srw_bert_backward: workspace too small");
  std::vector<BWOff> w;
  bert_weight_layout(d, w);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(a->weight_planes);
  const float* const* P = a->params;
  float* const* G = a->grads;
  auto F32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const int Tg = (int)d.Tg, D = d.D, impl = a->gemm_impl, acc = a->accumulate_grads ? 1 : 0, Lq = d.Lq;
  const uint32_t* dk = a->drop_seq_key;
  const int32_t* drw = a->drop_seq_row;
  const srw_bert_config& cf = *a->cfg;
  const int64_t DD = (int64_t)D * D;
  for (int l = 0; l < d.NL; ++l)
    SRW_REQUIRE(G[play(l, Y_KW)] == G[play(l, Y_QW)] + DD && G[play(l, Y_VW)] == G[play(l, Y_QW)] + 2 * DD && G[play(l, Y_KB)] == G[play(l, Y_QB)] + D &&
                    G[play(l, Y_VB)] == G[play(l, Y_QB)] + 2 * D,
                "srw_bert_backward: the gradients of query / key / value weights (and of their biases) must be contiguous (layer %d)", l);
  const int lay_hi = (a->layer_hi < 0 || a->layer_hi >= d.NL) ? d.NL - 1 : a->layer_hi;
  const int lay_lo = (a->layer_hi < 0 || a->layer_lo < 0) ? 0 : a->layer_lo;
  SRW_REQUIRE(lay_lo <= lay_hi, "srw_bert_backward: empty layer range [%d, %d]", lay_lo, lay_hi);
  float* dx = F32(L.dx);
  float* du = F32(L.bw.du);

  // ---- classifier + pooling ----
  if (lay_hi == d.NL - 1) {
    const DropParams dr = site_drop(dk, drw, 1 + 3 * d.NL, cf.p_pooled);
    const float* cls[4] = {P[pcls(d.NL, 2)], P[pcls(d.NL, 3)], P[pcls(d.NL, 4)], P[pcls(d.NL, 5)]};
    float* gcls[4] = {G[pcls(d.NL, 2)], G[pcls(d.NL, 3)], G[pcls(d.NL, 4)], G[pcls(d.NL, 5)]};
    SRW_TRY(enc_head_backward(a->dlogits, a->dfeat, d.Sg, Lq, D, d.C, dr, a->pool_len, cls, gcls, F32(L.feat), F32(L.z1), F32(L.a1), F32(L.dz1),
                              F32(L.dfeat), dx, acc, s));
  }
  const EncDims ed = enc_dims(d);
  for (int l = lay_hi; l >= lay_lo; --l) {
    const EncLayerW lw = layer_weights(P, wp, w, F32(L.qkv_bias), l, D);
    const EncDropSites ds = layer_sites(dk, drw, l, cf);
    EncLayerG lg;
    lg.qkv_w = G[play(l, Y_QW)]; lg.qkv_b = G[play(l, Y_QB)]; lg.ow = G[play(l, Y_OW)]; lg.ob = G[play(l, Y_OB)];
    lg.ln1w = G[play(l, Y_LN1W)]; lg.ln1b = G[play(l, Y_LN1B)]; lg.f1w = G[play(l, Y_F1W)]; lg.f1b = G[play(l, Y_F1B)];
    lg.f2w = G[play(l, Y_F2W)]; lg.f2b = G[play(l, Y_F2B)]; lg.ln2w = G[play(l, Y_LN2W)]; lg.ln2b = G[play(l, Y_LN2B)];
    SRW_TRY(enc_layer_backward(ws, ed, 0, d.Sg, L.lay[l], lw, lg, L.xp[l], dx, L.bw, F32(L.key_bias), reinterpret_cast<const int32_t*>(ws + L.kv_len),
                               0.125f, ds, acc, impl, s));
  }
  // ---- embeddings ----
  if (lay_lo == 0) {
    const DropParams dr = site_drop(dk, drw, 0, cf.p_hidden);
    if (dr.on) {
      enc_dropout_rows_kernel<<<148 * 8, 256, 0, s>>>(dx, Tg, D, Lq, dr);
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    srw_layernorm_bwd_args lb = {};
    lb.dy = dx; lb.lddy = D; lb.x = F32(L.e); lb.ldx = D; lb.rows = Tg; lb.cols = D; lb.gamma = P[E_LNW]; lb.mean = F32(L.mean0); lb.rstd = F32(L.rstd0);
    lb.dx = du; lb.lddx = D; lb.accumulate_dx = 0; lb.dgamma = G[E_LNW]; lb.dbeta = G[E_LNB]; lb.accumulate_dparams = acc; lb.workspace = F32(L.bw.ln_ws);
    SRW_TRY(srw_layernorm_bwd(&lb, s));
    if (!acc) SRW_CUDA(cudaMemsetAsync(G[E_WORD], 0, (size_t)d.V * D * 4, s));
    bert_word_grad_kernel<<<148 * 8, 256, 0, s>>>(du, a->input_ids, Tg, D, d.V, G[E_WORD]);
    g_launches++;
    SRW_LAUNCH_CHECK();
    const int64_t np = (int64_t)cf.max_position * D;
    bert_pos_grad_kernel<<<(int)cdiv64(np, 256), 256, 0, s>>>(du, d.Sg, Lq, D, cf.max_position, F32(L.bw.dy), G[E_POS], acc);   // bw.dy is free here: scratch
    g_launches++;
    SRW_LAUNCH_CHECK();
    bert_type_grad_kernel<<<cdiv(D, 128), 128, 0, s>>>(F32(L.bw.dy), Lq, D, cf.type_vocab, G[E_TYPE], acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  return SRW_OK;
}

extern "C" int srw_bert_forward(const srw_bert_fwd_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->cfg && a->params, "srw_bert_forward: null pointer");
  if (!graphs_enabled(s) || a->cfg->layers <= 0 || a->cfg->layers > 48) return bert_forward_body(a, s);
  KeyBuilder kb;
  kb.add((int)11); kb.add(*a->cfg); kb.add(s);
  for (int i = 0; i < bert_num_params(a->cfg); ++i) kb.add(a->params[i]);
  kb.add(a->weight_planes); kb.add(a->input_ids); kb.add(a->attention_mask); kb.add(a->batch); kb.add(a->seq_len); kb.add(a->grad_batch);
  kb.add(a->drop_seq_key); kb.add(a->drop_seq_row); kb.add(a->pool_len); kb.add(a->logits); kb.add(a->feat); kb.add(a->workspace); kb.add(a->workspace_bytes); kb.add(a->gemm_impl);
  return run_graphed(std::move(kb.k), s, [a](cudaStream_t st) { return bert_forward_body(a, st); });
}

extern "C" int srw_bert_backward(const srw_bert_bwd_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->cfg && a->params && a->grads, "srw_bert_backward: null pointer");
  if (!graphs_enabled(s) || a->cfg->layers <= 0 || a->cfg->layers > 48) return bert_backward_body(a, s);
  KeyBuilder kb;
  kb.add((int)12); kb.add(*a->cfg); kb.add(s);
  for (int i = 0; i < bert_num_params(a->cfg); ++i) { kb.add(a->params[i]); kb.add(a->grads[i]); }
  kb.add(a->weight_planes); kb.add(a->input_ids); kb.add(a->attention_mask); kb.add(a->batch); kb.add(a->seq_len); kb.add(a->grad_batch);
  kb.add(a->drop_seq_key); kb.add(a->drop_seq_row); kb.add(a->pool_len); kb.add(a->dlogits); kb.add(a->dfeat); kb.add(a->accumulate_grads); kb.add(a->workspace);
  kb.add(a->workspace_bytes); kb.add(a->gemm_impl); kb.add(a->layer_lo); kb.add(a->layer_hi);
  return run_graphed(std::move(kb.k), s, [a](cudaStream_t st) { return bert_backward_body(a, st); });
}
