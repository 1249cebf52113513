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
#include "srw_engine.cuh"

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

// x[t, :] = dropout mask * x[t, :] / keep, in place (the gradient entering a dropout whose mask is regenerated)
__global__ void dropout_rows_kernel(float* __restrict__ x, int rows, int D, int Lq, const DropParams dr) {
  const int64_t total4 = (int64_t)rows * D / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 4;
    const int row = (int)(e / D), c = (int)(e % D);
    const int sq = row / Lq, l = row % Lq;
    const uint32_t key = drop_site_key(dr.seq_key[sq], dr.site);
    const uint32_t base = ((uint32_t)dr.seq_row[sq] * (uint32_t)Lq + (uint32_t)l) * (uint32_t)D + (uint32_t)c;
    float4 v = *reinterpret_cast<float4*>(x + e);
    v.x = drop_kept(key, base, dr.thr24) ? v.x * dr.inv_keep : 0.f;
    v.y = drop_kept(key, base + 1, dr.thr24) ? v.y * dr.inv_keep : 0.f;
    v.z = drop_kept(key, base + 2, dr.thr24) ? v.z * dr.inv_keep : 0.f;
    v.w = drop_kept(key, base + 3, dr.thr24) ? v.w * dr.inv_keep : 0.f;
    *reinterpret_cast<float4*>(x + e) = v;
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

// feat[s, c] = mean_l dropout(x[(s, l), c])   (bert.py:35-37: drop, then mean over every position)
__global__ void bert_pool_fwd_kernel(const float* __restrict__ x, int S, int Lq, int D, const DropParams dr, float* __restrict__ feat,
                                     const int32_t* __restrict__ pool_len) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (c >= D) return;
  const int Lp = pool_len ? min(max(pool_len[s], 1), Lq) : Lq;
  uint32_t key = 0, base = 0;
  if (dr.on) {
    key = drop_site_key(dr.seq_key[s], dr.site);
    base = (uint32_t)dr.seq_row[s] * (uint32_t)Lq * (uint32_t)D + (uint32_t)c;
  }
  float acc = 0.f;
  for (int l = 0; l < Lp; ++l) {
    float v = x[((int64_t)s * Lq + l) * D + c];
    if (dr.on) v = drop_kept(key, base + (uint32_t)l * (uint32_t)D, dr.thr24) ? v * dr.inv_keep : 0.f;
    acc += v;
  }
  feat[(int64_t)s * D + c] = acc / (float)Lp;
}
// dx[(s, l), c] = dropout mask * dfeat[s, c] / (keep * L)
__global__ void bert_pool_bwd_kernel(const float* __restrict__ dfeat, int Sg, int Lq, int D, const DropParams dr, float* __restrict__ dx,
                                     const int32_t* __restrict__ pool_len) {
  const int64_t total = (int64_t)Sg * Lq * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const int64_t row = i / D;
    const int l = (int)(row % Lq), s = (int)(row / Lq);
    const int Lp = pool_len ? min(max(pool_len[s], 1), Lq) : Lq;
    float v = l < Lp ? dfeat[(int64_t)s * D + c] / (float)Lp : 0.f;
    if (dr.on) {
      const uint32_t key = drop_site_key(dr.seq_key[s], dr.site);
      const uint32_t idx = ((uint32_t)dr.seq_row[s] * (uint32_t)Lq + (uint32_t)l) * (uint32_t)D + (uint32_t)c;
      v = drop_kept(key, idx, dr.thr24) ? v * dr.inv_keep : 0.f;
    }
    dx[i] = v;
  }
}

// classifier (bert.py:16-20): z1 = feat Wc1^T + bc1, a1 = gelu(z1) (exact erf), logits = a1 Wc2^T + bc2.  One CTA per sequence.
__global__ void __launch_bounds__(256) bert_head_fwd_kernel(const float* __restrict__ feat, int D, int C, const float* __restrict__ W1, const float* __restrict__ b1,
                                                            const float* __restrict__ W2, const float* __restrict__ b2, float* __restrict__ z1_out,
                                                            float* __restrict__ a1_out, float* __restrict__ logits) {
  extern __shared__ float sm[];
  float* f = sm;        // [D]
  float* a1 = sm + D;   // [D]
  const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < D; d += 256) f[d] = feat[(int64_t)s * D + d];
  __syncthreads();
  for (int n = warp; n < D; n += 8) {
    const float* w = W1 + (int64_t)n * D;
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(f[k], w[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float z = acc + b1[n];
      z1_out[(int64_t)s * D + n] = z;
      const float g = 0.5f * z * (1.0f + erff(z * 0.70710678118654752440f));
      a1[n] = g;
      a1_out[(int64_t)s * D + n] = g;
    }
  }
  __syncthreads();
  for (int c = warp; c < C; c += 8) {
    const float* w = W2 + (int64_t)c * D;
    float acc = 0.f;
    for (int k = lane; k < D; k += 32) acc = fmaf(a1[k], w[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[(int64_t)s * C + c] = acc + b2[c];
  }
}
// per sequence: da1 = dlogits W2, dz1 = da1 gelu'(z1), dfeat = dz1 W1 (+ dfeat_in)
__global__ void __launch_bounds__(256) bert_head_bwd_rows_kernel(const float* __restrict__ dlogits, const float* __restrict__ dfeat_in, int D, int C,
                                                                 const float* __restrict__ W1, const float* __restrict__ W2, const float* __restrict__ z1,
                                                                 float* __restrict__ dz1_out, float* __restrict__ dfeat_out) {
  extern __shared__ float sm[];
  float* dl = sm;        // [C]
  float* dz = sm + C;    // [D]
  const int s = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += 256) dl[c] = dlogits[(int64_t)s * C + c];
  __syncthreads();
  for (int n = threadIdx.x; n < D; n += 256) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(dl[c], W2[(int64_t)c * D + n], acc);
    const float z = z1[(int64_t)s * D + n];
    const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    const float v = acc * (cdf + z * pdf);
    dz[n] = v;
    dz1_out[(int64_t)s * D + n] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < D; k += 256) {
    float acc = dfeat_in ? dfeat_in[(int64_t)s * D + k] : 0.f;
    for (int n = 0; n < D; ++n) acc = fmaf(dz[n], W1[(int64_t)n * D + k], acc);
    dfeat_out[(int64_t)s * D + k] = acc;
  }
}
// dW2[c, k] = sum_s dlogits[s, c] a1[s, k]; db2; dW1[n, k] = sum_s dz1[s, n] feat[s, k]; db1
__global__ void bert_head_bwd_params_kernel(const float* __restrict__ dlogits, const float* __restrict__ a1, const float* __restrict__ dz1,
                                            const float* __restrict__ feat, int Sg, int D, int C, float* __restrict__ dW1, float* __restrict__ db1,
                                            float* __restrict__ dW2, float* __restrict__ db2, int accumulate) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n1 = (int64_t)D * D, n2 = (int64_t)C * D;
  if (i < n1) {
    const int n = (int)(i / D), k = (int)(i % D);
    float acc = 0.f;
    for (int s = 0; s < Sg; ++s) acc = fmaf(dz1[(int64_t)s * D + n], feat[(int64_t)s * D + k], acc);
    dW1[i] = accumulate ? dW1[i] + acc : acc;
  } else if (i < n1 + n2) {
    const int64_t j = i - n1;
    const int c = (int)(j / D), k = (int)(j % D);
    float acc = 0.f;
    for (int s = 0; s < Sg; ++s) acc = fmaf(dlogits[(int64_t)s * C + c], a1[(int64_t)s * D + k], acc);
    dW2[j] = accumulate ? dW2[j] + acc : acc;
  } else if (i < n1 + n2 + D) {
    const int n = (int)(i - n1 - n2);
    float acc = 0.f;
    for (int s = 0; s < Sg; ++s) acc += dz1[(int64_t)s * D + n];
    db1[n] = accumulate ? db1[n] + acc : acc;
  } else if (i < n1 + n2 + D + C) {
    const int c = (int)(i - n1 - n2 - D);
    float acc = 0.f;
    for (int s = 0; s < Sg; ++s) acc += dlogits[(int64_t)s * C + c];
    db2[c] = accumulate ? db2[c] + acc : acc;
  }
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

struct BLayerBufs { int64_t qkv, o, lse, u1, mean1, rstd1, xm, xmp, z, h, u2, mean2, rstd2; };
struct BLayout {
  int64_t e, mean0, rstd0, key_bias, kv_len, qkv_bias;
  std::vector<int64_t> xf, xp;     // [NL + 1]: layer inputs / outputs, fp32 and planes
  std::vector<BLayerBufs> lay;
  int64_t feat, z1, a1;
  // backward scratch
  int64_t dx, du, dy, g, dz, d_o, dqkv, delta, dfeat, dz1, colsum_ws, colsum_ws2, ln_ws, ln_ws2, sk4[4];
  int64_t total;
};

static BLayout make_blayout(const BDims& d) {
  BLayout L;
  Carver c;
  const int64_t T = d.T, D = d.D, F = d.F;
  L.e = c.take(T * D * 4); L.mean0 = c.take(T * 4); L.rstd0 = c.take(T * 4);
  L.key_bias = c.take((int64_t)d.S * d.ldb * 4); L.kv_len = c.take((int64_t)d.S * 4);
  L.qkv_bias = c.take((int64_t)d.NL * 3 * D * 4);
  L.xf.resize(d.NL + 1); L.xp.resize(d.NL + 1); L.lay.resize(d.NL);
  for (int l = 0; l <= d.NL; ++l) { L.xf[l] = c.take(T * D * 4); L.xp[l] = c.take(T * D * 4); }
  for (int l = 0; l < d.NL; ++l) {
    BLayerBufs& b = L.lay[l];
    b.qkv = c.take(T * 3 * D * 4); b.o = c.take(T * D * 4); b.lse = c.take((int64_t)d.S * d.H * d.Lq * 4);
    b.u1 = c.take(T * D * 4); b.mean1 = c.take(T * 4); b.rstd1 = c.take(T * 4);
    b.xm = c.take(T * D * 4); b.xmp = c.take(T * D * 4);
    b.z = c.take(T * F * 4); b.h = c.take(T * F * 4);
    b.u2 = c.take(T * D * 4); b.mean2 = c.take(T * 4); b.rstd2 = c.take(T * 4);
  }
  L.feat = c.take((int64_t)d.S * D * 4); L.z1 = c.take((int64_t)d.S * D * 4); L.a1 = c.take((int64_t)d.S * D * 4);
  const int64_t Tg = std::max<int64_t>(d.Tg, 1), Sg = std::max(d.Sg, 1);
  L.dx = c.take(Tg * D * 4); L.du = c.take(Tg * D * 4); L.dy = c.take(Tg * D * 4); L.g = c.take(Tg * D * 4);
  L.dz = c.take(Tg * F * 4); L.d_o = c.take(Tg * D * 4); L.dqkv = c.take(Tg * 3 * D * 4);
  L.delta = c.take(Sg * d.H * d.Lq * 4); L.dfeat = c.take(Sg * D * 4); L.dz1 = c.take(Sg * D * 4);
  L.colsum_ws = c.take((int64_t)256 * std::max<int64_t>(3 * D, F) * 4);
  L.colsum_ws2 = c.take((int64_t)256 * std::max<int64_t>(3 * D, F) * 4);
  L.ln_ws = c.take((int64_t)3 * 256 * D * 4); L.ln_ws2 = c.take((int64_t)3 * 256 * D * 4);
  L.sk4[0] = c.take(splitk_for(d.D, d.F, Tg, nullptr) * 4);
  L.sk4[1] = c.take(splitk_for(d.F, d.D, Tg, nullptr) * 4);
  L.sk4[2] = c.take(splitk_for(d.D, d.D, Tg, nullptr) * 4);
  L.sk4[3] = c.take(splitk_for(3 * d.D, d.D, Tg, nullptr) * 4);
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
static srw_dropout site_spec(const uint32_t* key, const int32_t* row, uint32_t site, double p) {
  srw_dropout d;
  d.seq_key = key; d.seq_row = row; d.site = site; d.p = key ? p : 0.0;
  return d;
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
  const int T = (int)d.T, D = d.D, Fh = d.F, impl = a->gemm_impl, Lq = d.Lq;
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
  for (int l = 0; l < d.NL; ++l) {
    const BLayerBufs& b = L.lay[l];
    {
      Gemm g(T, 3 * D, D, impl);
      g.A(ws + L.xp[l], D, T, 0).Bm(wp + w[l].qkv, D, 3 * D, 0);
      g.g.epilogue = SRW_EPI_PLANES; g.g.bias = F32(L.qkv_bias) + (int64_t)l * 3 * D; g.g.out_planes = ws + b.qkv; g.g.ldp = 3 * D;
      g.g.out_plane_stride = (int64_t)T * 3 * D;
      SRW_TRY(g.run(s));
    }
    {
      srw_attn_fwd_args at = {};
      at.B = d.S; at.N = Lq; at.H = d.H; at.head_dim = 64; at.scale = 0.125f;
      at.qkv = ws + b.qkv; at.ld_qkv = 3 * D; at.qkv_plane_stride = (int64_t)T * 3 * D;
      at.o = ws + b.o; at.ld_o = D; at.o_plane_stride = (int64_t)T * D; at.lse = F32(b.lse);
      at.key_bias = F32(L.key_bias); at.ld_bias = d.ldb; at.kv_len = reinterpret_cast<const int32_t*>(ws + L.kv_len);
      at.drop = site_spec(dk, drw, 1 + 3 * l, cf.p_attn);
      SRW_TRY(srw_attn_fwd(&at, s));
    }
    {
      Gemm g(T, D, D, impl);   // u1 = x + dropout(o Wo^T + bo)
      g.A(ws + b.o, D, T, 0).Bm(wp + w[l].o, D, D, 0);
      g.g.epilogue = SRW_EPI_RESID; g.g.bias = P[play(l, Y_OB)]; g.g.resid = F32(L.xf[l]); g.g.ldr = D; g.g.out_f32 = F32(b.u1); g.g.ldo = D;
      g.g.drop = site_spec(dk, drw, 2 + 3 * l, cf.p_hidden); g.g.drop_rows_per_seq = Lq;
      SRW_TRY(g.run(s));
    }
    srw_layernorm_fwd_args ln = {};
    ln.x = F32(b.u1); ln.ldx = D; ln.rows = T; ln.cols = D; ln.eps = eps; ln.gamma = P[play(l, Y_LN1W)]; ln.beta = P[play(l, Y_LN1B)];
    ln.mean = F32(b.mean1); ln.rstd = F32(b.rstd1); ln.y_planes = ws + b.xmp; ln.ldp = D; ln.plane_stride = (int64_t)T * D;
    ln.y_f32 = F32(b.xm); ln.ldy = D;
    SRW_TRY(srw_layernorm_fwd(&ln, s));
    {
      Gemm g(T, Fh, D, impl);
      g.A(ws + b.xmp, D, T, 0).Bm(wp + w[l].f1, D, Fh, 0);
      g.g.epilogue = SRW_EPI_GELU; g.g.bias = P[play(l, Y_F1B)]; g.g.out_f32 = F32(b.z); g.g.ldo = Fh; g.g.out_planes = ws + b.h; g.g.ldp = Fh;
      g.g.out_plane_stride = (int64_t)T * Fh;
      SRW_TRY(g.run(s));
    }
    {
      Gemm g(T, D, Fh, impl);   // u2 = x_mid + dropout(h W2^T + b2)
      g.A(ws + b.h, Fh, T, 0).Bm(wp + w[l].f2, Fh, D, 0);
      g.g.epilogue = SRW_EPI_RESID; g.g.bias = P[play(l, Y_F2B)]; g.g.resid = F32(b.xm); g.g.ldr = D; g.g.out_f32 = F32(b.u2); g.g.ldo = D;
      g.g.drop = site_spec(dk, drw, 3 + 3 * l, cf.p_hidden); g.g.drop_rows_per_seq = Lq;
      SRW_TRY(g.run(s));
    }
    ln.x = F32(b.u2); ln.gamma = P[play(l, Y_LN2W)]; ln.beta = P[play(l, Y_LN2B)]; ln.mean = F32(b.mean2); ln.rstd = F32(b.rstd2);
    ln.y_planes = ws + L.xp[l + 1]; ln.y_f32 = F32(L.xf[l + 1]);
    SRW_TRY(srw_layernorm_fwd(&ln, s));
  }
  // ---- pooled features + classifier ----
  {
    const DropParams dr = site_drop(dk, drw, 1 + 3 * d.NL, cf.p_pooled);
    bert_pool_fwd_kernel<<<dim3(cdiv(D, 256), d.S), 256, 0, s>>>(F32(L.xf[d.NL]), d.S, Lq, D, dr, F32(L.feat), a->pool_len);
    g_launches++;
    SRW_LAUNCH_CHECK();
    bert_head_fwd_kernel<<<d.S, 256, 2 * D * sizeof(float), s>>>(F32(L.feat), D, d.C, P[pcls(d.NL, 2)], P[pcls(d.NL, 3)], P[pcls(d.NL, 4)], P[pcls(d.NL, 5)],
                                                                  F32(L.z1), F32(L.a1), a->logits);
    g_launches++;
    SRW_LAUNCH_CHECK();
    SRW_CUDA(cudaMemcpyAsync(a->feat, F32(L.feat), (size_t)d.S * D * 4, cudaMemcpyDeviceToDevice, s));
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
  const int T = (int)d.T, Tg = (int)d.Tg, D = d.D, Fh = d.F, impl = a->gemm_impl, acc = a->accumulate_grads ? 1 : 0, Lq = d.Lq;
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
  float* du = F32(L.du);

  // ---- classifier + pooling ----
  if (lay_hi == d.NL - 1) {
    bert_head_bwd_rows_kernel<<<d.Sg, 256, (d.C + D) * sizeof(float), s>>>(a->dlogits, a->dfeat, D, d.C, P[pcls(d.NL, 2)], P[pcls(d.NL, 4)], F32(L.z1),
                                                                            F32(L.dz1), F32(L.dfeat));
    g_launches++;
    SRW_LAUNCH_CHECK();
    const int64_t n = (int64_t)D * D + (int64_t)d.C * D + D + d.C;
    bert_head_bwd_params_kernel<<<(int)cdiv64(n, 256), 256, 0, s>>>(a->dlogits, F32(L.a1), F32(L.dz1), F32(L.feat), d.Sg, D, d.C, G[pcls(d.NL, 2)],
                                                                    G[pcls(d.NL, 3)], G[pcls(d.NL, 4)], G[pcls(d.NL, 5)], acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
    const DropParams dr = site_drop(dk, drw, 1 + 3 * d.NL, cf.p_pooled);
    bert_pool_bwd_kernel<<<148 * 8, 256, 0, s>>>(F32(L.dfeat), d.Sg, Lq, D, dr, dx, a->pool_len);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  const int ln_parts = srw_layernorm_bwd_nparts(Tg);
  for (int l = lay_hi; l >= lay_lo; --l) {
    const BLayerBufs& b = L.lay[l];
    srw_grad_fold_args fold_args = {};
    srw_grad_fold_args* fold = &fold_args;
    // ---- FFN branch: x_out = LN2(u2), u2 = x_mid + dropout(gelu(x_mid W1^T + b1) W2^T + b2) ----
    srw_layernorm_bwd_args lb = {};
    lb.dy = dx; lb.lddy = D; lb.x = F32(b.u2); lb.ldx = D; lb.rows = Tg; lb.cols = D; lb.gamma = P[play(l, Y_LN2W)];
    lb.mean = F32(b.mean2); lb.rstd = F32(b.rstd2); lb.dx = du; lb.lddx = D; lb.accumulate_dx = 0; lb.accumulate_dparams = acc;
    lb.workspace = F32(L.ln_ws);
    lb.dx_planes = ws + L.g; lb.ldp = D; lb.plane_stride = (int64_t)Tg * D; lb.rows_per_scale = 1;
    lb.drop = site_spec(dk, drw, 3 + 3 * l, cf.p_hidden); lb.drop_rows_per_seq = Lq;
    fold_colsum(fold, lb.workspace, ln_parts, 3 * (int64_t)D, D, G[play(l, Y_LN2W)], acc);
    fold_colsum(fold, lb.workspace + D, ln_parts, 3 * (int64_t)D, D, G[play(l, Y_LN2B)], acc);
    fold_colsum(fold, lb.workspace + 2 * D, ln_parts, 3 * (int64_t)D, D, G[play(l, Y_F2B)], acc);
    SRW_TRY(srw_layernorm_bwd(&lb, s));
    SRW_TRY(wgrad(D, Fh, Tg, ws + L.g, D, Tg, ws + b.h, Fh, T, F32(L.sk4[0]), G[play(l, Y_F2W)], Fh, acc, impl, s, fold));
    {
      Gemm g(Tg, Fh, D, impl);  // dz = (g W2) * gelu'(z)
      g.A(ws + L.g, D, Tg, 0).Bm(wp + w[l].f2, Fh, D, 1);
      g.g.epilogue = SRW_EPI_DGELU; g.g.aux = F32(b.z); g.g.ldaux = Fh; g.g.out_planes = ws + L.dz; g.g.ldp = Fh; g.g.out_plane_stride = (int64_t)Tg * Fh;
      SRW_TRY(g.run(s));
    }
    SRW_TRY(colsum_planes(ws + L.dz, Fh, Tg, Tg, Fh, G[play(l, Y_F1B)], acc, F32(L.colsum_ws), s, fold));
    SRW_TRY(wgrad(Fh, D, Tg, ws + L.dz, Fh, Tg, ws + b.xmp, D, T, F32(L.sk4[1]), G[play(l, Y_F1W)], D, acc, impl, s, fold));
    {
      Gemm g(Tg, D, Fh, impl);  // dy = du + dz W1   (gradient of x_mid: residual path + FFN path)
      g.A(ws + L.dz, Fh, Tg, 0).Bm(wp + w[l].f1, D, Fh, 1);
      g.g.epilogue = SRW_EPI_RESID; g.g.resid = du; g.g.ldr = D; g.g.out_f32 = F32(L.dy); g.g.ldo = D;
      SRW_TRY(g.run(s));
    }
    // ---- attention branch: x_mid = LN1(u1), u1 = x_in + dropout(attn(x_in) Wo^T + bo) ----
    lb.dy = F32(L.dy); lb.x = F32(b.u1); lb.gamma = P[play(l, Y_LN1W)]; lb.mean = F32(b.mean1); lb.rstd = F32(b.rstd1);
    lb.workspace = F32(L.ln_ws2);
    lb.drop = site_spec(dk, drw, 2 + 3 * l, cf.p_hidden);
    fold_colsum(fold, lb.workspace, ln_parts, 3 * (int64_t)D, D, G[play(l, Y_LN1W)], acc);
    fold_colsum(fold, lb.workspace + D, ln_parts, 3 * (int64_t)D, D, G[play(l, Y_LN1B)], acc);
    fold_colsum(fold, lb.workspace + 2 * D, ln_parts, 3 * (int64_t)D, D, G[play(l, Y_OB)], acc);
    SRW_TRY(srw_layernorm_bwd(&lb, s));
    SRW_TRY(wgrad(D, D, Tg, ws + L.g, D, Tg, ws + b.o, D, T, F32(L.sk4[2]), G[play(l, Y_OW)], D, acc, impl, s, fold));
    {
      Gemm g(Tg, D, D, impl);  // d_o = g Wo
      g.A(ws + L.g, D, Tg, 0).Bm(wp + w[l].o, D, D, 1);
      g.g.epilogue = SRW_EPI_PLANES; g.g.out_planes = ws + L.d_o; g.g.ldp = D; g.g.out_plane_stride = (int64_t)Tg * D;
      SRW_TRY(g.run(s));
    }
    {
      srw_attn_bwd_args at = {};
      at.B = d.Sg; at.N = Lq; at.H = d.H; at.head_dim = 64; at.scale = 0.125f;
      at.qkv = ws + b.qkv; at.ld_qkv = 3 * D; at.qkv_plane_stride = (int64_t)T * 3 * D;
      at.o = ws + b.o; at.ld_o = D; at.o_plane_stride = (int64_t)T * D;
      at.d_o = ws + L.d_o; at.ld_do = D; at.do_plane_stride = (int64_t)Tg * D;
      at.lse = F32(b.lse); at.delta = F32(L.delta);
      at.dqkv = ws + L.dqkv; at.ld_dqkv = 3 * D; at.dqkv_plane_stride = (int64_t)Tg * 3 * D;
      at.key_bias = F32(L.key_bias); at.ld_bias = d.ldb; at.kv_len = reinterpret_cast<const int32_t*>(ws + L.kv_len);
      at.drop = site_spec(dk, drw, 1 + 3 * l, cf.p_attn);
      SRW_TRY(srw_attn_bwd(&at, s));
    }
    SRW_TRY(colsum_planes(ws + L.dqkv, 3 * D, Tg, Tg, 3 * D, G[play(l, Y_QB)], acc, F32(L.colsum_ws2), s, fold));   // q | k | v biases are contiguous
    SRW_TRY(wgrad(3 * D, D, Tg, ws + L.dqkv, 3 * D, Tg, ws + L.xp[l], D, T, F32(L.sk4[3]), G[play(l, Y_QW)], D, acc, impl, s, fold));
    {
      Gemm g(Tg, D, 3 * D, impl);  // dx(layer input) = du + dqkv Wqkv
      g.A(ws + L.dqkv, 3 * D, Tg, 0).Bm(wp + w[l].qkv, D, 3 * D, 1);
      g.g.epilogue = SRW_EPI_RESID; g.g.resid = du; g.g.ldr = D; g.g.out_f32 = dx; g.g.ldo = D;
      SRW_TRY(g.run(s));
    }
    SRW_TRY(srw_grad_fold(fold, s));
  }
  // ---- embeddings ----
  if (lay_lo == 0) {
    const DropParams dr = site_drop(dk, drw, 0, cf.p_hidden);
    if (dr.on) {
      dropout_rows_kernel<<<148 * 8, 256, 0, s>>>(dx, Tg, D, Lq, dr);
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    srw_layernorm_bwd_args lb = {};
    lb.dy = dx; lb.lddy = D; lb.x = F32(L.e); lb.ldx = D; lb.rows = Tg; lb.cols = D; lb.gamma = P[E_LNW]; lb.mean = F32(L.mean0); lb.rstd = F32(L.rstd0);
    lb.dx = du; lb.lddx = D; lb.accumulate_dx = 0; lb.dgamma = G[E_LNW]; lb.dbeta = G[E_LNB]; lb.accumulate_dparams = acc; lb.workspace = F32(L.ln_ws);
    SRW_TRY(srw_layernorm_bwd(&lb, s));
    if (!acc) SRW_CUDA(cudaMemsetAsync(G[E_WORD], 0, (size_t)d.V * D * 4, s));
    bert_word_grad_kernel<<<148 * 8, 256, 0, s>>>(du, a->input_ids, Tg, D, d.V, G[E_WORD]);
    g_launches++;
    SRW_LAUNCH_CHECK();
    const int64_t np = (int64_t)cf.max_position * D;
    bert_pos_grad_kernel<<<(int)cdiv64(np, 256), 256, 0, s>>>(du, d.Sg, Lq, D, cf.max_position, F32(L.dy), G[E_POS], acc);   // L.dy is free here: scratch
    g_launches++;
    SRW_LAUNCH_CHECK();
    bert_type_grad_kernel<<<cdiv(D, 128), 128, 0, s>>>(F32(L.dy), Lq, D, cf.type_vocab, G[E_TYPE], acc);
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
