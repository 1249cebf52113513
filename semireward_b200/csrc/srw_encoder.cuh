// srw_encoder.cuh — the post-LN transformer encoder layer shared by the text and audio engines (srw_bert.cu, srw_hubert.cu), its
// backward, and the pooled classifier head both wrappers put on top (semilearn/nets/bert/bert.py:16-20,35-48 and
// semilearn/nets/hubert/hubert.py:17-21,44-49 are the same three lines: dropout -> mean over positions -> Linear-GELU-Linear).
//
//   forward:   [q|k|v] GEMM (packed [3D, D] weight) -> planes | attention (optional key-padding bias, dropout on the probabilities) |
//              Wo GEMM + dropout + residual -> u1 | LN -> x_mid | W1 GEMM + GELU (+ activation dropout) | W2 GEMM + dropout + residual
//              -> u2 | LN -> x_out                       (HF modeling_bert.py BertLayer; modeling_hubert.py HubertEncoderLayer)
//   backward:  mirrored; every LayerNorm backward hands du to the residual path and, through the dropout mask, to the branch.
// Both work on a RANGE of sequences [seq0, seq0 + nseq) of the launch: LayerDrop (HubertEncoder) skips a layer per model call, and
// one launch carries the three calls of `use_cat: False` as consecutive sequence ranges.
#pragma once
#include "srw_engine.cuh"

namespace srw {

struct EncDims { int S, Sg, Lq, D, H, F; int64_t T, Tg; int64_t ld_bias; };
struct EncLayerBufs { int64_t qkv, o, lse, u1, mean1, rstd1, xm, xmp, z, h, u2, mean2, rstd2; };   // byte offsets into the workspace
struct EncLayerW {          // one layer's weights: plane cache pointers + fp32 vectors
  const uint8_t *qkv, *o, *f1, *f2;
  const float *qkv_bias, *ob, *ln1w, *ln1b, *f1b, *f2b, *ln2w, *ln2b;
};
struct EncLayerG {          // gradient destinations; qkv_w = [3D, D] and qkv_b = [3D] are the packed q|k|v pieces
  float *qkv_w, *qkv_b, *ow, *ob, *ln1w, *ln1b, *f1w, *f1b, *f2w, *f2b, *ln2w, *ln2b;
};
struct EncDropSites {       // counter-dropout sites of the layer (include/srw.h: srw_dropout); p = 0 switches a site off
  const uint32_t* key; const int32_t* row;
  uint32_t attn, o, act, f2;
  double p_attn, p_hidden, p_act;
};
struct EncBwdScratch { int64_t du, dy, g, dz, d_o, dqkv, delta, colsum_ws, colsum_ws2, ln_ws, ln_ws2, sk4[4]; };

static inline srw_dropout enc_site(const EncDropSites& s, uint32_t site, double p, int seq0) {
  srw_dropout d;
  d.seq_key = s.key ? s.key + seq0 : nullptr; d.seq_row = s.row ? s.row + seq0 : nullptr; d.site = site; d.p = s.key ? p : 0.0;
  return d;
}

// x_out = layer(x_in) for sequences [seq0, seq0 + nseq).  xf / xp: fp32 activations and their planes ([T, D], plane stride T * D).
static inline int enc_layer_forward(uint8_t* ws, const EncDims& d, int seq0, int nseq, const EncLayerBufs& b, const EncLayerW& w, int64_t xf_in, int64_t xp_in,
                                    int64_t xf_out, int64_t xp_out, const float* key_bias, const int32_t* kv_len, float eps, float attn_scale,
                                    const EncDropSites& dr, int impl, cudaStream_t s) {
  auto F32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const int D = d.D, Fh = d.F, Lq = d.Lq;
  const int64_t T = d.T, r0 = (int64_t)seq0 * Lq;
  const int Tr = nseq * Lq;
  {
    Gemm g(Tr, 3 * D, D, impl);
    g.A(ws + xp_in + r0 * D * 2, D, T, 0).Bm(w.qkv, D, 3 * D, 0);
    g.g.epilogue = SRW_EPI_PLANES; g.g.bias = w.qkv_bias; g.g.out_planes = ws + b.qkv + r0 * 3 * D * 2; g.g.ldp = 3 * D;
    g.g.out_plane_stride = T * 3 * D;
    SRW_TRY(g.run(s));
  }
  {
    srw_attn_fwd_args at = {};
    at.B = nseq; at.N = Lq; at.H = d.H; at.head_dim = 64; at.scale = attn_scale;
    at.qkv = ws + b.qkv + r0 * 3 * D * 2; at.ld_qkv = 3 * D; at.qkv_plane_stride = T * 3 * D;
    at.o = ws + b.o + r0 * D * 2; at.ld_o = D; at.o_plane_stride = T * D; at.lse = F32(b.lse) + (int64_t)seq0 * d.H * Lq;
    if (key_bias) { at.key_bias = key_bias + (int64_t)seq0 * d.ld_bias; at.ld_bias = d.ld_bias; at.kv_len = kv_len ? kv_len + seq0 : nullptr; }
    at.drop = enc_site(dr, dr.attn, dr.p_attn, seq0);
    SRW_TRY(srw_attn_fwd(&at, s));
  }
  {
    Gemm g(Tr, D, D, impl);   // u1 = x + dropout(o Wo^T + bo)
    g.A(ws + b.o + r0 * D * 2, D, T, 0).Bm(w.o, D, D, 0);
    g.g.epilogue = SRW_EPI_RESID; g.g.bias = w.ob; g.g.resid = F32(xf_in) + r0 * D; g.g.ldr = D; g.g.out_f32 = F32(b.u1) + r0 * D; g.g.ldo = D;
    g.g.drop = enc_site(dr, dr.o, dr.p_hidden, seq0); g.g.drop_rows_per_seq = Lq;
    SRW_TRY(g.run(s));
  }
  srw_layernorm_fwd_args ln = {};
  ln.x = F32(b.u1) + r0 * D; ln.ldx = D; ln.rows = Tr; ln.cols = D; ln.eps = eps; ln.gamma = w.ln1w; ln.beta = w.ln1b;
  ln.mean = F32(b.mean1) + r0; ln.rstd = F32(b.rstd1) + r0; ln.y_planes = ws + b.xmp + r0 * D * 2; ln.ldp = D; ln.plane_stride = T * D;
  ln.y_f32 = F32(b.xm) + r0 * D; ln.ldy = D;
  SRW_TRY(srw_layernorm_fwd(&ln, s));
  {
    Gemm g(Tr, Fh, D, impl);
    g.A(ws + b.xmp + r0 * D * 2, D, T, 0).Bm(w.f1, D, Fh, 0);
    g.g.epilogue = SRW_EPI_GELU; g.g.bias = w.f1b; g.g.out_f32 = F32(b.z) + r0 * Fh; g.g.ldo = Fh; g.g.out_planes = ws + b.h + r0 * Fh * 2; g.g.ldp = Fh;
    g.g.out_plane_stride = T * Fh;
    g.g.drop = enc_site(dr, dr.act, dr.p_act, seq0); g.g.drop_rows_per_seq = Lq;
    SRW_TRY(g.run(s));
  }
  {
    Gemm g(Tr, D, Fh, impl);   // u2 = x_mid + dropout(h W2^T + b2)
    g.A(ws + b.h + r0 * Fh * 2, Fh, T, 0).Bm(w.f2, Fh, D, 0);
    g.g.epilogue = SRW_EPI_RESID; g.g.bias = w.f2b; g.g.resid = F32(b.xm) + r0 * D; g.g.ldr = D; g.g.out_f32 = F32(b.u2) + r0 * D; g.g.ldo = D;
    g.g.drop = enc_site(dr, dr.f2, dr.p_hidden, seq0); g.g.drop_rows_per_seq = Lq;
    SRW_TRY(g.run(s));
  }
  ln.x = F32(b.u2) + r0 * D; ln.gamma = w.ln2w; ln.beta = w.ln2b; ln.mean = F32(b.mean2) + r0; ln.rstd = F32(b.rstd2) + r0;
  ln.y_planes = ws + xp_out + r0 * D * 2; ln.y_f32 = F32(xf_out) + r0 * D;
  SRW_TRY(srw_layernorm_fwd(&ln, s));
  return SRW_OK;
}

// dx (in place, [Tg, D] fp32: gradient of the layer output on entry, of the layer input on exit) for sequences [seq0, seq0 + nseq) of
// the gradient rows; parameter gradients are written (acc == 0) or accumulated (acc != 0) into G.  The scratch buffers of B are
// sized for all Tg rows and indexed by the same row offsets, so disjoint ranges may be processed one after the other.
static inline int enc_layer_backward(uint8_t* ws, const EncDims& d, int seq0, int nseq, const EncLayerBufs& b, const EncLayerW& w, const EncLayerG& G,
                                     int64_t xp_in, float* dx_all, const EncBwdScratch& B, const float* key_bias, const int32_t* kv_len, float attn_scale,
                                     const EncDropSites& dr, int acc, int impl, cudaStream_t s) {
  auto F32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  const int D = d.D, Fh = d.F, Lq = d.Lq;
  const int64_t T = d.T, Tg = d.Tg, r0 = (int64_t)seq0 * Lq;
  const int Tr = nseq * Lq;
  float* dx = dx_all + r0 * D;
  float* du = F32(B.du) + r0 * D;
  uint8_t* gp = ws + B.g + r0 * D * 2;            // planes [Tg, D]
  uint8_t* dzp = ws + B.dz + r0 * Fh * 2;         // planes [Tg, F]
  uint8_t* dop = ws + B.d_o + r0 * D * 2;         // planes [Tg, D]
  uint8_t* dqkvp = ws + B.dqkv + r0 * 3 * D * 2;  // planes [Tg, 3D]
  const int ln_parts = srw_layernorm_bwd_nparts(Tr);
  srw_grad_fold_args fold_args = {};
  srw_grad_fold_args* fold = &fold_args;
  // ---- FFN branch: x_out = LN2(u2), u2 = x_mid + dropout(drop_act(gelu(x_mid W1^T + b1)) W2^T + b2) ----
  srw_layernorm_bwd_args lb = {};
  lb.dy = dx; lb.lddy = D; lb.x = F32(b.u2) + r0 * D; lb.ldx = D; lb.rows = Tr; lb.cols = D; lb.gamma = w.ln2w;
  lb.mean = F32(b.mean2) + r0; lb.rstd = F32(b.rstd2) + r0; lb.dx = du; lb.lddx = D; lb.accumulate_dx = 0; lb.accumulate_dparams = acc;
  lb.workspace = F32(B.ln_ws);
  lb.dx_planes = gp; lb.ldp = D; lb.plane_stride = Tg * D; lb.rows_per_scale = 1;
  lb.drop = enc_site(dr, dr.f2, dr.p_hidden, seq0); lb.drop_rows_per_seq = Lq;
  fold_colsum(fold, lb.workspace, ln_parts, 3 * (int64_t)D, D, G.ln2w, acc);
  fold_colsum(fold, lb.workspace + D, ln_parts, 3 * (int64_t)D, D, G.ln2b, acc);
  fold_colsum(fold, lb.workspace + 2 * D, ln_parts, 3 * (int64_t)D, D, G.f2b, acc);
  SRW_TRY(srw_layernorm_bwd(&lb, s));
  SRW_TRY(wgrad(D, Fh, Tr, gp, D, Tg, ws + b.h + r0 * Fh * 2, Fh, T, F32(B.sk4[0]), G.f2w, Fh, acc, impl, s, fold));
  {
    Gemm g(Tr, Fh, D, impl);  // dz = drop_act(g W2) * gelu'(z)
    g.A(gp, D, Tg, 0).Bm(w.f2, Fh, D, 1);
    g.g.epilogue = SRW_EPI_DGELU; g.g.aux = F32(b.z) + r0 * Fh; g.g.ldaux = Fh; g.g.out_planes = dzp; g.g.ldp = Fh; g.g.out_plane_stride = Tg * Fh;
    g.g.drop = enc_site(dr, dr.act, dr.p_act, seq0); g.g.drop_rows_per_seq = Lq;
    SRW_TRY(g.run(s));
  }
  SRW_TRY(colsum_planes(dzp, Fh, Tg, Tr, Fh, G.f1b, acc, F32(B.colsum_ws), s, fold));
  SRW_TRY(wgrad(Fh, D, Tr, dzp, Fh, Tg, ws + b.xmp + r0 * D * 2, D, T, F32(B.sk4[1]), G.f1w, D, acc, impl, s, fold));
  {
    Gemm g(Tr, D, Fh, impl);  // dy = du + dz W1   (gradient of x_mid: residual path + FFN path)
    g.A(dzp, Fh, Tg, 0).Bm(w.f1, D, Fh, 1);
    g.g.epilogue = SRW_EPI_RESID; g.g.resid = du; g.g.ldr = D; g.g.out_f32 = F32(B.dy) + r0 * D; g.g.ldo = D;
    SRW_TRY(g.run(s));
  }
  // ---- attention branch: x_mid = LN1(u1), u1 = x_in + dropout(attn(x_in) Wo^T + bo) ----
  lb.dy = F32(B.dy) + r0 * D; lb.x = F32(b.u1) + r0 * D; lb.gamma = w.ln1w; lb.mean = F32(b.mean1) + r0; lb.rstd = F32(b.rstd1) + r0;
  lb.workspace = F32(B.ln_ws2);
  lb.drop = enc_site(dr, dr.o, dr.p_hidden, seq0);
  fold_colsum(fold, lb.workspace, ln_parts, 3 * (int64_t)D, D, G.ln1w, acc);
  fold_colsum(fold, lb.workspace + D, ln_parts, 3 * (int64_t)D, D, G.ln1b, acc);
  fold_colsum(fold, lb.workspace + 2 * D, ln_parts, 3 * (int64_t)D, D, G.ob, acc);
  SRW_TRY(srw_layernorm_bwd(&lb, s));
  SRW_TRY(wgrad(D, D, Tr, gp, D, Tg, ws + b.o + r0 * D * 2, D, T, F32(B.sk4[2]), G.ow, D, acc, impl, s, fold));
  {
    Gemm g(Tr, D, D, impl);  // d_o = g Wo
    g.A(gp, D, Tg, 0).Bm(w.o, D, D, 1);
    g.g.epilogue = SRW_EPI_PLANES; g.g.out_planes = dop; g.g.ldp = D; g.g.out_plane_stride = Tg * D;
    SRW_TRY(g.run(s));
  }
  {
    srw_attn_bwd_args at = {};
    at.B = nseq; at.N = Lq; at.H = d.H; at.head_dim = 64; at.scale = attn_scale;
    at.qkv = ws + b.qkv + r0 * 3 * D * 2; at.ld_qkv = 3 * D; at.qkv_plane_stride = T * 3 * D;
    at.o = ws + b.o + r0 * D * 2; at.ld_o = D; at.o_plane_stride = T * D;
    at.d_o = dop; at.ld_do = D; at.do_plane_stride = Tg * D;
    at.lse = F32(b.lse) + (int64_t)seq0 * d.H * Lq; at.delta = F32(B.delta) + (int64_t)seq0 * d.H * Lq;
    at.dqkv = dqkvp; at.ld_dqkv = 3 * D; at.dqkv_plane_stride = Tg * 3 * D;
    if (key_bias) { at.key_bias = key_bias + (int64_t)seq0 * d.ld_bias; at.ld_bias = d.ld_bias; at.kv_len = kv_len ? kv_len + seq0 : nullptr; }
    at.drop = enc_site(dr, dr.attn, dr.p_attn, seq0);
    SRW_TRY(srw_attn_bwd(&at, s));
  }
  SRW_TRY(colsum_planes(dqkvp, 3 * D, Tg, Tr, 3 * D, G.qkv_b, acc, F32(B.colsum_ws2), s, fold));   // q | k | v biases are contiguous
  SRW_TRY(wgrad(3 * D, D, Tr, dqkvp, 3 * D, Tg, ws + xp_in + r0 * D * 2, D, T, F32(B.sk4[3]), G.qkv_w, D, acc, impl, s, fold));
  {
    Gemm g(Tr, D, 3 * D, impl);  // dx(layer input) = du + dqkv Wqkv
    g.A(dqkvp, 3 * D, Tg, 0).Bm(w.qkv, D, 3 * D, 1);
    g.g.epilogue = SRW_EPI_RESID; g.g.resid = du; g.g.ldr = D; g.g.out_f32 = dx; g.g.ldo = D;
    SRW_TRY(g.run(s));
  }
  SRW_TRY(srw_grad_fold(fold, s));
  return SRW_OK;
}

static inline void enc_take_bwd_scratch(Carver& c, const EncDims& d, EncBwdScratch& B) {
  const int64_t Tg = std::max<int64_t>(d.Tg, 1), Sg = std::max(d.Sg, 1), D = d.D, F = d.F;
  B.du = c.take(Tg * D * 4); B.dy = c.take(Tg * D * 4); B.g = c.take(Tg * D * 4);
  B.dz = c.take(Tg * F * 4); B.d_o = c.take(Tg * D * 4); B.dqkv = c.take(Tg * 3 * D * 4);
  B.delta = c.take(Sg * d.H * d.Lq * 4);
  B.colsum_ws = c.take((int64_t)256 * std::max<int64_t>(3 * D, F) * 4);
  B.colsum_ws2 = c.take((int64_t)256 * std::max<int64_t>(3 * D, F) * 4);
  B.ln_ws = c.take((int64_t)3 * 256 * D * 4); B.ln_ws2 = c.take((int64_t)3 * 256 * D * 4);
  B.sk4[0] = c.take(splitk_for(d.D, d.F, Tg, nullptr) * 4);
  B.sk4[1] = c.take(splitk_for(d.F, d.D, Tg, nullptr) * 4);
  B.sk4[2] = c.take(splitk_for(d.D, d.D, Tg, nullptr) * 4);
  B.sk4[3] = c.take(splitk_for(3 * d.D, d.D, Tg, nullptr) * 4);
}
static inline void enc_take_layer(Carver& c, const EncDims& d, EncLayerBufs& b) {
  const int64_t T = d.T, D = d.D, F = d.F;
  b.qkv = c.take(T * 3 * D * 4); b.o = c.take(T * D * 4); b.lse = c.take((int64_t)d.S * d.H * d.Lq * 4);
  b.u1 = c.take(T * D * 4); b.mean1 = c.take(T * 4); b.rstd1 = c.take(T * 4);
  b.xm = c.take(T * D * 4); b.xmp = c.take(T * D * 4);
  b.z = c.take(T * F * 4); b.h = c.take(T * F * 4);
  b.u2 = c.take(T * D * 4); b.mean2 = c.take(T * 4); b.rstd2 = c.take(T * 4);
}

// ------------------------------------------------------------------------------------------------
// pooled features + classifier head
// ------------------------------------------------------------------------------------------------
// feat[s, c] = mean_l dropout(x[(s, l), c])   (bert.py:35-37 / hubert.py:47-48: drop, then mean over every position)
static __global__ void enc_pool_fwd_kernel(const float* __restrict__ x, int S, int Lq, int D, const DropParams dr, float* __restrict__ feat,
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
static __global__ void enc_pool_bwd_kernel(const float* __restrict__ dfeat, int Sg, int Lq, int D, const DropParams dr, float* __restrict__ dx,
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

// classifier: z1 = feat Wc1^T + bc1, a1 = gelu(z1) (exact erf), logits = a1 Wc2^T + bc2.  One CTA per sequence.
static __global__ void __launch_bounds__(256) enc_head_fwd_kernel(const float* __restrict__ feat, int D, int C, const float* __restrict__ W1,
                                                                  const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                                                                  float* __restrict__ z1_out, float* __restrict__ a1_out, float* __restrict__ logits) {
  extern __shared__ float sm[];
  float* f = sm;        // [D]
  float* a1 = sm + D;   // [D]
  const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = threadIdx.x; d < D; d += 256) f[d] = feat[(int64_t)s * D + d];
  __syncthreads();
  for (int n = warp; n < D; n += 8) {
    const float* w = W1 + (int64_t)n * D;
    float acc = 0.f;
#pragma unroll 8
    for (int k = lane; k < D; k += 32) acc = fmaf(f[k], w[k], acc);   // unrolled: the weight loads of a row are in flight together
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
static __global__ void __launch_bounds__(256) enc_head_bwd_rows_kernel(const float* __restrict__ dlogits, const float* __restrict__ dfeat_in, int D, int C,
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
    float a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int n = 0;
    for (; n + 8 <= D; n += 8) {   // eight independent loads in flight per thread (the loop is a chain of L2 round trips otherwise)
      const float w0 = W1[(int64_t)n * D + k], w1 = W1[(int64_t)(n + 1) * D + k], w2 = W1[(int64_t)(n + 2) * D + k], w3 = W1[(int64_t)(n + 3) * D + k];
      const float w4 = W1[(int64_t)(n + 4) * D + k], w5 = W1[(int64_t)(n + 5) * D + k], w6 = W1[(int64_t)(n + 6) * D + k], w7 = W1[(int64_t)(n + 7) * D + k];
      acc = fmaf(dz[n], w0, acc); a1 = fmaf(dz[n + 1], w1, a1); a2 = fmaf(dz[n + 2], w2, a2); a3 = fmaf(dz[n + 3], w3, a3);
      acc = fmaf(dz[n + 4], w4, acc); a1 = fmaf(dz[n + 5], w5, a1); a2 = fmaf(dz[n + 6], w6, a2); a3 = fmaf(dz[n + 7], w7, a3);
    }
    for (; n < D; ++n) acc = fmaf(dz[n], W1[(int64_t)n * D + k], acc);
    dfeat_out[(int64_t)s * D + k] = (acc + a1) + (a2 + a3);
  }
}
// dW2[c, k] = sum_s dlogits[s, c] a1[s, k]; db2; dW1[n, k] = sum_s dz1[s, n] feat[s, k]; db1
static __global__ void enc_head_bwd_params_kernel(const float* __restrict__ dlogits, const float* __restrict__ a1, const float* __restrict__ dz1,
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

// x[t, :] = dropout mask * x[t, :] / keep, in place (the gradient entering a dropout whose mask is regenerated)
static __global__ void enc_dropout_rows_kernel(float* __restrict__ x, int rows, int D, int Lq, const DropParams dr) {
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

// pooled features + classifier forward: feat = mean(drop(x)); logits = head(feat).  cls = {W1, b1, W2, b2}
static inline int enc_head_forward(const float* x_last, int S, int Lq, int D, int C, const DropParams& dr, const int32_t* pool_len, const float* const cls[4],
                                   float* feat_ws, float* z1_ws, float* a1_ws, float* logits, float* feat_out, cudaStream_t s) {
  enc_pool_fwd_kernel<<<dim3(cdiv(D, 256), S), 256, 0, s>>>(x_last, S, Lq, D, dr, feat_ws, pool_len);
  g_launches++;
  SRW_LAUNCH_CHECK();
  enc_head_fwd_kernel<<<S, 256, 2 * D * sizeof(float), s>>>(feat_ws, D, C, cls[0], cls[1], cls[2], cls[3], z1_ws, a1_ws, logits);
  g_launches++;
  SRW_LAUNCH_CHECK();
  SRW_CUDA(cudaMemcpyAsync(feat_out, feat_ws, (size_t)S * D * 4, cudaMemcpyDeviceToDevice, s));
  return SRW_OK;
}
// -> dx [Sg * Lq, D] = gradient of the last hidden states; classifier gradients into gcls = {dW1, db1, dW2, db2}
static inline int enc_head_backward(const float* dlogits, const float* dfeat_in, int Sg, int Lq, int D, int C, const DropParams& dr, const int32_t* pool_len,
                                    const float* const cls[4], float* const gcls[4], const float* feat_ws, const float* z1_ws, const float* a1_ws,
                                    float* dz1_ws, float* dfeat_ws, float* dx, int acc, cudaStream_t s) {
  enc_head_bwd_rows_kernel<<<Sg, 256, (C + D) * sizeof(float), s>>>(dlogits, dfeat_in, D, C, cls[0], cls[2], z1_ws, dz1_ws, dfeat_ws);
  g_launches++;
  SRW_LAUNCH_CHECK();
  const int64_t n = (int64_t)D * D + (int64_t)C * D + D + C;
  enc_head_bwd_params_kernel<<<(int)cdiv64(n, 256), 256, 0, s>>>(dlogits, a1_ws, dz1_ws, feat_ws, Sg, D, C, gcls[0], gcls[1], gcls[2], gcls[3], acc);
  g_launches++;
  SRW_LAUNCH_CHECK();
  enc_pool_bwd_kernel<<<148 * 8, 256, 0, s>>>(dfeat_ws, Sg, Lq, D, dr, dx, pool_len);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

}  // namespace srw
