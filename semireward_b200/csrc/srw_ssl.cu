// srw_ssl — everything between the backbone outputs and the logit gradients of the SemiReward step, as single-CTA
// fused kernels (each replaces ~10-60 tiny PyTorch kernels plus host round trips of the reference):
//   srw_flexmatch_mask : softmax + argmax + FlexMatch threshold mask + selected_label scatter + class histogram update
//                        (srflexmatch/utils.py:23-63, hooks/pseudo_label.py:40, algorithmbase.py:332-333)
//   srw_ssl_loss       : supervised CE, reward mean-threshold mask2, masked consistency CE, total loss, d/dlogits
//                        (cross_entropy.py:11-31, consistency.py:13-45, srflexmatch.py:100-102,132,152,210)
//   srw_rewarder_fwd   : Rewarder.forward incl. the softmax over the 2B batch rows (semireward.py:52-72)
//   srw_generator_fwd  : Generator.forward + .long() (semireward.py:21-24, srflexmatch.py:157-158)
//   srw_rewarder_train : forward, both MSE losses, full backward and the Adam step (srflexmatch.py:173-208)
// All math is plain fp32 in a fixed, documented order (deterministic: no atomics, reductions in index order), which is
// what the bit-exact mask requirement needs.  HBM traffic is a few KB..MB; these kernels are latency-bound by design.
#include <atomic>
#include <mutex>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {
extern std::atomic<int64_t> g_launches;

constexpr int SSL_THREADS = 1024;
constexpr int MAX_ROWS = 2048;

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ int block_reduce_max_i(int v, int* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  int s = red[0];
  for (int w = 1; w < nw; ++w) s = max(s, red[w]);
  return s;
}

// warp-cooperative row softmax statistics: max (first index) and sum of exp(x - max)
__device__ __forceinline__ void warp_row_max_sum(const float* row, int C, int lane, float& m, float& sum) {
  m = -INFINITY;
  for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
  m = warp_max(m);
  sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(row[c] - m);
  sum = warp_sum(sum);
}

// ------------------------------------------------------------------------------------------------
// FlexMatch mask
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSL_THREADS) flexmatch_mask_kernel(const srw_flexmatch_mask_args a) {
  __shared__ int s_sel[MAX_ROWS];
  __shared__ int s_idx[MAX_ROWS];
  __shared__ int red_i[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int C = a.num_classes;
  for (int b = warp; b < a.B; b += nw) {
    const float* row = a.logits_w + (int64_t)b * a.ld_logits;
    float m, sum;
    warp_row_max_sum(row, C, lane, m, sum);
    // probs = exp(x - max) / sum ; argmax over the probabilities, first index on ties
    float best = -1.f;
    int best_i = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float pr = expf(row[c] - m) / sum;
      if (a.probs_w) a.probs_w[(int64_t)b * C + c] = pr;
      if (pr > best) { best = pr; best_i = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    if (lane == 0) {
      if (a.selected_label == nullptr) {   // FixedThresholdingHook (hooks/masking.py:42-57): stateless max_p >= p_cutoff
        a.mask[b] = best >= a.p_cutoff ? 1.0f : 0.0f;
        a.pseudo[b] = best_i;
        if (a.max_probs) a.max_probs[b] = best;
        continue;
      }
      const float acc = a.classwise_acc[best_i];
      const float thr = a.p_cutoff * (acc / (2.0f - acc));          // utils.py:52
      a.mask[b] = best >= thr ? 1.0f : 0.0f;
      a.pseudo[b] = best_i;
      if (a.max_probs) a.max_probs[b] = best;
      s_sel[b] = best >= a.p_cutoff ? 1 : 0;                         // utils.py:53
      s_idx[b] = best_i;
    }
  }
  if (a.selected_label == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    // selected_label[idx_ulb[select]] = max_idx[select]  (utils.py:60-61), histogram kept in step
    for (int b = 0; b < a.B; ++b) {
      if (!s_sel[b]) continue;
      const int64_t i = a.idx_ulb[b];
      if (i < 0 || i >= a.ulb_dest_len) continue;
      const int64_t old = a.selected_label[i];
      a.selected_label[i] = s_idx[b];
      a.hist[old + 1] -= 1;
      a.hist[s_idx[b] + 1] += 1;
    }
  }
  __syncthreads();
  // update(): classwise_acc[i] = count[i] / max(count)   (utils.py:23-35)
  int mx_all = 0, mx_pos = 0;
  for (int c = threadIdx.x; c <= C; c += blockDim.x) {
    const int h = a.hist[c];
    mx_all = max(mx_all, h);
    if (c > 0) mx_pos = max(mx_pos, h);
  }
  mx_all = block_reduce_max_i(mx_all, red_i);
  mx_pos = block_reduce_max_i(mx_pos, red_i);
  if (mx_all < a.ulb_dest_len) {
    const int denom = a.thresh_warmup ? mx_all : mx_pos;
    if (denom > 0)
      for (int c = threadIdx.x; c < C; c += blockDim.x) a.classwise_acc[c] = (float)((double)a.hist[c + 1] / (double)denom);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-wide dense helpers for the Rewarder / Generator (tiny matrices, fp32 FMA)
// ------------------------------------------------------------------------------------------------
enum { ACT_NONE = 0, ACT_RELU = 1 };

// out[r, n] = act(bias[n] + sum_k in[r, k] W[n, k]).  One warp per output COLUMN: the K / 32 weights of the row W[n, :] a lane needs are
// loaded once into registers (independent loads, one L2 round trip) and reused for every batch row; per output the products are
// summed in the same order as before (lane-strided partial sums over ascending k, then the xor-shuffle tree), so results are
// bit-identical to the one-warp-per-output form — but a warp now waits for memory once per column instead of once per output
// (these kernels are one CTA deep: the chain of dependent L2 round trips is what they cost).
template <int KR>
__device__ __forceinline__ void cta_linear_cols(const float* in, int64_t ld_in, int rows, const float* W, const float* bias, int N, float* out, int ld_out,
                                                int act) {
  constexpr int K = KR * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int n = warp; n < N; n += nw) {
    const float* w = W + (int64_t)n * K + lane;
    float wr[KR];
#pragma unroll
    for (int j = 0; j < KR; ++j) wr[j] = w[32 * j];
    const float bn = bias[n];
    for (int r = 0; r < rows; ++r) {
      const float* x = in + (int64_t)r * ld_in + lane;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < KR; ++j) acc = fmaf(x[32 * j], wr[j], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        acc += bn;
        out[(int64_t)r * ld_out + n] = (act == ACT_RELU) ? fmaxf(acc, 0.f) : acc;
      }
    }
  }
}
__device__ void cta_linear(const float* in, int64_t ld_in, int rows, int K, const float* W, const float* bias, int N, float* out, int ld_out,
                           int act) {
  switch (K) {
    case 64: cta_linear_cols<2>(in, ld_in, rows, W, bias, N, out, ld_out, act); return;
    case 128: cta_linear_cols<4>(in, ld_in, rows, W, bias, N, out, ld_out, act); return;
    case 256: cta_linear_cols<8>(in, ld_in, rows, W, bias, N, out, ld_out, act); return;
    case 384: cta_linear_cols<12>(in, ld_in, rows, W, bias, N, out, ld_out, act); return;
    case 768: cta_linear_cols<24>(in, ld_in, rows, W, bias, N, out, ld_out, act); return;
    default: break;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int o = warp; o < rows * N; o += nw) {
    const int r = o / N, n = o % N;
    const float* x = in + (int64_t)r * ld_in;
    const float* w = W + (int64_t)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(x[k], w[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += bias[n];
      out[(int64_t)r * ld_out + n] = (act == ACT_RELU) ? fmaxf(acc, 0.f) : acc;
    }
  }
}
// dX[r, k] = sum_n dY[r, n] W[n, k]: one thread per output, n ascending in ONE accumulator (the summation order of the parity tests);
// the weight loads of eight consecutive n are issued together so that the loop is N / 8 memory round trips, not N
__device__ void cta_linear_dx(const float* dY, int rows, int N, const float* W, int K, float* dX) {
  for (int o = threadIdx.x; o < rows * K; o += blockDim.x) {
    const int r = o / K, k = o % K;
    const float* dy = dY + r * N;
    const float* w = W + k;
    float acc = 0.f;
    int n = 0;
    for (; n + 8 <= N; n += 8) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = w[(int64_t)(n + j) * K];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(dy[n + j], wv[j], acc);
    }
    for (; n < N; ++n) acc = fmaf(dy[n], w[(int64_t)n * K], acc);
    dX[o] = acc;
  }
}
// dW[n, k] = sum_r dY[r, n] X[r, k];  db[n] = sum_r dY[r, n]
__device__ void cta_linear_dw(const float* dY, int rows, int N, const float* X, int64_t ldx, int K, float* dW, float* db) {
  for (int o = threadIdx.x; o < N * K; o += blockDim.x) {
    const int n = o / K, k = o % K;
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc = fmaf(dY[r * N + n], X[(int64_t)r * ldx + k], acc);
    dW[o] = acc;
  }
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < rows; ++r) acc += dY[r * N + n];
    db[n] = acc;
  }
}
// y = LN(x) over 128 columns (eps 1e-5, nn.LayerNorm default); keeps xhat and rstd for the backward
__device__ void cta_layernorm128(const float* x, int rows, const float* gamma, const float* beta, float* y, float* xhat, float* rstd) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < rows; r += nw) {
    const float* xr = x + r * 128;
    float v[4], s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] = xr[lane + 32 * j]; s += v[j]; }
    const float mean = warp_sum(s) * (1.0f / 128.0f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { v[j] -= mean; q += v[j] * v[j]; }
    const float rs = rsqrtf(warp_sum(q) * (1.0f / 128.0f) + 1e-5f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      const float xh = v[j] * rs;
      if (xhat) xhat[r * 128 + c] = xh;
      y[r * 128 + c] = xh * gamma[c] + beta[c];
    }
    if (lane == 0 && rstd) rstd[r] = rs;
  }
}
// dx = LN backward (rows x 128); dgamma/dbeta accumulated afterwards by cta_ln_param_grads
__device__ void cta_layernorm128_bwd(const float* dy, int rows, const float* gamma, const float* xhat, const float* rstd, float* dx) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < rows; r += nw) {
    float g[4], xh[4], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = lane + 32 * j;
      xh[j] = xhat[r * 128 + c];
      g[j] = dy[r * 128 + c] * gamma[c];
      s1 += g[j];
      s2 += g[j] * xh[j];
    }
    s1 = warp_sum(s1) * (1.0f / 128.0f);
    s2 = warp_sum(s2) * (1.0f / 128.0f);
    const float rs = rstd[r];
#pragma unroll
    for (int j = 0; j < 4; ++j) dx[r * 128 + lane + 32 * j] = rs * (g[j] - s1 - xh[j] * s2);
  }
}
__device__ void cta_ln_param_grads(const float* dy, const float* xhat, int rows, float* dgamma, float* dbeta) {
  for (int c = threadIdx.x; c < 128; c += blockDim.x) {
    float ag = 0.f, ab = 0.f;
    for (int r = 0; r < rows; ++r) {
      ag = fmaf(dy[r * 128 + c], xhat[r * 128 + c], ag);
      ab += dy[r * 128 + c];
    }
    dgamma[c] = ag;
    dbeta[c] = ab;
  }
}

// Rewarder parameter indices (state_dict order)
enum { R_FCW = 0, R_FCB, R_FNW, R_FNB, R_EMB, R_LNW, R_LNB, R_CAW, R_CAB, R_M1W, R_M1B, R_M2W, R_M2B, R_F1W, R_F1B, R_F2W, R_F2B, R_NUM };

struct RewPtrs { float* p[R_NUM]; };
struct RewSizes { int64_t n[R_NUM]; };

// workspace carve (floats), per batch B
struct RewWs {
  float *u, *fxh, *frs, *X /* [2B,128]: f then e */, *eraw, *exh, *ers, *att /* [2B] */, *wsm /* [2B] */, *ctx /* [128] */, *h, *h1, *h2, *h3, *r;
  float *dz4, *dh3, *dh2, *dh1, *dh, *dX, *du, *demb, *dctx, *dwv, *da;
};
__host__ __device__ inline int64_t rew_ws_floats(int B) {
  return (int64_t)B * (128 + 128 + 1 + 256 + 128 + 128 + 1 + 2 + 2 + 128 + 256 + 128 + 64 + 1 + 1 + 64 + 128 + 256 + 128 + 256 + 128 + 128 + 2 + 2) + 128 + 128 + 64;
}
__host__ __device__ inline RewWs rew_carve(float* w, int B) {
  RewWs s;
  auto take = [&](int64_t n) { float* p = w; w += n; return p; };
  s.u = take(B * 128); s.fxh = take(B * 128); s.frs = take(B); s.X = take(2 * B * 128); s.eraw = take(B * 128); s.exh = take(B * 128);
  s.ers = take(B); s.att = take(2 * B); s.wsm = take(2 * B); s.ctx = take(128); s.h = take(B * 128); s.h1 = take(B * 256);
  s.h2 = take(B * 128); s.h3 = take(B * 64); s.r = take(B);
  s.dz4 = take(B); s.dh3 = take(B * 64); s.dh2 = take(B * 128); s.dh1 = take(B * 256); s.dh = take(B * 128); s.dX = take(2 * B * 128);
  s.du = take(B * 128); s.demb = take(B * 128); s.dctx = take(128); s.dwv = take(2 * B); s.da = take(2 * B);
  return s;
}

// forward of Rewarder for the whole CTA; leaves every intermediate in ws.  semireward.py:52-72.
__device__ void rewarder_forward_cta(const RewPtrs& P, const RewWs& s, const float* feats, int64_t ld_feats, const int64_t* labels, int B,
                                     int D, int label_rows, float* red) {
  float* f = s.X;
  float* e = s.X + B * 128;
  cta_linear(feats, ld_feats, B, D, P.p[R_FCW], P.p[R_FCB], 128, s.u, 128, ACT_NONE);
  for (int o = threadIdx.x; o < B * 128; o += blockDim.x) {
    int64_t l = labels[o / 128];
    l = l < 0 ? 0 : (l >= label_rows ? label_rows - 1 : l);   // nn.Embedding would raise; clamp keeps the kernel in bounds
    s.eraw[o] = P.p[R_EMB][l * 128 + (o % 128)];
  }
  __syncthreads();
  cta_layernorm128(s.u, B, P.p[R_FNW], P.p[R_FNB], f, s.fxh, s.frs);
  cta_layernorm128(s.eraw, B, P.p[R_LNW], P.p[R_LNB], e, s.exh, s.ers);
  __syncthreads();
  cta_linear(s.X, 128, 2 * B, 128, P.p[R_CAW], P.p[R_CAB], 1, s.att, 1, ACT_NONE);
  __syncthreads();
  // softmax over the 2B rows (dim=0)
  float m = -INFINITY;
  for (int r = threadIdx.x; r < 2 * B; r += blockDim.x) m = fmaxf(m, s.att[r]);
  m = warp_max(m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int r = threadIdx.x; r < 2 * B; r += blockDim.x) {
    const float ev = expf(s.att[r] - m);
    s.wsm[r] = ev;
    sum += ev;
  }
  sum = block_reduce_sum(sum, red);
  for (int r = threadIdx.x; r < 2 * B; r += blockDim.x) s.wsm[r] = s.wsm[r] / sum;
  __syncthreads();
  for (int c = threadIdx.x; c < 128; c += blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < 2 * B; ++r) acc = fmaf(s.wsm[r], s.X[r * 128 + c], acc);
    s.ctx[c] = acc;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < B * 128; o += blockDim.x) s.h[o] = s.ctx[o % 128] + e[o];
  __syncthreads();
  cta_linear(s.h, 128, B, 128, P.p[R_M1W], P.p[R_M1B], 256, s.h1, 256, ACT_RELU);
  __syncthreads();
  cta_linear(s.h1, 256, B, 256, P.p[R_M2W], P.p[R_M2B], 128, s.h2, 128, ACT_NONE);
  __syncthreads();
  cta_linear(s.h2, 128, B, 128, P.p[R_F1W], P.p[R_F1B], 64, s.h3, 64, ACT_RELU);
  __syncthreads();
  cta_linear(s.h3, 64, B, 64, P.p[R_F2W], P.p[R_F2B], 1, s.r, 1, ACT_NONE);
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) s.r[b] = 1.0f / (1.0f + expf(-s.r[b]));
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// losses + dlogits
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSL_THREADS) ssl_loss_kernel(const srw_ssl_loss_args a, RewPtrs RP, int fused_rewarder, int use_smem) {
  extern __shared__ float dyn[];
  __shared__ float s_ce[2 * MAX_ROWS];   // [0,B_lb) supervised CE, [MAX_ROWS, +B_ulb) unsupervised CE
  __shared__ float s_w[MAX_ROWS];        // mask * mask2
  __shared__ float s_scalar[4];
  __shared__ float red_r[32];
  const float* rew = a.reward;
  if (fused_rewarder) {   // Rewarder.forward on (weak features, pseudo-labels) of the last sampling pass, intermediates on chip when they fit
    const RewWs rs = rew_carve(use_smem ? dyn : a.rew_workspace, a.B_ulb);
    rewarder_forward_cta(RP, rs, a.feats, a.ld_feats, a.pseudo, a.B_ulb, a.feature_dim, a.label_rows, red_r);
    rew = rs.r;
    if (a.reward_out)
      for (int b = threadIdx.x; b < a.B_ulb; b += blockDim.x) a.reward_out[b] = rs.r[b];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int C = a.num_classes;
  const int rows = a.B_lb + a.B_ulb;
  // mask2 = reward >= mean(reward)   (srflexmatch.py:100-101).  The sum is a fixed-order pairwise tree: samples that share
  // a pseudo-label have bit-identical rewards (the Rewarder sees features only through the batch context), and a batch
  // of identical rewards must compare equal to its own mean, which a pairwise sum guarantees for power-of-two batches
  // (a running sum r+r+r... rounds at 3r) and which is what torch's vectorised reduction does as well.
  if (rew) {
    float* tree = s_ce;   // scratch: the CE values are written later
    for (int b = threadIdx.x; b < a.B_ulb; b += blockDim.x) tree[b] = rew[b];
    __syncthreads();
    for (int stride = 1; stride < a.B_ulb; stride <<= 1) {
      for (int i = threadIdx.x * 2 * stride; i + stride < a.B_ulb; i += blockDim.x * 2 * stride) tree[i] += tree[i + stride];
      __syncthreads();
    }
    if (threadIdx.x == 0) s_scalar[0] = tree[0] / (float)a.B_ulb;
  } else if (threadIdx.x == 0) {
    s_scalar[0] = 0.f;
  }
  __syncthreads();
  for (int b = threadIdx.x; b < a.B_ulb; b += blockDim.x) {
    const float m2 = rew ? (rew[b] >= s_scalar[0] ? 1.0f : 0.0f) : 1.0f;
    if (a.mask2) a.mask2[b] = m2;
    s_w[b] = a.mask[b] * m2;
  }
  __syncthreads();
  for (int r = warp; r < rows; r += nw) {
    const bool lb = r < a.B_lb;
    const int b = lb ? r : r - a.B_lb;
    const float* row = (lb ? a.logits_lb : a.logits_s) + (int64_t)b * a.ld_logits;
    const int64_t tgt = lb ? a.y_lb[b] : a.pseudo[b];
    float m, sum;
    warp_row_max_sum(row, C, lane, m, sum);
    const float lse = m + logf(sum);
    if (lane == 0) s_ce[lb ? b : MAX_ROWS + b] = lse - row[tgt];   // -log_softmax[target]
    float* drow = lb ? a.dlogits_lb : a.dlogits_s;
    if (drow) {
      drow += (int64_t)b * a.ld_dlogits;
      const float w = lb ? 1.0f / (float)a.B_lb : a.lambda_u * s_w[b] / (float)a.B_ulb;
      for (int c = lane; c < C; c += 32) {
        const float pr = expf(row[c] - m) / sum;
        drow[c] = w * (pr - (c == tgt ? 1.0f : 0.0f));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sup = 0.f, unsup = 0.f, util = 0.f;
    for (int b = 0; b < a.B_lb; ++b) sup += s_ce[b];
    sup /= (float)a.B_lb;
    for (int b = 0; b < a.B_ulb; ++b) {
      unsup += s_ce[MAX_ROWS + b] * s_w[b];
      util += a.mask[b];
    }
    unsup /= (float)a.B_ulb;
    util /= (float)a.B_ulb;
    a.losses[0] = sup; a.losses[1] = unsup; a.losses[2] = sup + a.lambda_u * unsup; a.losses[3] = util;
  }
}

__global__ void __launch_bounds__(SSL_THREADS) rewarder_fwd_kernel(RewPtrs P, const float* feats, int64_t ld_feats, const int64_t* labels, int B,
                                                                  int D, int label_rows, float* reward, float* ws, int use_smem) {
  extern __shared__ float dyn[];
  __shared__ float red[32];
  const RewWs s = rew_carve(use_smem ? dyn : ws, B);
  rewarder_forward_cta(P, s, feats, ld_feats, labels, B, D, label_rows, red);
  for (int b = threadIdx.x; b < B; b += blockDim.x) reward[b] = s.r[b];
}

struct GenPtrs { const float* p[8]; };
__global__ void __launch_bounds__(SSL_THREADS) generator_fwd_kernel(GenPtrs P, const float* feats, int64_t ld_feats, int B, int D, int64_t* labels,
                                                                   float* ws) {
  float* a1 = ws; float* a2 = a1 + B * 256; float* a3 = a2 + B * 128; float* a4 = a3 + B * 64;
  cta_linear(feats, ld_feats, B, D, P.p[0], P.p[1], 256, a1, 256, ACT_RELU);
  __syncthreads();
  cta_linear(a1, 256, B, 256, P.p[2], P.p[3], 128, a2, 128, ACT_RELU);
  __syncthreads();
  cta_linear(a2, 128, B, 128, P.p[4], P.p[5], 64, a3, 64, ACT_RELU);
  __syncthreads();
  cta_linear(a3, 64, B, 64, P.p[6], P.p[7], 1, a4, 1, ACT_RELU);
  __syncthreads();
  for (int b = threadIdx.x; b < B; b += blockDim.x) labels[b] = (int64_t)a4[b];   // .long(): truncation toward zero
}

struct AdamScalars { float lr; double bc1, bc2_sqrt; };

// The online update runs as TWO launches:
//   rewarder_actgrad_kernel   one CTA: forward, both MSE losses and the backward down to every ACTIVATION gradient.  For the
//                             configs' batch sizes (B <= 22) all intermediates live in shared memory (the chain of ~25 dependent
//                             phases was bound by L2 round trips when they lived in the global workspace: 330-480 us); the
//                             intermediates the second launch needs are copied to the workspace at the end.
//   rewarder_wgrad_adam_kernel  many CTAs: one thread per PARAMETER element forms its gradient from those activations
//                             (sum over the <= 2B rows, same order as before) and applies torch.optim.Adam right away.
// Phase 1 stops after storing the gradients, phase 2 applies Adam to stored gradients (data-parallel exchange in between).
__global__ void __launch_bounds__(SSL_THREADS) rewarder_actgrad_kernel(RewPtrs P, const float* feats, int64_t ld_feats, const int64_t* gen_labels,
                                                                      const int64_t* true_labels, int B, int D, int label_rows, float* losses, float* ws,
                                                                      int loss_select, int use_smem) {
  extern __shared__ float dyn[];
  __shared__ float red[32];
  float* base = use_smem ? dyn : ws;
  const RewWs s = rew_carve(base, B);
  rewarder_forward_cta(P, s, feats, ld_feats, gen_labels, B, D, label_rows, red);
  // losses: generator_loss = MSE(r, 1), rewarder_loss = MSE(r, target), target = cos-sim of the two one-hots mapped to
  // (cos+1)/2 = 1 if equal else 0.5  (semireward.py:130-139, srflexmatch.py:195-199)
  float gl = 0.f, rl = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float r = s.r[b];
    const float tgt = (gen_labels[b] == true_labels[b]) ? 1.0f : 0.5f;
    gl += (r - 1.0f) * (r - 1.0f);
    rl += (r - tgt) * (r - tgt);
    const float dr = ((loss_select == 2 ? 0.f : 2.0f * (r - 1.0f)) + (loss_select == 1 ? 0.f : 2.0f * (r - tgt))) / (float)B;
    s.dz4[b] = dr * r * (1.0f - r);
  }
  gl = block_reduce_sum(gl, red);
  rl = block_reduce_sum(rl, red);
  if (threadIdx.x == 0) { losses[0] = gl / (float)B; losses[1] = rl / (float)B; }
  __syncthreads();
  for (int o = threadIdx.x; o < B * 64; o += blockDim.x) s.dh3[o] = s.h3[o] > 0.f ? s.dz4[o / 64] * P.p[R_F2W][o % 64] : 0.f;   // ffn_fc2, relu
  __syncthreads();
  cta_linear_dx(s.dh3, B, 64, P.p[R_F1W], 128, s.dh2);    // ffn_fc1
  __syncthreads();
  cta_linear_dx(s.dh2, B, 128, P.p[R_M2W], 256, s.dh1);   // mlp_fc2
  __syncthreads();
  for (int o = threadIdx.x; o < B * 256; o += blockDim.x) s.dh1[o] = s.h1[o] > 0.f ? s.dh1[o] : 0.f;
  __syncthreads();
  cta_linear_dx(s.dh1, B, 256, P.p[R_M1W], 128, s.dh);    // mlp_fc1
  __syncthreads();
  for (int c = threadIdx.x; c < 128; c += blockDim.x) {   // h = ctx + e
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += s.dh[b * 128 + c];
    s.dctx[c] = acc;
  }
  __syncthreads();
  {   // ctx = sum_r w_r X_r ; w = softmax(att) over rows
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = warp; r < 2 * B; r += nw) {
      float acc = 0.f;
      for (int c = lane; c < 128; c += 32) acc = fmaf(s.X[r * 128 + c], s.dctx[c], acc);
      acc = warp_sum(acc);
      if (lane == 0) s.dwv[r] = acc;
    }
  }
  __syncthreads();
  float dot = 0.f;
  for (int r = threadIdx.x; r < 2 * B; r += blockDim.x) dot += s.wsm[r] * s.dwv[r];
  dot = block_reduce_sum(dot, red);
  for (int r = threadIdx.x; r < 2 * B; r += blockDim.x) s.da[r] = s.wsm[r] * (s.dwv[r] - dot);
  __syncthreads();
  for (int o = threadIdx.x; o < 2 * B * 128; o += blockDim.x) {
    const int r = o / 128, c = o % 128;
    float v = s.wsm[r] * s.dctx[c] + s.da[r] * P.p[R_CAW][c];
    if (r >= B) v += s.dh[(r - B) * 128 + c];   // direct path h = ctx + e
    s.dX[o] = v;
  }
  __syncthreads();
  cta_layernorm128_bwd(s.dX, B, P.p[R_FNW], s.fxh, s.frs, s.du);                    // feature branch: f = LN(u), u = feats Wf^T + bf
  cta_layernorm128_bwd(s.dX + B * 128, B, P.p[R_LNW], s.exh, s.ers, s.demb);        // label branch: e = LN(Emb[label])
  __syncthreads();
  if (use_smem) {
    const int64_t n = rew_ws_floats(B);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) ws[i] = dyn[i];
  }
}

// one row per Rewarder tensor: how its gradient is formed from the activations rewarder_actgrad_kernel left in the workspace
enum { RG_LINEAR = 0, RG_BIAS = 1, RG_LN_GAMMA = 2, RG_EMB = 3 };
struct RewGradRow {
  int kind; int rows; int N, K;           // LINEAR: dW[n, k] = sum_r dY[r, n] X[r, k];  BIAS: db[n] = sum_r dY[r, n]
  const float* dY; const float* X; int64_t ldx;   // LN_GAMMA: dg[c] = sum_r dY[r, c] X[r, c] (K = 128);  EMB: dE[l, c] = sum_b [label_b == l] dY[b, c]
  int64_t first;                          // exclusive prefix sum of the element counts
};
struct RewGradTable { RewGradRow t[R_NUM]; int64_t total; };

__global__ void __launch_bounds__(256) rewarder_wgrad_adam_kernel(RewPtrs P, RewPtrs G, RewPtrs M, RewPtrs V, RewPtrs GA, int has_g_add,
                                                                  const __grid_constant__ RewGradTable tab, const int64_t* __restrict__ labels, int label_rows,
                                                                  AdamScalars ad, int phase) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= tab.total) return;
  int t = 0;
#pragma unroll 1
  while (t + 1 < R_NUM && i >= tab.t[t + 1].first) ++t;
  const RewGradRow& r = tab.t[t];
  const int64_t e = i - r.first;
  float gi;
  if (phase != 2) {
    float acc = 0.f;
    if (r.kind == RG_LINEAR) {
      const int n = (int)(e / r.K), k = (int)(e % r.K);
      for (int q = 0; q < r.rows; ++q) acc = fmaf(r.dY[q * r.N + n], r.X[(int64_t)q * r.ldx + k], acc);
    } else if (r.kind == RG_BIAS) {
      for (int q = 0; q < r.rows; ++q) acc += r.dY[q * r.N + (int)e];
    } else if (r.kind == RG_LN_GAMMA) {
      for (int q = 0; q < r.rows; ++q) acc = fmaf(r.dY[q * 128 + (int)e], r.X[q * 128 + (int)e], acc);
    } else {   // embedding rows: repeated labels summed in batch order
      const int l = (int)(e / 128), c = (int)(e % 128);
      for (int q = 0; q < r.rows; ++q) {
        int64_t lb = labels[q];
        lb = lb < 0 ? 0 : (lb >= label_rows ? label_rows - 1 : lb);
        if (lb == l) acc += r.dY[q * 128 + c];
      }
    }
    G.p[t][e] = acc;
    gi = acc;
    if (phase == 1) return;
  } else {
    gi = G.p[t][e];
    if (has_g_add) gi += GA.p[t][e];
  }
  // torch.optim.Adam (single-tensor math): m.lerp_(g, 1-b1); v = v*b2 + (1-b2) g g; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
  const float step_size = (float)((double)ad.lr / ad.bc1);
  const float bc2s = (float)ad.bc2_sqrt;
  const float m0 = M.p[t][e], v0 = V.p[t][e];
  const float mi = m0 + 0.1f * (gi - m0);
  const float vi = v0 * 0.999f + (0.001f * gi) * gi;
  M.p[t][e] = mi;
  V.p[t][e] = vi;
  const float denom = sqrtf(vi) / bc2s + 1e-8f;
  P.p[t][e] = P.p[t][e] - step_size * (mi / denom);
}

static void fill_ptrs(RewPtrs& dst, float* const* src) {
  for (int i = 0; i < R_NUM; ++i) dst.p[i] = src[i];
}

}  // namespace srw

using namespace srw;

extern "C" int srw_flexmatch_mask(const srw_flexmatch_mask_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->logits_w && a->pseudo && a->mask, "srw_flexmatch_mask: null pointer");
  SRW_REQUIRE(a->selected_label == nullptr || (a->idx_ulb && a->hist && a->classwise_acc), "srw_flexmatch_mask: FlexMatch state incomplete");
  SRW_REQUIRE(a->B > 0 && a->B <= MAX_ROWS && a->num_classes > 0, "srw_flexmatch_mask: 0 < B <= %d required (B=%d)", MAX_ROWS, a->B);
  flexmatch_mask_kernel<<<1, SSL_THREADS, 0, stream>>>(*a);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_ssl_loss(const srw_ssl_loss_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->logits_lb && a->logits_s && a->y_lb && a->pseudo && a->mask && a->losses, "srw_ssl_loss: null pointer");
  SRW_REQUIRE(a->B_lb > 0 && a->B_lb <= MAX_ROWS && a->B_ulb > 0 && a->B_ulb <= MAX_ROWS && a->num_classes > 0, "srw_ssl_loss: batch sizes must be in (0, %d]", MAX_ROWS);
  RewPtrs RP = {};
  int use_smem = 0;
  size_t dyn = 0;
  if (a->rp) {
    SRW_REQUIRE(a->feats && a->feature_dim > 0 && a->label_rows > 0 && a->rew_workspace, "srw_ssl_loss: the fused Rewarder needs feats, feature_dim, label_rows and rew_workspace");
    for (int i = 0; i < R_NUM; ++i) RP.p[i] = const_cast<float*>(a->rp[i]);
    const int64_t wsf = rew_ws_floats(a->B_ulb);
    use_smem = wsf * 4 <= 176 * 1024 ? 1 : 0;    // next to the kernel's 24 KB of static shared memory
    dyn = use_smem ? (size_t)wsf * 4 : 0;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(ssl_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 176 * 1024); });
    SRW_CUDA(attr_err);
  }
  ssl_loss_kernel<<<1, SSL_THREADS, dyn, stream>>>(*a, RP, a->rp ? 1 : 0, use_smem);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int64_t srw_rewarder_workspace_floats(int B, int feature_dim) {
  (void)feature_dim;
  if (B <= 0) return -1;
  return rew_ws_floats(B);
}

extern "C" int srw_rewarder_fwd(const srw_rewarder_fwd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->rp && a->feats && a->labels && a->reward && a->workspace, "srw_rewarder_fwd: null pointer");
  SRW_REQUIRE(a->B > 0 && a->B <= MAX_ROWS && a->feature_dim > 0 && a->label_rows > 0, "srw_rewarder_fwd: bad shape");
  RewPtrs P;
  for (int i = 0; i < R_NUM; ++i) P.p[i] = const_cast<float*>(a->rp[i]);
  const int64_t wsf = rew_ws_floats(a->B);
  const int use_smem = wsf * 4 <= 200 * 1024 ? 1 : 0;
  static std::once_flag once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [] { attr_err = cudaFuncSetAttribute(rewarder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
  SRW_CUDA(attr_err);
  rewarder_fwd_kernel<<<1, SSL_THREADS, use_smem ? (size_t)wsf * 4 : 0, stream>>>(P, a->feats, a->ld_feats, a->labels, a->B, a->feature_dim, a->label_rows, a->reward,
                                                                                a->workspace, use_smem);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_generator_fwd(const srw_generator_fwd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->gp && a->feats && a->labels && a->workspace, "srw_generator_fwd: null pointer");
  SRW_REQUIRE(a->B > 0 && a->B <= MAX_ROWS && a->feature_dim > 0, "srw_generator_fwd: bad shape");
  GenPtrs P;
  for (int i = 0; i < 8; ++i) P.p[i] = a->gp[i];
  generator_fwd_kernel<<<1, SSL_THREADS, 0, stream>>>(P, a->feats, a->ld_feats, a->B, a->feature_dim, a->labels, a->workspace);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_rewarder_train(const srw_rewarder_train_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->rp && a->g && a->m && a->v && a->feats && a->gen_labels && a->true_labels && a->losses && a->workspace, "srw_rewarder_train: null pointer");
  SRW_REQUIRE(a->B > 0 && a->B <= MAX_ROWS && a->feature_dim > 0 && a->label_rows > 0 && a->step >= 1, "srw_rewarder_train: bad shape / step");
  SRW_REQUIRE(a->loss_select >= 0 && a->loss_select <= 2, "srw_rewarder_train: loss_select must be 0, 1 or 2");
  RewPtrs P, G, M, V, GA = {};
  fill_ptrs(P, a->rp); fill_ptrs(G, a->g); fill_ptrs(M, a->m); fill_ptrs(V, a->v);
  if (a->g_add) fill_ptrs(GA, a->g_add);
  RewSizes N;
  const int64_t D = a->feature_dim, Lr = a->label_rows;
  const int64_t sizes[R_NUM] = {128 * D, 128, 128, 128, Lr * 128, 128, 128, 128, 1, 256 * 128, 256, 128 * 256, 128, 64 * 128, 64, 64, 1};
  for (int i = 0; i < R_NUM; ++i) N.n[i] = sizes[i];
  AdamScalars ad;
  ad.lr = a->lr;
  ad.bc1 = 1.0 - pow(0.9, (double)a->step);
  ad.bc2_sqrt = sqrt(1.0 - pow(0.999, (double)a->step));
  const int B = a->B;
  if (a->phase != 2) {
    const int64_t wsf = rew_ws_floats(B);
    const int use_smem = wsf * 4 <= 200 * 1024 ? 1 : 0;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(rewarder_actgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); });
    SRW_CUDA(attr_err);
    rewarder_actgrad_kernel<<<1, SSL_THREADS, use_smem ? (size_t)wsf * 4 : 0, stream>>>(P, a->feats, a->ld_feats, a->gen_labels, a->true_labels, B, a->feature_dim,
                                                                                     a->label_rows, a->losses, a->workspace, a->loss_select, use_smem);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // gradient recipes (workspace offsets of rew_carve)
  RewGradTable tab = {};
  {
    const RewWs c = rew_carve(a->workspace, B);
    float *fxh = c.fxh, *X = c.X, *exh = c.exh, *h = c.h, *h1 = c.h1, *h2 = c.h2, *h3 = c.h3, *dz4 = c.dz4, *dh3 = c.dh3, *dh2 = c.dh2, *dh1 = c.dh1,
          *dX = c.dX, *du = c.du, *demb = c.demb, *da = c.da;
    auto lin = [&](int t, int rows, int N_, int K_, const float* dY, const float* Xp, int64_t ldx) {
      tab.t[t].kind = RG_LINEAR; tab.t[t].rows = rows; tab.t[t].N = N_; tab.t[t].K = K_; tab.t[t].dY = dY; tab.t[t].X = Xp; tab.t[t].ldx = ldx;
    };
    auto bias = [&](int t, int rows, int N_, const float* dY) { tab.t[t].kind = RG_BIAS; tab.t[t].rows = rows; tab.t[t].N = N_; tab.t[t].K = 1; tab.t[t].dY = dY; };
    auto lng = [&](int t, int rows, const float* dY, const float* xh) { tab.t[t].kind = RG_LN_GAMMA; tab.t[t].rows = rows; tab.t[t].N = 128; tab.t[t].K = 128; tab.t[t].dY = dY; tab.t[t].X = xh; };
    lin(R_FCW, B, 128, a->feature_dim, du, a->feats, a->ld_feats); bias(R_FCB, B, 128, du);
    lng(R_FNW, B, dX, fxh); bias(R_FNB, B, 128, dX);
    tab.t[R_EMB].kind = RG_EMB; tab.t[R_EMB].rows = B; tab.t[R_EMB].N = 128; tab.t[R_EMB].K = 128; tab.t[R_EMB].dY = demb;
    lng(R_LNW, B, dX + B * 128, exh); bias(R_LNB, B, 128, dX + B * 128);
    lin(R_CAW, 2 * B, 1, 128, da, X, 128); bias(R_CAB, 2 * B, 1, da);
    lin(R_M1W, B, 256, 128, dh1, h, 128); bias(R_M1B, B, 256, dh1);
    lin(R_M2W, B, 128, 256, dh2, h1, 256); bias(R_M2B, B, 128, dh2);
    lin(R_F1W, B, 64, 128, dh3, h2, 128); bias(R_F1B, B, 64, dh3);
    lin(R_F2W, B, 1, 64, dz4, h3, 64); bias(R_F2B, B, 1, dz4);
    int64_t off = 0;
    for (int i = 0; i < R_NUM; ++i) { tab.t[i].first = off; off += N.n[i]; }
    tab.total = off;
  }
  rewarder_wgrad_adam_kernel<<<(int)cdiv64(tab.total, 256), 256, 0, stream>>>(P, G, M, V, GA, a->g_add ? 1 : 0, tab, a->gen_labels, a->label_rows, ad, a->phase);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
