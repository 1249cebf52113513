// srw_ssl2 — the FreeMatch and SoftMatch rows of the SemiReward step as single-CTA fused kernels:
//   srw_freematch_mask    : softmax + pseudo-labels + FreeMatchThresholdingHook.update/masking
//                           (semilearn/algorithms/freematch/utils.py:23-66, hooks/pseudo_label.py:40)
//   srw_freematch_entropy : the fairness entropy term of SRFreeMatch and its gradient w.r.t. the strong logits
//                           (semilearn/algorithms/srfreematch/srfreematch.py:12-44, 214-218)
//   srw_softmatch_mask    : softmax + DistAlignEMAHook.dist_align + SoftMatchWeightingHook.update/masking
//                           (semilearn/algorithms/hooks/dist_align.py:25-55, srsoftmatch/utils.py:31-77)
// Every EMA follows the reference's fp32 evaluation order (tensor * python-scalar products are fp32 products with the
// scalar rounded to fp32; no FMA contraction), reductions run in index order (deterministic).  The hook state lives on
// the device; nothing is read back to the host (the reference's two .item() syncs in SoftMatch disappear).
#include <atomic>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {
extern std::atomic<int64_t> g_launches;

constexpr int SSL2_THREADS = 1024;
constexpr int SSL2_MAX_ROWS = 2048;

__device__ __forceinline__ float blk_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < nw; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ float blk_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = red[0];
  for (int w = 1; w < nw; ++w) s = fmaxf(s, red[w]);
  return s;
}

// one warp: softmax of a row into `out` (may be NULL), returns (max prob, first argmax of the probabilities, first argmax of the logits)
__device__ __forceinline__ void warp_softmax_row(const float* row, int C, int lane, float* out, float& best_p, int& arg_p, int& arg_l) {
  float m = -INFINITY;
  int mi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float v = row[c];
    if (v > m) { m = v; mi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
  }
  float sum = 0.f;
  for (int c = lane; c < C; c += 32) sum += expf(row[c] - m);
  sum = warp_sum(sum);
  float bp = -1.f;
  int bi = 0x7fffffff;
  for (int c = lane; c < C; c += 32) {
    const float pr = expf(row[c] - m) / sum;
    if (out) out[c] = pr;
    if (pr > bp) { bp = pr; bi = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, bp, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > bp || (ob == bp && oi < bi)) { bp = ob; bi = oi; }
  }
  best_p = bp; arg_p = bi; arg_l = mi;
}

// ------------------------------------------------------------------------------------------------
// FreeMatch
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSL2_THREADS) freematch_mask_kernel(const srw_freematch_mask_args a, float m_f, float om_f) {
  extern __shared__ int s_hist[];                   // [C]
  __shared__ float s_maxp[SSL2_MAX_ROWS];           // local rows
  __shared__ int s_maxi[SSL2_MAX_ROWS];
  __shared__ float s_allp[SSL2_MAX_ROWS];           // rows update() sees (== local rows unless probs_all is given)
  __shared__ int s_alli[SSL2_MAX_ROWS];
  __shared__ float red[32];
  __shared__ float s_q[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int C = a.num_classes, B = a.B;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_hist[c] = 0;
  if (a.phase != 2) {   // softmax of the local rows
    for (int b = warp; b < B; b += nw) {
      float bp; int ap, al;
      warp_softmax_row(a.logits_w + (int64_t)b * a.ld_logits, C, lane, a.probs_w + (int64_t)b * C, bp, ap, al);
      if (lane == 0) {
        s_maxp[b] = bp; s_maxi[b] = ap;
        a.pseudo[b] = a.pseudo_from_probs ? ap : al;
        if (a.max_probs) a.max_probs[b] = bp;
      }
    }
    if (a.phase == 1) return;
  } else {              // phase 2: probs_w was written by phase 1; only max / argmax are needed again
    for (int b = warp; b < B; b += nw) {
      const float* pr = a.probs_w + (int64_t)b * C;
      float bp = -1.f; int bi = 0x7fffffff;
      for (int c = lane; c < C; c += 32) if (pr[c] > bp) { bp = pr[c]; bi = c; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, bp, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > bp || (ob == bp && oi < bi)) { bp = ob; bi = oi; }
      }
      if (lane == 0) { s_maxp[b] = bp; s_maxi[b] = bi; }
    }
  }
  // rows seen by update(): all ranks' probabilities when given (concat_all_gather, freematch/utils.py:25-26), else the local ones
  const float* up = a.probs_all ? a.probs_all : a.probs_w;
  const int BU = a.probs_all ? a.B_all : B;
  __syncthreads();
  if (a.probs_all) {
    for (int b = warp; b < BU; b += nw) {
      const float* pr = up + (int64_t)b * C;
      float bp = -1.f; int bi = 0x7fffffff;
      for (int c = lane; c < C; c += 32) if (pr[c] > bp) { bp = pr[c]; bi = c; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, bp, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > bp || (ob == bp && oi < bi)) { bp = ob; bi = oi; }
      }
      if (lane == 0) { s_allp[b] = bp; s_alli[b] = bi; }
    }
  } else {
    for (int b = threadIdx.x; b < B; b += blockDim.x) { s_allp[b] = s_maxp[b]; s_alli[b] = s_maxi[b]; }
  }
  __syncthreads();
  // ---- update(): time_p ----
  float stat;
  if (a.use_quantile) {
    // torch.quantile(x, 0.8): rank = 0.8f * (n - 1) in fp32, linear interpolation between the two neighbouring order statistics
    const float rank = __fmul_rn(0.8f, (float)(BU - 1));
    const int lo = (int)floorf(rank), hi = (int)ceilf(rank);
    for (int i = threadIdx.x; i < BU; i += blockDim.x) {
      const float v = s_allp[i];
      int r = 0;
      for (int j = 0; j < BU; ++j) {
        const float u = s_allp[j];
        r += (u < v || (u == v && j < i)) ? 1 : 0;
      }
      if (r == lo) s_q[0] = v;
      if (r == hi) s_q[1] = v;
    }
    __syncthreads();
    const float w = rank - (float)lo, lo_v = s_q[0], hi_v = s_q[1];
    // at::lerp: w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
    stat = w < 0.5f ? __fadd_rn(lo_v, __fmul_rn(w, __fsub_rn(hi_v, lo_v))) : __fsub_rn(hi_v, __fmul_rn(__fsub_rn(hi_v, lo_v), __fsub_rn(1.0f, w)));
  } else {
    float s = 0.f;
    if (threadIdx.x == 0) {
      for (int i = 0; i < BU; ++i) s += s_allp[i];
      s_q[0] = s / (float)BU;
    }
    __syncthreads();
    stat = s_q[0];
  }
  float time_p = __fadd_rn(__fmul_rn(*a.time_p, m_f), __fmul_rn(om_f, stat));
  if (a.clip_thresh) time_p = fminf(fmaxf(time_p, 0.0f), 0.95f);
  // ---- p_model, label_hist ----
  for (int b = threadIdx.x; b < BU; b += blockDim.x) atomicAdd(&s_hist[s_alli[b]], 1);
  __syncthreads();
  float pmax = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < BU; ++b) s += up[(int64_t)b * C + c];
    const float mean = s / (float)BU;
    const float pm = __fadd_rn(__fmul_rn(a.p_model[c], m_f), __fmul_rn(om_f, mean));
    a.p_model[c] = pm;
    pmax = fmaxf(pmax, pm);
    a.label_hist[c] = __fadd_rn(__fmul_rn(a.label_hist[c], m_f), __fmul_rn(om_f, (float)s_hist[c] / (float)BU));
  }
  pmax = blk_max(pmax, red);
  __syncthreads();   // p_model writes of this CTA are visible to all its threads
  // ---- masking(): max_p >= time_p * p_model[idx] / max(p_model), local rows ----
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float mod = a.p_model[s_maxi[b]] / pmax;
    a.mask[b] = s_maxp[b] >= __fmul_rn(time_p, mod) ? 1.0f : 0.0f;
  }
  __syncthreads();
  if (threadIdx.x == 0) *a.time_p = time_p;
}

// entropy term: loss = sum_c mpm[c] * log(mmn[c] + 1e-12) over the rows with mask != 0 (srfreematch.py:16-44)
__global__ void __launch_bounds__(SSL2_THREADS) freematch_entropy_kernel(const srw_freematch_entropy_args a) {
  extern __shared__ float sm[];
  const int C = a.num_classes, B = a.B;
  float* s_mean = sm;            // [C] sum -> mean of the selected rows' probabilities
  float* s_mpm = sm + C;         // [C] modulated p_model (normalised)
  float* s_g = sm + 2 * C;       // [C] d loss / d mean
  int* s_hist = reinterpret_cast<int*>(sm + 3 * C);   // [C]
  __shared__ int s_sel[SSL2_MAX_ROWS];
  __shared__ float red[32];
  __shared__ int s_n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int b = 0; b < B; ++b)
      if (a.mask[b] != 0.0f) s_sel[n++] = b;
    s_n = n;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) { s_mean[c] = 0.f; s_hist[c] = 0; }
  __syncthreads();
  const int n = s_n;
  if (n == 0) {   // "if mask.sum() > 0 ... else ent_loss = 0.0" (srfreematch.py:214-217)
    if (threadIdx.x == 0) a.losses[4] = 0.f;
    if (a.dlogits_s && !a.accumulate)
      for (int i = threadIdx.x; i < B * C; i += blockDim.x) a.dlogits_s[(int64_t)(i / C) * a.ld_dlogits + i % C] = 0.f;
    return;
  }
  // argmax histogram of the selected rows (integer atomics: deterministic)
  for (int i = warp; i < n; i += nw) {
    float bp; int ap, al;
    warp_softmax_row(a.logits_s + (int64_t)s_sel[i] * a.ld_logits, C, lane, nullptr, bp, ap, al);
    if (lane == 0) atomicAdd(&s_hist[ap], 1);
  }
  __syncthreads();
  // mean probability per class over the selected rows, summed in row order (thread = class); per-row softmax statistics first
  __shared__ float s_m[SSL2_MAX_ROWS], s_sum[SSL2_MAX_ROWS];
  for (int i = warp; i < n; i += nw) {
    const float* row = a.logits_s + (int64_t)s_sel[i] * a.ld_logits;
    float m = -INFINITY;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += expf(row[c] - m);
    s = warp_sum(s);
    if (lane == 0) { s_m[i] = m; s_sum[i] = s; }
  }
  __syncthreads();
  float part_mpm = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += expf(a.logits_s[(int64_t)s_sel[i] * a.ld_logits + c] - s_m[i]) / s_sum[i];
    s_mean[c] = s / (float)n;
    const float lh = a.label_hist[c];
    const float inv = lh == 0.f ? 0.f : 1.0f / lh;       // replace_inf_to_zero(1 / label_hist)
    const float v = a.p_model[c] * inv;
    s_mpm[c] = v;
    part_mpm += v;
  }
  const float sum_mpm = blk_sum(part_mpm, red);
  float part_S = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float hs = (float)s_hist[c] / (float)n;
    const float inv_s = s_hist[c] == 0 ? 0.f : 1.0f / hs;
    part_S += s_mean[c] * inv_s;
  }
  const float S = blk_sum(part_S, red);
  float part_loss = 0.f, part_A = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float hs = (float)s_hist[c] / (float)n;
    const float inv_s = s_hist[c] == 0 ? 0.f : 1.0f / hs;
    const float mpm = s_mpm[c] / sum_mpm;
    const float mmn = s_mean[c] * inv_s / S;
    part_loss += mpm * logf(mmn + 1e-12f);
    const float ac = mpm / (mmn + 1e-12f);
    s_g[c] = ac;                 // a[c]
    part_A += ac * mmn;
  }
  const float loss = blk_sum(part_loss, red);
  const float A = blk_sum(part_A, red);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float hs = (float)s_hist[c] / (float)n;
    const float inv_s = s_hist[c] == 0 ? 0.f : 1.0f / hs;
    s_g[c] = inv_s * (s_g[c] - A) / S / (float)n;    // d loss / d prob_s[i, c] for every selected row i
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.losses[4] = loss;
    a.losses[2] += a.lambda_e * loss;                // total = sup + lambda_u unsup + lambda_e ent (srfreematch.py:219)
  }
  if (!a.dlogits_s) return;
  if (!a.accumulate) {
    for (int i = threadIdx.x; i < B * C; i += blockDim.x) a.dlogits_s[(int64_t)(i / C) * a.ld_dlogits + i % C] = 0.f;
    __syncthreads();
  }
  for (int i = warp; i < n; i += nw) {
    const int b = s_sel[i];
    const float* row = a.logits_s + (int64_t)b * a.ld_logits;
    const float m = s_m[i], sum = s_sum[i];
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot += expf(row[c] - m) / sum * s_g[c];
    dot = warp_sum(dot);
    float* d = a.dlogits_s + (int64_t)b * a.ld_dlogits;
    for (int c = lane; c < C; c += 32) {
      const float pr = expf(row[c] - m) / sum;
      d[c] += a.lambda_e * pr * (s_g[c] - dot);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SoftMatch (+ DistAlign EMA)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSL2_THREADS) softmatch_mask_kernel(const srw_softmatch_mask_args a, float m_f, float om_f, double om_d) {
  __shared__ float s_maxp[SSL2_MAX_ROWS];
  __shared__ float s_stat[2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int C = a.num_classes, B = a.B;
  if (a.phase <= 1) {   // softmax of the local rows
    for (int b = warp; b < B; b += nw) {
      float bp; int ap, al;
      warp_softmax_row(a.logits_w + (int64_t)b * a.ld_logits, C, lane, a.probs_w + (int64_t)b * C, bp, ap, al);
      if (lane == 0) {
        s_maxp[b] = bp;
        a.pseudo[b] = a.pseudo_from_probs ? ap : al;
        if (a.phase == 1 && a.max_probs) a.max_probs[b] = bp;
      }
    }
    if (a.phase == 1) return;
    __syncthreads();
  }
  if (a.dist_align && (a.phase == 0 || a.phase == 2)) {
    // DistAlignEMAHook.update_p: p_model = mean(probs) on the first call, EMA afterwards (dist_align.py:44-48); under data
    // parallelism the mean runs over every rank's probabilities (probs_all, dist_align.py:40-42)
    const float* up = a.probs_all ? a.probs_all : a.probs_w;
    const int BU = a.probs_all ? a.B_all : B;
    const int first = *a.da_initialized == 0;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int b = 0; b < BU; ++b) s += up[(int64_t)b * C + c];
      const float mean = s / (float)BU;
      a.da_p_model[c] = first ? mean : __fadd_rn(__fmul_rn(a.da_p_model[c], m_f), __fmul_rn(mean, om_f));
    }
    __syncthreads();
    if (threadIdx.x == 0) *a.da_initialized = 1;
    // aligned = probs * (p_target + 1e-6) / (p_model + 1e-6), renormalised per row; only its row maximum is used
    for (int b = warp; b < B; b += nw) {
      const float* pr = a.probs_w + (int64_t)b * C;
      float s = 0.f, mx = -INFINITY;
      for (int c = lane; c < C; c += 32) {
        const float v = __fdiv_rn(__fmul_rn(pr[c], __fadd_rn(a.da_p_target[c], 1e-6f)), __fadd_rn(a.da_p_model[c], 1e-6f));
        if (a.probs_aligned) a.probs_aligned[(int64_t)b * C + c] = v;   // un-normalised; normalised below
        s += v;
        mx = fmaxf(mx, v);
      }
      s = warp_sum(s);
      mx = warp_max(mx);
      if (a.probs_aligned)
        for (int c = lane; c < C; c += 32) a.probs_aligned[(int64_t)b * C + c] /= s;
      if (lane == 0) {
        s_maxp[b] = mx / s;
        if (a.phase == 2 && a.max_probs) a.max_probs[b] = mx / s;
      }
    }
    if (a.phase == 2) return;
    __syncthreads();
  }
  if (a.phase == 3)
    for (int b = threadIdx.x; b < B; b += blockDim.x) s_maxp[b] = a.max_probs[b];
  __syncthreads();
  // SoftMatchWeightingHook.update (per_class = False): EMA of mean / unbiased variance of max_probs (of every rank's rows
  // under data parallelism: maxp_all, srsoftmatch/utils.py:33-34)
  if (threadIdx.x == 0) {
    const float* vals = a.maxp_all ? a.maxp_all : s_maxp;
    const int n = a.maxp_all ? a.n_all : B;
    float s = 0.f;
    for (int b = 0; b < n; ++b) s += vals[b];
    const float mu = s / (float)n;
    double q = 0.0;
    for (int b = 0; b < n; ++b) { const double d = (double)vals[b] - (double)mu; q += d * d; }
    const float var = n > 1 ? (float)(q / (double)(n - 1)) : NAN;   // torch.var(unbiased=True) of one element is nan
    // m * t + (1 - m) * x.item(): fp32 tensor product + python double product rounded to fp32 by the tensor add
    const float mu_t = __fadd_rn(__fmul_rn(m_f, *a.prob_max_mu_t), (float)(om_d * (double)mu));
    const float var_t = __fadd_rn(__fmul_rn(m_f, *a.prob_max_var_t), (float)(om_d * (double)var));
    *a.prob_max_mu_t = mu_t; *a.prob_max_var_t = var_t;
    s_stat[0] = mu_t; s_stat[1] = var_t;
  }
  __syncthreads();
  const float mu_t = s_stat[0];
  const float denom = __fdiv_rn(__fmul_rn(2.0f, s_stat[1]), (float)(a.n_sigma * a.n_sigma));
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float d = fminf(__fsub_rn(s_maxp[b], mu_t), 0.0f);
    a.mask[b] = expf(-__fdiv_rn(__fmul_rn(d, d), denom));
    if (a.max_probs && a.phase == 0) a.max_probs[b] = s_maxp[b];
  }
}

}  // namespace srw

using namespace srw;

extern "C" int srw_freematch_mask(const srw_freematch_mask_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->logits_w && a->time_p && a->p_model && a->label_hist && a->probs_w && a->pseudo && a->mask, "srw_freematch_mask: null pointer");
  SRW_REQUIRE(a->B > 0 && a->B <= SSL2_MAX_ROWS && a->num_classes > 0 && a->num_classes <= 8192, "srw_freematch_mask: 0 < B <= %d, C <= 8192 required", SSL2_MAX_ROWS);
  SRW_REQUIRE(a->phase >= 0 && a->phase <= 2 && (!a->probs_all || (a->B_all >= a->B && a->B_all <= SSL2_MAX_ROWS)), "srw_freematch_mask: bad phase / B_all");
  const float m_f = (float)a->momentum, om_f = (float)(1.0 - a->momentum);
  freematch_mask_kernel<<<1, SSL2_THREADS, a->num_classes * sizeof(int), stream>>>(*a, m_f, om_f);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_freematch_entropy(const srw_freematch_entropy_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->mask && a->logits_s && a->p_model && a->label_hist && a->losses, "srw_freematch_entropy: null pointer");
  SRW_REQUIRE(a->B > 0 && a->B <= SSL2_MAX_ROWS && a->num_classes > 0 && a->num_classes <= 8192, "srw_freematch_entropy: 0 < B <= %d, C <= 8192 required", SSL2_MAX_ROWS);
  static bool attr_set = false;
  const size_t smem = (size_t)a->num_classes * 4 * sizeof(float);
  if (!attr_set && smem > 24 * 1024) {
    SRW_CUDA(cudaFuncSetAttribute(freematch_entropy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 4 * (int)sizeof(float)));
    attr_set = true;
  }
  freematch_entropy_kernel<<<1, SSL2_THREADS, smem, stream>>>(*a);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_softmatch_mask(const srw_softmatch_mask_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->logits_w && a->prob_max_mu_t && a->prob_max_var_t && a->probs_w && a->pseudo && a->mask, "srw_softmatch_mask: null pointer");
  SRW_REQUIRE(!a->dist_align || (a->da_p_model && a->da_p_target && a->da_initialized), "srw_softmatch_mask: dist_align needs its state");
  SRW_REQUIRE(a->B > 0 && a->B <= SSL2_MAX_ROWS && a->num_classes > 0 && a->n_sigma > 0, "srw_softmatch_mask: 0 < B <= %d required", SSL2_MAX_ROWS);
  SRW_REQUIRE(a->phase >= 0 && a->phase <= 3 && (a->phase == 0 || a->max_probs), "srw_softmatch_mask: phases 1-3 need max_probs");
  SRW_REQUIRE((!a->probs_all || a->B_all >= a->B) && (!a->maxp_all || a->n_all >= a->B), "srw_softmatch_mask: gathered inputs smaller than the local batch");
  const float m_f = (float)a->momentum, om_f = (float)(1.0 - a->momentum);
  softmatch_mask_kernel<<<1, SSL2_THREADS, 0, stream>>>(*a, m_f, om_f, 1.0 - a->momentum);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
