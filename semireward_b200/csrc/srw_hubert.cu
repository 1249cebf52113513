// srw_hubert — ClassificationHubert.forward (semilearn/nets/hubert/hubert.py:24-50) and its backward as two native calls.
//
// The encoder arithmetic is Hugging Face transformers' HubertModel (un-vendored dependency of the reference, `transformers>=4.30.0`,
// 5.5.0 in this image; modeling_hubert.py: HubertFeatureEncoder / HubertFeatureProjection / HubertPositionalConvEmbedding /
// HubertEncoder with do_stable_layer_norm = False) restated on this library's kernels (include/srw.h has the op list).
//
// Convolutions as GEMMs without an im2col copy.  Activations of the conv stem are time-major planes [clip, frame, 512]; clip b's
// frames start at row b * tp (tp = padded frames per clip, chosen so that tp[l - 1] = stride[l] * tp[l] through the whole stem:
// 12800 / 6400 / ... / 200 for 64 000 samples).  Output frame t of a Conv1d(k, stride s) reads input frames s t .. s t + k - 1,
// which are k * 512 CONTIGUOUS elements starting at row s t: the im2col matrix is the input itself seen through a tensor map with
// row stride s * 512 < k * 512 (overlapping rows) — TMA does the gather, the GEMM kernel does not know.  The one padding frame per
// clip is computed (from finite garbage) and zeroed; valid frames never read it.  Backward: wgrad = the same view as the K-major
// operand; dgrad of a stride-2 conv splits into even / odd input frames, each a GEMM over dZ (the even one over an overlapping
// two-frame view of dZ), with gelu'(z) of the previous layer in the epilogue.
// The positional conv (k = 128, groups 16) uses a group-major zero-padded copy of h, [group][clip * P + frame][48]: frame t's
// receptive field is 128 * 48 contiguous elements, row stride 48 — again one GEMM per group over an overlapping view, split-K
// because K = 6144 exceeds the single-accumulator limit of srw_gemm.
// The first conv (1 input channel, k = 10) + GroupNorm + GELU is CUDA-core work bound by writing its 26 MB / clip output; it is
// recomputed from the waveform in the backward instead of being stored in fp32.
#include <cuda.h>

#include <vector>

#include "../../include/srw.h"
#include "srw_common.cuh"
#include "srw_encoder.cuh"

namespace srw {

constexpr int HC = 512;          // conv channels
constexpr int C0_FRAMES = 128;   // frames of the first conv handled by one CTA
constexpr int C0K = 10, C0S = 5;

// ------------------------------------------------------------------------------------------------
// first conv + GroupNorm + GELU
// ------------------------------------------------------------------------------------------------
// One CTA = (clip, chunk of C0_FRAMES frames), 256 threads, thread = 2 adjacent channels.  z[t, c] = sum_j w[c, j] x[5 t + j].
struct Conv0Ctx {
  float w0[C0K], w1[C0K];
  int c, t0, t1;
};
__device__ __forceinline__ void conv0_setup(Conv0Ctx& k, const float* __restrict__ wav, int64_t ld_wav, int samples, int n0, const float* __restrict__ W,
                                            float* xs) {
  const int s = blockIdx.y;
  k.c = threadIdx.x * 2;
  k.t0 = blockIdx.x * C0_FRAMES;
  k.t1 = min(k.t0 + C0_FRAMES, n0);
#pragma unroll
  for (int j = 0; j < C0K; ++j) { k.w0[j] = W[k.c * C0K + j]; k.w1[j] = W[(k.c + 1) * C0K + j]; }
  const int x0 = k.t0 * C0S, nx = (k.t1 - k.t0 - 1) * C0S + C0K;
  for (int i = threadIdx.x; i < nx; i += blockDim.x) xs[i] = (x0 + i < samples) ? wav[(int64_t)s * ld_wav + x0 + i] : 0.f;
  __syncthreads();
}
__device__ __forceinline__ void conv0_at(const Conv0Ctx& k, const float* xs, int t, float& z0, float& z1) {
  const float* x = xs + (t - k.t0) * C0S;
  float a = 0.f, b = 0.f;
#pragma unroll
  for (int j = 0; j < C0K; ++j) { const float xv = x[j]; a = fmaf(k.w0[j], xv, a); b = fmaf(k.w1[j], xv, b); }
  z0 = a; z1 = b;
}

// partial[(s, chunk), c] = (sum_t z, sum_t z^2) over the chunk's frames
__global__ void __launch_bounds__(256) hub_conv0_stats_kernel(const float* __restrict__ wav, int64_t ld_wav, int samples, int n0, const float* __restrict__ W,
                                                              float2* __restrict__ partial) {
  __shared__ float xs[C0_FRAMES * C0S + C0K];
  Conv0Ctx k;
  conv0_setup(k, wav, ld_wav, samples, n0, W, xs);
  float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
  for (int t = k.t0; t < k.t1; ++t) {
    float z0, z1;
    conv0_at(k, xs, t, z0, z1);
    s0 += z0; q0 = fmaf(z0, z0, q0); s1 += z1; q1 = fmaf(z1, z1, q1);
  }
  float2* dst = partial + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * HC + k.c;
  dst[0] = make_float2(s0, q0);
  dst[1] = make_float2(s1, q1);
}
// GroupNorm(512 groups) statistics of one (clip, channel): biased variance over the n0 frames (F.group_norm), partials folded in fp64
__global__ void hub_conv0_stats_finish_kernel(const float2* __restrict__ partial, int S, int chunks, int n0, float eps, float* __restrict__ mean,
                                              float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * HC) return;
  const int s = i / HC, c = i % HC;
  double a = 0.0, q = 0.0;
  for (int k = 0; k < chunks; ++k) {
    const float2 p = partial[((int64_t)s * chunks + k) * HC + c];
    a += p.x; q += p.y;
  }
  const double m = a / n0;
  const double var = fmax(q / n0 - m * m, 0.0);
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}
// y[(s, t), c] = gelu(gamma (z - mean) rstd + beta) as split planes; padding frames t in [n0, tp0) are written as zeros
__global__ void __launch_bounds__(256) hub_conv0_apply_kernel(const float* __restrict__ wav, int64_t ld_wav, int samples, int n0, int tp0, const float* __restrict__ W,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int64_t plane_stride) {
  __shared__ float xs[C0_FRAMES * C0S + C0K];
  Conv0Ctx k;
  const int s = blockIdx.y;
  conv0_setup(k, wav, ld_wav, samples, n0, W, xs);
  const float m0 = mean[s * HC + k.c], m1 = mean[s * HC + k.c + 1];
  const float a0 = rstd[s * HC + k.c] * gamma[k.c], a1 = rstd[s * HC + k.c + 1] * gamma[k.c + 1];
  const float b0 = beta[k.c], b1 = beta[k.c + 1];
  const int tend = min(blockIdx.x * C0_FRAMES + C0_FRAMES, tp0);
  for (int t = k.t0; t < tend; ++t) {
    uint32_t hi = 0, lo = 0;
    if (t < n0) {
      float z0, z1;
      conv0_at(k, xs, t, z0, z1);
      split2(gelu_f(fmaf(z0 - m0, a0, b0)), gelu_f(fmaf(z1 - m1, a1, b1)), hi, lo);
    }
    __nv_bfloat16* p = y + ((int64_t)s * tp0 + t) * HC + k.c;
    *reinterpret_cast<uint32_t*>(p) = hi;
    *reinterpret_cast<uint32_t*>(p + plane_stride) = lo;
  }
}
// backward, pass 1: du = dy gelu'(u); partial sums over the chunk of (du, du * nhat) per channel
__global__ void __launch_bounds__(256) hub_conv0_bwd_stats_kernel(const float* __restrict__ wav, int64_t ld_wav, int samples, int n0, int tp0,
                                                                  const float* __restrict__ W, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ dy,
                                                                  float2* __restrict__ partial) {
  __shared__ float xs[C0_FRAMES * C0S + C0K];
  Conv0Ctx k;
  const int s = blockIdx.y;
  conv0_setup(k, wav, ld_wav, samples, n0, W, xs);
  const float m0 = mean[s * HC + k.c], m1 = mean[s * HC + k.c + 1];
  const float r0 = rstd[s * HC + k.c], r1 = rstd[s * HC + k.c + 1];
  const float g0 = gamma[k.c], g1 = gamma[k.c + 1], b0 = beta[k.c], b1 = beta[k.c + 1];
  float a0 = 0.f, q0 = 0.f, a1 = 0.f, q1 = 0.f;
  for (int t = k.t0; t < k.t1; ++t) {
    float z0, z1;
    conv0_at(k, xs, t, z0, z1);
    const float2 d = *reinterpret_cast<const float2*>(dy + ((int64_t)s * tp0 + t) * HC + k.c);
    const float n0h = (z0 - m0) * r0, n1h = (z1 - m1) * r1;
    const float du0 = d.x * gelu_grad_f(fmaf(n0h, g0, b0)), du1 = d.y * gelu_grad_f(fmaf(n1h, g1, b1));
    a0 += du0; q0 = fmaf(du0, n0h, q0); a1 += du1; q1 = fmaf(du1, n1h, q1);
  }
  float2* dst = partial + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * HC + k.c;
  dst[0] = make_float2(a0, q0);
  dst[1] = make_float2(a1, q1);
}
// per (clip, channel): A = sum_t du, Q = sum_t du nhat (fp64 fold) -> ab[(s, c)] = (A / n0, Q / n0);  dbeta[c] (+)= sum_s A, dgamma[c] (+)= sum_s Q
__global__ void __launch_bounds__(256) hub_conv0_bwd_finish_kernel(const float2* __restrict__ partial, int Sg, int chunks, int n0, float2* __restrict__ ab,
                                                                   float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate) {
  // one warp per channel, lanes over the chunks of a clip (fp64), clips in order
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= HC) return;
  double ga = 0.0, gq = 0.0;
  for (int s = 0; s < Sg; ++s) {
    double a = 0.0, q = 0.0;
    for (int k = lane; k < chunks; k += 32) {
      const float2 p = partial[((int64_t)s * chunks + k) * HC + c];
      a += p.x; q += p.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
    if (lane == 0) ab[s * HC + c] = make_float2((float)(a / n0), (float)(q / n0));
    ga += a; gq += q;
  }
  if (lane == 0) {
    dbeta[c] = accumulate ? dbeta[c] + (float)ga : (float)ga;
    dgamma[c] = accumulate ? dgamma[c] + (float)gq : (float)gq;
  }
}
// backward, pass 2: dz = rstd gamma (du - A/n - nhat Q/n);  wpart[(s, chunk), c, j] = sum_t dz x[5 t + j]
__global__ void __launch_bounds__(256) hub_conv0_bwd_wgrad_kernel(const float* __restrict__ wav, int64_t ld_wav, int samples, int n0, int tp0,
                                                                  const float* __restrict__ W, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ dy,
                                                                  const float2* __restrict__ ab, float* __restrict__ wpart) {
  __shared__ float xs[C0_FRAMES * C0S + C0K];
  Conv0Ctx k;
  const int s = blockIdx.y;
  conv0_setup(k, wav, ld_wav, samples, n0, W, xs);
  const float m0 = mean[s * HC + k.c], m1 = mean[s * HC + k.c + 1];
  const float r0 = rstd[s * HC + k.c], r1 = rstd[s * HC + k.c + 1];
  const float g0 = gamma[k.c], g1 = gamma[k.c + 1], b0 = beta[k.c], b1 = beta[k.c + 1];
  const float2 ab0 = ab[s * HC + k.c], ab1 = ab[s * HC + k.c + 1];
  float acc0[C0K], acc1[C0K];
#pragma unroll
  for (int j = 0; j < C0K; ++j) acc0[j] = acc1[j] = 0.f;
  for (int t = k.t0; t < k.t1; ++t) {
    float z0, z1;
    conv0_at(k, xs, t, z0, z1);
    const float2 d = *reinterpret_cast<const float2*>(dy + ((int64_t)s * tp0 + t) * HC + k.c);
    const float n0h = (z0 - m0) * r0, n1h = (z1 - m1) * r1;
    const float du0 = d.x * gelu_grad_f(fmaf(n0h, g0, b0)), du1 = d.y * gelu_grad_f(fmaf(n1h, g1, b1));
    const float dz0 = r0 * g0 * (du0 - ab0.x - n0h * ab0.y), dz1 = r1 * g1 * (du1 - ab1.x - n1h * ab1.y);
    const float* x = xs + (t - k.t0) * C0S;
#pragma unroll
    for (int j = 0; j < C0K; ++j) { const float xv = x[j]; acc0[j] = fmaf(dz0, xv, acc0[j]); acc1[j] = fmaf(dz1, xv, acc1[j]); }
  }
  float* dst = wpart + (((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * HC + k.c) * C0K;
#pragma unroll
  for (int j = 0; j < C0K; ++j) { dst[j] = acc0[j]; dst[C0K + j] = acc1[j]; }
}
__global__ void hub_conv0_wgrad_finish_kernel(const float* __restrict__ wpart, int nparts, float* __restrict__ dW, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= HC * C0K) return;
  double a = 0.0;
  for (int p = 0; p < nparts; ++p) a += wpart[(int64_t)p * HC * C0K + i];
  dW[i] = accumulate ? dW[i] + (float)a : (float)a;
}

// ------------------------------------------------------------------------------------------------
// conv stem helpers (layers >= 1 run on srw_gemm)
// ------------------------------------------------------------------------------------------------
// planes[(b * tp + t), :] = 0 for t in [n, tp) of every clip, and for the `slack` rows behind the last clip
__global__ void hub_zero_pad_rows_kernel(__nv_bfloat16* __restrict__ planes, int64_t plane_stride, int S, int tp, int n, int slack, int cols) {
  const int per_clip = tp - n;
  const int total_rows = S * per_clip + slack;
  const int c8 = cols / 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (int64_t)total_rows * c8; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c8), c = (int)(i % c8) * 8;
    const int64_t row = r < S * per_clip ? (int64_t)(r / per_clip) * tp + n + r % per_clip : (int64_t)S * tp + (r - S * per_clip);
    *reinterpret_cast<uint4*>(planes + row * cols + c) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(planes + plane_stride + row * cols + c) = make_uint4(0, 0, 0, 0);
  }
}

// last conv output -> feature projection input: g = gelu(z) (compact rows), LayerNorm(512) -> planes.  One warp per frame.
__global__ void __launch_bounds__(256) hub_gelu_ln_fwd_kernel(const float* __restrict__ z, int S, int tp, int Fr, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float eps, float* __restrict__ g_out, float* __restrict__ mean_out,
                                                              float* __restrict__ rstd_out, __nv_bfloat16* __restrict__ yp, int64_t plane_stride) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= S * Fr) return;
  const int s = r / Fr, t = r % Fr;
  const float* zr = z + ((int64_t)s * tp + t) * HC;
  float4 v[4];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = j * 128 + lane * 4;
    const float4 a = *reinterpret_cast<const float4*>(zr + c);
    v[j] = make_float4(gelu_f(a.x), gelu_f(a.y), gelu_f(a.z), gelu_f(a.w));
    *reinterpret_cast<float4*>(g_out + (int64_t)r * HC + c) = v[j];
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(sum) * (1.0f / HC);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / HC) + eps);
  if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = j * 128 + lane * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    uint32_t h0, l0, h1, l1;
    split2(v[j].x * rstd * g.x + b.x, v[j].y * rstd * g.y + b.y, h0, l0);
    split2(v[j].z * rstd * g.z + b.z, v[j].w * rstd * g.w + b.w, h1, l1);
    __nv_bfloat16* hp = yp + (int64_t)r * HC + c;
    *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(hp + plane_stride) = make_uint2(l0, l1);
  }
}
// dz[(s, t), :] = dg[(s, t) compact, :] * gelu'(z[(s, t), :]) as planes in the padded row layout; padding frames zero
__global__ void hub_gelu_bwd_scatter_kernel(const float* __restrict__ dg, const float* __restrict__ z, int Sg, int tp, int Fr, __nv_bfloat16* __restrict__ dzp,
                                            int64_t plane_stride) {
  const int64_t total4 = (int64_t)Sg * tp * HC / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 4;
    const int64_t row = e / HC;
    const int c = (int)(e % HC), s = (int)(row / tp), t = (int)(row % tp);
    uint32_t h0 = 0, l0 = 0, h1 = 0, l1 = 0;
    if (t < Fr) {
      const float4 d = *reinterpret_cast<const float4*>(dg + ((int64_t)s * Fr + t) * HC + c);
      const float4 zz = *reinterpret_cast<const float4*>(z + e);
      split2(d.x * gelu_grad_f(zz.x), d.y * gelu_grad_f(zz.y), h0, l0);
      split2(d.z * gelu_grad_f(zz.z), d.w * gelu_grad_f(zz.w), h1, l1);
    }
    *reinterpret_cast<uint2*>(dzp + e) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(dzp + plane_stride + e) = make_uint2(l0, l1);
  }
}
// conv weight gradient: the wgrad GEMM leaves split-K partials of dWf[co, (j, ci)]; the parameter is stored [co, ci, j]
__global__ void hub_conv_wgrad_finish_kernel(const float* __restrict__ ws, int split, int k, float* __restrict__ dW, int accumulate) {
  const int64_t n = (int64_t)HC * HC * k;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i / (HC * k)), rem = (int)(i % (HC * k));
    const int j = rem / HC, ci = rem % HC;
    float a = 0.f;
    for (int p = 0; p < split; ++p) a += ws[(int64_t)p * n + i];
    float* d = dW + ((int64_t)co * HC + ci) * k + j;
    *d = accumulate ? *d + a : a;
  }
}

// ------------------------------------------------------------------------------------------------
// feature projection tail, positional conv
// ------------------------------------------------------------------------------------------------
// h = specaugment(dropout(h)) in place, and its group-major zero-padded copy xg[g][s * P + front + t][c % gc] (planes) for the
// positional conv (the padding rows were zeroed by a memset)
__global__ void hub_featproj_post_kernel(float* __restrict__ h, int S, int Fr, int D, const DropParams dr, const uint8_t* __restrict__ mask_time,
                                         const float* __restrict__ mse, __nv_bfloat16* __restrict__ xg, int64_t plane_stride, int64_t rows_per_group, int P,
                                         int front, int gc) {
  const int64_t total4 = (int64_t)S * Fr * D / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 4;
    const int64_t row = e / D;
    const int c = (int)(e % D), s = (int)(row / Fr), t = (int)(row % Fr);
    float4 v = *reinterpret_cast<const float4*>(h + e);
    if (dr.on) {
      const uint32_t key = drop_site_key(dr.seq_key[s], dr.site);
      const uint32_t base = ((uint32_t)dr.seq_row[s] * (uint32_t)Fr + (uint32_t)t) * (uint32_t)D + (uint32_t)c;
      v.x = drop_kept(key, base, dr.thr24) ? v.x * dr.inv_keep : 0.f;
      v.y = drop_kept(key, base + 1, dr.thr24) ? v.y * dr.inv_keep : 0.f;
      v.z = drop_kept(key, base + 2, dr.thr24) ? v.z * dr.inv_keep : 0.f;
      v.w = drop_kept(key, base + 3, dr.thr24) ? v.w * dr.inv_keep : 0.f;
    }
    if (mask_time && mask_time[row]) v = *reinterpret_cast<const float4*>(mse + c);
    *reinterpret_cast<float4*>(h + e) = v;
    uint32_t h0, l0, h1, l1;
    split2(v.x, v.y, h0, l0);
    split2(v.z, v.w, h1, l1);
    const int g = c / gc, cl = c % gc;
    __nv_bfloat16* p = xg + ((int64_t)g * rows_per_group + (int64_t)s * P + front + t) * gc + cl;
    *reinterpret_cast<uint2*>(p) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(p + plane_stride) = make_uint2(l0, l1);
  }
}

// pz = bias + sum of the split-K partials of the 16 group GEMMs; hsum = h + gelu(pz); x = dropout(LN(hsum)).  One warp per frame.
// ws layout: [g][split][M = S * P][gc].  D = 128 * J4.
template <int J4>
__global__ void __launch_bounds__(256) hub_pos_finish_fwd_kernel(const float* __restrict__ ws, int split, int64_t M, int P, int gc, const float* __restrict__ bias,
                                                                 const float* __restrict__ h, int S, int Fr, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps, float* __restrict__ pz_out, float* __restrict__ hsum_out,
                                                                 float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ xf,
                                                                 __nv_bfloat16* __restrict__ xp, int64_t plane_stride, const DropParams dr) {
  constexpr int D = J4 * 128;
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= S * Fr) return;
  const int s = r / Fr, t = r % Fr;
  const int64_t mrow = (int64_t)s * P + t;
  float4 v[J4];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    const int c = j * 128 + lane * 4;
    const int g = c / gc, cl = c % gc;
    float4 a = *reinterpret_cast<const float4*>(bias + c);
    for (int p = 0; p < split; ++p) {
      const float4 w = *reinterpret_cast<const float4*>(ws + (((int64_t)g * split + p) * M + mrow) * gc + cl);
      a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w;
    }
    *reinterpret_cast<float4*>(pz_out + (int64_t)r * D + c) = a;
    const float4 hh = *reinterpret_cast<const float4*>(h + (int64_t)r * D + c);
    v[j] = make_float4(hh.x + gelu_f(a.x), hh.y + gelu_f(a.y), hh.z + gelu_f(a.z), hh.w + gelu_f(a.w));
    *reinterpret_cast<float4*>(hsum_out + (int64_t)r * D + c) = v[j];
    sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(sum) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (lane == 0) { mean_out[r] = mean; rstd_out[r] = rstd; }
  uint32_t key = 0, base = 0;
  if (dr.on) {
    key = drop_site_key(dr.seq_key[s], dr.site);
    base = ((uint32_t)dr.seq_row[s] * (uint32_t)Fr + (uint32_t)t) * (uint32_t)D;
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
    *reinterpret_cast<float4*>(xf + (int64_t)r * D + c) = y;
    uint32_t h0, l0, h1, l1;
    split2(y.x, y.y, h0, l0);
    split2(y.z, y.w, h1, l1);
    __nv_bfloat16* hp = xp + (int64_t)r * D + c;
    *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(hp + plane_stride) = make_uint2(l0, l1);
  }
}

// backward into the positional branch: dpz = dhsum * gelu'(pz) -> group-major padded planes dzg[g][s * P + front + t][c % gc]
__global__ void hub_pos_bwd_prep_kernel(const float* __restrict__ dhsum, const float* __restrict__ pz, int Sg, int Fr, int D, __nv_bfloat16* __restrict__ dzg,
                                        int64_t plane_stride, int64_t rows_per_group, int P, int front, int gc, float* __restrict__ dpz_out) {
  const int64_t total4 = (int64_t)Sg * Fr * D / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 4;
    const int64_t row = e / D;
    const int c = (int)(e % D), s = (int)(row / Fr), t = (int)(row % Fr);
    const float4 d = *reinterpret_cast<const float4*>(dhsum + e);
    const float4 z = *reinterpret_cast<const float4*>(pz + e);
    const float4 v = make_float4(d.x * gelu_grad_f(z.x), d.y * gelu_grad_f(z.y), d.z * gelu_grad_f(z.z), d.w * gelu_grad_f(z.w));
    *reinterpret_cast<float4*>(dpz_out + e) = v;
    uint32_t h0, l0, h1, l1;
    split2(v.x, v.y, h0, l0);
    split2(v.z, v.w, h1, l1);
    const int g = c / gc, cl = c % gc;
    __nv_bfloat16* p = dzg + ((int64_t)g * rows_per_group + (int64_t)s * P + front + t) * gc + cl;
    *reinterpret_cast<uint2*>(p) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(p + plane_stride) = make_uint2(l0, l1);
  }
}
// out[c] (+)= sum over rows of x[r, c]   (small row counts: pos-conv bias, projection bias)
__global__ void hub_colsum_kernel(const float* __restrict__ x, int rows, int cols, float* __restrict__ out, int accumulate) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), part = threadIdx.x >> 5;
  float a = 0.f;
  if (c < cols)
    for (int r = part; r < rows; r += 8) a += x[(int64_t)r * cols + c];
  red[part][threadIdx.x & 31] = a;
  __syncthreads();
  if (part == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) t += red[p][threadIdx.x & 31];
    out[c] = accumulate ? out[c] + t : t;
  }
}
// dh = dhsum + sum of the split-K partials of the 16 dgrad GEMMs; SpecAugment frames pass their gradient to masked_spec_embed instead
// (partial column sums in mse_part[blockIdx.x][D]); the feature-projection dropout mask; -> dh fp32 + planes.  One warp per frame.
template <int J4>
__global__ void __launch_bounds__(256) hub_pos_finish_bwd_kernel(const float* __restrict__ ws, int split, int64_t M, int P, int gc, const float* __restrict__ dhsum,
                                                                 int Sg, int Fr, const uint8_t* __restrict__ mask_time, const DropParams dr, float* __restrict__ dh,
                                                                 __nv_bfloat16* __restrict__ dhp, int64_t plane_stride, float* __restrict__ mse_part) {
  constexpr int D = J4 * 128;
  __shared__ float msum[8][D];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x * 8 + warp;
  const bool valid = r < Sg * Fr;
  const int s = valid ? r / Fr : 0, t = valid ? r % Fr : 0;
  const int64_t mrow = (int64_t)s * P + t;
  const bool masked = valid && mask_time && mask_time[r];
  uint32_t key = 0, base = 0;
  if (dr.on && valid) {
    key = drop_site_key(dr.seq_key[s], dr.site);
    base = ((uint32_t)dr.seq_row[s] * (uint32_t)Fr + (uint32_t)t) * (uint32_t)D;
  }
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    const int c = j * 128 + lane * 4;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const int g = c / gc, cl = c % gc;
      a = *reinterpret_cast<const float4*>(dhsum + (int64_t)r * D + c);
      for (int p = 0; p < split; ++p) {
        const float4 w = *reinterpret_cast<const float4*>(ws + (((int64_t)g * split + p) * M + mrow) * gc + cl);
        a.x += w.x; a.y += w.y; a.z += w.z; a.w += w.w;
      }
    }
    *reinterpret_cast<float4*>(&msum[warp][c]) = masked ? a : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!valid) continue;
    if (masked) a = make_float4(0.f, 0.f, 0.f, 0.f);
    else if (dr.on) {
      a.x = drop_kept(key, base + c, dr.thr24) ? a.x * dr.inv_keep : 0.f;
      a.y = drop_kept(key, base + c + 1, dr.thr24) ? a.y * dr.inv_keep : 0.f;
      a.z = drop_kept(key, base + c + 2, dr.thr24) ? a.z * dr.inv_keep : 0.f;
      a.w = drop_kept(key, base + c + 3, dr.thr24) ? a.w * dr.inv_keep : 0.f;
    }
    *reinterpret_cast<float4*>(dh + (int64_t)r * D + c) = a;
    uint32_t h0, l0, h1, l1;
    split2(a.x, a.y, h0, l0);
    split2(a.z, a.w, h1, l1);
    __nv_bfloat16* hp = dhp + (int64_t)r * D + c;
    *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(hp + plane_stride) = make_uint2(l0, l1);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += 256) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += msum[w][c];
    mse_part[(int64_t)blockIdx.x * D + c] = a;
  }
}

// weight norm of the positional conv (torch._weight_norm(v, g, dim = 2)): w[o, i, j] = g[j] v[o, i, j] / ||v[:, :, j]||
__global__ void __launch_bounds__(256) hub_pos_norm_kernel(const float* __restrict__ v, int n_oi, int K, float* __restrict__ norm) {
  __shared__ float red[8];
  const int j = blockIdx.x;
  float a = 0.f;
  for (int i = threadIdx.x; i < n_oi; i += 256) { const float x = v[(int64_t)i * K + j]; a = fmaf(x, x, a); }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    norm[j] = sqrtf(t);
  }
}
// forward operand  wf[g][o % gc][j * gc + i]      = w[o, i, j]          (K index = (tap, in-channel): the overlapping row view of xg)
// dgrad operand    wd[g][i][j' * gc + o % gc]     = w[o, i, K - 1 - j'] (K index = (flipped tap, out-channel): the view of dzg)
__global__ void hub_pos_relayout_kernel(const float* __restrict__ v, const float* __restrict__ gain, const float* __restrict__ norm, int D, int gc, int K,
                                        __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd, int64_t plane_stride) {
  const int64_t n = (int64_t)D * gc * K;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx / ((int64_t)gc * K)), rem = (int)(idx % ((int64_t)gc * K));
    const int i = rem / K, j = rem % K;
    const float w = gain[j] * v[idx] / norm[j];
    __nv_bfloat16 hi, lo;
    split_bf16(w, hi, lo);
    const int g = o / gc, ol = o % gc;
    const int64_t f = ((int64_t)g * gc + ol) * ((int64_t)K * gc) + (int64_t)j * gc + i;
    const int64_t d = ((int64_t)g * gc + i) * ((int64_t)K * gc) + (int64_t)(K - 1 - j) * gc + ol;
    wf[f] = hi; wf[f + plane_stride] = lo;
    wd[d] = hi; wd[d + plane_stride] = lo;
  }
}
// weight-norm backward.  The wgrad GEMMs leave split-K partials of dw[g][o % gc][(j, i)].  Per tap j:
//   dot[j] = sum_{o,i} dw v;  dg[j] = dot / norm;  dv = g / norm * (dw - v * dot / norm^2)
__global__ void __launch_bounds__(256) hub_pos_wn_dot_kernel(const float* __restrict__ ws, int split, const float* __restrict__ v, int D, int gc, int K,
                                                             float* __restrict__ dw_out, float* __restrict__ dot) {
  __shared__ float red[8];
  const int j = blockIdx.x;
  const int n_oi = D * gc;
  const int64_t per_group = (int64_t)gc * K * gc;   // elements of one group's [gc, K * gc] matrix
  float a = 0.f;
  for (int oi = threadIdx.x; oi < n_oi; oi += 256) {
    const int o = oi / gc, i = oi % gc, g = o / gc, ol = o % gc;
    float dw = 0.f;
    for (int p = 0; p < split; ++p) dw += ws[((int64_t)g * split + p) * per_group + (int64_t)ol * K * gc + (int64_t)j * gc + i];
    dw_out[(int64_t)oi * K + j] = dw;
    a = fmaf(dw, v[(int64_t)oi * K + j], a);
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    dot[j] = t;
  }
}
__global__ void hub_pos_wn_apply_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ gain, const float* __restrict__ norm,
                                        const float* __restrict__ dot, int64_t n, int K, float* __restrict__ dv, float* __restrict__ dgain, int accumulate) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % K);
    const float nj = norm[j];
    const float val = gain[j] / nj * (dw[idx] - v[idx] * dot[j] / (nj * nj));
    dv[idx] = accumulate ? dv[idx] + val : val;
    if (idx < K) {
      const float dgv = dot[idx] / norm[idx];
      dgain[idx] = accumulate ? dgain[idx] + dgv : dgv;
    }
  }
}

// conv weights [co, ci, k] -> forward operand wf[co][j * 512 + ci]; for k = 3 also the even-frame dgrad operand
// we[ci][0 .. 511] = w[co, ci, 2] (previous output frame), we[ci][512 + co] = w[co, ci, 0] (this output frame)
__global__ void hub_conv_relayout_kernel(const float* __restrict__ w, int k, __nv_bfloat16* __restrict__ wf, int64_t wf_ps, __nv_bfloat16* __restrict__ we,
                                         int64_t we_ps) {
  const int64_t n = (int64_t)HC * HC * k;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx / (HC * k)), rem = (int)(idx % (HC * k));
    const int ci = rem / k, j = rem % k;
    __nv_bfloat16 hi, lo;
    split_bf16(w[idx], hi, lo);
    const int64_t f = (int64_t)co * (k * HC) + (int64_t)j * HC + ci;
    wf[f] = hi; wf[f + wf_ps] = lo;
    if (we && j != 1) {
      const int64_t e = (int64_t)ci * (2 * HC) + (j == 2 ? 0 : HC) + co;
      we[e] = hi; we[e + we_ps] = lo;
    }
  }
}
struct HubQkvBiasPtrs { const float* p[3 * 48]; };
__global__ void hub_pack_qkv_bias_kernel(const __grid_constant__ HubQkvBiasPtrs P, int layers, int D, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= layers * 3 * D) return;
  out[i] = P.p[i / D][i % D];
}
// d masked_spec_embed (+)= column sums of the per-CTA partials of hub_pos_finish_bwd_kernel
__global__ void hub_mse_grad_kernel(const float* __restrict__ part, int nparts, int D, float* __restrict__ out, int accumulate, int any_mask) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float a = 0.f;
  if (any_mask)
    for (int p = 0; p < nparts; ++p) a += part[(int64_t)p * D + c];
  out[c] = accumulate ? out[c] + a : a;
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
struct HDims {
  int S, Sg, Ts, NC, D, H, F, NL, C, PK, PG, GC;
  int k[SRW_HUBERT_MAX_CONV], st[SRW_HUBERT_MAX_CONV], n[SRW_HUBERT_MAX_CONV], tp[SRW_HUBERT_MAX_CONV];
  int Fr, P, chunks0;
  int64_t T, Tg;
};
constexpr int SLACK = 8;   // zero rows behind the last clip of every conv activation (the padding frame of the last clip reads past it)

static int make_hdims(const srw_hubert_config* c, int batch, int samples, int grad_batch, HDims& d) {
  SRW_REQUIRE(c && batch > 0 && grad_batch >= 0 && grad_batch <= batch, "srw_hubert: bad batch (%d, grad %d)", batch, grad_batch);
  SRW_REQUIRE(c->hidden == 384 || c->hidden == 768 || c->hidden == 1024, "srw_hubert: hidden must be 384, 768 or 1024 (got %d)", c->hidden);
  SRW_REQUIRE(c->heads * 64 == c->hidden, "srw_hubert: heads * 64 must equal hidden");
  SRW_REQUIRE(c->intermediate % 64 == 0 && c->layers > 0 && c->layers <= 48 && c->num_classes > 0, "srw_hubert: bad intermediate / layers / num_classes");
  SRW_REQUIRE(c->conv_dim == HC && c->num_conv >= 2 && c->num_conv <= SRW_HUBERT_MAX_CONV, "srw_hubert: conv_dim must be 512, 2 <= num_conv <= 8");
  SRW_REQUIRE(c->conv_kernel[0] == C0K && c->conv_stride[0] == C0S, "srw_hubert: the first conv must be k = 10, stride 5");
  for (int l = 1; l < c->num_conv; ++l)
    SRW_REQUIRE(c->conv_stride[l] == 2 && (c->conv_kernel[l] == 2 || c->conv_kernel[l] == 3), "srw_hubert: conv layer %d must be stride 2 with k = 2 or 3", l);
  SRW_REQUIRE(c->pos_groups > 0 && c->hidden % c->pos_groups == 0 && (c->hidden / c->pos_groups) % 8 == 0 && c->pos_kernel % 2 == 0 && c->pos_kernel >= 2,
              "srw_hubert: positional conv needs an even kernel and hidden / groups a multiple of 8");
  d.S = batch; d.Sg = grad_batch; d.Ts = samples; d.NC = c->num_conv; d.D = c->hidden; d.H = c->heads; d.F = c->intermediate; d.NL = c->layers;
  d.C = c->num_classes; d.PK = c->pos_kernel; d.PG = c->pos_groups; d.GC = c->hidden / c->pos_groups;
  int n = samples;
  for (int l = 0; l < d.NC; ++l) {
    d.k[l] = c->conv_kernel[l]; d.st[l] = c->conv_stride[l];
    SRW_REQUIRE(n >= d.k[l], "srw_hubert: clips of %d samples are too short for the conv stem", samples);
    n = (n - d.k[l]) / d.st[l] + 1;
    d.n[l] = n;
  }
  d.Fr = d.n[d.NC - 1];
  SRW_REQUIRE(d.Fr >= 16 && d.Fr <= 512, "srw_hubert: %d frames per clip; the attention kernels take 16 .. 512 (about 0.33 s .. 10.2 s of audio)", d.Fr);
  // padded frames per clip: tp[l - 1] = stride[l] * tp[l], every tp[l] > n[l] (at least one padding frame keeps the even/odd dgrad views simple)
  int tpl = 0;
  for (int l = 0; l < d.NC; ++l) {
    int64_t prod = 1;
    for (int j = l + 1; j < d.NC; ++j) prod *= d.st[j];
    tpl = std::max<int>(tpl, (int)cdiv64(d.n[l] + 1, prod));
  }
  d.tp[d.NC - 1] = tpl;
  for (int l = d.NC - 1; l > 0; --l) d.tp[l - 1] = d.st[l] * d.tp[l];
  d.P = (d.Fr + d.PK - 1 + 7) / 8 * 8;
  d.chunks0 = cdiv(d.tp[0], C0_FRAMES);
  d.T = (int64_t)batch * d.Fr; d.Tg = (int64_t)grad_batch * d.Fr;
  SRW_REQUIRE((int64_t)batch * d.H * d.Fr * (int64_t)d.Fr < (int64_t)1 << 32 && (int64_t)d.Fr * d.F < (int64_t)1 << 26,
              "srw_hubert: batch too large for the 32-bit dropout counters");
  return SRW_OK;
}

static EncDims hub_enc_dims(const HDims& d) {
  EncDims e;
  e.S = d.S; e.Sg = d.Sg; e.Lq = d.Fr; e.D = d.D; e.H = d.H; e.F = d.F; e.T = d.T; e.Tg = d.Tg; e.ld_bias = 0;
  return e;
}

struct HLayout {
  int64_t part0, mean0, rstd0;
  int64_t y[SRW_HUBERT_MAX_CONV], y_rows[SRW_HUBERT_MAX_CONV];   // planes [y_rows, 512] (y_rows = S * tp + SLACK)
  int64_t z[SRW_HUBERT_MAX_CONV];                                 // fp32 pre-activations of layers >= 1, [S * tp, 512]
  int64_t g_last, lnp, mean_fp, rstd_fp, h0, xg, xg_rows, posws, pos_split, pz, hsum, mean_e, rstd_e;
  std::vector<int64_t> xf, xp;
  std::vector<EncLayerBufs> lay;
  int64_t feat, z1, a1;
  // backward
  int64_t dx, dfeat, dz1, dhsum, dpz, dzg, dzg_rows, posw_ws, posw_split, dw_pos, dot_pos, dh0, dh0p, mse_part, mse_nparts, dln, dg_last;
  int64_t dzbuf[2], dz_rows, dy0, convw_ws, projw_ws, bpart0, ab0, wpart0;
  EncBwdScratch bw;
  int64_t total;
};

static int pos_fwd_split(const HDims& d, int64_t M) {
  int split = 1;
  splitk_for((int)std::min<int64_t>(M, 1 << 30), d.GC, (int64_t)d.PK * d.GC, &split);
  const int min_split = cdiv(d.PK * d.GC, 4096);
  return std::max(split, min_split);
}

static HLayout make_hlayout(const HDims& d) {
  HLayout L;
  Carver c;
  const int64_t S = d.S, D = d.D, T = d.T;
  L.part0 = c.take(S * d.chunks0 * HC * 8); L.mean0 = c.take(S * HC * 4); L.rstd0 = c.take(S * HC * 4);
  for (int l = 0; l < d.NC; ++l) {
    L.y_rows[l] = S * d.tp[l] + SLACK;
    L.y[l] = c.take(L.y_rows[l] * HC * 4);
    L.z[l] = l > 0 ? c.take(S * d.tp[l] * HC * 4) : 0;
  }
  L.g_last = c.take(T * HC * 4); L.lnp = c.take(T * HC * 4); L.mean_fp = c.take(T * 4); L.rstd_fp = c.take(T * 4);
  L.h0 = c.take(T * D * 4);
  L.xg_rows = S * d.P + d.PK;
  L.xg = c.take((int64_t)d.PG * L.xg_rows * d.GC * 4);
  L.pos_split = pos_fwd_split(d, S * d.P);
  L.posws = c.take((int64_t)d.PG * L.pos_split * S * d.P * d.GC * 4);
  L.pz = c.take(T * D * 4); L.hsum = c.take(T * D * 4); L.mean_e = c.take(T * 4); L.rstd_e = c.take(T * 4);
  const EncDims ed = hub_enc_dims(d);
  L.xf.resize(d.NL + 1); L.xp.resize(d.NL + 1); L.lay.resize(d.NL);
  for (int l = 0; l <= d.NL; ++l) { L.xf[l] = c.take(T * D * 4); L.xp[l] = c.take(T * D * 4); }
  for (int l = 0; l < d.NL; ++l) enc_take_layer(c, ed, L.lay[l]);
  L.feat = c.take(S * D * 4); L.z1 = c.take(S * D * 4); L.a1 = c.take(S * D * 4);
  // ---- backward ----
  const int64_t Sg = std::max(d.Sg, 1), Tg = std::max<int64_t>(d.Tg, 1);
  L.dx = c.take(Tg * D * 4); L.dfeat = c.take(Sg * D * 4); L.dz1 = c.take(Sg * D * 4);
  enc_take_bwd_scratch(c, ed, L.bw);
  L.dhsum = c.take(Tg * D * 4); L.dpz = c.take(Tg * D * 4);
  L.dzg_rows = Sg * d.P + d.PK;
  L.dzg = c.take((int64_t)d.PG * L.dzg_rows * d.GC * 4);
  {
    int split = 1;
    splitk_for(d.GC, d.PK * d.GC, Sg * d.P, &split);
    L.posw_split = split;
    L.posw_ws = c.take((int64_t)d.PG * split * d.GC * d.PK * d.GC * 4);
  }
  L.dw_pos = c.take((int64_t)D * d.GC * d.PK * 4); L.dot_pos = c.take((int64_t)d.PK * 4);
  L.dh0 = c.take(Tg * D * 4); L.dh0p = c.take(Tg * D * 4);
  L.mse_nparts = cdiv64(Tg, 8);
  L.mse_part = c.take(L.mse_nparts * D * 4);
  L.dln = c.take(Tg * HC * 4); L.dg_last = c.take(Tg * HC * 4);
  L.dz_rows = Sg * d.tp[1] + SLACK + 1;                      // + 1: a zero row in front (frame -1 of the first clip in the even dgrad view)
  L.dzbuf[0] = c.take(L.dz_rows * HC * 4); L.dzbuf[1] = c.take(L.dz_rows * HC * 4);
  L.dy0 = c.take(Sg * d.tp[0] * HC * 4);
  int64_t wsmax = 0;
  for (int l = 1; l < d.NC; ++l) wsmax = std::max(wsmax, splitk_for(HC, d.k[l] * HC, Sg * d.tp[l], nullptr));
  L.convw_ws = c.take(wsmax * 4);
  L.projw_ws = c.take(splitk_for(d.D, HC, Tg, nullptr) * 4);
  L.bpart0 = c.take(Sg * d.chunks0 * HC * 8); L.ab0 = c.take(Sg * HC * 8); L.wpart0 = c.take(Sg * d.chunks0 * HC * C0K * 4);
  L.total = c.off;
  return L;
}

// parameter indices (ClassificationHubert.state_dict() order, include/srw.h)
enum { HP_MSE = 0, HP_C0W, HP_GNW, HP_GNB, HP_CONV1 /* .. HP_CONV1 + NC - 2 */ };
static inline int hp_after_conv(const HDims& d, int which) { return HP_CONV1 + (d.NC - 1) + which; }   // 0 fp.ln.w, 1 fp.ln.b, 2 proj.w, 3 proj.b, 4 pos.bias, 5 pos.g, 6 pos.v, 7 enc.ln.w, 8 enc.ln.b
enum { FP_LNW = 0, FP_LNB, FP_W, FP_B, POS_B, POS_G, POS_V, ENC_LNW, ENC_LNB, HP_FRONT_TAIL };
enum { HY_KW = 0, HY_KB, HY_VW, HY_VB, HY_QW, HY_QB, HY_OW, HY_OB, HY_LN1W, HY_LN1B, HY_F1W, HY_F1B, HY_F2W, HY_F2B, HY_LN2W, HY_LN2B };
static inline int hp_layer(const HDims& d, int l, int which) { return hp_after_conv(d, HP_FRONT_TAIL) + 16 * l + which; }
static inline int hp_cls(const HDims& d, int which) { return hp_after_conv(d, HP_FRONT_TAIL) + 16 * d.NL + which; }   // 0 cls0.w, 1 cls0.b, 2 cls2.w, 3 cls2.b
static int hub_num_params(const srw_hubert_config* c) { return HP_CONV1 + (c->num_conv - 1) + HP_FRONT_TAIL + 16 * c->layers + 4; }

struct HWOff {
  int64_t convf[SRW_HUBERT_MAX_CONV], conve[SRW_HUBERT_MAX_CONV];
  int64_t proj, posf, posd, pos_norm, qkv_bias;
  std::vector<int64_t> qkv, o, f1, f2;
  int64_t total;
};
static HWOff hub_weight_layout(const HDims& d) {
  HWOff w;
  Carver c;
  for (int l = 1; l < d.NC; ++l) {
    w.convf[l] = c.take((int64_t)HC * d.k[l] * HC * 4);
    w.conve[l] = d.k[l] == 3 ? c.take((int64_t)HC * 2 * HC * 4) : -1;
  }
  w.proj = c.take((int64_t)d.D * HC * 4);
  w.posf = c.take((int64_t)d.D * d.GC * d.PK * 4);
  w.posd = c.take((int64_t)d.D * d.GC * d.PK * 4);
  w.pos_norm = c.take((int64_t)d.PK * 4);
  w.qkv_bias = c.take((int64_t)d.NL * 3 * d.D * 4);
  w.qkv.resize(d.NL); w.o.resize(d.NL); w.f1.resize(d.NL); w.f2.resize(d.NL);
  for (int l = 0; l < d.NL; ++l) {
    w.qkv[l] = c.take((int64_t)3 * d.D * d.D * 4);
    w.o[l] = c.take((int64_t)d.D * d.D * 4);
    w.f1[l] = c.take((int64_t)d.F * d.D * 4);
    w.f2[l] = c.take((int64_t)d.D * d.F * 4);
  }
  w.total = c.off;
  return w;
}

static EncLayerW hub_layer_weights(const float* const* P, const uint8_t* wp, const HWOff& w, const HDims& d, int l) {
  EncLayerW lw;
  lw.qkv = wp + w.qkv[l]; lw.o = wp + w.o[l]; lw.f1 = wp + w.f1[l]; lw.f2 = wp + w.f2[l];
  lw.qkv_bias = reinterpret_cast<const float*>(wp + w.qkv_bias) + (int64_t)l * 3 * d.D;
  lw.ob = P[hp_layer(d, l, HY_OB)]; lw.ln1w = P[hp_layer(d, l, HY_LN1W)]; lw.ln1b = P[hp_layer(d, l, HY_LN1B)];
  lw.f1b = P[hp_layer(d, l, HY_F1B)]; lw.f2b = P[hp_layer(d, l, HY_F2B)]; lw.ln2w = P[hp_layer(d, l, HY_LN2W)]; lw.ln2b = P[hp_layer(d, l, HY_LN2B)];
  return lw;
}
static EncDropSites hub_layer_sites(const uint32_t* key, const int32_t* row, int l, const srw_hubert_config& cf) {
  EncDropSites ds;
  ds.key = key; ds.row = row; ds.attn = 2 + 4 * l; ds.o = 3 + 4 * l; ds.act = 4 + 4 * l; ds.f2 = 5 + 4 * l;
  ds.p_attn = cf.p_attn; ds.p_hidden = cf.p_hidden; ds.p_act = cf.p_act;
  return ds;
}
static DropParams hub_site_drop(const uint32_t* key, const int32_t* row, uint32_t site, double p) {
  srw_dropout dd;
  dd.seq_key = key; dd.seq_row = row; dd.site = site; dd.p = p;
  return make_drop(dd);
}

// LayerDrop: the runs of consecutive clips (within [0, limit)) that execute layer l
struct Run { int seq0, nseq; };
static int layer_runs(int num_segments, const int32_t* segment_start, const uint8_t* layer_skip, int layers, int l, int batch, int limit, std::vector<Run>& out) {
  out.clear();
  if (num_segments <= 0 || !segment_start || !layer_skip) {
    if (limit > 0) out.push_back({0, limit});
    return SRW_OK;
  }
  SRW_REQUIRE(segment_start[0] == 0 && segment_start[num_segments] == batch, "srw_hubert: segment_start must run from 0 to batch");
  for (int g = 0; g < num_segments; ++g) {
    const int a = segment_start[g], b = std::min(segment_start[g + 1], limit);
    SRW_REQUIRE(segment_start[g + 1] >= a, "srw_hubert: segment_start must be non-decreasing");
    if (b <= a || layer_skip[(int64_t)g * layers + l]) continue;
    if (!out.empty() && out.back().seq0 + out.back().nseq == a) out.back().nseq += b - a;
    else out.push_back({a, b - a});
  }
  return SRW_OK;
}

}  // namespace srw

using namespace srw;

extern "C" int srw_hubert_frames(const srw_hubert_config* c, int samples) {
  HDims d;
  if (make_hdims(c, 1, samples, 0, d)) return -1;
  return d.Fr;
}

extern "C" int64_t srw_hubert_weight_planes_bytes(const srw_hubert_config* c) {
  HDims d;
  if (!c) return -1;
  int samples = 400;
  for (int tries = 0; tries < 12 && make_hdims(c, 1, samples, 0, d); ++tries) samples *= 2;   // any valid clip length: the cache does not depend on it
  if (make_hdims(c, 1, samples, 0, d)) return -1;
  return hub_weight_layout(d).total;
}

extern "C" int64_t srw_hubert_workspace_bytes(const srw_hubert_config* c, int batch, int samples, int grad_batch) {
  HDims d;
  if (make_hdims(c, batch, samples, grad_batch, d)) return -1;
  return make_hlayout(d).total;
}

// what: 1 = front end (re-laid-out conv weights, projection, weight-normalised positional taps, packed q|k|v biases), 2 = the encoder
// matrices (q|k|v, out_proj, FFN), 3 = both
static int hubert_prepare(const srw_hubert_config* c, const float* const* P, void* weight_planes, int what, cudaStream_t s) {
  SRW_REQUIRE(c && P && weight_planes, "srw_hubert_prepare_weights: null pointer");
  HDims d;
  int samples = 400;
  for (int tries = 0; tries < 12 && make_hdims(c, 1, samples, 0, d); ++tries) samples *= 2;
  SRW_TRY(make_hdims(c, 1, samples, 0, d));
  const HWOff w = hub_weight_layout(d);
  uint8_t* base = reinterpret_cast<uint8_t*>(weight_planes);
  auto BF = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(base + off); };
  const int D = d.D, F = d.F;
  if (what & 1) {
    for (int l = 1; l < d.NC; ++l) {
      const int k = d.k[l];
      hub_conv_relayout_kernel<<<148 * 4, 256, 0, s>>>(P[HP_CONV1 + l - 1], k, BF(w.convf[l]), (int64_t)HC * k * HC, k == 3 ? BF(w.conve[l]) : nullptr,
                                                       (int64_t)HC * 2 * HC);
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    SRW_TRY(split_to(P[hp_after_conv(d, FP_W)], HC, d.D, HC, base + w.proj, HC, nullptr, 1, s));
    float* norm = reinterpret_cast<float*>(base + w.pos_norm);
    hub_pos_norm_kernel<<<d.PK, 256, 0, s>>>(P[hp_after_conv(d, POS_V)], d.D * d.GC, d.PK, norm);
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_pos_relayout_kernel<<<148 * 4, 256, 0, s>>>(P[hp_after_conv(d, POS_V)], P[hp_after_conv(d, POS_G)], norm, d.D, d.GC, d.PK, BF(w.posf), BF(w.posd),
                                                    (int64_t)d.D * d.GC * d.PK);
    g_launches++;
    SRW_LAUNCH_CHECK();
    HubQkvBiasPtrs bp = {};
    for (int l = 0; l < d.NL; ++l) {
      bp.p[3 * l] = P[hp_layer(d, l, HY_QB)]; bp.p[3 * l + 1] = P[hp_layer(d, l, HY_KB)]; bp.p[3 * l + 2] = P[hp_layer(d, l, HY_VB)];
    }
    const int nb = d.NL * 3 * D;
    hub_pack_qkv_bias_kernel<<<cdiv(nb, 256), 256, 0, s>>>(bp, d.NL, D, reinterpret_cast<float*>(base + w.qkv_bias));
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  if (what & 2) {
    for (int l = 0; l < d.NL; ++l) {
      const int order[3] = {HY_QW, HY_KW, HY_VW};   // packed rows: q | k | v (the attention kernels' column order)
      for (int i = 0; i < 3; ++i) {
        srw_split_args a = {};
        a.x = P[hp_layer(d, l, order[i])]; a.ldx = D; a.rows = D; a.cols = D; a.rows_per_scale = 1;
        a.planes = base + w.qkv[l] + (int64_t)i * D * D * 2; a.ldp = D; a.plane_stride = (int64_t)3 * D * D;
        SRW_TRY(srw_split_planes(&a, s));
      }
      SRW_TRY(split_to(P[hp_layer(d, l, HY_OW)], D, D, D, base + w.o[l], D, nullptr, 1, s));
      SRW_TRY(split_to(P[hp_layer(d, l, HY_F1W)], D, F, D, base + w.f1[l], D, nullptr, 1, s));
      SRW_TRY(split_to(P[hp_layer(d, l, HY_F2W)], F, D, F, base + w.f2[l], F, nullptr, 1, s));
    }
  }
  return SRW_OK;
}

extern "C" int srw_hubert_prepare_weights(const srw_hubert_config* c, const float* const* P, void* weight_planes, void* stream_) {
  return hubert_prepare(c, P, weight_planes, 3, reinterpret_cast<cudaStream_t>(stream_));
}
extern "C" int srw_hubert_prepare_front(const srw_hubert_config* c, const float* const* P, void* weight_planes, void* stream_) {
  return hubert_prepare(c, P, weight_planes, 1, reinterpret_cast<cudaStream_t>(stream_));
}

extern "C" int srw_hubert_weight_plane_slot(const srw_hubert_config* c, int param_index, int64_t* byte_offset, int* cols, int* ldp, int64_t* plane_stride) {
  SRW_REQUIRE(c && byte_offset && cols && ldp && plane_stride, "srw_hubert_weight_plane_slot: null pointer");
  HDims d;
  int samples = 400;
  for (int tries = 0; tries < 12 && make_hdims(c, 1, samples, 0, d); ++tries) samples *= 2;
  SRW_TRY(make_hdims(c, 1, samples, 0, d));
  const HWOff w = hub_weight_layout(d);
  const int rel = param_index - hp_layer(d, 0, 0);
  if (rel >= 0 && rel < 16 * d.NL) {
    const int l = rel / 16, which = rel % 16;
    const int64_t D = d.D, F = d.F;
    int64_t off = -1, cc = 0, ps = 0;
    if (which == HY_QW) { off = w.qkv[l]; cc = D; ps = 3 * D * D; }
    else if (which == HY_KW) { off = w.qkv[l] + D * D * 2; cc = D; ps = 3 * D * D; }        // rows [D, 2D) of the packed [3D, D] hi plane (bf16)
    else if (which == HY_VW) { off = w.qkv[l] + 2 * D * D * 2; cc = D; ps = 3 * D * D; }
    else if (which == HY_OW) { off = w.o[l]; cc = D; ps = D * D; }
    else if (which == HY_F1W) { off = w.f1[l]; cc = D; ps = F * D; }
    else if (which == HY_F2W) { off = w.f2[l]; cc = F; ps = D * F; }
    if (off >= 0) {
      *byte_offset = off; *cols = (int)cc; *ldp = (int)cc; *plane_stride = ps;
      return SRW_OK;
    }
  }
  set_last_error("srw_hubert_weight_plane_slot: parameter %d has no per-parameter planes", param_index);
  return SRW_ERR_ARG;
}

static int hubert_forward_body(const srw_hubert_fwd_args* a, cudaStream_t s) {
  SRW_REQUIRE(a && a->cfg && a->params && a->weight_planes && a->wav && a->logits && a->feat && a->workspace, "srw_hubert_forward: null pointer");
  SRW_REQUIRE((a->drop_seq_key == nullptr) == (a->drop_seq_row == nullptr), "srw_hubert_forward: drop_seq_key and drop_seq_row go together");
  HDims d;
  SRW_TRY(make_hdims(a->cfg, a->batch, a->samples, a->grad_batch, d));
  SRW_REQUIRE(a->ld_wav >= a->samples, "srw_hubert_forward: ld_wav < samples");
  const HLayout L = make_hlayout(d);
  SRW_REQUIRE(a->workspace_bytes >= L.total, "srw_hubert_forward: workspace too small (%lld < %lld)", (long long)a->workspace_bytes, (long long)L.total);
  SRW_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 1023) == 0, "srw_hubert_forward: workspace must be 1024-byte aligned");
  const HWOff w = hub_weight_layout(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(a->weight_planes);
  const float* const* P = a->params;
  auto F32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  auto BF = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  const int S = d.S, D = d.D, impl = a->gemm_impl, Fr = d.Fr, T = (int)d.T;
  const float eps = a->cfg->ln_eps;
  const uint32_t* dk = a->drop_seq_key;
  const int32_t* drw = a->drop_seq_row;
  const srw_hubert_config& cf = *a->cfg;

  // ---- conv 0 + GroupNorm + GELU ----
  {
    const dim3 grid(d.chunks0, S);
    hub_conv0_stats_kernel<<<dim3(cdiv(d.n[0], C0_FRAMES), S), 256, 0, s>>>(a->wav, a->ld_wav, d.Ts, d.n[0], P[HP_C0W], reinterpret_cast<float2*>(ws + L.part0));
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_conv0_stats_finish_kernel<<<cdiv(S * HC, 256), 256, 0, s>>>(reinterpret_cast<const float2*>(ws + L.part0), S, cdiv(d.n[0], C0_FRAMES), d.n[0], eps,
                                                                   F32(L.mean0), F32(L.rstd0));
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_conv0_apply_kernel<<<grid, 256, 0, s>>>(a->wav, a->ld_wav, d.Ts, d.n[0], d.tp[0], P[HP_C0W], F32(L.mean0), F32(L.rstd0), P[HP_GNW], P[HP_GNB], BF(L.y[0]),
                                                L.y_rows[0] * HC);
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_zero_pad_rows_kernel<<<8, 256, 0, s>>>(BF(L.y[0]), L.y_rows[0] * HC, S, d.tp[0], d.tp[0], SLACK, HC);   // only the slack rows: the kernel above wrote the padding frames
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // ---- conv layers 1 .. NC-1: one GEMM each over the overlapping row view ----
  for (int l = 1; l < d.NC; ++l) {
    const int k = d.k[l], M = S * d.tp[l];
    Gemm g(M, HC, k * HC, impl);
    g.g.a = ws + L.y[l - 1]; g.g.lda = (int64_t)d.st[l] * HC; g.g.a_plane_stride = L.y_rows[l - 1] * HC; g.g.a_mn_major = 0;
    g.g.b = wp + w.convf[l]; g.g.ldb = (int64_t)k * HC; g.g.b_plane_stride = (int64_t)HC * k * HC; g.g.b_mn_major = 0;
    g.g.epilogue = SRW_EPI_GELU; g.g.out_f32 = F32(L.z[l]); g.g.ldo = HC; g.g.out_planes = ws + L.y[l]; g.g.ldp = HC; g.g.out_plane_stride = L.y_rows[l] * HC;
    SRW_TRY(g.run(s));
    hub_zero_pad_rows_kernel<<<cdiv(S * (d.tp[l] - d.n[l]) + SLACK, 4), 256, 0, s>>>(BF(L.y[l]), L.y_rows[l] * HC, S, d.tp[l], d.n[l], SLACK, HC);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // ---- feature projection ----
  {
    const int last = d.NC - 1;
    hub_gelu_ln_fwd_kernel<<<cdiv(T, 8), 256, 0, s>>>(F32(L.z[last]), S, d.tp[last], Fr, P[hp_after_conv(d, FP_LNW)], P[hp_after_conv(d, FP_LNB)], eps,
                                                      F32(L.g_last), F32(L.mean_fp), F32(L.rstd_fp), BF(L.lnp), (int64_t)T * HC);
    g_launches++;
    SRW_LAUNCH_CHECK();
    Gemm g(T, D, HC, impl);
    g.A(ws + L.lnp, HC, T, 0).Bm(wp + w.proj, HC, D, 0);
    g.g.epilogue = SRW_EPI_F32; g.g.bias = P[hp_after_conv(d, FP_B)]; g.g.out_f32 = F32(L.h0); g.g.ldo = D;
    SRW_TRY(g.run(s));
    SRW_CUDA(cudaMemsetAsync(ws + L.xg, 0, (size_t)d.PG * L.xg_rows * d.GC * 4, s));
    const DropParams dr = hub_site_drop(dk, drw, 0, cf.p_feat_proj);
    hub_featproj_post_kernel<<<148 * 4, 256, 0, s>>>(F32(L.h0), S, Fr, D, dr, a->mask_time, P[HP_MSE], BF(L.xg), (int64_t)d.PG * L.xg_rows * d.GC, L.xg_rows, d.P,
                                                     d.PK / 2, d.GC);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // ---- positional conv: one split-K GEMM per group, then + residual, LayerNorm, dropout ----
  {
    const int64_t M = (int64_t)S * d.P;
    const int K = d.PK * d.GC, split = (int)L.pos_split;
    for (int g_ = 0; g_ < d.PG; ++g_) {
      Gemm g((int)M, d.GC, K, impl);
      g.g.a = ws + L.xg + (int64_t)g_ * L.xg_rows * d.GC * 2; g.g.lda = d.GC; g.g.a_plane_stride = (int64_t)d.PG * L.xg_rows * d.GC; g.g.a_mn_major = 0;
      g.g.b = wp + w.posf + (int64_t)g_ * d.GC * K * 2; g.g.ldb = K; g.g.b_plane_stride = (int64_t)d.D * d.GC * d.PK; g.g.b_mn_major = 0;
      g.g.epilogue = SRW_EPI_SPLITK; g.g.split_k = split; g.g.workspace = F32(L.posws) + (int64_t)g_ * split * M * d.GC;
      SRW_TRY(g.run(s));
    }
    const DropParams dr = hub_site_drop(dk, drw, 1, cf.p_hidden);
#define SRW_POSF(J)                                                                                                                                       \
  hub_pos_finish_fwd_kernel<J><<<cdiv(T, 8), 256, 0, s>>>(F32(L.posws), split, M, d.P, d.GC, P[hp_after_conv(d, POS_B)], F32(L.h0), S, Fr,                 \
                                                          P[hp_after_conv(d, ENC_LNW)], P[hp_after_conv(d, ENC_LNB)], eps, F32(L.pz), F32(L.hsum), F32(L.mean_e), \
                                                          F32(L.rstd_e), F32(L.xf[0]), BF(L.xp[0]), (int64_t)T * D, dr)
    if (D == 384) SRW_POSF(3); else if (D == 768) SRW_POSF(6); else SRW_POSF(8);
#undef SRW_POSF
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // ---- encoder layers (LayerDrop: per call) ----
  const EncDims ed = hub_enc_dims(d);
  std::vector<Run> runs;
  for (int l = 0; l < d.NL; ++l) {
    const EncLayerW lw = hub_layer_weights(P, wp, w, d, l);
    const EncDropSites ds = hub_layer_sites(dk, drw, l, cf);
    SRW_TRY(layer_runs(a->num_segments, a->segment_start, a->layer_skip, d.NL, l, S, S, runs));
    int covered = 0;
    for (const Run& r : runs) {
      if (r.seq0 > covered) {   // skipped clips in front of this run: the layer is the identity for them
        const int64_t r0 = (int64_t)covered * Fr, nr = (int64_t)(r.seq0 - covered) * Fr;
        SRW_CUDA(cudaMemcpyAsync(F32(L.xf[l + 1]) + r0 * D, F32(L.xf[l]) + r0 * D, (size_t)nr * D * 4, cudaMemcpyDeviceToDevice, s));
        for (int pl = 0; pl < 2; ++pl)
          SRW_CUDA(cudaMemcpyAsync(ws + L.xp[l + 1] + ((int64_t)pl * T + r0) * D * 2, ws + L.xp[l] + ((int64_t)pl * T + r0) * D * 2, (size_t)nr * D * 2,
                                   cudaMemcpyDeviceToDevice, s));
      }
      SRW_TRY(enc_layer_forward(ws, ed, r.seq0, r.nseq, L.lay[l], lw, L.xf[l], L.xp[l], L.xf[l + 1], L.xp[l + 1], nullptr, nullptr, eps, 0.125f, ds, impl, s));
      covered = r.seq0 + r.nseq;
    }
    if (covered < S) {
      const int64_t r0 = (int64_t)covered * Fr, nr = (int64_t)(S - covered) * Fr;
      SRW_CUDA(cudaMemcpyAsync(F32(L.xf[l + 1]) + r0 * D, F32(L.xf[l]) + r0 * D, (size_t)nr * D * 4, cudaMemcpyDeviceToDevice, s));
      for (int pl = 0; pl < 2; ++pl)
        SRW_CUDA(cudaMemcpyAsync(ws + L.xp[l + 1] + ((int64_t)pl * T + r0) * D * 2, ws + L.xp[l] + ((int64_t)pl * T + r0) * D * 2, (size_t)nr * D * 2,
                                 cudaMemcpyDeviceToDevice, s));
    }
  }
  // ---- pooled features + classifier ----
  {
    const DropParams dr = hub_site_drop(dk, drw, 2 + 4 * d.NL, cf.p_pooled);
    const float* cls[4] = {P[hp_cls(d, 0)], P[hp_cls(d, 1)], P[hp_cls(d, 2)], P[hp_cls(d, 3)]};
    SRW_TRY(enc_head_forward(F32(L.xf[d.NL]), S, Fr, D, d.C, dr, nullptr, cls, F32(L.feat), F32(L.z1), F32(L.a1), a->logits, a->feat, s));
  }
  return SRW_OK;
}

static int hubert_backward_body(const srw_hubert_bwd_args* a, cudaStream_t s) {
  SRW_REQUIRE(a && a->cfg && a->params && a->weight_planes && a->wav && a->dlogits && a->grads && a->workspace, "srw_hubert_backward: null pointer");
  HDims d;
  SRW_TRY(make_hdims(a->cfg, a->batch, a->samples, a->grad_batch, d));
  SRW_REQUIRE(d.Sg > 0, "srw_hubert_backward: grad_batch must be > 0");
  const HLayout L = make_hlayout(d);
  SRW_REQUIRE(a->workspace_bytes >= L.total, "srw_hubert_backward: workspace too small");
  const HWOff w = hub_weight_layout(d);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  const uint8_t* wp = reinterpret_cast<const uint8_t*>(a->weight_planes);
  const float* const* P = a->params;
  float* const* G = a->grads;
  auto F32 = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
  auto BF = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
  const int S = d.S, Sg = d.Sg, D = d.D, impl = a->gemm_impl, acc = a->accumulate_grads ? 1 : 0, Fr = d.Fr, Tg = (int)d.Tg, T = (int)d.T;
  const uint32_t* dk = a->drop_seq_key;
  const int32_t* drw = a->drop_seq_row;
  const srw_hubert_config& cf = *a->cfg;
  const int64_t DD = (int64_t)D * D;
  for (int l = 0; l < d.NL; ++l)
    SRW_REQUIRE(G[hp_layer(d, l, HY_KW)] == G[hp_layer(d, l, HY_QW)] + DD && G[hp_layer(d, l, HY_VW)] == G[hp_layer(d, l, HY_QW)] + 2 * DD &&
                    G[hp_layer(d, l, HY_KB)] == G[hp_layer(d, l, HY_QB)] + D && G[hp_layer(d, l, HY_VB)] == G[hp_layer(d, l, HY_QB)] + 2 * D,
                "srw_hubert_backward: the gradients of q_proj / k_proj / v_proj weights (and of their biases) must be contiguous in that order (layer %d)", l);
  float* dx = F32(L.dx);
  (void)S; (void)T;

  // ---- classifier + pooling ----
  {
    const DropParams dr = hub_site_drop(dk, drw, 2 + 4 * d.NL, cf.p_pooled);
    const float* cls[4] = {P[hp_cls(d, 0)], P[hp_cls(d, 1)], P[hp_cls(d, 2)], P[hp_cls(d, 3)]};
    float* gcls[4] = {G[hp_cls(d, 0)], G[hp_cls(d, 1)], G[hp_cls(d, 2)], G[hp_cls(d, 3)]};
    SRW_TRY(enc_head_backward(a->dlogits, a->dfeat, Sg, Fr, D, d.C, dr, nullptr, cls, gcls, F32(L.feat), F32(L.z1), F32(L.a1), F32(L.dz1), F32(L.dfeat), dx,
                              acc, s));
  }
  // ---- encoder layers ----
  const EncDims ed = hub_enc_dims(d);
  std::vector<Run> runs;
  for (int l = d.NL - 1; l >= 0; --l) {
    const EncLayerW lw = hub_layer_weights(P, wp, w, d, l);
    const EncDropSites ds = hub_layer_sites(dk, drw, l, cf);
    EncLayerG lg;
    lg.qkv_w = G[hp_layer(d, l, HY_QW)]; lg.qkv_b = G[hp_layer(d, l, HY_QB)]; lg.ow = G[hp_layer(d, l, HY_OW)]; lg.ob = G[hp_layer(d, l, HY_OB)];
    lg.ln1w = G[hp_layer(d, l, HY_LN1W)]; lg.ln1b = G[hp_layer(d, l, HY_LN1B)]; lg.f1w = G[hp_layer(d, l, HY_F1W)]; lg.f1b = G[hp_layer(d, l, HY_F1B)];
    lg.f2w = G[hp_layer(d, l, HY_F2W)]; lg.f2b = G[hp_layer(d, l, HY_F2B)]; lg.ln2w = G[hp_layer(d, l, HY_LN2W)]; lg.ln2b = G[hp_layer(d, l, HY_LN2B)];
    SRW_TRY(layer_runs(a->num_segments, a->segment_start, a->layer_skip, d.NL, l, d.S, Sg, runs));
    int racc = acc;
    for (const Run& r : runs) {   // clips that skipped the layer keep dx unchanged (the identity)
      SRW_TRY(enc_layer_backward(ws, ed, r.seq0, r.nseq, L.lay[l], lw, lg, L.xp[l], dx, L.bw, nullptr, nullptr, 0.125f, ds, racc, impl, s));
      racc = 1;
    }
    if (runs.empty() && !acc) {   // every gradient clip skipped this layer: its parameters get zero gradient this step
      const int which[16] = {HY_KW, HY_KB, HY_VW, HY_VB, HY_QW, HY_QB, HY_OW, HY_OB, HY_LN1W, HY_LN1B, HY_F1W, HY_F1B, HY_F2W, HY_F2B, HY_LN2W, HY_LN2B};
      const int64_t numel[16] = {DD, D, DD, D, DD, D, DD, D, D, D, (int64_t)d.F * D, d.F, (int64_t)D * d.F, D, D, D};
      for (int i = 0; i < 16; ++i) SRW_CUDA(cudaMemsetAsync(G[hp_layer(d, l, which[i])], 0, (size_t)numel[i] * 4, s));
    }
  }
  // ---- encoder input: x = dropout(LN(hsum)), hsum = h + gelu(pz) ----
  {
    const DropParams dr = hub_site_drop(dk, drw, 1, cf.p_hidden);
    if (dr.on) {
      enc_dropout_rows_kernel<<<148 * 8, 256, 0, s>>>(dx, Tg, D, Fr, dr);
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    srw_layernorm_bwd_args lb = {};
    lb.dy = dx; lb.lddy = D; lb.x = F32(L.hsum); lb.ldx = D; lb.rows = Tg; lb.cols = D; lb.gamma = P[hp_after_conv(d, ENC_LNW)];
    lb.mean = F32(L.mean_e); lb.rstd = F32(L.rstd_e); lb.dx = F32(L.dhsum); lb.lddx = D; lb.accumulate_dx = 0;
    lb.dgamma = G[hp_after_conv(d, ENC_LNW)]; lb.dbeta = G[hp_after_conv(d, ENC_LNB)]; lb.accumulate_dparams = acc; lb.workspace = F32(L.bw.ln_ws);
    SRW_TRY(srw_layernorm_bwd(&lb, s));
  }
  // ---- positional conv backward ----
  {
    const int K = d.PK * d.GC, front_d = d.PK / 2 - 1;
    SRW_CUDA(cudaMemsetAsync(ws + L.dzg, 0, (size_t)d.PG * L.dzg_rows * d.GC * 4, s));
    hub_pos_bwd_prep_kernel<<<148 * 4, 256, 0, s>>>(F32(L.dhsum), F32(L.pz), Sg, Fr, D, BF(L.dzg), (int64_t)d.PG * L.dzg_rows * d.GC, L.dzg_rows, d.P, front_d,
                                                    d.GC, F32(L.dpz));
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_colsum_kernel<<<cdiv(D, 32), 256, 0, s>>>(F32(L.dpz), Tg, D, G[hp_after_conv(d, POS_B)], acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
    const int64_t M = (int64_t)Sg * d.P;
    const int split = pos_fwd_split(d, M);
    SRW_REQUIRE((int64_t)split * M <= L.pos_split * (int64_t)d.S * d.P, "srw_hubert_backward: positional dgrad workspace");
    for (int g_ = 0; g_ < d.PG; ++g_) {
      Gemm g((int)M, d.GC, K, impl);   // dh[(s, t), group cols] = window of dzg rows (t .. t + 127 in padded coordinates) x flipped taps
      g.g.a = ws + L.dzg + (int64_t)g_ * L.dzg_rows * d.GC * 2; g.g.lda = d.GC; g.g.a_plane_stride = (int64_t)d.PG * L.dzg_rows * d.GC; g.g.a_mn_major = 0;
      g.g.b = wp + w.posd + (int64_t)g_ * d.GC * K * 2; g.g.ldb = K; g.g.b_plane_stride = (int64_t)d.D * d.GC * d.PK; g.g.b_mn_major = 0;
      g.g.epilogue = SRW_EPI_SPLITK; g.g.split_k = split; g.g.workspace = F32(L.posws) + (int64_t)g_ * split * M * d.GC;
      SRW_TRY(g.run(s));
    }
    // weight gradient: dw[g][o][(j, i)] = sum over rows (s * P + t) of dz[row + front_d][o] * xg[row + j][i]
    const int wsplit = (int)L.posw_split;
    for (int g_ = 0; g_ < d.PG; ++g_) {
      Gemm g(d.GC, K, (int)M, impl);
      g.g.a = ws + L.dzg + ((int64_t)g_ * L.dzg_rows + front_d) * d.GC * 2; g.g.lda = d.GC; g.g.a_plane_stride = (int64_t)d.PG * L.dzg_rows * d.GC; g.g.a_mn_major = 1;
      g.g.b = ws + L.xg + (int64_t)g_ * L.xg_rows * d.GC * 2; g.g.ldb = d.GC; g.g.b_plane_stride = (int64_t)d.PG * L.xg_rows * d.GC; g.g.b_mn_major = 1;
      g.g.epilogue = SRW_EPI_SPLITK; g.g.split_k = wsplit; g.g.workspace = F32(L.posw_ws) + (int64_t)g_ * wsplit * d.GC * K;
      SRW_TRY(g.run(s));
    }
    const float* norm = reinterpret_cast<const float*>(wp + w.pos_norm);
    hub_pos_wn_dot_kernel<<<d.PK, 256, 0, s>>>(F32(L.posw_ws), wsplit, P[hp_after_conv(d, POS_V)], D, d.GC, d.PK, F32(L.dw_pos), F32(L.dot_pos));
    g_launches++;
    SRW_LAUNCH_CHECK();
    const int64_t nv = (int64_t)D * d.GC * d.PK;
    hub_pos_wn_apply_kernel<<<148 * 4, 256, 0, s>>>(F32(L.dw_pos), P[hp_after_conv(d, POS_V)], P[hp_after_conv(d, POS_G)], norm, F32(L.dot_pos), nv, d.PK,
                                                    G[hp_after_conv(d, POS_V)], G[hp_after_conv(d, POS_G)], acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
    const DropParams dr = hub_site_drop(dk, drw, 0, cf.p_feat_proj);
#define SRW_POSB(J)                                                                                                                                    \
  hub_pos_finish_bwd_kernel<J><<<(int)L.mse_nparts, 256, 0, s>>>(F32(L.posws), split, M, d.P, d.GC, F32(L.dhsum), Sg, Fr, a->mask_time, dr, F32(L.dh0), \
                                                                 BF(L.dh0p), (int64_t)Tg * D, F32(L.mse_part))
    if (D == 384) SRW_POSB(3); else if (D == 768) SRW_POSB(6); else SRW_POSB(8);
#undef SRW_POSB
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_mse_grad_kernel<<<cdiv(D, 128), 128, 0, s>>>(F32(L.mse_part), (int)L.mse_nparts, D, G[HP_MSE], acc, a->mask_time ? 1 : 0);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  // ---- feature projection backward ----
  {
    hub_colsum_kernel<<<cdiv(D, 32), 256, 0, s>>>(F32(L.dh0), Tg, D, G[hp_after_conv(d, FP_B)], acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
    SRW_TRY(wgrad(D, HC, Tg, ws + L.dh0p, D, Tg, ws + L.lnp, HC, d.T, F32(L.projw_ws), G[hp_after_conv(d, FP_W)], HC, acc, impl, s));
    Gemm g(Tg, HC, D, impl);   // d(ln out) = dh Wp
    g.A(ws + L.dh0p, D, Tg, 0).Bm(wp + w.proj, HC, D, 1);
    g.g.epilogue = SRW_EPI_F32; g.g.out_f32 = F32(L.dln); g.g.ldo = HC;
    SRW_TRY(g.run(s));
    srw_layernorm_bwd_args lb = {};
    lb.dy = F32(L.dln); lb.lddy = HC; lb.x = F32(L.g_last); lb.ldx = HC; lb.rows = Tg; lb.cols = HC; lb.gamma = P[hp_after_conv(d, FP_LNW)];
    lb.mean = F32(L.mean_fp); lb.rstd = F32(L.rstd_fp); lb.dx = F32(L.dg_last); lb.lddx = HC; lb.accumulate_dx = 0;
    lb.dgamma = G[hp_after_conv(d, FP_LNW)]; lb.dbeta = G[hp_after_conv(d, FP_LNB)]; lb.accumulate_dparams = acc; lb.workspace = F32(L.bw.ln_ws);
    SRW_TRY(srw_layernorm_bwd(&lb, s));
  }
  // ---- conv stem backward ----
  {
    const int last = d.NC - 1;
    // dz buffers: [1 zero row][Sg * tp rows][slack]; element offset of row 0 = HC
    const int64_t dz_ps = L.dz_rows * HC;
    for (int i = 0; i < 2; ++i) {
      SRW_CUDA(cudaMemsetAsync(ws + L.dzbuf[i], 0, (size_t)HC * 2, s));
      SRW_CUDA(cudaMemsetAsync(ws + L.dzbuf[i] + dz_ps * 2, 0, (size_t)HC * 2, s));
    }
    int cur = last & 1;
    hub_gelu_bwd_scatter_kernel<<<148 * 4, 256, 0, s>>>(F32(L.dg_last), F32(L.z[last]), Sg, d.tp[last], Fr, BF(L.dzbuf[cur]) + HC, dz_ps);
    g_launches++;
    SRW_LAUNCH_CHECK();
    for (int l = last; l >= 1; --l) {
      const int k = d.k[l];
      const int64_t M = (int64_t)Sg * d.tp[l];
      const uint8_t* dz = ws + L.dzbuf[cur] + HC * 2;     // row 0 of dz_l
      // weight gradient: dWf[co, (j, ci)] = sum over rows of dz[row, co] * y_{l-1}[view row, (j, ci)]
      {
        int split = 1;
        splitk_for(HC, k * HC, M, &split);
        Gemm g(HC, k * HC, (int)M, impl);
        g.g.a = dz; g.g.lda = HC; g.g.a_plane_stride = dz_ps; g.g.a_mn_major = 1;
        g.g.b = ws + L.y[l - 1]; g.g.ldb = (int64_t)d.st[l] * HC; g.g.b_plane_stride = L.y_rows[l - 1] * HC; g.g.b_mn_major = 1;
        g.g.epilogue = SRW_EPI_SPLITK; g.g.split_k = split; g.g.workspace = F32(L.convw_ws);
        SRW_TRY(g.run(s));
        hub_conv_wgrad_finish_kernel<<<148 * 4, 256, 0, s>>>(F32(L.convw_ws), split, k, G[HP_CONV1 + l - 1], acc);
        g_launches++;
        SRW_LAUNCH_CHECK();
      }
      // input gradient.  Output rows u of the GEMMs = pairs of input frames (2u, 2u + 1): the previous layer's buffers seen with row stride 2 * 512.
      const bool to_conv0 = (l == 1);
      uint8_t* nxt = ws + L.dzbuf[cur ^ 1] + HC * 2;
      auto set_out = [&](Gemm& g, int half) {   // half 0: even frames, 1: odd frames, -1: both (N = 1024)
        const int64_t off = half > 0 ? HC : 0;
        if (to_conv0) {
          g.g.epilogue = SRW_EPI_F32; g.g.out_f32 = F32(L.dy0) + off; g.g.ldo = 2 * HC;
        } else {
          g.g.epilogue = SRW_EPI_DGELU; g.g.aux = F32(L.z[l - 1]) + off; g.g.ldaux = 2 * HC;
          g.g.out_planes = nxt + off * 2; g.g.ldp = 2 * HC; g.g.out_plane_stride = dz_ps;
        }
      };
      if (k == 2) {
        Gemm g((int)M, 2 * HC, HC, impl);   // [dx(2u) | dx(2u + 1)] = dz[u] * [W_0 | W_1]
        g.g.a = dz; g.g.lda = HC; g.g.a_plane_stride = dz_ps; g.g.a_mn_major = 0;
        g.g.b = wp + w.convf[l]; g.g.ldb = 2 * HC; g.g.b_plane_stride = (int64_t)HC * k * HC; g.g.b_mn_major = 1;
        set_out(g, -1);
        SRW_TRY(g.run(s));
      } else {
        {
          Gemm g((int)M, HC, 2 * HC, impl);   // dx(2u) = [dz[u - 1] | dz[u]] * [W_2 ; W_0]: two-frame overlapping view starting one row early
          g.g.a = dz - HC * 2; g.g.lda = HC; g.g.a_plane_stride = dz_ps; g.g.a_mn_major = 0;
          g.g.b = wp + w.conve[l]; g.g.ldb = 2 * HC; g.g.b_plane_stride = (int64_t)HC * 2 * HC; g.g.b_mn_major = 0;
          set_out(g, 0);
          SRW_TRY(g.run(s));
        }
        {
          Gemm g((int)M, HC, HC, impl);       // dx(2u + 1) = dz[u] * W_1
          g.g.a = dz; g.g.lda = HC; g.g.a_plane_stride = dz_ps; g.g.a_mn_major = 0;
          g.g.b = wp + w.convf[l] + HC * 2; g.g.ldb = (int64_t)k * HC; g.g.b_plane_stride = (int64_t)HC * k * HC; g.g.b_mn_major = 1;
          set_out(g, 1);
          SRW_TRY(g.run(s));
        }
      }
      cur ^= 1;
    }
    // conv 0: GroupNorm + GELU + the 10-tap conv, recomputed from the waveform
    const dim3 grid(cdiv(d.n[0], C0_FRAMES), Sg);
    hub_conv0_bwd_stats_kernel<<<grid, 256, 0, s>>>(a->wav, a->ld_wav, d.Ts, d.n[0], d.tp[0], P[HP_C0W], F32(L.mean0), F32(L.rstd0), P[HP_GNW], P[HP_GNB],
                                                    F32(L.dy0), reinterpret_cast<float2*>(ws + L.bpart0));
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_conv0_bwd_finish_kernel<<<cdiv(HC, 8), 256, 0, s>>>(reinterpret_cast<const float2*>(ws + L.bpart0), Sg, grid.x, d.n[0],
                                                              reinterpret_cast<float2*>(ws + L.ab0), G[HP_GNW], G[HP_GNB], acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_conv0_bwd_wgrad_kernel<<<grid, 256, 0, s>>>(a->wav, a->ld_wav, d.Ts, d.n[0], d.tp[0], P[HP_C0W], F32(L.mean0), F32(L.rstd0), P[HP_GNW], P[HP_GNB],
                                                    F32(L.dy0), reinterpret_cast<const float2*>(ws + L.ab0), F32(L.wpart0));
    g_launches++;
    SRW_LAUNCH_CHECK();
    hub_conv0_wgrad_finish_kernel<<<cdiv(HC * C0K, 256), 256, 0, s>>>(F32(L.wpart0), Sg * (int)grid.x, G[HP_C0W], acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  return SRW_OK;
}

static bool hub_skips(int num_segments, const uint8_t* layer_skip, int layers) {
  if (num_segments <= 0 || !layer_skip) return false;
  for (int64_t i = 0; i < (int64_t)num_segments * layers; ++i)
    if (layer_skip[i]) return true;
  return false;
}

extern "C" int srw_hubert_forward(const srw_hubert_fwd_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->cfg && a->params, "srw_hubert_forward: null pointer");
  // a LayerDrop pattern changes the launch sequence from step to step: those calls run eagerly
  if (!graphs_enabled(s) || a->cfg->layers <= 0 || a->cfg->layers > 48 || hub_skips(a->num_segments, a->layer_skip, a->cfg->layers)) return hubert_forward_body(a, s);
  KeyBuilder kb;
  kb.add((int)21); kb.add(*a->cfg); kb.add(s);
  for (int i = 0; i < hub_num_params(a->cfg); ++i) kb.add(a->params[i]);
  kb.add(a->weight_planes); kb.add(a->wav); kb.add(a->ld_wav); kb.add(a->batch); kb.add(a->samples); kb.add(a->grad_batch); kb.add(a->mask_time);
  kb.add(a->drop_seq_key); kb.add(a->drop_seq_row); kb.add(a->logits); kb.add(a->feat); kb.add(a->workspace); kb.add(a->workspace_bytes); kb.add(a->gemm_impl);
  return run_graphed(std::move(kb.k), s, [a](cudaStream_t st) { return hubert_forward_body(a, st); });
}

extern "C" int srw_hubert_backward(const srw_hubert_bwd_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->cfg && a->params && a->grads, "srw_hubert_backward: null pointer");
  if (!graphs_enabled(s) || a->cfg->layers <= 0 || a->cfg->layers > 48 || hub_skips(a->num_segments, a->layer_skip, a->cfg->layers)) return hubert_backward_body(a, s);
  KeyBuilder kb;
  kb.add((int)22); kb.add(*a->cfg); kb.add(s);
  for (int i = 0; i < hub_num_params(a->cfg); ++i) { kb.add(a->params[i]); kb.add(a->grads[i]); }
  kb.add(a->weight_planes); kb.add(a->wav); kb.add(a->ld_wav); kb.add(a->batch); kb.add(a->samples); kb.add(a->grad_batch); kb.add(a->mask_time);
  kb.add(a->drop_seq_key); kb.add(a->drop_seq_row); kb.add(a->dlogits); kb.add(a->dfeat); kb.add(a->accumulate_grads); kb.add(a->workspace);
  kb.add(a->workspace_bytes); kb.add(a->gemm_impl);
  return run_graphed(std::move(kb.k), s, [a](cudaStream_t st) { return hubert_backward_body(a, st); });
}
