// srw_wrn — WideResNet.forward (semilearn/nets/wrn/wrn.py:118-146; BasicBlock :30-54) and its backward as two native calls.
//
// Layout: every activation is NHWC with a one-pixel zero border, flattened to "position rows": image n, padded pixel (y', x') with
// y', x' in [0, H + 2) is row (n (H+2) + y') (W+2) + x' of a [rows, C] matrix.  GEMM operands are split planes of that matrix with
// W + 3 zero rows in front and behind; convolution outputs and gradients are fp32 [rows, C].
//   3x3, stride 1, padding 1:  out[row] = sum_{dy} in[row + (dy - 1)(W+2) - 1 .. + 3 C) * W[:, dy, :, :]   — three K segments, each
//       3 C CONTIGUOUS elements of the input (dx and channel run together in NHWC), kernel row dy = (W+2) rows further down: ONE
//       srw_gemm with a_seg_k = 3 C, a_seg_rows = W + 2 over an overlapping row view (lda = C); no im2col buffer.  Border rows of the
//       output are garbage and are masked by every consumer (statistics skip them, operand writers zero them).
//   dgrad = the same GEMM over the output gradient with flipped kernels; wgrad = three MN-major GEMMs (one per kernel row) with
//       the rows as the reduction dimension (split-K), folded and permuted into [co, ci, 3, 3] by one kernel.
//   3x3 stride 2 (first layer of block2 / block3) = the stride-1 GEMM at full resolution followed by a 2x sub-sampling (4x the FLOPs
//       of two of the 25 convolutions, no second code path); its backward scatters the gradient into a zero full-resolution tensor.
//   1x1 shortcut convolutions are plain GEMMs over the position rows (stride 2: over a sub-sampled copy of the input planes).
// BatchNorm couples all rows of the launch (wrn.py:33,37,97: train-mode batch statistics): per layer a two-stage column reduction
// (fp32 partials per CTA, folded in fp64) gives mean / biased variance, advances the running statistics (momentum 0.001, unbiased
// variance) and the next pass normalises + LeakyReLU(0.1) + splits into the next convolution's operand planes.  The backward needs
// sum(du) and sum(du * xhat) over all rows the same way, so the weak rows of the batch take part in it (9 B F).
#include <cuda.h>

#include <vector>

#include "../../include/srw.h"
#include "srw_common.cuh"
#include "srw_engine.cuh"

namespace srw {

struct Geo { int Hs, Wp, Pp; };   // spatial size, padded width, positions per image
__device__ __forceinline__ bool geo_valid(const Geo& g, int64_t row, int& n, int& y, int& x) {
  n = (int)(row / g.Pp);
  const int p = (int)(row % g.Pp);
  y = p / g.Wp; x = p % g.Wp;
  return y >= 1 && y <= g.Hs && x >= 1 && x <= g.Hs;
}
__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// image batch NCHW fp32 -> position-row planes with 8 channels (3 real + 5 zero: 16-byte rows for TMA), zero border
__global__ void wrn_input_planes_kernel(const float* __restrict__ x, int N, Geo g, __nv_bfloat16* __restrict__ planes, int64_t plane_stride) {
  const int64_t rows = (int64_t)N * g.Pp;
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    int n, y, xx;
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    if (geo_valid(g, r, n, y, xx)) {
      const int64_t hw = (int64_t)g.Hs * g.Hs;
      const float* px = x + (int64_t)n * 3 * hw + (int64_t)(y - 1) * g.Hs + (xx - 1);
      split2(px[0], px[hw], h[0], l[0]);
      split2(px[2 * hw], 0.f, h[1], l[1]);
    }
    *reinterpret_cast<uint4*>(planes + r * 8) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(planes + plane_stride + r * 8) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics: column sums over the valid rows, two stages
// ------------------------------------------------------------------------------------------------
constexpr int BN_ROWS_PER_CTA = 1024;
// MODE 0: partial = (sum x, sum x^2).  MODE 1 (backward): du = dy * lrelu'(u), u = (x - mean) rstd gamma + beta; partial = (sum du, sum du xhat)
template <int MODE>
__global__ void __launch_bounds__(256) wrn_bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t rows, int C, Geo g,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float slope, float* __restrict__ partial) {
  __shared__ float4 red[2][256];
  const int tpr = C / 4;                        // threads per row (4 .. 32)
  const int rl = threadIdx.x / tpr, q = threadIdx.x % tpr, rpi = 256 / tpr;
  const int c = q * 4;
  const int64_t r0 = (int64_t)blockIdx.x * BN_ROWS_PER_CTA, r1 = min(rows, r0 + BN_ROWS_PER_CTA);
  float4 m4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = m4, g4 = m4, b4 = m4;
  if (MODE == 1) {
    m4 = *reinterpret_cast<const float4*>(mean + c);
    a4 = *reinterpret_cast<const float4*>(rstd + c);
    g4 = *reinterpret_cast<const float4*>(gamma + c);
    b4 = *reinterpret_cast<const float4*>(beta + c);
  }
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), t = s;
  constexpr int U = 4;    // rows in flight per thread: the kernel is a pure HBM stream, one 16-byte load per row would leave it latency-bound
  for (int64_t r = r0 + rl; r < r1; r += (int64_t)U * rpi) {
    float4 v[U], d[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t rr = r + (int64_t)u * rpi;
      int n, y, xx;
      ok[u] = rr < r1 && geo_valid(g, rr, n, y, xx);
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f); d[u] = v[u];
      if (ok[u]) {
        v[u] = *reinterpret_cast<const float4*>(x + rr * C + c);
        if (MODE == 1) d[u] = *reinterpret_cast<const float4*>(dy + rr * C + c);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      if (MODE == 0) {
        s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w;
        t.x = fmaf(v[u].x, v[u].x, t.x); t.y = fmaf(v[u].y, v[u].y, t.y); t.z = fmaf(v[u].z, v[u].z, t.z); t.w = fmaf(v[u].w, v[u].w, t.w);
      } else {
        const float hx = (v[u].x - m4.x) * a4.x, hy = (v[u].y - m4.y) * a4.y, hz = (v[u].z - m4.z) * a4.z, hw = (v[u].w - m4.w) * a4.w;
        const float dx_ = fmaf(hx, g4.x, b4.x) > 0.f ? d[u].x : d[u].x * slope, dy_ = fmaf(hy, g4.y, b4.y) > 0.f ? d[u].y : d[u].y * slope;
        const float dz_ = fmaf(hz, g4.z, b4.z) > 0.f ? d[u].z : d[u].z * slope, dw_ = fmaf(hw, g4.w, b4.w) > 0.f ? d[u].w : d[u].w * slope;
        s.x += dx_; s.y += dy_; s.z += dz_; s.w += dw_;
        t.x = fmaf(dx_, hx, t.x); t.y = fmaf(dy_, hy, t.y); t.z = fmaf(dz_, hz, t.z); t.w = fmaf(dw_, hw, t.w);
      }
    }
  }
  red[0][threadIdx.x] = s; red[1][threadIdx.x] = t;
  __syncthreads();
  if (rl == 0) {
    for (int j = 1; j < rpi; ++j) {
      const float4 a = red[0][j * tpr + q], b = red[1][j * tpr + q];
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
      t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
    }
    float* dst = partial + (int64_t)blockIdx.x * 2 * C;
    *reinterpret_cast<float4*>(dst + c) = s;
    *reinterpret_cast<float4*>(dst + C + c) = t;
  }
}
// forward finish: mean, rstd of the batch (training) or of the running statistics (eval); running statistics advanced `1 + repeats`
// times in training (F.batch_norm: running = (1 - m) running + m stat, unbiased variance; `repeats` = further identical passes of
// the same batch, the deterministic sampling passes of stage 2)
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// one warp per channel: lanes stride over the per-CTA partials (fp64), shuffle tree
__device__ __forceinline__ void bn_fold_partials(const float* __restrict__ partial, int nparts, int C, int c, double& s, double& q) {
  const int lane = threadIdx.x & 31;
  double a = 0.0, b = 0.0;
  for (int p = lane; p < nparts; p += 32) { a += partial[(int64_t)p * 2 * C + c]; b += partial[(int64_t)p * 2 * C + C + c]; }
  s = warp_sum_f64(a); q = warp_sum_f64(b);
}
// SyncBatchNorm: sums[c] = sum, sums[C + c] = sum of squares of THIS rank (all-reduced by the host before the finish kernel reads them);
// backward form: also the LOCAL dgamma / dbeta
__global__ void wrn_bn_fold_kernel(const float* __restrict__ partial, int nparts, int C, float* __restrict__ sums, float* __restrict__ dgamma,
                                   float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  bn_fold_partials(partial, nparts, C, c, s, q);
  if ((threadIdx.x & 31) != 0) return;
  sums[c] = (float)s; sums[C + c] = (float)q;
  if (dgamma) { dgamma[c] = accumulate ? dgamma[c] + (float)q : (float)q; dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s; }
}
__global__ void wrn_bn_finish_kernel(const float* __restrict__ partial, int nparts, int C, double n, float eps, float momentum, int training, int repeats,
                                     float* __restrict__ running_mean, float* __restrict__ running_var, int64_t* __restrict__ num_batches_tracked,
                                     float* __restrict__ mean_out, float* __restrict__ rstd_out, const float* __restrict__ sums) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const bool lead = (threadIdx.x & 31) == 0;
  if (!training) {
    if (lead) {
      mean_out[c] = running_mean[c];
      rstd_out[c] = 1.0f / sqrtf(running_var[c] + eps);
    }
    return;
  }
  double s, q;
  if (sums) { s = sums[c]; q = sums[C + c]; }        // global sums (SyncBatchNorm)
  else bn_fold_partials(partial, nparts, C, c, s, q);
  if (!lead) return;
  const double m = s / n;
  const double var = fmax(q / n - m * m, 0.0);
  const float mf = (float)m, vf = (float)var;
  mean_out[c] = mf;
  rstd_out[c] = 1.0f / sqrtf(vf + eps);
  const float unbiased = (float)(var * (n / (n - 1.0)));
  float rm = running_mean[c], rv = running_var[c];
  for (int i = 0; i <= repeats; ++i) {
    rm = (1.0f - momentum) * rm + momentum * mf;
    rv = (1.0f - momentum) * rv + momentum * unbiased;
  }
  running_mean[c] = rm; running_var[c] = rv;
  if (c == 0 && num_batches_tracked) *num_batches_tracked += 1 + repeats;
}
// backward finish: dgamma (+)= sum du xhat, dbeta (+)= sum du; coef = (sum du / n, sum du xhat / n)
__global__ void wrn_bn_bwd_finish_kernel(const float* __restrict__ partial, int nparts, int C, double n, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                         int accumulate, float* __restrict__ coef, const float* __restrict__ sums) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  if (sums) { s = sums[c]; q = sums[C + c]; }
  else bn_fold_partials(partial, nparts, C, c, s, q);
  if ((threadIdx.x & 31) != 0) return;
  if (dgamma) { dgamma[c] = accumulate ? dgamma[c] + (float)q : (float)q; dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s; }
  coef[c] = (float)(s / n); coef[C + c] = (float)(q / n);
}

// y = lrelu((x - mean) rstd gamma + beta) as operand planes with zero border (raw != 0: y = x, the un-normalised input of the
// first convolution of block2 / block3, wrn.py:51)
__global__ void wrn_bn_act_planes_kernel(const float* __restrict__ x, int64_t rows, int C, Geo g, const float* __restrict__ mean, const float* __restrict__ rstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta, float slope, int raw,
                                         __nv_bfloat16* __restrict__ planes, int64_t plane_stride) {
  const int c4 = C / 4;
  const int64_t total = rows * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i % c4) * 4;
    int n, y, xx;
    uint32_t h0 = 0, l0 = 0, h1 = 0, l1 = 0;
    if (geo_valid(g, r, n, y, xx)) {
      float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
      if (!raw) {
        const float4 m4 = *reinterpret_cast<const float4*>(mean + c), a4 = *reinterpret_cast<const float4*>(rstd + c);
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + c), b4 = *reinterpret_cast<const float4*>(beta + c);
        v.x = lrelu(fmaf((v.x - m4.x) * a4.x, g4.x, b4.x), slope); v.y = lrelu(fmaf((v.y - m4.y) * a4.y, g4.y, b4.y), slope);
        v.z = lrelu(fmaf((v.z - m4.z) * a4.z, g4.z, b4.z), slope); v.w = lrelu(fmaf((v.w - m4.w) * a4.w, g4.w, b4.w), slope);
      }
      split2(v.x, v.y, h0, l0);
      split2(v.z, v.w, h1, l1);
    }
    *reinterpret_cast<uint2*>(planes + r * C + c) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(planes + plane_stride + r * C + c) = make_uint2(l0, l1);
  }
}
// BatchNorm + LeakyReLU backward: dx = gamma rstd (du - coef0 - xhat coef1) (+ resid), du = dy lrelu'(u).  Outputs fp32 and / or
// operand planes with zero border.  identity != 0: dx = dy (+ resid) — only the plane conversion / residual add.
__global__ void wrn_bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, int64_t rows, int C, Geo g, const float* __restrict__ mean,
                                        const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
                                        const float* __restrict__ coef, int identity, const float* __restrict__ resid, float* __restrict__ out_f32,
                                        __nv_bfloat16* __restrict__ planes, int64_t plane_stride) {
  const int c4 = C / 4;
  const int64_t total = rows * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i % c4) * 4;
    int n, y, xx;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (geo_valid(g, r, n, y, xx)) {
      const float4 d = *reinterpret_cast<const float4*>(dy + r * C + c);
      if (identity) {
        o = d;
      } else {
        const float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
        const float4 m4 = *reinterpret_cast<const float4*>(mean + c), a4 = *reinterpret_cast<const float4*>(rstd + c);
        const float4 g4 = *reinterpret_cast<const float4*>(gamma + c), b4 = *reinterpret_cast<const float4*>(beta + c);
        const float4 k0 = *reinterpret_cast<const float4*>(coef + c), k1 = *reinterpret_cast<const float4*>(coef + C + c);
        const float hx = (v.x - m4.x) * a4.x, hy = (v.y - m4.y) * a4.y, hz = (v.z - m4.z) * a4.z, hw = (v.w - m4.w) * a4.w;
        const float ux = fmaf(hx, g4.x, b4.x) > 0.f ? d.x : d.x * slope, uy = fmaf(hy, g4.y, b4.y) > 0.f ? d.y : d.y * slope;
        const float uz = fmaf(hz, g4.z, b4.z) > 0.f ? d.z : d.z * slope, uw = fmaf(hw, g4.w, b4.w) > 0.f ? d.w : d.w * slope;
        o.x = g4.x * a4.x * (ux - k0.x - hx * k1.x); o.y = g4.y * a4.y * (uy - k0.y - hy * k1.y);
        o.z = g4.z * a4.z * (uz - k0.z - hz * k1.z); o.w = g4.w * a4.w * (uw - k0.w - hw * k1.w);
      }
      if (resid) {
        const float4 rr = *reinterpret_cast<const float4*>(resid + r * C + c);
        o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
      }
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * C + c) = o;
    if (planes) {
      uint32_t h0, l0, h1, l1;
      split2(o.x, o.y, h0, l0);
      split2(o.z, o.w, h1, l1);
      *reinterpret_cast<uint2*>(planes + r * C + c) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(planes + plane_stride + r * C + c) = make_uint2(l0, l1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stride 2: sub-sampling (forward) and zero up-sampling (backward) between a full-resolution geometry gf and its half gh
// ------------------------------------------------------------------------------------------------
// dst (half, fp32 or planes) [n, y', x'] = src (full) [n, 2 (y' - 1) + 1, 2 (x' - 1) + 1]; border rows zero
__global__ void wrn_subsample_kernel(const float* __restrict__ src_f32, const __nv_bfloat16* __restrict__ src_planes, int64_t src_ps, int N, int C, Geo gf, Geo gh,
                                     float* __restrict__ dst_f32, __nv_bfloat16* __restrict__ dst_planes, int64_t dst_ps) {
  const int c4 = C / 4;
  const int64_t total = (int64_t)N * gh.Pp * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i % c4) * 4;
    int n, y, xx;
    const bool ok = geo_valid(gh, r, n, y, xx);
    const int64_t sr = (int64_t)n * gf.Pp + (int64_t)(2 * (y - 1) + 1) * gf.Wp + (2 * (xx - 1) + 1);
    if (dst_f32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) v = *reinterpret_cast<const float4*>(src_f32 + sr * C + c);
      *reinterpret_cast<float4*>(dst_f32 + r * C + c) = v;
    } else {
      uint2 h = make_uint2(0, 0), l = h;
      if (ok) { h = *reinterpret_cast<const uint2*>(src_planes + sr * C + c); l = *reinterpret_cast<const uint2*>(src_planes + src_ps + sr * C + c); }
      *reinterpret_cast<uint2*>(dst_planes + r * C + c) = h;
      *reinterpret_cast<uint2*>(dst_planes + dst_ps + r * C + c) = l;
    }
  }
}
// full-resolution planes = the half-resolution gradient at the sampled pixels, zero elsewhere
__global__ void wrn_upsample_planes_kernel(const float* __restrict__ src, int N, int C, Geo gf, Geo gh, __nv_bfloat16* __restrict__ planes, int64_t plane_stride) {
  const int c4 = C / 4;
  const int64_t total = (int64_t)N * gf.Pp * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i % c4) * 4;
    int n, y, xx;
    uint32_t h0 = 0, l0 = 0, h1 = 0, l1 = 0;
    if (geo_valid(gf, r, n, y, xx) && ((y - 1) & 1) == 0 && ((xx - 1) & 1) == 0) {
      const int64_t sr = (int64_t)n * gh.Pp + (int64_t)((y - 1) / 2 + 1) * gh.Wp + ((xx - 1) / 2 + 1);
      const float4 v = *reinterpret_cast<const float4*>(src + sr * C + c);
      split2(v.x, v.y, h0, l0);
      split2(v.z, v.w, h1, l1);
    }
    *reinterpret_cast<uint2*>(planes + r * C + c) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(planes + plane_stride + r * C + c) = make_uint2(l0, l1);
  }
}
// dst (full, fp32) [sampled pixels] += src (half, fp32)
__global__ void wrn_add_upsampled_kernel(const float* __restrict__ src, int N, int C, Geo gf, Geo gh, float* __restrict__ dst) {
  const int c4 = C / 4;
  const int64_t total = (int64_t)N * gh.Pp * c4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i % c4) * 4;
    int n, y, xx;
    if (!geo_valid(gh, r, n, y, xx)) continue;
    const int64_t dr = (int64_t)n * gf.Pp + (int64_t)(2 * (y - 1) + 1) * gf.Wp + (2 * (xx - 1) + 1);
    const float4 v = *reinterpret_cast<const float4*>(src + r * C + c);
    float4 d = *reinterpret_cast<float4*>(dst + dr * C + c);
    d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w;
    *reinterpret_cast<float4*>(dst + dr * C + c) = d;
  }
}

// ------------------------------------------------------------------------------------------------
// weights
// ------------------------------------------------------------------------------------------------
// conv weight [co, ci, 3, 3] -> forward operand  wf[co][dy * kpad + dx * cin_p + ci]            (kpad = ceil(3 cin_p / 64) * 64, zeros in the padding)
//                               dgrad operand    wd[ci][dy * kpad_o + dx * cout + co] = w[co, ci, 2 - dy, 2 - dx]
__global__ void wrn_conv_relayout_kernel(const float* __restrict__ w, int cout, int cin, int cin_p, int kpad, int kpad_o, __nv_bfloat16* __restrict__ wf,
                                         int64_t wf_ps, __nv_bfloat16* __restrict__ wd, int64_t wd_ps) {
  const int64_t n = (int64_t)cout * cin * 9;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx / (cin * 9)), rem = (int)(idx % (cin * 9));
    const int ci = rem / 9, dy = (rem % 9) / 3, dx = rem % 3;
    __nv_bfloat16 hi, lo;
    split_bf16(w[idx], hi, lo);
    const int64_t f = (int64_t)co * 3 * kpad + (int64_t)dy * kpad + dx * cin_p + ci;
    wf[f] = hi; wf[f + wf_ps] = lo;
    if (wd) {
      const int64_t d = (int64_t)ci * 3 * kpad_o + (int64_t)(2 - dy) * kpad_o + (2 - dx) * cout + co;
      wd[d] = hi; wd[d + wd_ps] = lo;
    }
  }
}
// the three per-kernel-row wgrad GEMMs leave split-K partials ws[dy][split][co][dx * cin_p + ci]; dW[co, ci, dy, dx] (+)= their sum
__global__ void wrn_conv_wgrad_finish_kernel(const float* __restrict__ ws, int split, int cout, int cin, int cin_p, float* __restrict__ dW, int accumulate) {
  const int64_t n = (int64_t)cout * cin * 9;
  const int64_t per = (int64_t)cout * 3 * cin_p;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx / (cin * 9)), rem = (int)(idx % (cin * 9));
    const int ci = rem / 9, dy = (rem % 9) / 3, dx = rem % 3;
    float a = 0.f;
    for (int p = 0; p < split; ++p) a += ws[((int64_t)dy * split + p) * per + (int64_t)co * 3 * cin_p + dx * cin_p + ci];
    dW[idx] = accumulate ? dW[idx] + a : a;
  }
}

// ------------------------------------------------------------------------------------------------
// pooled head
// ------------------------------------------------------------------------------------------------
// feat[n, c] = mean over the H x W pixels of lrelu(bn(x))   (wrn.py:143-145 + adaptive_avg_pool2d)
__global__ void wrn_pool_fwd_kernel(const float* __restrict__ x, int C, Geo g, const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, float slope, float* __restrict__ feat) {
  const int n = blockIdx.x, c = threadIdx.x;
  if (c >= C) return;
  const float m = mean[c], a = rstd[c], gm = gamma[c], b = beta[c];
  float acc = 0.f;
  for (int y = 1; y <= g.Hs; ++y)
    for (int xx = 1; xx <= g.Hs; ++xx) {
      const float v = x[((int64_t)n * g.Pp + (int64_t)y * g.Wp + xx) * C + c];
      acc += lrelu(fmaf((v - m) * a, gm, b), slope);
    }
  feat[(int64_t)n * C + c] = acc / (float)(g.Hs * g.Hs);
}
// dy[(n, pixel), c] = dfeat[n, c] / (H W) on the valid pixels (the gradient entering the final LeakyReLU)
__global__ void wrn_pool_bwd_kernel(const float* __restrict__ dfeat, int64_t rows, int C, Geo g, float* __restrict__ dy) {
  const int c4 = C / 4;
  const float inv = 1.0f / (float)(g.Hs * g.Hs);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * c4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i % c4) * 4;
    int n, y, xx;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (geo_valid(g, r, n, y, xx)) {
      v = *reinterpret_cast<const float4*>(dfeat + (int64_t)n * C + c);
      v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    }
    *reinterpret_cast<float4*>(dy + r * C + c) = v;
  }
}
// logits[n, k] = feat[n, :] . W[k, :] + b[k]     (wrn.py:100,135)
__global__ void wrn_fc_fwd_kernel(const float* __restrict__ feat, int N, int D, int K, const float* __restrict__ W, const float* __restrict__ b,
                                  float* __restrict__ logits) {
  extern __shared__ float f[];
  const int n = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) f[d] = feat[(int64_t)n * D + d];
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float a = 0.f;
    for (int d = 0; d < D; ++d) a = fmaf(f[d], W[(int64_t)k * D + d], a);
    logits[(int64_t)n * K + k] = a + b[k];
  }
}
// dfeat[n, d] = sum_k dlogits[n, k] W[k, d] (+ dfeat_in); rows >= grad_rows have no logit gradient (the weak rows of the batch)
__global__ void wrn_fc_bwd_rows_kernel(const float* __restrict__ dlogits, int grad_rows, const float* __restrict__ dfeat_in, int D, int K,
                                       const float* __restrict__ W, float* __restrict__ dfeat) {
  const int n = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = (dfeat_in && n < grad_rows) ? dfeat_in[(int64_t)n * D + d] : 0.f;
    if (n < grad_rows)
      for (int k = 0; k < K; ++k) a = fmaf(dlogits[(int64_t)n * K + k], W[(int64_t)k * D + d], a);
    dfeat[(int64_t)n * D + d] = a;
  }
}
// dW[k, d] (+)= sum_n dlogits[n, k] feat[n, d];  db[k] (+)= sum_n dlogits[n, k]
__global__ void wrn_fc_bwd_params_kernel(const float* __restrict__ dlogits, const float* __restrict__ feat, int grad_rows, int D, int K,
                                         float* __restrict__ dW, float* __restrict__ db, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * D) {
    const int k = i / D, d = i % D;
    float a = 0.f;
    for (int n = 0; n < grad_rows; ++n) a = fmaf(dlogits[(int64_t)n * K + k], feat[(int64_t)n * D + d], a);
    dW[i] = accumulate ? dW[i] + a : a;
  } else if (i < K * D + K) {
    const int k = i - K * D;
    float a = 0.f;
    for (int n = 0; n < grad_rows; ++n) a += dlogits[(int64_t)n * K + k];
    db[k] = accumulate ? db[k] + a : a;
  }
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
constexpr int WRN_MAX_BLOCKS = 64;
struct WBlock { int cin, cout, stride, abr, stage_in, stage_out; };
struct WDims {
  int N, C, img, nb;                 // batch rows, classes, image size, number of BasicBlocks
  int ch[4];
  Geo geo[3];
  int64_t M[3];                      // position rows per stage
  std::vector<WBlock> blocks;
  float momentum, slope;
};
static inline int kpad_of(int c) { return cdiv(3 * c, 64) * 64; }
static inline int64_t plane_rows(const WDims& d, int st) { return d.M[st] + 2 * (d.geo[st].Wp + 1) + 1; }
static inline int front_rows(const WDims& d, int st) { return d.geo[st].Wp + 1; }

static int make_wdims(const srw_wrn_config* c, int batch, WDims& d) {
  SRW_REQUIRE(c && batch > 0, "srw_wrn: bad batch");
  SRW_REQUIRE(c->depth >= 10 && (c->depth - 4) % 6 == 0 && c->num_classes > 0, "srw_wrn: depth must be 6 n + 4");
  SRW_REQUIRE(c->widen == 1 || c->widen == 2 || c->widen == 4 || c->widen == 8, "srw_wrn: widen_factor must be 1, 2, 4 or 8 (got %d)", c->widen);
  SRW_REQUIRE(c->img_size >= 8 && c->img_size % 4 == 0 && c->img_size <= 224, "srw_wrn: img_size must be a multiple of 4 in [8, 224]");
  const int n = (c->depth - 4) / 6;
  SRW_REQUIRE(3 * n <= WRN_MAX_BLOCKS, "srw_wrn: too deep");
  d.N = batch; d.C = c->num_classes; d.img = c->img_size; d.nb = 3 * n;
  d.ch[0] = 16; d.ch[1] = 16 * c->widen; d.ch[2] = 32 * c->widen; d.ch[3] = 64 * c->widen;
  for (int s = 0; s < 3; ++s) {
    d.geo[s].Hs = c->img_size >> s; d.geo[s].Wp = d.geo[s].Hs + 2; d.geo[s].Pp = d.geo[s].Wp * d.geo[s].Wp;
    d.M[s] = (int64_t)batch * d.geo[s].Pp;
    SRW_REQUIRE(d.M[s] < ((int64_t)1 << 31) - 4096, "srw_wrn: batch too large (position rows must fit 31 bits)");
  }
  d.blocks.clear();
  for (int b = 0; b < 3; ++b)
    for (int i = 0; i < n; ++i) {
      WBlock k;
      k.cin = i == 0 ? d.ch[b] : d.ch[b + 1]; k.cout = d.ch[b + 1]; k.stride = (i == 0 && b > 0) ? 2 : 1; k.abr = b == 0;
      k.stage_out = b; k.stage_in = (i == 0 && b > 0) ? b - 1 : b;
      d.blocks.push_back(k);
    }
  d.momentum = c->bn_momentum; d.slope = c->slope;
  return SRW_OK;
}

// parameter indices (WideResNet.named_parameters() order): conv1.w, conv1.b, per block {bn1.w, bn1.b, conv1.w, bn2.w, bn2.b, conv2.w[, convShortcut.w]},
// bn1.w, bn1.b, classifier.w, classifier.b.  BatchNorm buffers: per block bn1 then bn2, then the final bn1.
struct WIdx {
  std::vector<int> bn1w, conv1, bn2w, conv2, sc;   // per block (sc = -1 when cin == cout)
  int fbnw, fcw, fcb, num;
};
static WIdx make_widx(const WDims& d) {
  WIdx x;
  int i = 2;
  for (const WBlock& b : d.blocks) {
    x.bn1w.push_back(i); x.conv1.push_back(i + 2); x.bn2w.push_back(i + 3); x.conv2.push_back(i + 5);
    i += 6;
    if (b.cin != b.cout) { x.sc.push_back(i); i += 1; } else x.sc.push_back(-1);
  }
  x.fbnw = i; x.fcw = i + 2; x.fcb = i + 3; x.num = i + 4;
  return x;
}

struct WWOff { int64_t stem; std::vector<int64_t> c1f, c1d, c2f, c2d, sc; int64_t total; };
static WWOff wrn_weight_layout(const WDims& d) {
  WWOff w;
  Carver c;
  w.stem = c.take((int64_t)16 * 3 * kpad_of(8) * 4);
  for (const WBlock& b : d.blocks) {
    w.c1f.push_back(c.take((int64_t)b.cout * 3 * kpad_of(b.cin) * 4));
    w.c1d.push_back(c.take((int64_t)b.cin * 3 * kpad_of(b.cout) * 4));
    w.c2f.push_back(c.take((int64_t)b.cout * 3 * kpad_of(b.cout) * 4));
    w.c2d.push_back(c.take((int64_t)b.cout * 3 * kpad_of(b.cout) * 4));
    w.sc.push_back(b.cin != b.cout ? c.take((int64_t)b.cout * b.cin * 4) : -1);
  }
  w.total = c.off;
  return w;
}

struct WBlockBufs { int64_t o1p, c1, c1full, o2p, out, sc, xsp, mean1, rstd1, mean2, rstd2; };
struct WLayout {
  int64_t xin_p, x0;                       // input planes (8 channels), stem output fp32
  std::vector<WBlockBufs> blk;
  int64_t meanf, rstdf, feat, partial;
  // backward
  int64_t dcur[2], dplanes, dtmp, dtmp2, dsmall, dfeat, coef, wgrad_ws, dpartial;
  int64_t total;
};
static int64_t max_rows_ch(const WDims& d) {   // largest rows * channels of any activation
  int64_t m = 0;
  for (const WBlock& b : d.blocks) m = std::max<int64_t>(m, std::max(d.M[b.stage_in] * std::max(b.cin, b.stride == 2 ? b.cout : 0), d.M[b.stage_out] * b.cout));
  return std::max<int64_t>(m, d.M[0] * 16);
}
static int64_t max_plane_bytes(const WDims& d) {
  int64_t m = 0;
  for (const WBlock& b : d.blocks) {
    m = std::max<int64_t>(m, plane_rows(d, b.stage_in) * std::max(b.cin, b.stride == 2 ? b.cout : 0) * 4);
    m = std::max<int64_t>(m, plane_rows(d, b.stage_out) * b.cout * 4);
  }
  return std::max<int64_t>(m, plane_rows(d, 0) * 16 * 4);
}
static WLayout make_wlayout(const WDims& d) {
  WLayout L;
  Carver c;
  L.xin_p = c.take(plane_rows(d, 0) * 8 * 4);
  L.x0 = c.take(d.M[0] * 16 * 4);
  int maxc = 16;
  for (const WBlock& b : d.blocks) {
    WBlockBufs k = {};
    const int si = b.stage_in, so = b.stage_out;
    k.o1p = c.take(plane_rows(d, si) * b.cin * 4);            // act(bn1(x)) planes, or the raw x planes (block2 / block3 first layers)
    k.c1 = c.take(d.M[so] * b.cout * 4);
    k.c1full = b.stride == 2 ? c.take(d.M[si] * b.cout * 4) : -1;
    k.o2p = c.take(plane_rows(d, so) * b.cout * 4);
    k.out = c.take(d.M[so] * b.cout * 4);
    k.sc = b.cin != b.cout ? c.take(d.M[so] * b.cout * 4) : -1;
    k.xsp = (b.cin != b.cout && b.stride == 2) ? c.take(plane_rows(d, so) * b.cin * 4) : -1;
    k.mean1 = c.take(b.cin * 4); k.rstd1 = c.take(b.cin * 4); k.mean2 = c.take(b.cout * 4); k.rstd2 = c.take(b.cout * 4);
    L.blk.push_back(k);
    maxc = std::max(maxc, b.cout);
  }
  L.meanf = c.take(maxc * 4); L.rstdf = c.take(maxc * 4);
  L.feat = c.take((int64_t)d.N * d.ch[3] * 4);
  const int64_t nparts = cdiv64(d.M[0], BN_ROWS_PER_CTA);
  L.partial = c.take(nparts * 2 * maxc * 4);
  // backward
  const int64_t big = max_rows_ch(d) * 4;
  L.dcur[0] = c.take(big); L.dcur[1] = c.take(big); L.dtmp = c.take(big); L.dtmp2 = c.take(big); L.dsmall = c.take(big);
  L.dplanes = c.take(max_plane_bytes(d));
  L.dfeat = c.take((int64_t)d.N * d.ch[3] * 4);
  L.coef = c.take(2 * maxc * 4);
  int64_t wsmax = 0;
  for (const WBlock& b : d.blocks) {
    wsmax = std::max(wsmax, 3 * splitk_for(b.cout, 3 * b.cin, d.M[b.stage_in], nullptr));
    wsmax = std::max(wsmax, 3 * splitk_for(b.cout, 3 * b.cout, d.M[b.stage_out], nullptr));
    if (b.cin != b.cout) wsmax = std::max(wsmax, splitk_for(b.cout, b.cin, d.M[b.stage_out], nullptr));
  }
  wsmax = std::max(wsmax, 3 * splitk_for(16, 24, d.M[0], nullptr));
  L.wgrad_ws = c.take(wsmax * 4);
  L.dpartial = c.take(nparts * 2 * maxc * 4);
  L.total = c.off;
  return L;
}

// ---- host helpers -------------------------------------------------------------------------------
struct Ctx {
  const WDims& d; uint8_t* ws; const uint8_t* wp; int impl; cudaStream_t s;
  srw_allreduce_sum_fn sync_fn; void* sync_ctx; float* sync_buf; int world;
  float* F32(int64_t off) const { return reinterpret_cast<float*>(ws + off); }
  __nv_bfloat16* BF(int64_t off) const { return reinterpret_cast<__nv_bfloat16*>(ws + off); }
};
static inline int grid_for(int64_t items) { return (int)std::min<int64_t>(148 * 8, std::max<int64_t>(1, cdiv64(items, 256))); }

// zero the slack rows in front of and behind the position rows of a planes buffer (both planes)
static int zero_slack(const Ctx& k, int64_t planes_off, int st, int C) {
  const int64_t pr = plane_rows(k.d, st), fr = front_rows(k.d, st);
  for (int pl = 0; pl < 2; ++pl) {
    uint8_t* base = k.ws + planes_off + (int64_t)pl * pr * C * 2;
    SRW_CUDA(cudaMemsetAsync(base, 0, (size_t)fr * C * 2, k.s));
    SRW_CUDA(cudaMemsetAsync(base + (fr + k.d.M[st]) * C * 2, 0, (size_t)(pr - fr - k.d.M[st]) * C * 2, k.s));
  }
  return SRW_OK;
}
static inline __nv_bfloat16* pos_rows(const Ctx& k, int64_t planes_off, int st, int C) { return k.BF(planes_off) + (int64_t)front_rows(k.d, st) * C; }

// out[M rows of stage st, cout] = conv3x3(in planes [., cin]) with the operand `w` ([cout][3 * kpad(cin)]); epilogue F32 (+ bias) or RESID
static int conv3x3(const Ctx& k, int64_t in_planes_off, int st, int cin, const uint8_t* w, int cout, float* out, const float* bias, const float* resid) {
  const int kp = kpad_of(cin);
  Gemm g((int)k.d.M[st], cout, 3 * kp, k.impl);
  g.g.a = k.ws + in_planes_off; g.g.lda = cin; g.g.a_plane_stride = plane_rows(k.d, st) * cin; g.g.a_mn_major = 0;
  g.g.a_seg_k = 3 * cin; g.g.a_seg_rows = k.d.geo[st].Wp;
  g.g.b = w; g.g.ldb = 3 * kp; g.g.b_plane_stride = (int64_t)cout * 3 * kp; g.g.b_mn_major = 0;
  g.g.epilogue = resid ? SRW_EPI_RESID : SRW_EPI_F32; g.g.bias = bias; g.g.resid = resid; g.g.ldr = cout; g.g.out_f32 = out; g.g.ldo = cout;
  return g.run(k.s);
}
// dW[cout, cin, 3, 3] (+)= sum over the rows of stage st of dout[row, co] * in[view row, (dy, dx, ci)]
static int conv3x3_wgrad(const Ctx& k, int64_t dout_planes_off, int64_t in_planes_off, int st, int cin_p, int cin, int cout, float* ws, float* dW, int acc) {
  int split = 1;
  const int64_t M = k.d.M[st];
  splitk_for(cout, 3 * cin_p, M, &split);
  const int Wp = k.d.geo[st].Wp;
  for (int dy = 0; dy < 3; ++dy) {
    Gemm g(cout, 3 * cin_p, (int)M, k.impl);
    g.g.a = pos_rows(k, dout_planes_off, st, cout); g.g.lda = cout; g.g.a_plane_stride = plane_rows(k.d, st) * cout; g.g.a_mn_major = 1;
    g.g.b = k.BF(in_planes_off) + (int64_t)dy * Wp * cin_p; g.g.ldb = cin_p; g.g.b_plane_stride = plane_rows(k.d, st) * cin_p; g.g.b_mn_major = 1;
    g.g.epilogue = SRW_EPI_SPLITK; g.g.split_k = split; g.g.workspace = ws + (int64_t)dy * split * cout * 3 * cin_p;
    SRW_TRY(g.run(k.s));
  }
  wrn_conv_wgrad_finish_kernel<<<grid_for((int64_t)cout * cin * 9), 256, 0, k.s>>>(ws, split, cout, cin, cin_p, dW, acc);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
static int bn_stats(const Ctx& k, const float* x, int st, int C, float eps, int training, int repeats, float* rm, float* rv, int64_t* nbt, float* partial,
                    float* mean, float* rstd) {
  const int64_t M = k.d.M[st];
  const int nparts = (int)cdiv64(M, BN_ROWS_PER_CTA);
  if (training) {
    wrn_bn_partial_kernel<0><<<nparts, 256, 0, k.s>>>(x, nullptr, M, C, k.d.geo[st], nullptr, nullptr, nullptr, nullptr, 0.f, partial);
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  double n = (double)k.d.N * k.d.geo[st].Hs * k.d.geo[st].Hs;
  const float* sums = nullptr;
  if (training && k.sync_fn) {   // SyncBatchNorm: local sums -> all-reduce over the ranks (host-enqueued on this stream) -> global statistics
    wrn_bn_fold_kernel<<<cdiv(C, 8), 256, 0, k.s>>>(partial, nparts, C, k.sync_buf, nullptr, nullptr, 0);
    g_launches++;
    SRW_LAUNCH_CHECK();
    SRW_REQUIRE(k.sync_fn(k.sync_ctx, k.sync_buf, 2 * C) == 0, "srw_wrn: the SyncBatchNorm all-reduce callback failed");
    sums = k.sync_buf;
    n *= k.world;
  }
  wrn_bn_finish_kernel<<<cdiv(C, 8), 256, 0, k.s>>>(partial, nparts, C, n, eps, k.d.momentum, training, repeats, rm, rv, nbt, mean, rstd, sums);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
static int bn_act_planes(const Ctx& k, const float* x, int st, int C, const float* mean, const float* rstd, const float* gamma, const float* beta, int raw,
                         int64_t planes_off) {
  SRW_TRY(zero_slack(k, planes_off, st, C));
  wrn_bn_act_planes_kernel<<<grid_for(k.d.M[st] * C / 4), 256, 0, k.s>>>(x, k.d.M[st], C, k.d.geo[st], mean, rstd, gamma, beta, k.d.slope, raw,
                                                                        pos_rows(k, planes_off, st, C), plane_rows(k.d, st) * C);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
// BatchNorm + LeakyReLU backward over stage st: dy, x -> (dgamma, dbeta), dx (+ resid) as fp32 and / or planes (planes_off >= 0)
static int bn_backward(const Ctx& k, const float* dy, const float* x, int st, int C, const float* mean, const float* rstd, const float* gamma, const float* beta,
                       float* dgamma, float* dbeta, int acc, float* partial, float* coef, const float* resid, float* out_f32, int64_t planes_off) {
  const int64_t M = k.d.M[st];
  const int nparts = (int)cdiv64(M, BN_ROWS_PER_CTA);
  wrn_bn_partial_kernel<1><<<nparts, 256, 0, k.s>>>(x, dy, M, C, k.d.geo[st], mean, rstd, gamma, beta, k.d.slope, partial);
  g_launches++;
  SRW_LAUNCH_CHECK();
  double n = (double)k.d.N * k.d.geo[st].Hs * k.d.geo[st].Hs;
  if (k.sync_fn) {   // dgamma / dbeta from the LOCAL sums (DDP averages them with the other gradients), dx from the global ones
    wrn_bn_fold_kernel<<<cdiv(C, 8), 256, 0, k.s>>>(partial, nparts, C, k.sync_buf, dgamma, dbeta, acc);
    g_launches++;
    SRW_LAUNCH_CHECK();
    SRW_REQUIRE(k.sync_fn(k.sync_ctx, k.sync_buf, 2 * C) == 0, "srw_wrn: the SyncBatchNorm all-reduce callback failed");
    n *= k.world;
    wrn_bn_bwd_finish_kernel<<<cdiv(C, 8), 256, 0, k.s>>>(partial, nparts, C, n, nullptr, nullptr, 0, coef, k.sync_buf);
  } else {
    wrn_bn_bwd_finish_kernel<<<cdiv(C, 8), 256, 0, k.s>>>(partial, nparts, C, n, dgamma, dbeta, acc, coef, nullptr);
  }
  g_launches++;
  SRW_LAUNCH_CHECK();
  if (planes_off >= 0) SRW_TRY(zero_slack(k, planes_off, st, C));
  wrn_bn_bwd_apply_kernel<<<grid_for(M * C / 4), 256, 0, k.s>>>(dy, x, M, C, k.d.geo[st], mean, rstd, gamma, beta, k.d.slope, coef, 0, resid, out_f32,
                                                                planes_off >= 0 ? pos_rows(k, planes_off, st, C) : nullptr, plane_rows(k.d, st) * C);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
// planes (zero border, zero slack) of an fp32 gradient tensor
static int grad_planes(const Ctx& k, const float* dy, int st, int C, int64_t planes_off) {
  SRW_TRY(zero_slack(k, planes_off, st, C));
  wrn_bn_bwd_apply_kernel<<<grid_for(k.d.M[st] * C / 4), 256, 0, k.s>>>(dy, nullptr, k.d.M[st], C, k.d.geo[st], nullptr, nullptr, nullptr, nullptr, 0.f, nullptr, 1,
                                                                        nullptr, nullptr, pos_rows(k, planes_off, st, C), plane_rows(k.d, st) * C);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

}  // namespace srw

using namespace srw;

extern "C" int64_t srw_wrn_weight_planes_bytes(const srw_wrn_config* c) {
  WDims d;
  if (make_wdims(c, 1, d)) return -1;
  return wrn_weight_layout(d).total;
}
extern "C" int64_t srw_wrn_workspace_bytes(const srw_wrn_config* c, int batch) {
  WDims d;
  if (make_wdims(c, batch, d)) return -1;
  return make_wlayout(d).total;
}
extern "C" int srw_wrn_num_params(const srw_wrn_config* c) {
  WDims d;
  if (make_wdims(c, 1, d)) return -1;
  return make_widx(d).num;
}

extern "C" int srw_wrn_prepare_weights(const srw_wrn_config* c, const float* const* P, void* weight_planes, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(c && P && weight_planes, "srw_wrn_prepare_weights: null pointer");
  WDims d;
  SRW_TRY(make_wdims(c, 1, d));
  const WWOff w = wrn_weight_layout(d);
  const WIdx ix = make_widx(d);
  uint8_t* base = reinterpret_cast<uint8_t*>(weight_planes);
  SRW_CUDA(cudaMemsetAsync(base, 0, (size_t)w.total, s));   // the K padding of every operand must be zero
  auto BF = [&](int64_t off) { return reinterpret_cast<__nv_bfloat16*>(base + off); };
  auto relayout = [&](const float* wt, int cout, int cin, int cin_p, int64_t f_off, int64_t d_off) -> int {
    const int kp = kpad_of(cin_p), kpo = kpad_of(cout);
    wrn_conv_relayout_kernel<<<grid_for((int64_t)cout * cin * 9), 256, 0, s>>>(wt, cout, cin, cin_p, kp, kpo, BF(f_off), (int64_t)cout * 3 * kp,
                                                                              d_off >= 0 ? BF(d_off) : nullptr, (int64_t)cin_p * 3 * kpo);
    g_launches++;
    SRW_LAUNCH_CHECK();
    return SRW_OK;
  };
  SRW_TRY(relayout(P[0], 16, 3, 8, w.stem, -1));
  for (int b = 0; b < d.nb; ++b) {
    const WBlock& k = d.blocks[b];
    SRW_TRY(relayout(P[ix.conv1[b]], k.cout, k.cin, k.cin, w.c1f[b], w.c1d[b]));
    SRW_TRY(relayout(P[ix.conv2[b]], k.cout, k.cout, k.cout, w.c2f[b], w.c2d[b]));
    if (ix.sc[b] >= 0) SRW_TRY(split_to(P[ix.sc[b]], k.cin, k.cout, k.cin, base + w.sc[b], k.cin, nullptr, 1, s));
  }
  return SRW_OK;
}

static int wrn_forward_body(const srw_wrn_fwd_args* a, cudaStream_t s) {
  SRW_REQUIRE(a && a->cfg && a->params && a->bn_running_mean && a->bn_running_var && a->weight_planes && a->x && a->logits && a->feat && a->workspace,
              "srw_wrn_forward: null pointer");
  WDims d;
  SRW_TRY(make_wdims(a->cfg, a->batch, d));
  const WLayout L = make_wlayout(d);
  SRW_REQUIRE(a->workspace_bytes >= L.total, "srw_wrn_forward: workspace too small (%lld < %lld)", (long long)a->workspace_bytes, (long long)L.total);
  SRW_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 1023) == 0, "srw_wrn_forward: workspace must be 1024-byte aligned");
  SRW_REQUIRE(a->stat_repeats >= 0, "srw_wrn_forward: stat_repeats < 0");
  const WWOff w = wrn_weight_layout(d);
  const WIdx ix = make_widx(d);
  SRW_REQUIRE(!a->sync_fn || (a->sync_buf && a->world_size >= 1), "srw_wrn_forward: sync_fn needs sync_buf and world_size");
  const Ctx k = {d, reinterpret_cast<uint8_t*>(a->workspace), reinterpret_cast<const uint8_t*>(a->weight_planes), a->gemm_impl, s,
                 a->sync_fn, a->sync_ctx, a->sync_buf, a->world_size};
  const float* const* P = a->params;
  const int tr = a->training ? 1 : 0, rep = a->stat_repeats;
  auto NBT = [&](int i) { return a->bn_num_batches_tracked ? a->bn_num_batches_tracked[i] : nullptr; };

  // ---- stem: conv1 (3 -> 16, bias) ----
  SRW_TRY(zero_slack(k, L.xin_p, 0, 8));
  wrn_input_planes_kernel<<<grid_for(d.M[0]), 256, 0, s>>>(a->x, d.N, d.geo[0], pos_rows(k, L.xin_p, 0, 8), plane_rows(d, 0) * 8);
  g_launches++;
  SRW_LAUNCH_CHECK();
  SRW_TRY(conv3x3(k, L.xin_p, 0, 8, k.wp + w.stem, 16, k.F32(L.x0), P[1], nullptr));
  int64_t x_off = L.x0;
  // ---- blocks ----
  for (int b = 0; b < d.nb; ++b) {
    const WBlock& bk = d.blocks[b];
    const WBlockBufs& B = L.blk[b];
    const int si = bk.stage_in, so = bk.stage_out;
    const bool equal = bk.cin == bk.cout;
    const float* x = k.F32(x_off);
    // bn1 statistics always advance (wrn.py:46-50: the module runs even when its output is dropped)
    SRW_TRY(bn_stats(k, x, si, bk.cin, 1e-5f, tr, rep, a->bn_running_mean[2 * b], a->bn_running_var[2 * b], NBT(2 * b), k.F32(L.partial), k.F32(B.mean1),
                     k.F32(B.rstd1)));
    const bool raw = !equal && !bk.abr;
    SRW_TRY(bn_act_planes(k, x, si, bk.cin, k.F32(B.mean1), k.F32(B.rstd1), P[ix.bn1w[b]], P[ix.bn1w[b] + 1], raw ? 1 : 0, B.o1p));
    // conv1
    if (bk.stride == 1) {
      SRW_TRY(conv3x3(k, B.o1p, si, bk.cin, k.wp + w.c1f[b], bk.cout, k.F32(B.c1), nullptr, nullptr));
    } else {
      SRW_TRY(conv3x3(k, B.o1p, si, bk.cin, k.wp + w.c1f[b], bk.cout, k.F32(B.c1full), nullptr, nullptr));
      wrn_subsample_kernel<<<grid_for(d.M[so] * bk.cout / 4), 256, 0, s>>>(k.F32(B.c1full), nullptr, 0, d.N, bk.cout, d.geo[si], d.geo[so], k.F32(B.c1), nullptr, 0);
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    SRW_TRY(bn_stats(k, k.F32(B.c1), so, bk.cout, 1e-5f, tr, rep, a->bn_running_mean[2 * b + 1], a->bn_running_var[2 * b + 1], NBT(2 * b + 1), k.F32(L.partial),
                     k.F32(B.mean2), k.F32(B.rstd2)));
    SRW_TRY(bn_act_planes(k, k.F32(B.c1), so, bk.cout, k.F32(B.mean2), k.F32(B.rstd2), P[ix.bn2w[b]], P[ix.bn2w[b] + 1], 0, B.o2p));
    // shortcut
    const float* sc = x;
    if (!equal) {
      const __nv_bfloat16* ain = pos_rows(k, B.o1p, si, bk.cin);
      int64_t a_ps = plane_rows(d, si) * bk.cin;
      if (bk.stride == 2) {
        SRW_TRY(zero_slack(k, B.xsp, so, bk.cin));
        wrn_subsample_kernel<<<grid_for(d.M[so] * bk.cin / 4), 256, 0, s>>>(nullptr, pos_rows(k, B.o1p, si, bk.cin), plane_rows(d, si) * bk.cin, d.N, bk.cin,
                                                                          d.geo[si], d.geo[so], nullptr, pos_rows(k, B.xsp, so, bk.cin), plane_rows(d, so) * bk.cin);
        g_launches++;
        SRW_LAUNCH_CHECK();
        ain = pos_rows(k, B.xsp, so, bk.cin); a_ps = plane_rows(d, so) * bk.cin;
      }
      Gemm g((int)d.M[so], bk.cout, bk.cin, k.impl);
      g.g.a = ain; g.g.lda = bk.cin; g.g.a_plane_stride = a_ps; g.g.a_mn_major = 0;
      g.g.b = k.wp + w.sc[b]; g.g.ldb = bk.cin; g.g.b_plane_stride = (int64_t)bk.cout * bk.cin; g.g.b_mn_major = 0;
      g.g.epilogue = SRW_EPI_F32; g.g.out_f32 = k.F32(B.sc); g.g.ldo = bk.cout;
      SRW_TRY(g.run(s));
      sc = k.F32(B.sc);
    }
    // conv2 + residual
    SRW_TRY(conv3x3(k, B.o2p, so, bk.cout, k.wp + w.c2f[b], bk.cout, k.F32(B.out), nullptr, sc));
    x_off = B.out;
  }
  // ---- final BatchNorm (eps 1e-3) + LeakyReLU + average pool + classifier ----
  const int Cf = d.ch[3];
  SRW_TRY(bn_stats(k, k.F32(x_off), 2, Cf, 1e-3f, tr, rep, a->bn_running_mean[2 * d.nb], a->bn_running_var[2 * d.nb], NBT(2 * d.nb), k.F32(L.partial),
                   k.F32(L.meanf), k.F32(L.rstdf)));
  wrn_pool_fwd_kernel<<<d.N, Cf, 0, s>>>(k.F32(x_off), Cf, d.geo[2], k.F32(L.meanf), k.F32(L.rstdf), P[ix.fbnw], P[ix.fbnw + 1], d.slope, k.F32(L.feat));
  g_launches++;
  SRW_LAUNCH_CHECK();
  wrn_fc_fwd_kernel<<<d.N, 128, Cf * sizeof(float), s>>>(k.F32(L.feat), d.N, Cf, d.C, P[ix.fcw], P[ix.fcb], a->logits);
  g_launches++;
  SRW_LAUNCH_CHECK();
  SRW_CUDA(cudaMemcpyAsync(a->feat, k.F32(L.feat), (size_t)d.N * Cf * 4, cudaMemcpyDeviceToDevice, s));
  return SRW_OK;
}

static int wrn_backward_body(const srw_wrn_bwd_args* a, cudaStream_t s) {
  SRW_REQUIRE(a && a->cfg && a->params && a->weight_planes && a->dlogits && a->grads && a->workspace, "srw_wrn_backward: null pointer");
  WDims d;
  SRW_TRY(make_wdims(a->cfg, a->batch, d));
  SRW_REQUIRE(a->grad_rows > 0 && a->grad_rows <= a->batch, "srw_wrn_backward: 0 < grad_rows <= batch");
  const WLayout L = make_wlayout(d);
  SRW_REQUIRE(a->workspace_bytes >= L.total, "srw_wrn_backward: workspace too small");
  const WWOff w = wrn_weight_layout(d);
  const WIdx ix = make_widx(d);
  SRW_REQUIRE(!a->sync_fn || (a->sync_buf && a->world_size >= 1), "srw_wrn_backward: sync_fn needs sync_buf and world_size");
  const Ctx k = {d, reinterpret_cast<uint8_t*>(a->workspace), reinterpret_cast<const uint8_t*>(a->weight_planes), a->gemm_impl, s,
                 a->sync_fn, a->sync_ctx, a->sync_buf, a->world_size};
  const float* const* P = a->params;
  float* const* G = a->grads;
  const int acc = a->accumulate_grads ? 1 : 0;
  const int Cf = d.ch[3];
  float* partial = k.F32(L.dpartial);
  float* coef = k.F32(L.coef);

  // ---- classifier, pool, final BatchNorm ----
  wrn_fc_bwd_rows_kernel<<<d.N, 128, 0, s>>>(a->dlogits, a->grad_rows, a->dfeat, Cf, d.C, P[ix.fcw], k.F32(L.dfeat));
  g_launches++;
  SRW_LAUNCH_CHECK();
  wrn_fc_bwd_params_kernel<<<cdiv(d.C * Cf + d.C, 256), 256, 0, s>>>(a->dlogits, k.F32(L.feat), a->grad_rows, Cf, d.C, G[ix.fcw], G[ix.fcb], acc);
  g_launches++;
  SRW_LAUNCH_CHECK();
  wrn_pool_bwd_kernel<<<grid_for(d.M[2] * Cf / 4), 256, 0, s>>>(k.F32(L.dfeat), d.M[2], Cf, d.geo[2], k.F32(L.dtmp));
  g_launches++;
  SRW_LAUNCH_CHECK();
  int cur = 0;
  const int64_t last_out = L.blk[d.nb - 1].out;
  SRW_TRY(bn_backward(k, k.F32(L.dtmp), k.F32(last_out), 2, Cf, k.F32(L.meanf), k.F32(L.rstdf), P[ix.fbnw], P[ix.fbnw + 1], G[ix.fbnw], G[ix.fbnw + 1], acc, partial,
                      coef, nullptr, k.F32(L.dcur[cur]), -1));
  // ---- blocks, last to first ----
  for (int b = d.nb - 1; b >= 0; --b) {
    const WBlock& bk = d.blocks[b];
    const WBlockBufs& B = L.blk[b];
    const int si = bk.stage_in, so = bk.stage_out;
    const bool equal = bk.cin == bk.cout;
    const int64_t x_off = b == 0 ? L.x0 : L.blk[b - 1].out;
    float* dout = k.F32(L.dcur[cur]);            // gradient of the block output [M_out, cout]
    float* dx = k.F32(L.dcur[cur ^ 1]);          // gradient of the block input  [M_in, cin]
    // conv2: weight gradient, input gradient
    SRW_TRY(grad_planes(k, dout, so, bk.cout, L.dplanes));
    SRW_TRY(conv3x3_wgrad(k, L.dplanes, B.o2p, so, bk.cout, bk.cout, bk.cout, k.F32(L.wgrad_ws), G[ix.conv2[b]], acc));
    SRW_TRY(conv3x3(k, L.dplanes, so, bk.cout, k.wp + w.c2d[b], bk.cout, k.F32(L.dtmp), nullptr, nullptr));          // d o2
    // shortcut convolution (reads the same planes of dout)
    if (!equal) {
      const int64_t in_p = bk.stride == 2 ? B.xsp : B.o1p;
      const int st_in = bk.stride == 2 ? so : si;   // the shortcut's GEMM rows live at the output resolution
      SRW_TRY(wgrad(bk.cout, bk.cin, d.M[so], pos_rows(k, L.dplanes, so, bk.cout), bk.cout, plane_rows(d, so), pos_rows(k, in_p, st_in, bk.cin), bk.cin,
                    plane_rows(d, st_in), k.F32(L.wgrad_ws), G[ix.sc[b]], bk.cin, acc, k.impl, s));
      Gemm g((int)d.M[so], bk.cin, bk.cout, k.impl);   // d(shortcut input) = dout Wsc
      g.g.a = pos_rows(k, L.dplanes, so, bk.cout); g.g.lda = bk.cout; g.g.a_plane_stride = plane_rows(d, so) * bk.cout; g.g.a_mn_major = 0;
      g.g.b = k.wp + w.sc[b]; g.g.ldb = bk.cin; g.g.b_plane_stride = (int64_t)bk.cout * bk.cin; g.g.b_mn_major = 1;
      g.g.epilogue = SRW_EPI_F32; g.g.out_f32 = k.F32(L.dsmall); g.g.ldo = bk.cin;
      SRW_TRY(g.run(s));
    }
    // bn2 + LeakyReLU backward -> d c1 (planes; at full resolution with zeros between the samples for a stride-2 conv1)
    if (bk.stride == 1) {
      SRW_TRY(bn_backward(k, k.F32(L.dtmp), k.F32(B.c1), so, bk.cout, k.F32(B.mean2), k.F32(B.rstd2), P[ix.bn2w[b]], P[ix.bn2w[b] + 1], G[ix.bn2w[b]],
                          G[ix.bn2w[b] + 1], acc, partial, coef, nullptr, nullptr, L.dplanes));
    } else {
      SRW_TRY(bn_backward(k, k.F32(L.dtmp), k.F32(B.c1), so, bk.cout, k.F32(B.mean2), k.F32(B.rstd2), P[ix.bn2w[b]], P[ix.bn2w[b] + 1], G[ix.bn2w[b]],
                          G[ix.bn2w[b] + 1], acc, partial, coef, nullptr, k.F32(L.dtmp2), -1));
      SRW_TRY(zero_slack(k, L.dplanes, si, bk.cout));
      wrn_upsample_planes_kernel<<<grid_for(d.M[si] * bk.cout / 4), 256, 0, s>>>(k.F32(L.dtmp2), d.N, bk.cout, d.geo[si], d.geo[so], pos_rows(k, L.dplanes, si, bk.cout),
                                                                                plane_rows(d, si) * bk.cout);
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    // conv1: weight gradient (operand = o1p: act(bn1(x)) or raw x), input gradient at the input resolution
    SRW_TRY(conv3x3_wgrad(k, L.dplanes, B.o1p, si, bk.cin, bk.cin, bk.cout, k.F32(L.wgrad_ws), G[ix.conv1[b]], acc));
    if (equal) {
      SRW_TRY(conv3x3(k, L.dplanes, si, bk.cout, k.wp + w.c1d[b], bk.cin, k.F32(L.dtmp), nullptr, nullptr));         // d o1
      // x feeds bn1 (-> conv1) and the identity shortcut: dx = dout + bn1_backward(d o1)
      SRW_TRY(bn_backward(k, k.F32(L.dtmp), k.F32(x_off), si, bk.cin, k.F32(B.mean1), k.F32(B.rstd1), P[ix.bn1w[b]], P[ix.bn1w[b] + 1], G[ix.bn1w[b]],
                          G[ix.bn1w[b] + 1], acc, partial, coef, dout, dx, -1));
    } else if (bk.abr) {
      // x -> bn1 -> act = xin feeds conv1 AND the shortcut conv (same resolution): d xin = dgrad(conv1) + dsmall
      SRW_TRY(conv3x3(k, L.dplanes, si, bk.cout, k.wp + w.c1d[b], bk.cin, k.F32(L.dtmp), nullptr, k.F32(L.dsmall)));
      SRW_TRY(bn_backward(k, k.F32(L.dtmp), k.F32(x_off), si, bk.cin, k.F32(B.mean1), k.F32(B.rstd1), P[ix.bn1w[b]], P[ix.bn1w[b] + 1], G[ix.bn1w[b]],
                          G[ix.bn1w[b] + 1], acc, partial, coef, nullptr, dx, -1));
    } else {
      // raw x feeds conv1 (stride 2) and the stride-2 shortcut conv; bn1's output is unused: no gradient for its parameters
      SRW_TRY(conv3x3(k, L.dplanes, si, bk.cout, k.wp + w.c1d[b], bk.cin, dx, nullptr, nullptr));
      wrn_add_upsampled_kernel<<<grid_for(d.M[so] * bk.cin / 4), 256, 0, s>>>(k.F32(L.dsmall), d.N, bk.cin, d.geo[si], d.geo[so], dx);
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    cur ^= 1;
  }
  // ---- stem ----
  {
    float* dout = k.F32(L.dcur[cur]);
    SRW_TRY(grad_planes(k, dout, 0, 16, L.dplanes));
    SRW_TRY(conv3x3_wgrad(k, L.dplanes, L.xin_p, 0, 8, 3, 16, k.F32(L.wgrad_ws), G[0], acc));
    // bias gradient = column sums over the valid rows
    const int nparts = (int)cdiv64(d.M[0], BN_ROWS_PER_CTA);
    wrn_bn_partial_kernel<0><<<nparts, 256, 0, s>>>(dout, nullptr, d.M[0], 16, d.geo[0], nullptr, nullptr, nullptr, nullptr, 0.f, partial);
    g_launches++;
    SRW_LAUNCH_CHECK();
    wrn_bn_bwd_finish_kernel<<<2, 256, 0, s>>>(partial, nparts, 16, 1.0, nullptr, nullptr, 0, coef, nullptr);   // coef[0 .. 15] = column sums (n = 1)
    g_launches++;
    SRW_LAUNCH_CHECK();
    if (acc) {
      srw_splitk_reduce_args r = {};
      r.workspace = coef; r.split_k = 1; r.M = 1; r.N = 16; r.out = G[1]; r.ldo = 16; r.accumulate = 1;
      SRW_TRY(srw_splitk_reduce(&r, s));
    } else {
      SRW_CUDA(cudaMemcpyAsync(G[1], coef, 16 * 4, cudaMemcpyDeviceToDevice, s));
    }
  }
  return SRW_OK;
}

extern "C" int srw_wrn_forward(const srw_wrn_fwd_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->cfg && a->params, "srw_wrn_forward: null pointer");
  if (!graphs_enabled(s) || a->sync_fn) return wrn_forward_body(a, s);
  const int np = srw_wrn_num_params(a->cfg);
  SRW_REQUIRE(np > 0, "srw_wrn_forward: bad config");
  const int nbn = 2 * ((a->cfg->depth - 4) / 6) * 3 + 1;
  KeyBuilder kb;
  kb.add((int)31); kb.add(*a->cfg); kb.add(s);
  for (int i = 0; i < np; ++i) kb.add(a->params[i]);
  for (int i = 0; i < nbn; ++i) { kb.add(a->bn_running_mean[i]); kb.add(a->bn_running_var[i]); kb.add(a->bn_num_batches_tracked ? a->bn_num_batches_tracked[i] : nullptr); }
  kb.add(a->weight_planes); kb.add(a->x); kb.add(a->batch); kb.add(a->training); kb.add(a->stat_repeats); kb.add(a->logits); kb.add(a->feat);
  kb.add(a->workspace); kb.add(a->workspace_bytes); kb.add(a->gemm_impl);
  return run_graphed(std::move(kb.k), s, [a](cudaStream_t st) { return wrn_forward_body(a, st); });
}

extern "C" int srw_wrn_backward(const srw_wrn_bwd_args* a, void* stream_) {
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->cfg && a->params && a->grads, "srw_wrn_backward: null pointer");
  if (!graphs_enabled(s) || a->sync_fn) return wrn_backward_body(a, s);
  const int np = srw_wrn_num_params(a->cfg);
  SRW_REQUIRE(np > 0, "srw_wrn_backward: bad config");
  KeyBuilder kb;
  kb.add((int)32); kb.add(*a->cfg); kb.add(s);
  for (int i = 0; i < np; ++i) { kb.add(a->params[i]); kb.add(a->grads[i]); }
  kb.add(a->weight_planes); kb.add(a->batch); kb.add(a->grad_rows); kb.add(a->dlogits); kb.add(a->dfeat); kb.add(a->accumulate_grads); kb.add(a->workspace);
  kb.add(a->workspace_bytes); kb.add(a->gemm_impl);
  return run_graphed(std::move(kb.k), s, [a](cudaStream_t st) { return wrn_backward_body(a, st); });
}
