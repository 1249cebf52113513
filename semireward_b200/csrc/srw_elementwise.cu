// srw_elementwise — HBM-bound kernels around the GEMMs: split-plane conversion, split-K reduce, column sums (bias
// gradients), LayerNorm forward/backward (vit.py:164-165,282), patch embedding (vit.py:39-44, 279-280) and the
// final-norm + head (vit.py:282, 298, 304) with their backward.  All fp32 math, coalesced row-major accesses,
// deterministic two-stage reductions (no float atomics).
#include <atomic>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {
extern std::atomic<int64_t> g_launches;

// ------------------------------------------------------------------------------------------------
// split planes (+ optional transposed copy)
// ------------------------------------------------------------------------------------------------
__global__ void split_planes_kernel(const float* __restrict__ x, int64_t ldx, int rows, int cols,
                                    const float* __restrict__ row_scale, int rows_per_scale,
                                    __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t ps,
                                    __nv_bfloat16* __restrict__ planes_t, int64_t ldpt, int64_t pst) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = x[(int64_t)r * ldx + c];
      if (row_scale) v *= row_scale[r / rows_per_scale];
      if (planes) {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        planes[(int64_t)r * ldp + c] = h;
        planes[(int64_t)r * ldp + c + ps] = l;
      }
    }
    tile[i][threadIdx.x] = v;
  }
  if (planes_t == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;  // output row = c, output col = r
    if (c < cols && r < rows) {
      __nv_bfloat16 h, l;
      split_bf16(tile[threadIdx.x][i], h, l);
      planes_t[(int64_t)c * ldpt + r] = h;
      planes_t[(int64_t)c * ldpt + r + pst] = l;
    }
  }
}

// Fast path (no transposed copy): planes = split(scale(r) * x), optionally with per-block column partial sums
// (the bias gradient of the layer whose output gradient this is).  128 columns per block-column (float4 per lane),
// 8 warps stride the rows of the block's row range.
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, int64_t ldx, int rows, int cols,
                                                         const float* __restrict__ row_scale, int rows_per_scale,
                                                         __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t ps,
                                                         float* __restrict__ partial /* [gridDim.y, cols] or NULL */) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + lane * 4;
  const int rows_per_block = (rows + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < cols) {
    for (int r = r_begin + warp; r < r_end; r += 8) {
      float4 v = *reinterpret_cast<const float4*>(x + (int64_t)r * ldx + c);
      if (row_scale) {
        const float sc = row_scale[r / rows_per_scale];
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      }
      uint32_t h0, l0, h1, l1;
      split2(v.x, v.y, h0, l0);
      split2(v.z, v.w, h1, l1);
      __nv_bfloat16* hp = planes + (int64_t)r * ldp + c;
      *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(hp + ps) = make_uint2(l0, l1);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (partial == nullptr) return;
  red[warp][lane] = acc;
  __syncthreads();
  if (warp == 0 && c < cols) {
    float4 s = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) { s.x += red[w][lane].x; s.y += red[w][lane].y; s.z += red[w][lane].z; s.w += red[w][lane].w; }
    *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.y * cols + c) = s;
  }
}

// column partial sums of split planes: 8 bf16 (one uint4 per plane) per lane, 256 columns per block-column
__global__ void __launch_bounds__(256) colsum_planes_kernel(const __nv_bfloat16* __restrict__ planes, int64_t ldp, int64_t ps, int rows, int cols,
                                                            float* __restrict__ partial /* [gridDim.y, cols] */) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][32][9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int rows_per_block = (rows + gridDim.y - 1) / gridDim.y;
  const int r_begin = blockIdx.y * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (c < cols) {
#pragma unroll 4   // eight 16-byte loads in flight per lane; the sums keep their row order
    for (int r = r_begin + warp; r < r_end; r += 8) {
      const uint4 h = *reinterpret_cast<const uint4*>(planes + (int64_t)r * ldp + c);
      const uint4 l = *reinterpret_cast<const uint4*>(planes + (int64_t)r * ldp + c + ps);
      const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[2 * j] += bf16_lo_f(hh[j]) + bf16_lo_f(ll[j]);
        acc[2 * j + 1] += bf16_hi_f(hh[j]) + bf16_hi_f(ll[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[warp][lane][j] = acc[j];
  __syncthreads();
  if (warp == 0 && c < cols) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][lane][j];
      partial[(int64_t)blockIdx.y * cols + c + j] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// split-K reduce
// ------------------------------------------------------------------------------------------------
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int split, int M, int N, float* __restrict__ out, int64_t ldo,
                                     int accumulate) {
  pdl_trigger();
  pdl_wait();
  const int64_t total4 = (int64_t)M * N / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 4;
    const int m = (int)(e / N), n = (int)(e % N);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int s = 0; s < split; ++s) {
      const float4 v = *reinterpret_cast<const float4*>(ws + (int64_t)s * M * N + e);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    float4* dst = reinterpret_cast<float4*>(out + (int64_t)m * ldo + n);
    if (accumulate) {
      const float4 o = *dst;
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    *dst = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// column sums: out[c] (+)= sum_r scale(r) * x[r,c]
// ------------------------------------------------------------------------------------------------
__global__ void colsum_stage1_kernel(const float* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ planes,
                                     int64_t ldp, int64_t ps, const float* __restrict__ row_scale, int rows_per_scale, int rows,
                                     int cols, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (c < cols) {
    for (int r = blockIdx.y * 8 + ty; r < rows; r += 8 * gridDim.y) {
      float v = x ? x[(int64_t)r * ldx + c] : plane_value(planes, ps, (int64_t)r * ldp + c);
      if (row_scale) v *= row_scale[r / rows_per_scale];
      acc += v;
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    partial[(int64_t)blockIdx.y * cols + c] = s;
  }
}
// out_y[c] (+)= sum_p partial[p * stride_p + y * stride_y + c]  for y = blockIdx.y; 8 warps split the partials, then a
// shared-memory fold in fixed order (deterministic).  Used by colsum (y = 0) and LayerNorm dgamma / dbeta (y = 0, 1).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int nparts, int64_t stride_p, int64_t stride_y,
                                                              int cols, float* __restrict__ out0, float* __restrict__ out1, int accumulate,
                                                              float* __restrict__ out2 = nullptr, int accumulate2 = 0) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const float* src = partial + blockIdx.y * stride_y;
  float acc = 0.f;
  if (c < cols) {
#pragma unroll 4
    for (int p = ty; p < nparts; p += 8) acc += src[(int64_t)p * stride_p + c];
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][tx];
    float* out = blockIdx.y == 0 ? out0 : (blockIdx.y == 1 ? out1 : out2);
    const int acc_flag = blockIdx.y == 2 ? accumulate2 : accumulate;
    out[c] = acc_flag ? out[c] + s : s;
  }
}

// ------------------------------------------------------------------------------------------------
// grad fold: every deferred reduction of one transformer block's backward in one launch
// ------------------------------------------------------------------------------------------------
struct FoldParams {
  int n_splitk, n_colsum;
  int first_block[13];   // blocks [first_block[i], first_block[i+1]) work on problem i (split-K problems first)
  srw_splitk_reduce_args sk[4];
  srw_fold_colsum cs[8];
};

__global__ void __launch_bounds__(256) grad_fold_kernel(const __grid_constant__ FoldParams fp) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[8][33];
  const int nprob = fp.n_splitk + fp.n_colsum;
  int prob = 0;
  while (prob + 1 < nprob && (int)blockIdx.x >= fp.first_block[prob + 1]) ++prob;
  const int local = blockIdx.x - fp.first_block[prob], nblk = fp.first_block[prob + 1] - fp.first_block[prob];
  if (prob < fp.n_splitk) {
    // same order as splitk_reduce_kernel: slices 0 .. split-1, then the previous value
    const srw_splitk_reduce_args& a = fp.sk[prob];
    const int64_t MN = (int64_t)a.M * a.N, total4 = MN / 4;
    for (int64_t i = local * 256 + threadIdx.x; i < total4; i += (int64_t)nblk * 256) {
      const int64_t e = i * 4;
      const int m = (int)(e / a.N), n = (int)(e % a.N);
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4   // four slices' loads in flight, added in slice order
      for (int s = 0; s < a.split_k; ++s) {
        const float4 v = *reinterpret_cast<const float4*>(a.workspace + (int64_t)s * MN + e);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      float4* dst = reinterpret_cast<float4*>(a.out + (int64_t)m * a.ldo + n);
      if (a.accumulate) {
        const float4 o = *dst;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      *dst = acc;
    }
  } else {
    // same order as reduce_partials_kernel: 8 warps split the partials, then a fixed-order fold
    const srw_fold_colsum& a = fp.cs[prob - fp.n_splitk];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = local * 32 + tx;
    float acc = 0.f;
    if (c < a.cols) {
#pragma unroll 4
      for (int p = ty; p < a.nparts; p += 8) acc += a.partial[(int64_t)p * a.stride_p + c];
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < a.cols) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += red[i][tx];
      a.out[c] = a.accumulate ? a.out[c] + s : s;
    }
  }
}

__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int nparts, int cols, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int p = 0; p < nparts; ++p) s += partial[(int64_t)p * cols + c];
  out[c] = accumulate ? out[c] + s : s;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm forward: one warp per row
// ------------------------------------------------------------------------------------------------
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, int64_t ldx, int rows, int cols, float eps,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, __nv_bfloat16* __restrict__ yp, int64_t ldp, int64_t ps,
                                     float* __restrict__ yf, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * ldx;
  float s = 0.f;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.f;
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float a = v.x - mean, b = v.y - mean, cc = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (cc * cc + d * d);
  }
  const float var = warp_sum(q) / (float)cols;
  const float rstd = rsqrtf(var + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  for (int c = lane * 4; c < cols; c += 128) {
    const float4 v = *reinterpret_cast<const float4*>(xr + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 y;
    y.x = (v.x - mean) * rstd * g.x + b.x;
    y.y = (v.y - mean) * rstd * g.y + b.y;
    y.z = (v.z - mean) * rstd * g.z + b.z;
    y.w = (v.w - mean) * rstd * g.w + b.w;
    if (yf) *reinterpret_cast<float4*>(yf + (int64_t)row * ldy + c) = y;
    if (yp) {
      uint32_t h0, l0, h1, l1;
      split2(y.x, y.y, h0, l0);
      split2(y.z, y.w, h1, l1);
      __nv_bfloat16* hp = yp + (int64_t)row * ldp + c;
      *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(hp + ps) = make_uint2(l0, l1);
    }
  }
}

// Register-resident variant for cols == J * 128: the row is read from HBM exactly once.
template <int J>
__global__ void __launch_bounds__(256) layernorm_fwd_reg_kernel(const float* __restrict__ x, int64_t ldx, int rows, float eps,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                __nv_bfloat16* __restrict__ yp, int64_t ldp, int64_t ps, float* __restrict__ yf,
                                                                int64_t ldy) {
  pdl_trigger();
  pdl_wait();
  constexpr int COLS = J * 128;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + (int64_t)row * ldx;
  float4 v[J];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    v[j] = *reinterpret_cast<const float4*>(xr + j * 128 + lane * 4);
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  const float mean = warp_sum(s) * (1.0f / COLS);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / COLS) + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int c = j * 128 + lane * 4;
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    float4 y;
    y.x = v[j].x * rstd * g.x + b.x;
    y.y = v[j].y * rstd * g.y + b.y;
    y.z = v[j].z * rstd * g.z + b.z;
    y.w = v[j].w * rstd * g.w + b.w;
    if (yf) *reinterpret_cast<float4*>(yf + (int64_t)row * ldy + c) = y;
    if (yp) {
      uint32_t h0, l0, h1, l1;
      split2(y.x, y.y, h0, l0);
      split2(y.z, y.w, h1, l1);
      __nv_bfloat16* hp = yp + (int64_t)row * ldp + c;
      *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(hp + ps) = make_uint2(l0, l1);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  Block = 8 warps, each warp walks rows (stride 8*gridDim.x); per-lane column partials of
// dgamma/dbeta stay in registers (J = cols/32 columns per lane, column = lane + 32*j -> coalesced), reduced across
// the block through shared memory into partial[blockIdx.x][2][cols]; a second kernel folds the partials.
// ------------------------------------------------------------------------------------------------
template <int J>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x,
                                                            int64_t ldx, int rows, const float* __restrict__ gamma,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            float* __restrict__ dx, int64_t lddx, int accumulate_dx,
                                                            float* __restrict__ partial, __nv_bfloat16* __restrict__ dxp, int64_t ldp, int64_t ps,
                                                            const float* __restrict__ row_scale, int rows_per_scale) {
  pdl_trigger();
  pdl_wait();
  constexpr int COLS = J * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float pg[J], pb[J], gm[J], pc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) { pg[j] = 0.f; pb[j] = 0.f; pc[j] = 0.f; gm[j] = gamma[lane + 32 * j]; }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += 8 * gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    const float* dyr = dy + (int64_t)row * lddy;
    const float* xr = x + (int64_t)row * ldx;
    float g[J], xh[J];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float d = dyr[lane + 32 * j];
      xh[j] = (xr[lane + 32 * j] - mu) * rs;
      g[j] = d * gm[j];
      s1 += g[j];
      s2 += g[j] * xh[j];
      pg[j] += d * xh[j];
      pb[j] += d;
    }
    s1 = warp_sum(s1) * (1.0f / COLS);
    s2 = warp_sum(s2) * (1.0f / COLS);
    float* dxr = dx + (int64_t)row * lddx;
    const float sc = (dxp && row_scale) ? row_scale[row / rows_per_scale] : 1.0f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      float v = rs * (g[j] - s1 - xh[j] * s2);
      if (accumulate_dx) v += dxr[lane + 32 * j];
      dxr[lane + 32 * j] = v;
      if (dxp) {   // the next GEMM's operand: planes of scale(row) * dx, and its column sums (that layer's bias gradient)
        const float u = v * sc;
        __nv_bfloat16 h, l;
        split_bf16(u, h, l);
        dxp[(int64_t)row * ldp + lane + 32 * j] = h;
        dxp[(int64_t)row * ldp + lane + 32 * j + ps] = l;
        pc[j] += u;
      }
    }
  }
  __shared__ float red[8][COLS];
  const int npass = dxp ? 3 : 2;
  for (int pass = 0; pass < npass; ++pass) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < J; ++j) red[warp][lane + 32 * j] = pass == 0 ? pg[j] : (pass == 1 ? pb[j] : pc[j]);
    __syncthreads();
    for (int c = threadIdx.x; c < COLS; c += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][c];
      partial[((int64_t)blockIdx.x * 3 + pass) * COLS + c] = s;
    }
  }
}

// Vectorised variant for cols = 128 * J4 (384, 768, 1024): lane owns 4 consecutive columns per 128-column group, so dy / x /
// dx move as float4 and the optional plane hand-over as 8-byte hi / lo stores.  Same arithmetic and reduction order per column.
template <int J4>
__global__ void __launch_bounds__(256) layernorm_bwd_vec_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ x,
                                                                int64_t ldx, int rows, const float* __restrict__ gamma,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                float* __restrict__ dx, int64_t lddx, int accumulate_dx,
                                                                float* __restrict__ partial, __nv_bfloat16* __restrict__ dxp, int64_t ldp,
                                                                int64_t ps, const float* __restrict__ row_scale, int rows_per_scale,
                                                                const DropParams dr, int drop_rows_per_seq) {
  pdl_trigger();
  pdl_wait();
  constexpr int COLS = J4 * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 pg[J4], pb[J4], pc[J4], gm[J4];
#pragma unroll
  for (int j = 0; j < J4; ++j) {
    pg[j] = pb[j] = pc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    gm[j] = *reinterpret_cast<const float4*>(gamma + j * 128 + lane * 4);
  }
#pragma unroll 2   // a warp walks ~2 rows (4112 rows over 256 x 8 warps): both rows' loads go out together
  for (int row = blockIdx.x * 8 + warp; row < rows; row += 8 * gridDim.x) {
    const float mu = mean[row], rs = rstd[row];
    const float* dyr = dy + (int64_t)row * lddy;
    const float* xr = x + (int64_t)row * ldx;
    float4 g[J4], xh[J4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < J4; ++j) {
      const float4 d = *reinterpret_cast<const float4*>(dyr + j * 128 + lane * 4);
      const float4 xv = *reinterpret_cast<const float4*>(xr + j * 128 + lane * 4);
      xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[j] = make_float4(d.x * gm[j].x, d.y * gm[j].y, d.z * gm[j].z, d.w * gm[j].w);
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * xh[j].x + g[j].y * xh[j].y) + (g[j].z * xh[j].z + g[j].w * xh[j].w);
      pg[j].x += d.x * xh[j].x; pg[j].y += d.y * xh[j].y; pg[j].z += d.z * xh[j].z; pg[j].w += d.w * xh[j].w;
      pb[j].x += d.x; pb[j].y += d.y; pb[j].z += d.z; pb[j].w += d.w;
    }
    s1 = warp_sum(s1) * (1.0f / COLS);
    s2 = warp_sum(s2) * (1.0f / COLS);
    float* dxr = dx + (int64_t)row * lddx;
    const float sc = (dxp && row_scale) ? row_scale[row / rows_per_scale] : 1.0f;
#pragma unroll
    for (int j = 0; j < J4; ++j) {
      float4 v = make_float4(rs * (g[j].x - s1 - xh[j].x * s2), rs * (g[j].y - s1 - xh[j].y * s2), rs * (g[j].z - s1 - xh[j].z * s2),
                             rs * (g[j].w - s1 - xh[j].w * s2));
      float4* dst = reinterpret_cast<float4*>(dxr + j * 128 + lane * 4);
      if (accumulate_dx) {
        const float4 o = *dst;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *dst = v;
      if (dxp) {   // the next GEMM's operand: planes of scale(row) * dx, and its column sums (that layer's bias gradient)
        float4 u = make_float4(v.x * sc, v.y * sc, v.z * sc, v.w * sc);
        if (dr.on) {   // post-LN encoders: the gradient entering the dropout in front of the residual add (same bits as the forward)
          const int sq = row / drop_rows_per_seq;
          const uint32_t key = drop_site_key(dr.seq_key[sq], dr.site);
          const uint32_t base = ((uint32_t)dr.seq_row[sq] * (uint32_t)drop_rows_per_seq + (uint32_t)(row - sq * drop_rows_per_seq)) * (uint32_t)COLS +
                                (uint32_t)(j * 128 + lane * 4);
          u.x = drop_kept(key, base, dr.thr24) ? u.x * dr.inv_keep : 0.f;
          u.y = drop_kept(key, base + 1, dr.thr24) ? u.y * dr.inv_keep : 0.f;
          u.z = drop_kept(key, base + 2, dr.thr24) ? u.z * dr.inv_keep : 0.f;
          u.w = drop_kept(key, base + 3, dr.thr24) ? u.w * dr.inv_keep : 0.f;
        }
        uint32_t h0, l0, h1, l1;
        split2(u.x, u.y, h0, l0);
        split2(u.z, u.w, h1, l1);
        __nv_bfloat16* hp = dxp + (int64_t)row * ldp + j * 128 + lane * 4;
        *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(hp + ps) = make_uint2(l0, l1);
        pc[j].x += u.x; pc[j].y += u.y; pc[j].z += u.z; pc[j].w += u.w;
      }
    }
  }
  __shared__ float4 red[8][COLS / 4];
  const int npass = dxp ? 3 : 2;
  for (int pass = 0; pass < npass; ++pass) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < J4; ++j) red[warp][j * 32 + lane] = pass == 0 ? pg[j] : (pass == 1 ? pb[j] : pc[j]);
    __syncthreads();
    for (int c4 = threadIdx.x; c4 < COLS / 4; c4 += 256) {
      float4 sacc = red[0][c4];
#pragma unroll
      for (int w = 1; w < 8; ++w) { sacc.x += red[w][c4].x; sacc.y += red[w][c4].y; sacc.z += red[w][c4].z; sacc.w += red[w][c4].w; }
      *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.x * 3 + pass) * COLS + c4 * 4) = sacc;
    }
  }
}

__global__ void layernorm_bwd_params_kernel(const float* __restrict__ partial, int nparts, int cols, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float sg = 0.f, sb = 0.f;
  for (int p = 0; p < nparts; ++p) {
    sg += partial[((int64_t)p * 2 + 0) * cols + c];
    sb += partial[((int64_t)p * 2 + 1) * cols + c];
  }
  dgamma[c] = accumulate ? dgamma[c] + sg : sg;
  dbeta[c] = accumulate ? dbeta[c] + sb : sb;
}


// x[i] *= *scale unless *scale == 1 (then no memory is touched)
__global__ void scale_inplace_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ scale) {
  const float sc = *scale;
  if (sc == 1.0f) return;
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(x)[i];
    v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
    reinterpret_cast<float4*>(x)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) x[n4 * 4 + threadIdx.x] *= sc;
}
}  // namespace srw

using namespace srw;

extern "C" int srw_split_planes(const srw_split_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->x && a->rows > 0 && a->cols > 0 && (a->planes || a->planes_t), "srw_split_planes: bad args");
  const bool fast = a->planes && !a->planes_t && a->cols % 4 == 0 && a->ldx % 4 == 0 && a->ldp % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(a->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->planes) & 7) == 0 && a->plane_stride % 4 == 0;
  if (fast) {
    const int nparts = a->colsum_out ? std::min(64, cdiv(a->rows, 32)) : std::min(std::max(1, 592 / cdiv(a->cols, 128)), cdiv(a->rows, 8));
    SRW_REQUIRE(!a->colsum_out || a->colsum_workspace, "srw_split_planes: colsum_out needs colsum_workspace (>= 64 * cols floats)");
    SRW_CUDA(launch_pdl(split_rows_kernel, dim3(dim3(cdiv(a->cols, 128), nparts)), dim3(256), 0, stream, a->x, a->ldx, a->rows, a->cols, a->row_scale,
                                                                          a->rows_per_scale > 0 ? a->rows_per_scale : 1,
                                                                          reinterpret_cast<__nv_bfloat16*>(a->planes), a->ldp, a->plane_stride,
                                                                          a->colsum_out ? a->colsum_workspace : nullptr));
    g_launches++;
    SRW_LAUNCH_CHECK();
    if (a->colsum_out) {
      SRW_CUDA(launch_pdl(reduce_partials_kernel, dim3(dim3(cdiv(a->cols, 32), 1)), dim3(256), 0, stream, a->colsum_workspace, nparts, a->cols, 0, a->cols, a->colsum_out, nullptr,
                                                                               a->colsum_accumulate, nullptr, 0));
      g_launches++;
      SRW_LAUNCH_CHECK();
    }
    return SRW_OK;
  }
  SRW_REQUIRE(!a->colsum_out, "srw_split_planes: fused column sums need the aligned non-transposed path");
  dim3 grid(cdiv(a->cols, 32), cdiv(a->rows, 32)), block(32, 8);
  split_planes_kernel<<<grid, block, 0, stream>>>(a->x, a->ldx, a->rows, a->cols, a->row_scale, a->rows_per_scale > 0 ? a->rows_per_scale : 1,
                                                 reinterpret_cast<__nv_bfloat16*>(a->planes), a->ldp, a->plane_stride,
                                                 reinterpret_cast<__nv_bfloat16*>(a->planes_t), a->ldpt, a->plane_stride_t);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_splitk_reduce(const srw_splitk_reduce_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->workspace && a->out && a->split_k >= 1 && a->N % 4 == 0 && a->ldo % 4 == 0, "srw_splitk_reduce: bad args");
  const int64_t total4 = (int64_t)a->M * a->N / 4;
  const int blocks = (int)std::min<int64_t>(cdiv64(total4, 256), 148 * 8);
  SRW_CUDA(launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, stream, a->workspace, a->split_k, a->M, a->N, a->out, a->ldo, a->accumulate));
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_colsum(const srw_colsum_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && (a->x || a->planes) && a->workspace && a->rows > 0 && a->cols > 0, "srw_colsum: bad args");
  const int nparts = srw_colsum_nparts(a->rows);
  if (a->planes && !a->row_scale && a->cols % 8 == 0 && a->ldp % 8 == 0 && a->plane_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(a->planes) & 15) == 0) {
    SRW_CUDA(launch_pdl(colsum_planes_kernel, dim3(dim3(cdiv(a->cols, 256), nparts)), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(a->planes), a->ldp, a->plane_stride,
                                                                              a->rows, a->cols, a->workspace));
    g_launches++;
    SRW_LAUNCH_CHECK();
    if (!a->out) return SRW_OK;   // the caller folds the partials (srw_grad_fold)
    SRW_CUDA(launch_pdl(reduce_partials_kernel, dim3(dim3(cdiv(a->cols, 32), 1)), dim3(256), 0, stream, a->workspace, nparts, a->cols, 0, a->cols, a->out, nullptr, a->accumulate, nullptr, 0));
    g_launches++;
    SRW_LAUNCH_CHECK();
    return SRW_OK;
  }
  dim3 grid(cdiv(a->cols, 32), nparts);
  colsum_stage1_kernel<<<grid, 256, 0, stream>>>(a->x, a->ldx, reinterpret_cast<const __nv_bfloat16*>(a->planes), a->ldp, a->plane_stride,
                                                a->row_scale, a->rows_per_scale > 0 ? a->rows_per_scale : 1, a->rows, a->cols, a->workspace);
  g_launches++;
  SRW_LAUNCH_CHECK();
  if (!a->out) return SRW_OK;
  SRW_CUDA(launch_pdl(reduce_partials_kernel, dim3(dim3(cdiv(a->cols, 32), 1)), dim3(256), 0, stream, a->workspace, nparts, a->cols, 0, a->cols, a->out, nullptr, a->accumulate, nullptr, 0));
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_colsum_nparts(int rows) { return std::min(64, cdiv(rows, 32)); }
extern "C" int srw_layernorm_bwd_nparts(int rows) { return std::min(256, cdiv(rows, 8)); }

extern "C" int srw_grad_fold(const srw_grad_fold_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->n_splitk >= 0 && a->n_splitk <= 4 && a->n_colsum >= 0 && a->n_colsum <= 8 && a->n_splitk + a->n_colsum > 0, "srw_grad_fold: bad counts");
  FoldParams fp = {};
  fp.n_splitk = a->n_splitk; fp.n_colsum = a->n_colsum;
  int nb = 0, k = 0;
  for (int i = 0; i < a->n_splitk; ++i, ++k) {
    const srw_splitk_reduce_args& r = a->splitk[i];
    SRW_REQUIRE(r.workspace && r.out && r.split_k >= 1 && r.N % 4 == 0 && r.ldo % 4 == 0, "srw_grad_fold: bad split-K problem %d", i);
    fp.sk[i] = r;
    fp.first_block[k] = nb;
    nb += (int)std::min<int64_t>(cdiv64((int64_t)r.M * r.N / 4, 256), 148 * 2);
  }
  for (int i = 0; i < a->n_colsum; ++i, ++k) {
    const srw_fold_colsum& r = a->colsum[i];
    SRW_REQUIRE(r.partial && r.out && r.nparts >= 1 && r.cols > 0, "srw_grad_fold: bad column-sum problem %d", i);
    fp.cs[i] = r;
    fp.first_block[k] = nb;
    nb += cdiv(r.cols, 32);
  }
  fp.first_block[k] = nb;
  SRW_CUDA(launch_pdl(grad_fold_kernel, dim3(nb), dim3(256), 0, stream, fp));
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_layernorm_fwd(const srw_layernorm_fwd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->x && a->gamma && a->beta && a->rows > 0 && a->cols > 0 && a->cols % 4 == 0 && a->ldx % 4 == 0,
              "srw_layernorm_fwd: bad args (cols %% 4 == 0 required)");
  SRW_REQUIRE(!a->y_planes || a->ldp % 4 == 0, "srw_layernorm_fwd: ldp %% 4");
  SRW_REQUIRE(!a->y_f32 || a->ldy % 4 == 0, "srw_layernorm_fwd: ldy %% 4");
#define SRW_LN_FWD(J)                                                                                                                   \
  SRW_CUDA(launch_pdl(layernorm_fwd_reg_kernel<J>, dim3(cdiv(a->rows, 8)), dim3(256), 0, stream, a->x, a->ldx, a->rows, a->eps, a->gamma, a->beta, a->mean, a->rstd,       \
                                                                   reinterpret_cast<__nv_bfloat16*>(a->y_planes), a->ldp, a->plane_stride, \
                                                                   a->y_f32, a->ldy))
  if (a->cols == 384) SRW_LN_FWD(3);
  else if (a->cols == 768) SRW_LN_FWD(6);
  else if (a->cols == 1024) SRW_LN_FWD(8);
  else
    layernorm_fwd_kernel<<<cdiv(a->rows, 8), 256, 0, stream>>>(a->x, a->ldx, a->rows, a->cols, a->eps, a->gamma, a->beta, a->mean, a->rstd,
                                                              reinterpret_cast<__nv_bfloat16*>(a->y_planes), a->ldp, a->plane_stride,
                                                              a->y_f32, a->ldy);
#undef SRW_LN_FWD
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_layernorm_bwd(const srw_layernorm_bwd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->dy && a->x && a->gamma && a->mean && a->rstd && a->dx && a->workspace && a->rows > 0, "srw_layernorm_bwd: bad args");
  SRW_REQUIRE(!a->dx_planes || (a->ldp > 0 && a->plane_stride > 0), "srw_layernorm_bwd: dx_planes needs ldp / plane_stride");
  SRW_REQUIRE(a->cols % 32 == 0 && a->cols <= 1024, "srw_layernorm_bwd: cols must be a multiple of 32 and <= 1024 (cols=%d)", a->cols);
  const int nblocks = srw_layernorm_bwd_nparts(a->rows);
  const DropParams dr = make_drop(a->drop);
#define SRW_LN_BWD(J)                                                                                                             \
  case J:                                                                                                                         \
    SRW_CUDA(launch_pdl(layernorm_bwd_kernel<J>, dim3(nblocks), dim3(256), 0, stream, a->dy, a->lddy, a->x, a->ldx, a->rows, a->gamma, a->mean, a->rstd, a->dx, \
                                                         a->lddx, a->accumulate_dx, a->workspace, reinterpret_cast<__nv_bfloat16*>(a->dx_planes), a->ldp,   \
                                                         a->plane_stride, a->row_scale, a->rows_per_scale > 0 ? a->rows_per_scale : 1));                                \
    break;
#define SRW_LN_BWD_VEC(J4)                                                                                                          \
  SRW_CUDA(launch_pdl(layernorm_bwd_vec_kernel<J4>, dim3(nblocks), dim3(256), 0, stream, a->dy, a->lddy, a->x, a->ldx, a->rows, a->gamma,   \
                      a->mean, a->rstd, a->dx, a->lddx, a->accumulate_dx, a->workspace, reinterpret_cast<__nv_bfloat16*>(a->dx_planes), a->ldp, \
                      a->plane_stride, a->row_scale, a->rows_per_scale > 0 ? a->rows_per_scale : 1, dr, a->drop_rows_per_seq > 0 ? a->drop_rows_per_seq : 1))
  const bool vec_ok = a->lddy % 4 == 0 && a->ldx % 4 == 0 && a->lddx % 4 == 0 && (!a->dx_planes || (a->ldp % 4 == 0 && a->plane_stride % 4 == 0)) &&
                      ((reinterpret_cast<uintptr_t>(a->dy) | reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->dx) |
                        reinterpret_cast<uintptr_t>(a->gamma)) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->dx_planes) & 7) == 0;
  SRW_REQUIRE(!dr.on || (vec_ok && a->dx_planes && (a->cols == 384 || a->cols == 768 || a->cols == 1024)),
              "srw_layernorm_bwd: the dropout hand-over needs dx_planes and the vectorised path (cols 384 / 768 / 1024, 16-byte aligned)");
  if (vec_ok && a->cols == 384) { SRW_LN_BWD_VEC(3); }
  else if (vec_ok && a->cols == 768) { SRW_LN_BWD_VEC(6); }
  else if (vec_ok && a->cols == 1024) { SRW_LN_BWD_VEC(8); }
  else
  switch (a->cols / 32) {
    SRW_LN_BWD(2) SRW_LN_BWD(4) SRW_LN_BWD(6) SRW_LN_BWD(8) SRW_LN_BWD(12) SRW_LN_BWD(16) SRW_LN_BWD(24) SRW_LN_BWD(32)
    default:
      set_last_error("srw_layernorm_bwd: unsupported cols=%d", a->cols);
      return SRW_ERR_UNSUPPORTED;
  }
#undef SRW_LN_BWD
#undef SRW_LN_BWD_VEC
  g_launches++;
  SRW_LAUNCH_CHECK();
  if (a->dgamma && a->dbeta) {
    SRW_REQUIRE(!a->colsum_out || a->dx_planes, "srw_layernorm_bwd: colsum_out needs dx_planes");
    SRW_CUDA(launch_pdl(reduce_partials_kernel, dim3(cdiv(a->cols, 32), (a->dx_planes && a->colsum_out) ? 3 : 2), dim3(256), 0, stream, a->workspace, nblocks,
                        3 * (int64_t)a->cols, a->cols, a->cols, a->dgamma, a->dbeta, a->accumulate_dparams, a->colsum_out, a->colsum_accumulate));
    g_launches++;
    SRW_LAUNCH_CHECK();
  }
  return SRW_OK;
}

extern "C" int srw_scale_inplace(float* x, int64_t n, const float* scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(x && scale && n > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "srw_scale_inplace: bad args (x must be 16-byte aligned)");
  scale_inplace_kernel<<<148 * 8, 256, 0, stream>>>(x, n, scale);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
