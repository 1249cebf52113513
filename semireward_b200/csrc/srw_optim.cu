// srw_optim — multi-tensor AdamW / Adam in ONE launch (replaces optimizer.step() + model.zero_grad() of
// param_update.py:36-40 for the optimizer built by get_optimizer, build.py:193-224).  HBM-bound: per element it reads
// p, g, m, v and writes p, m, v (+ the split-bf16 planes of p for the ViT engine's weight cache) = 28-32 bytes.
// Each CTA handles one SRW_ADAMW_BLOCK_ELEMS chunk of one tensor, found by binary search in the device-resident table;
// float4 accesses, grid = total chunks.
#include <atomic>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {
extern std::atomic<int64_t> g_launches;

struct AdamWScalars {
  double lr_factor, beta1, beta2, eps, bc1, bc2_sqrt;
  int decoupled, num_tensors;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float decay_mul, float wd_l2, float w1, float b2, float w2,
                                          float bc2s, float eps, float step_size) {
  if (wd_l2 != 0.f) g = fmaf(wd_l2, p, g);       // Adam (coupled L2): grad += wd * p
  p *= decay_mul;                                 // AdamW: p *= 1 - lr * wd   (1.0 otherwise)
  m = m + w1 * (g - m);                           // lerp
  v = v * b2 + (w2 * g) * g;
  const float denom = sqrtf(v) / bc2s + eps;
  p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(256) adamw_kernel(const srw_adamw_row* __restrict__ table, const AdamWScalars sc) {
  // binary search: last row with first_block <= blockIdx.x
  int lo = 0, hi = sc.num_tensors - 1;
  const int64_t blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const srw_adamw_row row = table[lo];
  const int64_t base = (blk - row.first_block) * SRW_ADAMW_BLOCK_ELEMS;
  const int64_t end = min(row.numel, base + SRW_ADAMW_BLOCK_ELEMS);
  const double lr = row.lr * sc.lr_factor;                        // LambdaLR: group lr = base lr * factor (double math)
  const float step_size = (float)(lr / sc.bc1);
  const float decay_mul = (sc.decoupled && row.weight_decay != 0.0) ? (float)(1.0 - lr * row.weight_decay) : 1.0f;
  const float wd_l2 = sc.decoupled ? 0.f : (float)row.weight_decay;
  const float w1 = (float)(1.0 - sc.beta1), b2 = (float)sc.beta2, w2 = (float)(1.0 - sc.beta2);
  const float bc2s = (float)sc.bc2_sqrt, eps = (float)sc.eps;
  __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(row.planes);
  const bool vec = ((row.numel & 3) == 0) && ((reinterpret_cast<uintptr_t>(row.param) | reinterpret_cast<uintptr_t>(row.grad) |
                                                reinterpret_cast<uintptr_t>(row.exp_avg) | reinterpret_cast<uintptr_t>(row.exp_avg_sq)) & 15) == 0;
  if (vec) {
    for (int64_t i = base + threadIdx.x * 4; i < end; i += 256 * 4) {
      float4 p = *reinterpret_cast<float4*>(row.param + i);
      const float4 g = *reinterpret_cast<const float4*>(row.grad + i);
      float4 m = *reinterpret_cast<float4*>(row.exp_avg + i);
      float4 v = *reinterpret_cast<float4*>(row.exp_avg_sq + i);
      adam_elem(p.x, g.x, m.x, v.x, decay_mul, wd_l2, w1, b2, w2, bc2s, eps, step_size);
      adam_elem(p.y, g.y, m.y, v.y, decay_mul, wd_l2, w1, b2, w2, bc2s, eps, step_size);
      adam_elem(p.z, g.z, m.z, v.z, decay_mul, wd_l2, w1, b2, w2, bc2s, eps, step_size);
      adam_elem(p.w, g.w, m.w, v.w, decay_mul, wd_l2, w1, b2, w2, bc2s, eps, step_size);
      *reinterpret_cast<float4*>(row.param + i) = p;
      *reinterpret_cast<float4*>(row.exp_avg + i) = m;
      *reinterpret_cast<float4*>(row.exp_avg_sq + i) = v;
      if (planes) {
        const float pv[4] = {p.x, p.y, p.z, p.w};
        if (row.cols == row.ldp) {
          uint32_t h0, l0, h1, l1;
          split2(pv[0], pv[1], h0, l0);
          split2(pv[2], pv[3], h1, l1);
          *reinterpret_cast<uint2*>(planes + i) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(planes + i + row.plane_stride) = make_uint2(l0, l1);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int64_t e = i + j, r = e / row.cols, c = e % row.cols;
            __nv_bfloat16 h, l;
            split_bf16(pv[j], h, l);
            planes[r * row.ldp + c] = h;
            planes[r * row.ldp + c + row.plane_stride] = l;
          }
        }
      }
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += 256) {
      float p = row.param[i], m = row.exp_avg[i], v = row.exp_avg_sq[i];
      adam_elem(p, row.grad[i], m, v, decay_mul, wd_l2, w1, b2, w2, bc2s, eps, step_size);
      row.param[i] = p; row.exp_avg[i] = m; row.exp_avg_sq[i] = v;
      if (planes) {
        const int64_t r = i / row.cols, c = i % row.cols;
        __nv_bfloat16 h, l;
        split_bf16(p, h, l);
        planes[r * row.ldp + c] = h;
        planes[r * row.ldp + c + row.plane_stride] = l;
      }
    }
  }
}
// torch.optim.SGD single-tensor math (momentum, dampening 0, nesterov, coupled weight decay), same chunk-per-CTA table walk
__global__ void __launch_bounds__(256) sgd_kernel(const srw_adamw_row* __restrict__ table, int num_tensors, double lr_factor, float momentum, int nesterov,
                                                  int first_step) {
  int lo = 0, hi = num_tensors - 1;
  const int64_t blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const srw_adamw_row row = table[lo];
  const int64_t base = (blk - row.first_block) * SRW_ADAMW_BLOCK_ELEMS;
  const int64_t end = min(row.numel, base + SRW_ADAMW_BLOCK_ELEMS);
  const float lr = (float)(row.lr * lr_factor), wd = (float)row.weight_decay;
  for (int64_t i = base + threadIdx.x; i < end; i += 256) {
    float p = row.param[i], g = row.grad[i];
    if (wd != 0.f) g = fmaf(wd, p, g);
    float upd = g;
    if (momentum != 0.f) {
      const float buf = first_step ? g : fmaf(momentum, row.exp_avg[i], g);
      row.exp_avg[i] = buf;
      upd = nesterov ? fmaf(momentum, buf, g) : buf;
    }
    row.param[i] = p - lr * upd;
  }
}
// EMA of the parameters (EMA.update, semilearn/core/utils/misc.py:152-155): shadow = (1 - d) * p + d * shadow, two rounded
// products and one rounded sum like the reference's tensor expression (bit-exact).  Same chunk-per-CTA table walk as AdamW.
__global__ void __launch_bounds__(256) ema_kernel(const srw_ema_row* __restrict__ table, int num_tensors, float d, float omd) {
  int lo = 0, hi = num_tensors - 1;
  const int64_t blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].first_block <= blk) lo = mid; else hi = mid - 1;
  }
  const srw_ema_row row = table[lo];
  const int64_t base = (blk - row.first_block) * SRW_ADAMW_BLOCK_ELEMS;
  const int64_t end = min(row.numel, base + SRW_ADAMW_BLOCK_ELEMS);
  const bool vec = ((row.numel & 3) == 0) && ((reinterpret_cast<uintptr_t>(row.param) | reinterpret_cast<uintptr_t>(row.shadow)) & 15) == 0;
  if (vec) {
    for (int64_t i = base + threadIdx.x * 4; i < end; i += 256 * 4) {
      const float4 p = *reinterpret_cast<const float4*>(row.param + i);
      float4 s = *reinterpret_cast<float4*>(row.shadow + i);
      s.x = __fadd_rn(__fmul_rn(omd, p.x), __fmul_rn(d, s.x)); s.y = __fadd_rn(__fmul_rn(omd, p.y), __fmul_rn(d, s.y));
      s.z = __fadd_rn(__fmul_rn(omd, p.z), __fmul_rn(d, s.z)); s.w = __fadd_rn(__fmul_rn(omd, p.w), __fmul_rn(d, s.w));
      *reinterpret_cast<float4*>(row.shadow + i) = s;
    }
  } else {
    for (int64_t i = base + threadIdx.x; i < end; i += 256) row.shadow[i] = __fadd_rn(__fmul_rn(omd, row.param[i]), __fmul_rn(d, row.shadow[i]));
  }
}
}  // namespace srw

using namespace srw;

extern "C" int srw_ema_step(const srw_ema_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->table && a->num_tensors > 0 && a->total_blocks > 0, "srw_ema_step: bad args");
  ema_kernel<<<(unsigned)a->total_blocks, 256, 0, stream>>>(a->table, a->num_tensors, (float)a->decay, (float)(1.0 - a->decay));
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_sgd_step(const srw_sgd_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->table && a->num_tensors > 0 && a->total_blocks > 0, "srw_sgd_step: bad args");
  sgd_kernel<<<(unsigned)a->total_blocks, 256, 0, stream>>>(reinterpret_cast<const srw_adamw_row*>(a->table), a->num_tensors, a->lr_factor, (float)a->momentum,
                                                           a->nesterov, a->first_step);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

extern "C" int srw_adamw_step(const srw_adamw_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->table && a->num_tensors > 0 && a->total_blocks > 0 && a->step >= 1, "srw_adamw_step: bad args");
  AdamWScalars sc;
  sc.lr_factor = a->lr_factor; sc.beta1 = a->beta1; sc.beta2 = a->beta2; sc.eps = a->eps;
  sc.bc1 = 1.0 - pow(a->beta1, (double)a->step);
  sc.bc2_sqrt = sqrt(1.0 - pow(a->beta2, (double)a->step));
  sc.decoupled = a->decoupled; sc.num_tensors = a->num_tensors;
  void* prof = prof_begin(SRW_PROF_ADAMW, 0.0, 32.0 * (double)a->total_blocks * SRW_ADAMW_BLOCK_ELEMS, stream);
  adamw_kernel<<<(unsigned)a->total_blocks, 256, 0, stream>>>(a->table, sc);
  prof_end(prof, stream);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
