// srw_augment — the reference's image input pipeline on the device: `transform_weak` / `transform_strong` of get_cifar
// (semilearn/datasets/cv_datasets/cifar.py:34-49) as BasicDataset.__getitem__ applies them per sample
// (semilearn/datasets/cv_datasets/datasetbase.py:74-115), with RandAugment(3, 5) + Cutout
// (semilearn/datasets/augmentation/randaugment.py:16-206).  The reference runs three PIL pipelines per sample in DataLoader
// workers and ships fp32 tensors over PCIe; here the uint8 dataset lives in HBM (CIFAR-100: 150 MB), the host sends only the
// random DECISIONS (232 B per sample) and one CTA per sample produces the normalised fp32 NCHW row in place of the batch.
//
// Byte / integer work, bit-exact with Pillow 12.2 + torchvision 0.26 (oracle/augment_oracle.py restates and pins them):
//   reflect-pad + crop + hflip (gather) -> up to 3 ops -> Cutout rectangle -> (v / 255 - mean) / std.
// An image (H*W*3 bytes, <= 48 KB) stays in shared memory for the whole chain: two buffers (ops with spatial support write
// the other one), per-band histograms and look-up tables next to them.  HBM traffic = 3 B read + 12 B written per pixel.
// Every float / double expression uses the round-to-nearest intrinsics: nvcc must not contract a*b+c into an FMA where Pillow's C
// or Python evaluates two rounded operations.
#include <cuda_runtime.h>

#include <atomic>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {
extern std::atomic<int64_t> g_launches;

constexpr int AUG_THREADS = 256;

struct AugScratch {
  int hist[3][256];
  unsigned char lut[3][256];
  int sum;            // luma sum (Contrast)
  int pad_[3];
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
  i = i < 0 ? -i : i;
  return i >= n ? 2 * (n - 1) - i : i;
}

// convert('L'): ITU-R 601-2 luma in 16-bit fixed point, rounded
__device__ __forceinline__ int luma8(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

// libImaging Blend.c: (UINT8)((int)in1 + alpha * ((int)in2 - (int)in1)), C float arithmetic, clipped when extrapolating
__device__ __forceinline__ unsigned char blend8(int deg, int img, float alpha, bool interp) {
  const float t = __fadd_rn((float)deg, __fmul_rn(alpha, (float)(img - deg)));
  if (interp) return (unsigned char)(int)t;
  return t <= 0.f ? 0 : (t >= 255.f ? 255 : (unsigned char)(int)t);
}

__device__ void build_histogram(const unsigned char* img, int nbytes, AugScratch* sc) {
  for (int i = threadIdx.x; i < 768; i += AUG_THREADS) (&sc->hist[0][0])[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) atomicAdd(&sc->hist[i % 3][img[i]], 1);
  __syncthreads();
}

__device__ void apply_lut(unsigned char* img, int nbytes, const AugScratch* sc) {
  for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) img[i] = sc->lut[i % 3][img[i]];
  __syncthreads();
}

// ImageOps.autocontrast(cutoff = 0): per band, stretch [lowest, highest occupied level] to [0, 255]; Python float arithmetic
// (scale = 255.0 / (hi - lo); offset = -lo * scale; int(ix * scale + offset), clamped).
__device__ void op_autocontrast(unsigned char* img, int nbytes, AugScratch* sc) {
  build_histogram(img, nbytes, sc);
  __shared__ int lohi[3][2];
  if (threadIdx.x < 3) {
    const int* h = sc->hist[threadIdx.x];
    int lo = 0, hi = 255;
    while (lo < 255 && h[lo] == 0) ++lo;
    while (hi > 0 && h[hi] == 0) --hi;
    lohi[threadIdx.x][0] = lo; lohi[threadIdx.x][1] = hi;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 768; i += AUG_THREADS) {
    const int c = i >> 8, ix = i & 255, lo = lohi[c][0], hi = lohi[c][1];
    int v = ix;
    if (hi > lo) {
      const double scale = __ddiv_rn(255.0, (double)(hi - lo));
      const double offset = __dmul_rn(-(double)lo, scale);
      v = (int)__dadd_rn(__dmul_rn((double)ix, scale), offset);      // int(): truncation toward zero
      v = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
    sc->lut[c][ix] = (unsigned char)v;
  }
  __syncthreads();
  apply_lut(img, nbytes, sc);
}

// ImageOps.equalize: per band, integer arithmetic: step = (pixels - count of the highest occupied level) // 255;
// lut[i] = (step // 2 + sum_{j < i} h[j]) // step, clipped at 255 by Image.point; identity when <= 1 level or step == 0.
__device__ void op_equalize(unsigned char* img, int nbytes, AugScratch* sc) {
  build_histogram(img, nbytes, sc);
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    const int* h = sc->hist[c];
    int levels = 0, last = 0, total = 0;
    for (int i = 0; i < 256; ++i)
      if (h[i]) { ++levels; last = h[i]; total += h[i]; }
    const int step = (total - last) / 255;
    if (levels <= 1 || step == 0) {
      for (int i = 0; i < 256; ++i) sc->lut[c][i] = (unsigned char)i;
    } else {
      int n = step / 2;
      for (int i = 0; i < 256; ++i) {
        const int v = n / step;
        sc->lut[c][i] = (unsigned char)(v > 255 ? 255 : v);
        n += h[i];
      }
    }
  }
  __syncthreads();
  apply_lut(img, nbytes, sc);
}

// Image.transform(AFFINE, NEAREST), fill 0 (libImaging Geometry.c).  a[1] == a[3] == 0: the axis-aligned path (source column and
// row from coordinates ACCUMULATED in doubles, negative = outside); otherwise 16.16 fixed point with the pixel centre folded
// into the offsets.
__device__ void op_affine(const unsigned char* src, unsigned char* dst, int S, const double* a, AugScratch* sc) {
  const int npix = S * S;
  if (a[1] == 0.0 && a[3] == 0.0) {
    // libImaging walks xo += a[0] / yo += a[4] once per column / row: the same running sums, one thread per column and per row
    int* xin = sc->hist[0];                // [S] source column of every output column (S <= 128)
    int* yin = sc->hist[1];
    if (threadIdx.x < 2 * S) {
      const bool is_y = threadIdx.x >= S;
      const int k = is_y ? threadIdx.x - S : threadIdx.x;
      const double step = is_y ? a[4] : a[0];
      double o = __dadd_rn(is_y ? a[5] : a[2], __dmul_rn(step, 0.5));
      for (int j = 0; j < k; ++j) o = __dadd_rn(o, step);
      (is_y ? yin : xin)[k] = o < 0.0 ? -1 : (int)o;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < npix; p += AUG_THREADS) {
      const int y = p / S, x = p - y * S;
      const int xi = xin[x], yi = yin[y];
      const bool ok = xi >= 0 && xi < S && yi >= 0 && yi < S;
      const int q = ok ? (yi * S + xi) * 3 : 0;
      dst[p * 3 + 0] = ok ? src[q + 0] : 0;
      dst[p * 3 + 1] = ok ? src[q + 1] : 0;
      dst[p * 3 + 2] = ok ? src[q + 2] : 0;
    }
  } else {
    auto fix = [](double v) { return (int)floor(__dadd_rn(__dmul_rn(v, 65536.0), 0.5)); };
    const int a0 = fix(a[0]), a1 = fix(a[1]), a3 = fix(a[3]), a4 = fix(a[4]);
    const int a2 = fix(__dadd_rn(__dadd_rn(a[2], __dmul_rn(a[0], 0.5)), __dmul_rn(a[1], 0.5)));
    const int a5 = fix(__dadd_rn(__dadd_rn(a[5], __dmul_rn(a[3], 0.5)), __dmul_rn(a[4], 0.5)));
    for (int p = threadIdx.x; p < npix; p += AUG_THREADS) {
      const int y = p / S, x = p - y * S;
      const int xi = (a2 + a1 * y + a0 * x) >> 16, yi = (a5 + a4 * y + a3 * x) >> 16;
      const bool ok = xi >= 0 && xi < S && yi >= 0 && yi < S;
      const int q = ok ? (yi * S + xi) * 3 : 0;
      dst[p * 3 + 0] = ok ? src[q + 0] : 0;
      dst[p * 3 + 1] = ok ? src[q + 1] : 0;
      dst[p * 3 + 2] = ok ? src[q + 2] : 0;
    }
  }
  __syncthreads();
}

// ImageEnhance.Sharpness: blend(image.filter(SMOOTH), image, alpha).  SMOOTH = (1 1 1 / 1 5 1 / 1 1 1) / 13 on the interior, border
// copied; the rounded float sum of Filter.c equals (2 sum + 13) / 26 in integers (sum / 13 is never within float error of a half).
__device__ void op_sharpness(const unsigned char* src, unsigned char* dst, int S, float alpha, bool interp) {
  const int nbytes = S * S * 3, row = S * 3;
  for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) {
    const int y = i / row, xb = i - y * row, x = xb / 3;
    const int v = src[i];
    int deg = v;
    if (y > 0 && y < S - 1 && x > 0 && x < S - 1) {
      const int s = src[i - row - 3] + src[i - row] + src[i - row + 3] + src[i - 3] + 5 * v + src[i + 3] + src[i + row - 3] + src[i + row] +
                    src[i + row + 3];
      deg = (2 * s + 13) / 26;
    }
    dst[i] = blend8(deg, v, alpha, interp);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(AUG_THREADS)
augment_kernel(const srw_augment_args A) {
  extern __shared__ __align__(16) unsigned char aug_smem[];
  const int S = A.img_size, npix = S * S, nbytes = npix * 3;
  const int buf_bytes = (nbytes + 15) & ~15;
  unsigned char* buf0 = aug_smem;
  unsigned char* buf1 = aug_smem + buf_bytes;
  AugScratch* sc = reinterpret_cast<AugScratch*>(aug_smem + 2 * buf_bytes);
  const srw_aug_sample& smp = A.samples[blockIdx.x];

  // ---- F.pad(reflect) + crop + hflip as one gather from the resident dataset ----
  {
    const unsigned char* src = A.src + (size_t)smp.src_index * nbytes;
    const int top = smp.crop_top - A.padding, left = smp.crop_left - A.padding, flip = smp.flip;
    for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) {
      const int p = i / 3, c = i - p * 3, y = p / S, x = p - y * S;
      const int ys = reflect_idx(y + top, S), xs = reflect_idx((flip ? S - 1 - x : x) + left, S);
      buf0[i] = src[(ys * S + xs) * 3 + c];
    }
  }
  __syncthreads();
  unsigned char* cur = buf0;
  unsigned char* alt = buf1;

  // ---- RandAugment ops (uniform per CTA) ----
  for (int k = 0; k < smp.n_ops; ++k) {
    const srw_aug_op_desc& op = smp.ops[k];
    const bool interp = op.alpha >= 0.f && op.alpha <= 1.f;
    switch (op.op) {
      case SRW_AUG_AUTOCONTRAST: op_autocontrast(cur, nbytes, sc); break;
      case SRW_AUG_EQUALIZE: op_equalize(cur, nbytes, sc); break;
      case SRW_AUG_POSTERIZE: {          // ImageOps.posterize: keep the top `bits` bits
        const unsigned char mask = (unsigned char)~((1 << (8 - op.ival)) - 1);
        for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) cur[i] &= mask;
        __syncthreads();
      } break;
      case SRW_AUG_SOLARIZE: {           // ImageOps.solarize: invert levels >= threshold (ival = ceil(threshold))
        for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) { const int v = cur[i]; cur[i] = (unsigned char)(v < op.ival ? v : 255 - v); }
        __syncthreads();
      } break;
      case SRW_AUG_BRIGHTNESS: {         // blend(black, image, alpha)
        for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) cur[i] = blend8(0, cur[i], op.alpha, interp);
        __syncthreads();
      } break;
      case SRW_AUG_COLOR: {              // blend(grey version, image, alpha)
        for (int p = threadIdx.x; p < npix; p += AUG_THREADS) {
          const int r = cur[p * 3], g = cur[p * 3 + 1], b = cur[p * 3 + 2], l = luma8(r, g, b);
          cur[p * 3] = blend8(l, r, op.alpha, interp); cur[p * 3 + 1] = blend8(l, g, op.alpha, interp); cur[p * 3 + 2] = blend8(l, b, op.alpha, interp);
        }
        __syncthreads();
      } break;
      case SRW_AUG_CONTRAST: {           // blend(constant int(mean luma + 0.5), image, alpha)
        if (threadIdx.x == 0) sc->sum = 0;
        __syncthreads();
        int part = 0;
        for (int p = threadIdx.x; p < npix; p += AUG_THREADS) part += luma8(cur[p * 3], cur[p * 3 + 1], cur[p * 3 + 2]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&sc->sum, part);
        __syncthreads();
        const int mean = (int)__dadd_rn(__ddiv_rn((double)sc->sum, (double)npix), 0.5);
        for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) cur[i] = blend8(mean, cur[i], op.alpha, interp);
        __syncthreads();
      } break;
      case SRW_AUG_SHARPNESS: {
        op_sharpness(cur, alt, S, op.alpha, interp);
        unsigned char* t = cur; cur = alt; alt = t;
      } break;
      case SRW_AUG_ROTATE: case SRW_AUG_SHEAR_X: case SRW_AUG_SHEAR_Y: case SRW_AUG_TRANSLATE_X: case SRW_AUG_TRANSLATE_Y: {
        if (!op.identity) {
          op_affine(cur, alt, S, op.a, sc);
          unsigned char* t = cur; cur = alt; alt = t;
        }
      } break;
      default: break;                    // SRW_AUG_IDENTITY
    }
  }

  // ---- Cutout: ImageDraw.rectangle, both corners inclusive, clipped ----
  if (smp.cut_x1 >= smp.cut_x0 && smp.cut_y1 >= smp.cut_y0) {
    const int x0 = max(smp.cut_x0, 0), y0 = max(smp.cut_y0, 0), x1 = min(smp.cut_x1, S - 1), y1 = min(smp.cut_y1, S - 1);
    for (int p = threadIdx.x; p < npix; p += AUG_THREADS) {
      const int y = p / S, x = p - y * S;
      if (x >= x0 && x <= x1 && y >= y0 && y <= y1) { cur[p * 3] = 125; cur[p * 3 + 1] = 123; cur[p * 3 + 2] = 114; }
    }
    __syncthreads();
  }

  // ---- ToTensor (true division by 255) + Normalize, HWC bytes -> CHW floats, coalesced stores ----
  float* out = A.out + (size_t)blockIdx.x * nbytes;
  for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) {
    const int c = i / npix, p = i - c * npix;
    const float t = __fdiv_rn((float)cur[p * 3 + c], 255.f);
    out[i] = __fdiv_rn(__fsub_rn(t, A.mean[c]), A.std[c]);
  }
  if (A.out_u8) {
    unsigned char* o8 = A.out_u8 + (size_t)blockIdx.x * nbytes;
    for (int i = threadIdx.x; i < nbytes; i += AUG_THREADS) o8[i] = cur[i];
  }
}
}  // namespace srw

using namespace srw;

extern "C" int srw_augment_batch(const srw_augment_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->src && a->samples && a->out && a->n > 0 && a->n_src > 0, "srw_augment_batch: null pointer or empty batch");
  SRW_REQUIRE(a->img_size >= 3 && a->img_size <= 128, "srw_augment_batch: image size %d outside [3, 128] (two copies must fit in shared memory)", a->img_size);
  SRW_REQUIRE(a->padding >= 0 && a->padding < a->img_size, "srw_augment_batch: reflect padding %d needs padding < size %d", a->padding, a->img_size);
  const int nbytes = a->img_size * a->img_size * 3;
  const size_t smem = 2 * (size_t)((nbytes + 15) & ~15) + sizeof(AugScratch);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    SRW_CUDA(cudaFuncSetAttribute(augment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  augment_kernel<<<(unsigned)a->n, AUG_THREADS, smem, stream>>>(*a);
  g_launches++;
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}
