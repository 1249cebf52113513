// tcgen05.mma probe for the attention kernels (diagnostic; not part of the library).
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o semireward_b200/lib/mma_probe scripts/mma_probe.cu -lcuda
// 1. cycles per M=128 x N x K=16 bf16 MMA as a function of N, of the B layout (K- / MN-major), of where A comes from
//    (shared-memory descriptor vs TMEM) and of the number of accumulators the chain rotates over;
// 2. a functional check of the "P through TMEM" plan: P written in place of S as packed bf16 hi|lo column groups with
//    tcgen05.st, consumed as the A operand from TMEM, against B = [V_hi | V_lo] (one N = 128 MN-major operand whose two
//    64-wide chunks are LBO apart) and B = V_hi.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../semireward_b200/csrc/srw_common.cuh"

using namespace srw;

struct Variant {
  int n, b_mn, ts, nacc, reps;
};

// one CTA, 160 threads: warps 0-3 idle helpers (TMEM alloc), warp 4 issues
__global__ void __launch_bounds__(160, 1) timing_kernel(Variant v, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  // A tile: 128 rows x 64 bf16 K-major at 0 (16 KB); B region at 32 KB (up to 256 rows K-major = 32 KB, or MN-major chunks 16 KB apart)
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;   // bf16 1.0
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 4 && elect_one()) {
    const uint64_t da = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t db = v.b_mn ? umma_smem_desc(smem_u32(smem) + 32768, 16384, 1024) : umma_smem_desc(smem_u32(smem) + 32768, 16, 1024);
    const uint32_t idesc = umma_idesc_bf16(v.n, 0, v.b_mn);
    const uint32_t bstep = v.b_mn ? 128u : 2u;
    const uint32_t tm_a = tmem + 448;
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {   // rep 0, 1 warm up; rep 2 is reported
      const long long t0 = clock64();
      for (int i = 0; i < v.reps; ++i) {
        const int kk = i & 3;
        const uint32_t acc = tmem + (uint32_t)((i % v.nacc) * v.n);
        if (v.ts) umma_bf16_ts(acc, tm_a + kk * 8, db + kk * bstep, idesc, 1u);
        else umma_bf16(acc, da + kk * 2, db + kk * bstep, idesc, 1u);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// same measurement with everything known at compile time and the issue loop fully unrolled (what the real kernels do):
// separates the tensor pipe's own per-instruction cost from the cost of the issuing thread's address arithmetic
template <int N, int BMN, int TS, int NACC, int REPS>
__global__ void __launch_bounds__(160, 1) timing_unrolled_kernel(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 4 && elect_one()) {
    const uint64_t da = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t db = BMN ? umma_smem_desc(smem_u32(smem) + 32768, 16384, 1024) : umma_smem_desc(smem_u32(smem) + 32768, 16, 1024);
    constexpr uint32_t idesc = umma_idesc_bf16(N, 0, BMN);
    constexpr uint32_t bstep = BMN ? 128u : 2u;
    const uint32_t tm_a = tmem + 448;
    uint32_t phase = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
#pragma unroll
      for (int i = 0; i < REPS; ++i) {
        constexpr int dummy = 0;
        const int kk = i & 3;
        const uint32_t acc = tmem + (uint32_t)((i % NACC) * N);
        if (TS) umma_bf16_ts(acc, tm_a + kk * 8, db + kk * bstep, idesc, 1u);
        else umma_bf16(acc, da + kk * 2, db + kk * bstep, idesc, 1u);
        (void)dummy;
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N, int BMN, int TS, int NACC, int REPS>
int run_unrolled(const char* name, long long* d_out) {
  const int SMEM = 100 * 1024;
  if (cudaFuncSetAttribute(timing_unrolled_kernel<N, BMN, TS, NACC, REPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) return 1;
  timing_unrolled_kernel<N, BMN, TS, NACC, REPS><<<1, 160, SMEM>>>(d_out);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: launch failed\n", name); return 1; }
  long long h[2];
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  printf("%-40s %10d %12.1f %12.1f   (unrolled)\n", name, REPS, (double)h[0] / REPS, (double)h[1] / REPS);
  return 0;
}

// TMEM read / write bandwidth: `nwarps` warps (warp w -> lane quarter w & 3, column interleave w >> 2) sweep 256 columns
// `reps` times with x16 loads (mode 0: wait after every load; mode 1: two loads in flight; mode 2: load + store back)
__global__ void __launch_bounds__(512, 1) tmem_bw_kernel(int nwarps, int mode, int reps, long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const int q = warp & 3, part = warp >> 2, nparts = nwarps / 4;
  const uint32_t base = tmem + ((uint32_t)(q * 32) << 16);
  uint32_t keep = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int rep = 0; rep < reps; ++rep) {
      if (mode == 1) {
        for (int sc = part; sc < 16; sc += 2 * nparts) {
          uint32_t a[16], b[16];
          tmem_ld_32x32b_x16(base + sc * 16, a);
          tmem_ld_32x32b_x16(base + ((sc + nparts) & 15) * 16, b);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) keep ^= a[j] + b[j];
        }
      } else {
        for (int sc = part; sc < 16; sc += nparts) {
          uint32_t a[16];
          tmem_ld_32x32b_x16(base + sc * 16, a);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) keep ^= a[j];
          if (mode == 2) {
            tmem_st_32x32b_x16(base + sc * 16, a);
            tmem_st_wait();
          }
        }
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  if (keep == 0x12345u) out[1] = keep;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

__host__ __device__ inline uint32_t sw_off(int r, int g) { return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((g ^ (r & 7)) << 4)); }
__host__ __device__ inline int ph_val(int r, int k) { return ((r * 7 + k * 3) % 7) - 3; }
__host__ __device__ inline int pl_val(int r, int k) { return ((r * 5 + k * 11) % 5) - 2; }
__host__ __device__ inline int vh_val(int k, int d) { return ((k * 3 + d * 5) % 7) - 3; }
__host__ __device__ inline int vl_val(int k, int d) { return ((k * 13 + d * 2) % 5) - 2; }

__device__ inline uint16_t bf16_bits(int v) { return (uint16_t)(__float_as_uint((float)v) >> 16); }   // small ints are exact

// functional: 64 keys.  packing 0: two K elements per 32-bit column (low half = even k); packing 1: one per column (low half)
__global__ void __launch_bounds__(160, 1) functional_kernel(int packing, float* out /* [128][128] */) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t plane = 272 * 128;   // as in the attention kernel: V_lo plane NP*128 bytes after V_hi
  // V planes, MN-major SWIZZLE_128B: row = key, 128 B = 64 head dims
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int k = i / 64, d = i % 64;
    const uint32_t off = sw_off(k, d >> 3) + (d & 7) * 2;
    *reinterpret_cast<uint16_t*>(smem + off) = bf16_bits(vh_val(k, d));
    *reinterpret_cast<uint16_t*>(smem + plane + off) = bf16_bits(vl_val(k, d));
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t TM_P = tmem, TM_D = tmem + 384;
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int g = 0; g < 4; ++g) {   // 16-key groups: columns [16g, 16g+8) = hi, [16g+8, 16g+16) = lo
      uint32_t v[16];
      if (packing == 0) {
        for (int j = 0; j < 8; ++j) {
          const int k = g * 16 + 2 * j;
          v[j] = (uint32_t)bf16_bits(ph_val(r, k)) | ((uint32_t)bf16_bits(ph_val(r, k + 1)) << 16);
          v[8 + j] = (uint32_t)bf16_bits(pl_val(r, k)) | ((uint32_t)bf16_bits(pl_val(r, k + 1)) << 16);
        }
        tmem_st_32x32b_x16(TM_P + lane_addr + g * 16, v);
      } else {   // one element per column: hi group in columns [32g, 32g+16), lo in [32g+16, 32g+32)
        for (int j = 0; j < 16; ++j) v[j] = bf16_bits(ph_val(r, g * 16 + j));
        tmem_st_32x32b_x16(TM_P + lane_addr + g * 32, v);
        for (int j = 0; j < 16; ++j) v[j] = bf16_bits(pl_val(r, g * 16 + j));
        tmem_st_32x32b_x16(TM_P + lane_addr + g * 32 + 16, v);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4 && elect_one()) {
    const uint64_t dv2 = umma_smem_desc(smem_u32(smem), plane, 1024);   // N = 128: chunk 0 = V_hi, chunk 1 = V_lo (LBO = plane)
    const uint64_t dv1 = umma_smem_desc(smem_u32(smem), 1024, 1024);    // N = 64: V_hi
    const uint32_t id2 = umma_idesc_bf16(128, 0, 1), id1 = umma_idesc_bf16(64, 0, 1);
    const int gs = packing == 0 ? 16 : 32, lo = packing == 0 ? 8 : 16;
    for (int kk = 0; kk < 4; ++kk) {
      umma_bf16_ts(TM_D, TM_P + kk * gs, dv2 + kk * 128, id2, kk > 0 ? 1u : 0u);          // [Ph Vh | Ph Vl]
      umma_bf16_ts(TM_D + 64, TM_P + kk * gs + lo, dv1 + kk * 128, id1, 1u);               // + Pl Vh on the right half
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    const int r = warp * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    for (int c = 0; c < 8; ++c) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(TM_D + lane_addr + c * 16, v);
      tmem_ld_wait();
      for (int j = 0; j < 16; ++j) out[r * 128 + c * 16 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

#define CK(x)                                                                        \
  do {                                                                               \
    cudaError_t e = (x);                                                             \
    if (e != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main() {
  const int SMEM = 100 * 1024;
  CK(cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  CK(cudaFuncSetAttribute(functional_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  long long* d_out;
  CK(cudaMalloc(&d_out, 16));
  struct Named { const char* name; Variant v; };
  const Named vs[] = {
      {"SS  N=256 B K-major  1 acc", {256, 0, 0, 1, 48}},
      {"SS  N=128 B K-major  1 acc", {128, 0, 0, 1, 48}},
      {"SS  N=64  B K-major  1 acc", {64, 0, 0, 1, 48}},
      {"SS  N=16  B K-major  1 acc", {16, 0, 0, 1, 48}},
      {"SS  N=64  B MN-major 1 acc", {64, 1, 0, 1, 48}},
      {"SS  N=64  B MN-major 2 acc", {64, 1, 0, 2, 48}},
      {"SS  N=64  B MN-major 4 acc", {64, 1, 0, 4, 48}},
      {"SS  N=128 B MN-major 1 acc", {128, 1, 0, 1, 48}},
      {"SS  N=128 B MN-major 2 acc", {128, 1, 0, 2, 48}},
      {"SS  N=256 B MN-major 1 acc", {256, 1, 0, 1, 48}},
      {"TS  N=64  B MN-major 1 acc", {64, 1, 1, 1, 48}},
      {"TS  N=64  B MN-major 4 acc", {64, 1, 1, 4, 48}},
      {"TS  N=128 B MN-major 1 acc", {128, 1, 1, 1, 48}},
      {"TS  N=128 B MN-major 2 acc", {128, 1, 1, 2, 48}},
      {"TS  N=256 B K-major  1 acc", {256, 0, 1, 1, 48}},
      {"TS  N=64  B K-major  1 acc", {64, 0, 1, 1, 48}},
      {"SS  N=64  B MN-major 1 acc, 12 MMAs", {64, 1, 0, 1, 12}},
      {"TS  N=64  B MN-major 1 acc, 12 MMAs", {64, 1, 1, 1, 12}},
      {"SS  N=256 B K-major  1 acc, 12 MMAs", {256, 0, 0, 1, 12}},
  };
  printf("%-40s %10s %12s %12s\n", "variant (M=128, K=16, bf16)", "MMAs", "issue clk/MMA", "total clk/MMA");
  for (const Named& nv : vs) {
    timing_kernel<<<1, 160, SMEM>>>(nv.v, d_out);
    CK(cudaDeviceSynchronize());
    long long h[2];
    CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
    printf("%-40s %10d %12.1f %12.1f\n", nv.name, nv.v.reps, (double)h[0] / nv.v.reps, (double)h[1] / nv.v.reps);
  }


  run_unrolled<256, 0, 0, 1, 48>("SS  N=256 B K-major  1 acc", d_out);
  run_unrolled<128, 0, 0, 1, 48>("SS  N=128 B K-major  1 acc", d_out);
  run_unrolled<64, 0, 0, 1, 48>("SS  N=64  B K-major  1 acc", d_out);
  run_unrolled<16, 0, 0, 1, 48>("SS  N=16  B K-major  1 acc", d_out);
  run_unrolled<64, 1, 0, 1, 48>("SS  N=64  B MN-major 1 acc", d_out);
  run_unrolled<64, 1, 0, 4, 48>("SS  N=64  B MN-major 4 acc", d_out);
  run_unrolled<128, 1, 0, 1, 48>("SS  N=128 B MN-major 1 acc", d_out);
  run_unrolled<256, 1, 0, 1, 48>("SS  N=256 B MN-major 1 acc", d_out);
  run_unrolled<64, 1, 1, 1, 48>("TS  N=64  B MN-major 1 acc", d_out);
  run_unrolled<64, 1, 1, 4, 48>("TS  N=64  B MN-major 4 acc", d_out);
  run_unrolled<128, 1, 1, 1, 48>("TS  N=128 B MN-major 1 acc", d_out);
  run_unrolled<256, 1, 1, 1, 48>("TS  N=256 B MN-major 1 acc", d_out);
  run_unrolled<256, 0, 1, 1, 48>("TS  N=256 B K-major  1 acc", d_out);
  run_unrolled<64, 0, 1, 1, 48>("TS  N=64  B K-major  1 acc", d_out);
  run_unrolled<64, 1, 0, 1, 12>("SS  N=64  B MN-major 1 acc, 12 MMAs", d_out);
  run_unrolled<64, 1, 1, 1, 12>("TS  N=64  B MN-major 1 acc, 12 MMAs", d_out);
  run_unrolled<128, 1, 1, 1, 12>("TS  N=128 B MN-major 1 acc, 12 MMAs", d_out);
  for (int mode = 0; mode < 3; ++mode)
    for (int nw = 4; nw <= 16; nw *= 2) {
      const int reps = 20;
      tmem_bw_kernel<<<1, 512>>>(nw, mode, reps, d_out);
      CK(cudaDeviceSynchronize());
      long long h[2];
      CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
      const double bytes = 128.0 * 256 * 4 * reps;
      printf("TMEM %s, %2d warps: %.1f B/clk (%lld clk for %d sweeps of 128 KB)\n",
             mode == 0 ? "ld x16, wait each      " : mode == 1 ? "ld x16, two in flight  " : "ld x16 + st x16 in place", nw, bytes / h[0], h[0], reps);
    }
  float* d_f;
  CK(cudaMalloc(&d_f, 128 * 128 * 4));
  std::vector<float> hf(128 * 128);
  for (int packing = 0; packing < 2; ++packing) {
    CK(cudaMemset(d_f, 0, 128 * 128 * 4));
    functional_kernel<<<1, 160, SMEM>>>(packing, d_f);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hf.data(), d_f, 128 * 128 * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    int bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int d = 0; d < 64; ++d) {
        int hh = 0, mix = 0;
        for (int k = 0; k < 64; ++k) {
          hh += ph_val(r, k) * vh_val(k, d);
          mix += ph_val(r, k) * vl_val(k, d) + pl_val(r, k) * vh_val(k, d);
        }
        const double e0 = fabs(hf[r * 128 + d] - hh), e1 = fabs(hf[r * 128 + 64 + d] - mix);
        if (e0 > 0 || e1 > 0) ++bad;
        worst = fmax(worst, fmax(e0, e1));
      }
    printf("functional (P in TMEM as A, B = [V_hi|V_lo] N=128 + V_hi N=64), packing %d: %s  (mismatching entries %d of 8192, max |err| %.1f)\n",
           packing, bad == 0 ? "EXACT" : "MISMATCH", bad, worst);
    if (bad) {
      printf("  row 1 got:      ");
      for (int d = 0; d < 8; ++d) printf("%7.1f", hf[128 + d]);
      printf(" | ");
      for (int d = 64; d < 72; ++d) printf("%7.1f", hf[128 + d]);
      printf("\n");
    }
  }
  return 0;
}
