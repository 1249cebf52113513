// semireward_b200 — shared device/host helpers (sm_100a only).
//
// Data convention used by every tensor-core operand in this library ("split planes"):
//   an fp32 matrix X[rows, cols] that feeds an MMA is stored as TWO bf16 planes, hi then lo, back to back:
//       hi = bf16_rn(X),  lo = bf16_rn(X - float(hi))          (|X - hi - lo| <= 2^-18 |X|)
//   planes[0 .. rows*cols) = hi, planes[rows*cols .. 2*rows*cols) = lo.  Same 4 bytes/element as fp32.
//   A product is evaluated as  hi*hi + hi*lo + lo*hi  with fp32 accumulation in TMEM ("bf16x3"), which keeps
//   ~16 mantissa bits per operand: that is what the reference's fp32 parity gate (1e-3 on logits/loss, bit-exact
//   masks; BASELINE.json north_star) needs and single-pass TF32/bf16 cannot give (SURVEY.md §7 "Hard parts").
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>

#include "../../include/srw.h"

struct CUtensorMap_st;

namespace srw {

// ------------------------------------------------------------------------------------------------
// error plumbing (C-ABI functions return 0 / negative code and never throw)
// ------------------------------------------------------------------------------------------------

void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define SRW_CUDA(expr)                                                    \
  do {                                                                    \
    cudaError_t _e = (expr);                                              \
    if (_e != cudaSuccess) return ::srw::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define SRW_LAUNCH_CHECK() SRW_CUDA(cudaGetLastError())

#define SRW_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      ::srw::set_last_error(__VA_ARGS__);      \
      return SRW_ERR_ARG;               \
    }                                          \
  } while (0)

// bench-only device timing of a kernel class (srw_profile_enable); no-ops when disabled
void* prof_begin(int cls, double flops, double bytes, cudaStream_t s);
void prof_end(void* handle, cudaStream_t s);

// Host side of PDL: launch `kernel` allowing it to overlap the tail of its stream predecessor (srw_set_pdl_mode / SRW_PDL=0
// turn the attribute off; the kernels' pdl_wait() is then a no-op).  Only for kernels that call pdl_wait().
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// split-plane helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// split two floats -> packed hi pair, packed lo pair (element 0 in the low half).  Two packed conversions
// (cvt.rn.bf16x2.f32), two bit ops and two subtractions: 3 instructions per element.
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float r0 = x0 - __uint_as_float(hi << 16);
  const float r1 = x1 - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0, r1);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float bf16_lo_f(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

__device__ __forceinline__ float plane_value(const __nv_bfloat16* __restrict__ planes, int64_t plane_stride, int64_t idx) {
  return __bfloat162float(planes[idx]) + __bfloat162float(planes[idx + plane_stride]);
}

// erf-GELU (nn.GELU default; reference vit.py:223) and its derivative.  erf is evaluated branch-free with Abramowitz-Stegun
// 7.1.26 (|error| < 5e-7 in fp32, about twice the fp32 rounding of gelu itself and far below the 1e-3 parity gate): the
// fused GEMM epilogues are instruction-bound, and libdevice erff costs ~3x as many instructions.  exp(-x^2/2) is shared by
// the erf tail and the Gaussian density of the derivative.
// The fused GEMM epilogues are instruction-issue bound (ncu source view: ~40 SASS instructions per element before this
// form), so the special functions are the raw MUFU approximations (rcp.approx / ex2.approx, <= 2 ulp, no IEEE slow-path
// branches or denormal fix-ups) and every constant is folded: 4 FMUL + 7 FFMA + 2 MUFU per gelu().
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// erf_abs = erf(|x| / sqrt 2) in [0, 1), gauss = exp(-x^2 / 2)
__device__ __forceinline__ void gelu_parts(float x, float& erf_abs, float& gauss) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, fabsf(x), 1.0f));
  gauss = ex2_approx((x * x) * (-0.5f * 1.4426950408889634f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  erf_abs = fmaf(-(poly * t), gauss, 1.0f);
}
__device__ __forceinline__ float gelu_f(float x) {
  float e, g;
  gelu_parts(x, e, g);
  const float hx = 0.5f * x;
  return fmaf(fabsf(hx), e, hx);                            // 0.5 x (1 + sign(x) erf(|x|/sqrt 2))
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  float e, g;
  gelu_parts(x, e, g);
  return fmaf(x * 0.39894228040143267794f, g, fmaf(0.5f, copysignf(e, x), 0.5f));
}

// ------------------------------------------------------------------------------------------------
// counter-based dropout (srw_dropout, include/srw.h): element `idx` of dropout site `site` of the sequence with stream key
// `seq_key` is KEPT iff the top 24 bits of lowbias32(idx + site_key) are below keep * 2^24.  No RNG state: the forward
// epilogues and every backward kernel regenerate the same bits, and the CPU oracle (oracle/bert_oracle.py) restates it.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t drop_site_key(uint32_t seq_key, uint32_t site) { return lowbias32(seq_key + site * 0x9E3779B9U); }
__device__ __forceinline__ bool drop_kept(uint32_t site_key, uint32_t idx, uint32_t thr24) { return (lowbias32(idx + site_key) >> 8) < thr24; }
struct DropParams {          // kernel-side view of srw_dropout
  const uint32_t* seq_key; const int32_t* seq_row;
  uint32_t site, thr24; float inv_keep; int on;
};
static inline DropParams make_drop(const srw_dropout& d) {
  DropParams p;
  p.seq_key = d.seq_key; p.seq_row = d.seq_row; p.site = d.site;
  p.on = (d.p > 0.0 && d.seq_key != nullptr) ? 1 : 0;
  const double keep = 1.0 - d.p;
  p.thr24 = (uint32_t)(keep * 16777216.0);
  p.inv_keep = p.on ? 1.0f / (float)keep : 1.0f;
  return p;
}

// ------------------------------------------------------------------------------------------------
// warp / block reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA, tcgen05 (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a descriptor / phase bug must surface as a trapped kernel (an error code at the next sync),
// never as a hung GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("srw: mbarrier wait timeout (block %d,%d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Programmatic dependent launch (PDL).  Kernels of the ViT engine are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: pdl_trigger() at the top lets the NEXT kernel of the stream be
// scheduled as soon as every CTA of this one has started, so its launch latency and prologue (barrier init, TMEM
// allocation, descriptor prefetch) hide under this kernel's tail; pdl_wait() blocks until the PREVIOUS kernel has
// completed and its writes are visible, and must precede the first global-memory access.  Both are no-ops for launches
// without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (bf16 in, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A is read from tensor memory (lane == row, 32-bit column j holds K elements 2j
// (low half) and 2j+1, so a K = 16 step is 8 columns).  scripts/mma_probe.cu: an M128 N64 K16 MMA costs 35 clk this way
// against 50 clk with A from shared memory (the 4 KB A read is what bounds the small-N shapes).
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t), columns [col, col+32)
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 lanes x 16 consecutive 32-bit columns, registers -> tensor memory (thread t writes row lane base + t)
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor (sm_100 format, version=1), SWIZZLE_128B.
//   start address >>4 in [0,14), LBO>>4 in [16,30), SBO>>4 in [32,46), version(=1) in [46,48), layout type in [61,64).
// K-major tile  [rows][64 bf16] (128 B rows, 8-row / 1024 B swizzle atoms stacked along rows):
//   LBO ignored (1), SBO = 1024 B between 8-row groups; advancing 16 elements along K adds 32 B to the start address.
// MN-major tile [k rows][64 bf16 of MN] (128 B per k row, 8 k-rows per 1024 B atom), 64-wide MN chunks `lbo_bytes` apart:
//   LBO = distance between MN chunks, SBO = 1024 B between 8-k groups; advancing 16 along K adds 2048 B.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for Blackwell
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// instruction descriptor for kind::f16, bf16 x bf16 -> fp32, M=128
int make_plane_tmap(::CUtensorMap_st* out, const void* base, int64_t inner, int64_t outer, int64_t ld, int64_t plane_stride,
                    int box_outer, int box_planes);

int gemm_pick_bn(int M, int N, int splits, int ctas = 148);
int gemm2_pick_bn(int M, int N, int splits, bool b_mn, int ctas = 148, int k_per_cta = 0);

__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

}  // namespace srw
