// srw_gemm — D[M,N] = A[M,K] * B[N,K]^T for the reference's nn.Linear call sites (vit.py:93,105,70,73) and their
// dgrad / wgrad, on sm_100a tensor cores:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared memory -> tcgen05.mma kind::f16 (bf16 x bf16, fp32 accumulate in
//   TMEM) -> tcgen05.ld -> fused epilogue -> global.
// Operands are "split planes" (srw_common.cuh): per 64-wide K block the kernel stages A_hi, A_lo, B_hi, B_lo once and
// issues three MMAs (hi*hi, hi*lo, lo*hi) into the same accumulator.
//
// Persistent kernel: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x + i*gridDim.x (n fastest, so the CTAs
// running together share A row-blocks and the whole weight matrix through L2).  Tile = 128 x BN, BN in {192, 128, 64}
// (192 divides every ViT width: 384/1152/1536, 768/2304/3072).  256 threads:
//   warp 0  : TMA producer (one elected lane)          warp 2 : TMEM allocator / deallocator
//   warp 1  : MMA issuer  (one elected lane)           warps 4-15: epilogue (12 warps: the fused epilogues — erf GELU, hi/lo
//                                                       split, residual — are instruction-bound, not memory-bound)
// The fp32 accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i (TMEM -> registers ->
// fused op -> global) overlaps the MMAs of tile i+1.  One accumulator per tile; the host keeps K per CTA <= 4096 (split-K
// beyond) because the tensor core's fp32 accumulate truncates and long single-accumulator sums lose the cross terms.
// A SIMT kernel over the same operands and the same epilogue code is kept as the on-device verification twin
// (srw_gemm_args.impl = SRW_GEMM_SIMT); it is a debugging aid, never the default path.
#include <cuda.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {

extern std::atomic<int64_t> g_launches;

constexpr int BM = 128, BK = 64;
constexpr int PLANE_TILE_BYTES = 128 * BK * 2;         // 16 KiB: one 128 x 64 bf16 tile
constexpr int A_STAGE_BYTES = 2 * PLANE_TILE_BYTES;    // A_hi, A_lo
constexpr int TMEM_COLS = 512;                         // two accumulator buffers of 256 columns
constexpr int MAX_K_PER_CTA = 4096;   // ViT-B / BERT-base fc2: K = 3072
__host__ __device__ constexpr int gemm_stage_bytes(int bn) { return A_STAGE_BYTES + bn * BK * 2 * 2; }
__host__ __device__ constexpr int gemm_stages(int bn) { return bn == 64 ? 3 : 2; }
constexpr int EPI_STAGE_LD = 36;                         // floats per staged row (32 + 4 pad: 16 B aligned, conflict-free)
constexpr int EPI_WARPS = 12;                            // 3 warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int GEMM_THREADS = 128 + 32 * EPI_WARPS;
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * EPI_STAGE_LD * 4;  // one 32 x 32 fp32 transpose buffer per epilogue warp
__host__ __device__ constexpr int gemm_smem_bytes(int bn) {
  return gemm_stages(bn) * gemm_stage_bytes(bn) + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_STAGE_BYTES;
}

unsigned long long* g_gemm_trace = nullptr;

struct EpiParams {
  int M, N;
  int epilogue;
  const float* bias;
  const float* resid; int64_t ldr;
  const float* row_scale; int rows_per_scale;
  const float* aux; int64_t ldaux;
  float* out_f32; int64_t ldo;
  __nv_bfloat16* out_planes; int64_t ldp; int64_t out_plane_stride;
  float* workspace;  // split-K partials [split, M, N]
  DropParams drop; int drop_rows_per_seq;   // SRW_EPI_RESID: dropout between the dense layer and the residual add
};

// Apply the fused epilogue to NV consecutive accumulator columns of one row.  NV is 32 (tcgen05 path: one tcgen05.ld
// chunk) or 4 (SIMT twin); col0 % NV == 0, N % 4 == 0 (checked on the host).
template <int NV>
__device__ __forceinline__ void epilogue_store(const EpiParams& p, int row, int col0, float (&v)[NV], int split) {
  if (row >= p.M) return;
  const int epi = p.epilogue;
  if (epi == SRW_EPI_SPLITK) {
    float* dst = p.workspace + ((int64_t)split * p.M + row) * p.N + col0;
#pragma unroll
    for (int j = 0; j < NV; j += 4)
      if (col0 + j < p.N) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    return;
  }
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < NV; j += 4)
      if (col0 + j < p.N) {
        const float4 b = *reinterpret_cast<const float4*>(p.bias + col0 + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
  }
  if (epi == SRW_EPI_RESID) {
    const float s = p.row_scale ? p.row_scale[row / p.rows_per_scale] : 1.0f;
    const float* r = p.resid + (int64_t)row * p.ldr + col0;
    float* dst = p.out_f32 + (int64_t)row * p.ldo + col0;
    if (p.drop.on) {
      const int sq = row / p.drop_rows_per_seq;
      const uint32_t key = drop_site_key(p.drop.seq_key[sq], p.drop.site);
      const uint32_t base = ((uint32_t)p.drop.seq_row[sq] * (uint32_t)p.drop_rows_per_seq + (uint32_t)(row - sq * p.drop_rows_per_seq)) * (uint32_t)p.N + (uint32_t)col0;
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] = drop_kept(key, base + j, p.drop.thr24) ? v[j] * p.drop.inv_keep : 0.f;
    }
#pragma unroll
    for (int j = 0; j < NV; j += 4)
      if (col0 + j < p.N) {
        const float4 rr = *reinterpret_cast<const float4*>(r + j);
        *reinterpret_cast<float4*>(dst + j) =
            make_float4(rr.x + s * v[j], rr.y + s * v[j + 1], rr.z + s * v[j + 2], rr.w + s * v[j + 3]);
      }
    return;
  }
  if (epi == SRW_EPI_F32 || epi == SRW_EPI_GELU) {
    float* dst = p.out_f32 + (int64_t)row * p.ldo + col0;
#pragma unroll
    for (int j = 0; j < NV; j += 4)
      if (col0 + j < p.N) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    if (epi == SRW_EPI_F32) return;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = gelu_f(v[j]);
  }
  if (epi == SRW_EPI_DGELU) {
    const float* z = p.aux + (int64_t)row * p.ldaux + col0;
#pragma unroll
    for (int j = 0; j < NV; j += 4)
      if (col0 + j < p.N) {
        const float4 zz = *reinterpret_cast<const float4*>(z + j);
        v[j] *= gelu_grad_f(zz.x); v[j + 1] *= gelu_grad_f(zz.y); v[j + 2] *= gelu_grad_f(zz.z); v[j + 3] *= gelu_grad_f(zz.w);
      }
  }
  if (p.drop.on) {   // activation dropout (HubertFeedForward.intermediate_dropout): h = drop(gelu(z)); backward: the same mask on the incoming gradient
    const int sq = row / p.drop_rows_per_seq;
    const uint32_t key = drop_site_key(p.drop.seq_key[sq], p.drop.site);
    const uint32_t base = ((uint32_t)p.drop.seq_row[sq] * (uint32_t)p.drop_rows_per_seq + (uint32_t)(row - sq * p.drop_rows_per_seq)) * (uint32_t)p.N + (uint32_t)col0;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = drop_kept(key, base + j, p.drop.thr24) ? v[j] * p.drop.inv_keep : 0.f;
  }
  // planes out (SRW_EPI_PLANES, SRW_EPI_GELU, SRW_EPI_DGELU)
  __nv_bfloat16* hi = p.out_planes + (int64_t)row * p.ldp + col0;
  __nv_bfloat16* lo = hi + p.out_plane_stride;
#pragma unroll
  for (int j = 0; j < NV; j += 4)
    if (col0 + j < p.N) {
      uint32_t h0, l0, h1, l1;
      split2(v[j], v[j + 1], h0, l0);
      split2(v[j + 2], v[j + 3], h1, l1);
      *reinterpret_cast<uint2*>(hi + j) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(lo + j) = make_uint2(l0, l1);
    }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 kernel
// ------------------------------------------------------------------------------------------------
// Epilogue of one tile for one warp.  TMEM gives each lane one accumulator ROW; global memory wants lanes along the
// COLUMNS.  Every 32 x 32 chunk is transposed through a padded per-warp smem buffer so that 8 lanes cover one 128 B row
// segment (4 rows per access): all bias / residual / aux reads and all stores are fully coalesced.
//
// Interior tiles take drain_full<>: the epilogue kind is a template parameter and there is no bounds logic, so the eight
// row-group iterations of a chunk are straight-line code; the residual / aux operands of all eight are requested BEFORE
// the TMEM load and the transpose, which hides their L2 latency (the epilogues are latency- and issue-bound, not
// bandwidth-bound: 12 warps, 3 per scheduler).
template <int BN, int EPI, bool ACTDROP>
__device__ __forceinline__ void drain_full(const EpiParams& ep, uint32_t acc_addr, float* st, int q, int part, int lane, int m0, int n0, int split,
                                           int nchunks = BN / 32) {
  const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
  const int row_first = m0 + q * 32 + sub_r;       // this lane's rows: row_first + 4 i
  float scale[8];
  uint32_t dkey[8], dbase[8];
  // RESID: run-time switch (BERT / HuBERT output dropout).  GELU / DGELU (HuBERT's activation dropout): only in the ACTDROP kernels —
  // the default kernels keep exactly the code of the drop-free epilogues (mask state in registers cost them spills and 2.6 % of a ViT step)
  const bool dropping = (EPI == SRW_EPI_RESID && ep.drop.on) || (ACTDROP && (EPI == SRW_EPI_GELU || EPI == SRW_EPI_DGELU));
  if (EPI == SRW_EPI_RESID) {
#pragma unroll
    for (int i = 0; i < 8; ++i) scale[i] = ep.row_scale ? ep.row_scale[(row_first + 4 * i) / ep.rows_per_scale] : 1.0f;
  }
  if (dropping) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_first + 4 * i, sq = row / ep.drop_rows_per_seq;
      dkey[i] = drop_site_key(ep.drop.seq_key[sq], ep.drop.site);
      dbase[i] = ((uint32_t)ep.drop.seq_row[sq] * (uint32_t)ep.drop_rows_per_seq + (uint32_t)(row - sq * ep.drop_rows_per_seq)) * (uint32_t)ep.N;
    }
  }
#pragma unroll 1
  for (int c = part; c < nchunks; c += EPI_WARPS / 4) {
    const int col = n0 + c * 32 + sub_c;
    float4 pre[8];
    if (EPI == SRW_EPI_RESID) {
      const float* r = ep.resid + (int64_t)row_first * ep.ldr + col;
#pragma unroll
      for (int i = 0; i < 8; ++i) pre[i] = *reinterpret_cast<const float4*>(r + (int64_t)(4 * i) * ep.ldr);
    } else if (EPI == SRW_EPI_DGELU) {
      const float* z = ep.aux + (int64_t)row_first * ep.ldaux + col;
#pragma unroll
      for (int i = 0; i < 8; ++i) pre[i] = *reinterpret_cast<const float4*>(z + (int64_t)(4 * i) * ep.ldaux);
    }
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EPI != SRW_EPI_SPLITK && ep.bias != nullptr) b4 = *reinterpret_cast<const float4*>(ep.bias + col);
    uint32_t r[32];
    tmem_ld_32x32b_x32(acc_addr + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(st + lane * EPI_STAGE_LD + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_first + 4 * i;
      float4 x = *reinterpret_cast<const float4*>(st + (4 * i + sub_r) * EPI_STAGE_LD + sub_c);
      x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
      if (EPI == SRW_EPI_SPLITK) {
        *reinterpret_cast<float4*>(ep.workspace + ((int64_t)split * ep.M + row) * ep.N + col) = x;
        continue;
      }
      if (EPI == SRW_EPI_RESID) {
        const float sc = scale[i];
        if (dropping) {
          const uint32_t b0 = dbase[i] + (uint32_t)col;
          x.x = drop_kept(dkey[i], b0, ep.drop.thr24) ? x.x * ep.drop.inv_keep : 0.f;
          x.y = drop_kept(dkey[i], b0 + 1, ep.drop.thr24) ? x.y * ep.drop.inv_keep : 0.f;
          x.z = drop_kept(dkey[i], b0 + 2, ep.drop.thr24) ? x.z * ep.drop.inv_keep : 0.f;
          x.w = drop_kept(dkey[i], b0 + 3, ep.drop.thr24) ? x.w * ep.drop.inv_keep : 0.f;
        }
        *reinterpret_cast<float4*>(ep.out_f32 + (int64_t)row * ep.ldo + col) =
            make_float4(fmaf(sc, x.x, pre[i].x), fmaf(sc, x.y, pre[i].y), fmaf(sc, x.z, pre[i].z), fmaf(sc, x.w, pre[i].w));
        continue;
      }
      if (EPI == SRW_EPI_F32 || EPI == SRW_EPI_GELU) {
        *reinterpret_cast<float4*>(ep.out_f32 + (int64_t)row * ep.ldo + col) = x;
        if (EPI == SRW_EPI_F32) continue;
        x.x = gelu_f(x.x); x.y = gelu_f(x.y); x.z = gelu_f(x.z); x.w = gelu_f(x.w);
      }
      if (EPI == SRW_EPI_DGELU) {
        x.x *= gelu_grad_f(pre[i].x); x.y *= gelu_grad_f(pre[i].y); x.z *= gelu_grad_f(pre[i].z); x.w *= gelu_grad_f(pre[i].w);
      }
      if ((EPI == SRW_EPI_GELU || EPI == SRW_EPI_DGELU) && dropping) {
        const uint32_t b0 = dbase[i] + (uint32_t)col;
        x.x = drop_kept(dkey[i], b0, ep.drop.thr24) ? x.x * ep.drop.inv_keep : 0.f;
        x.y = drop_kept(dkey[i], b0 + 1, ep.drop.thr24) ? x.y * ep.drop.inv_keep : 0.f;
        x.z = drop_kept(dkey[i], b0 + 2, ep.drop.thr24) ? x.z * ep.drop.inv_keep : 0.f;
        x.w = drop_kept(dkey[i], b0 + 3, ep.drop.thr24) ? x.w * ep.drop.inv_keep : 0.f;
      }
      uint32_t h0, l0, h1, l1;
      split2(x.x, x.y, h0, l0);
      split2(x.z, x.w, h1, l1);
      __nv_bfloat16* hp = ep.out_planes + (int64_t)row * ep.ldp + col;
      *reinterpret_cast<uint2*>(hp) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(hp + ep.out_plane_stride) = make_uint2(l0, l1);
    }
    __syncwarp();
  }
}

template <int BN, bool ACTDROP>
__device__ __forceinline__ void drain_accumulator(const EpiParams& ep, uint32_t acc_addr, float* st, int q, int part, int lane, int m0, int n0,
                                                  int split, bool has_k) {
  // full rows and a whole number of 32-column chunks (the last N tile of a narrow output, e.g. the 32- and 128-channel convolutions
  // of srw_wrn.cu under a 64- / 192-wide tile, takes the straight-line path over the chunks it has)
  const int ncols = min(BN, ep.N - n0);
  if (has_k && m0 + BM <= ep.M && (ncols & 31) == 0) {
    const int nch = ncols >> 5;
    switch (ep.epilogue) {
      case SRW_EPI_F32: drain_full<BN, SRW_EPI_F32, false>(ep, acc_addr, st, q, part, lane, m0, n0, split, nch); return;
      case SRW_EPI_PLANES: drain_full<BN, SRW_EPI_PLANES, false>(ep, acc_addr, st, q, part, lane, m0, n0, split, nch); return;
      case SRW_EPI_GELU: drain_full<BN, SRW_EPI_GELU, ACTDROP>(ep, acc_addr, st, q, part, lane, m0, n0, split, nch); return;
      case SRW_EPI_RESID: drain_full<BN, SRW_EPI_RESID, false>(ep, acc_addr, st, q, part, lane, m0, n0, split, nch); return;
      case SRW_EPI_DGELU: drain_full<BN, SRW_EPI_DGELU, ACTDROP>(ep, acc_addr, st, q, part, lane, m0, n0, split, nch); return;
      default: drain_full<BN, SRW_EPI_SPLITK, false>(ep, acc_addr, st, q, part, lane, m0, n0, split, nch); return;
    }
  }
  // edge tiles (partial in M or N) and empty K ranges: generic bounds-checked path
  const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
#pragma unroll 1
  for (int c = part; c < BN / 32; c += EPI_WARPS / 4) {
    if (n0 + c * 32 >= ep.N) break;
    if (has_k) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(acc_addr + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(st + lane * EPI_STAGE_LD + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(st + lane * EPI_STAGE_LD + j) = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
      const int rr = i * 4 + sub_r;
      const float4 x = *reinterpret_cast<const float4*>(st + rr * EPI_STAGE_LD + sub_c);
      float v[4] = {x.x, x.y, x.z, x.w};
      epilogue_store<4>(ep, m0 + q * 32 + rr, n0 + c * 32 + sub_c, v, split);
    }
    __syncwarp();
  }
}

// Debug timeline (srw_gemm_set_trace): per CTA 16 clock64() stamps — 0 entry, 1 setup done, 2 first TMA issued, 3 first stage
// landed, 4+2i accumulator of the CTA's i-th tile complete, 5+2i its epilogue done, 15 exit.  NULL in production.
struct TcParams {
  unsigned long long* trace;
  int K;            // reduction length
  int kb_per_split; // k blocks (of 64) handled by one split
  int a_mn, b_mn;   // operand majors
  int m_tiles, n_tiles, splits;
  int a_seg_kb;     // > 0: K-major A is read in K segments of a_seg_kb blocks; segment s starts a_seg_rows * s rows further down
  int a_seg_rows;   //      (implicit-GEMM 3x3 convolution over a zero-bordered NHWC tensor, srw_wrn.cu)
};

template <int BN, bool ACTDROP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16x3_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                           const TcParams tp, const EpiParams ep) {
  constexpr int STAGES = gemm_stages(BN);
  constexpr int STAGE_BYTES = gemm_stage_bytes(BN);
  constexpr int B_PLANE_BYTES = BN * BK * 2;            // one K-major B plane
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);

  pdl_trigger();
  unsigned long long* tr = tp.trace ? tp.trace + (size_t)blockIdx.x * 16 : nullptr;
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb_total = (tp.K + BK - 1) / BK;
  const int tiles_mn = tp.m_tiles * tp.n_tiles;
  const int total_tiles = tiles_mn * tp.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], EPI_WARPS);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above is CTA-local; operands / epilogue inputs come from the previous kernel
  if (tr && threadIdx.x == 0) tr[1] = clock64();

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      uint32_t it = 0;   // running k-block counter across tiles -> stage / phase
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int split = t / tiles_mn, rem = t % tiles_mn;
        const int m0 = (rem / tp.n_tiles) * BM, n0 = (rem % tp.n_tiles) * BN;
        const int kb_begin = split * tp.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + tp.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
          if (tr && it == 0) tr[2] = clock64();
          const int k0 = kb * BK;
          if (!tp.a_mn) {
            if (tp.a_seg_kb > 0) {
              const int seg = kb / tp.a_seg_kb;
              tma_load_3d(st, &tmap_a, &full_bar[s], (kb - seg * tp.a_seg_kb) * BK, m0 + seg * tp.a_seg_rows, 0);
            } else {
              tma_load_3d(st, &tmap_a, &full_bar[s], k0, m0, 0);                     // [2][128][64]
            }
          } else {
            tma_load_3d(st, &tmap_a, &full_bar[s], m0, k0, 0);                       // [2][64 k][64 mn] chunk 0
            tma_load_3d(st + PLANE_TILE_BYTES, &tmap_a, &full_bar[s], m0 + 64, k0, 0);  // chunk 1
          }
          uint8_t* sb = st + A_STAGE_BYTES;
          if (!tp.b_mn) {
            tma_load_3d(sb, &tmap_b, &full_bar[s], k0, n0, 0);                       // [2][BN][64]
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_3d(sb + c * PLANE_TILE_BYTES, &tmap_b, &full_bar[s], n0 + c * 64, k0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_bf16(BN, tp.a_mn, tp.b_mn);
      // per-operand descriptor geometry
      const uint32_t a_lo_off = tp.a_mn ? (PLANE_TILE_BYTES / 2) : PLANE_TILE_BYTES;  // lo plane offset inside the operand
      const uint32_t b_lo_off = tp.b_mn ? (PLANE_TILE_BYTES / 2) : B_PLANE_BYTES;
      const uint32_t a_lbo = tp.a_mn ? PLANE_TILE_BYTES : 16, b_lbo = tp.b_mn ? PLANE_TILE_BYTES : 16;
      const uint32_t a_kstep = (tp.a_mn ? 2048 : 32) >> 4, b_kstep = (tp.b_mn ? 2048 : 32) >> 4;
      // base descriptors of stage 0, built once: the issuing thread must not spend more cycles rebuilding descriptors
      // than the MMAs take (BN = 128: 64 clk each); a stage / K step only adds (bytes >> 4) to the start-address field
      const uint32_t s0 = smem_u32(smem);
      const uint64_t dA_hi = umma_smem_desc(s0, a_lbo, 1024), dA_lo = umma_smem_desc(s0 + a_lo_off, a_lbo, 1024);
      const uint64_t dB_hi = umma_smem_desc(s0 + A_STAGE_BYTES, b_lbo, 1024), dB_lo = umma_smem_desc(s0 + A_STAGE_BYTES + b_lo_off, b_lbo, 1024);
      uint32_t it = 0, tile_iter = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tile_iter) {
        const int split = t / tiles_mn;
        const int kb_begin = split * tp.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + tp.kb_per_split);
        const uint32_t as = tile_iter & 1;
        mbar_wait(&acc_empty[as], ((tile_iter >> 1) & 1) ^ 1);   // epilogue has drained this accumulator buffer
        tc_fence_after();
        const uint32_t acc_addr = tmem_base + as * 256;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          if (tr && it == 0) tr[3] = clock64();
          const uint32_t so = s * (STAGE_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t a_hi = dA_hi + so + kk * a_kstep, a_lo = dA_lo + so + kk * a_kstep;
            const uint64_t b_hi = dB_hi + so + kk * b_kstep, b_lo = dB_lo + so + kk * b_kstep;
            umma_bf16(acc_addr, a_lo, b_hi, idesc, (kb > kb_begin || kk > 0) ? 1u : 0u);  // small terms first
            umma_bf16(acc_addr, a_hi, b_lo, idesc, 1u);
            umma_bf16(acc_addr, a_hi, b_hi, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs have read it
        }
        umma_commit(&acc_full[as]);    // accumulator of this tile complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int q = warp & 3;           // TMEM lane quarter this warp may access (hardware rule: warp id % 4)
    const int part = (warp - 4) >> 2;  // which of the EPI_WARPS/4 column interleaves
    uint32_t tile_iter = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tile_iter) {
      const int split = t / tiles_mn, rem = t % tiles_mn;
      const int m0 = (rem / tp.n_tiles) * BM, n0 = (rem % tp.n_tiles) * BN;
      const int kb_begin = split * tp.kb_per_split;
      const bool has_k = kb_begin < num_kb_total;
      const uint32_t as = tile_iter & 1;
      mbar_wait(&acc_full[as], (tile_iter >> 1) & 1);
      tc_fence_after();
      if (tr && warp == 4 && lane == 0 && tile_iter < 5) tr[4 + 2 * tile_iter] = clock64();
      drain_accumulator<BN, ACTDROP>(ep, tmem_base + as * 256, epi_stage + (warp - 4) * (32 * EPI_STAGE_LD), q, part, lane, m0, n0, split, has_k);
      tc_fence_before();
      __syncwarp();
      if (tr && warp == 4 && lane == 0 && tile_iter < 5) tr[5 + 2 * tile_iter] = clock64();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tr && threadIdx.x == 0) tr[15] = clock64();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// 2-CTA kernel: a CTA pair (cluster of 2 on one TPC) computes a 256 x BN tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows of N); the pair's MMA reads both halves, so
// B crosses L2 -> SM once per pair instead of once per CTA, and a stage shrinks from 32+BN/4 KB... to 32 KB + BN*128 B,
// which buys a third pipeline stage.  Roles per CTA as above; only the leader CTA (rank 0) issues MMAs, its commits
// arrive (multicast) on both CTAs' barriers; both CTAs' TMA loads complete on the LEADER's full barrier.
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int gemm2_stage_bytes(int bn) { return A_STAGE_BYTES + (bn / 2) * BK * 2 * 2; }
__host__ __device__ constexpr int gemm2_stages(int bn) { return bn == 256 ? 2 : 3; }
__host__ __device__ constexpr int gemm2_smem_bytes(int bn) { return gemm2_stages(bn) * gemm2_stage_bytes(bn) + 1024 + 256 + EPI_STAGE_BYTES; }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to the mbarrier at `bar_cluster_addr` (the leader CTA's barrier)
__device__ __forceinline__ void tma_load_3d_2cta(void* smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this smem offset in BOTH CTAs of the pair once all prior MMAs of this thread are complete
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16_m256(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

template <int BN, bool ACTDROP>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_bf16x3_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TcParams tp,
                            const EpiParams ep) {
  constexpr int STAGES = gemm2_stages(BN);
  constexpr int STAGE_BYTES = gemm2_stage_bytes(BN);
  constexpr int BH = BN / 2;                            // rows of B staged by each CTA
  constexpr int B_PLANE_BYTES = BH * BK * 2;            // one K-major plane of the half tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;   // [2]
  uint64_t* acc_empty = acc_full + 2;        // [2]  (waited on by the leader only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* epi_stage = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);

  pdl_trigger();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_kb_total = (tp.K + BK - 1) / BK;
  const int tiles_mn = tp.m_tiles * tp.n_tiles;       // m_tiles counts 256-row pair tiles here
  const int total_tiles = tiles_mn * tp.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 2 * EPI_WARPS);   // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta(tmem_slot, TMEM_COLS);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = pair; t < total_tiles; t += num_pairs) {
        const int split = t / tiles_mn, rem = t % tiles_mn;
        const int m0 = (rem / tp.n_tiles) * 256 + (int)rank * BM, n0 = (rem % tp.n_tiles) * BN + (int)rank * BH;
        const int kb_begin = split * tp.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + tp.kb_per_split);
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* st = smem + s * STAGE_BYTES;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);   // this CTA's bytes + the peer's
          const int k0 = kb * BK;
          if (!tp.a_mn) {
            tma_load_3d_2cta(st, &tmap_a, leader_full, k0, m0, 0);
          } else {
            tma_load_3d_2cta(st, &tmap_a, leader_full, m0, k0, 0);
            tma_load_3d_2cta(st + PLANE_TILE_BYTES, &tmap_a, leader_full, m0 + 64, k0, 0);
          }
          uint8_t* sb = st + A_STAGE_BYTES;
          if (!tp.b_mn) {
            tma_load_3d_2cta(sb, &tmap_b, leader_full, k0, n0, 0);                    // [2][BH][64]
          } else {
#pragma unroll
            for (int c = 0; c < BH / 64; ++c) tma_load_3d_2cta(sb + c * PLANE_TILE_BYTES, &tmap_b, leader_full, n0 + c * 64, k0, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_bf16_m256(BN, tp.a_mn, tp.b_mn);
      const uint32_t a_lo_off = tp.a_mn ? (PLANE_TILE_BYTES / 2) : PLANE_TILE_BYTES;
      const uint32_t b_lo_off = tp.b_mn ? (PLANE_TILE_BYTES / 2) : B_PLANE_BYTES;
      const uint32_t a_lbo = tp.a_mn ? PLANE_TILE_BYTES : 16, b_lbo = tp.b_mn ? PLANE_TILE_BYTES : 16;
      const uint32_t a_kstep = (tp.a_mn ? 2048 : 32) >> 4, b_kstep = (tp.b_mn ? 2048 : 32) >> 4;
      const uint32_t s0 = smem_u32(smem);   // stage-0 base descriptors built once (see the 1-CTA kernel)
      const uint64_t dA_hi = umma_smem_desc(s0, a_lbo, 1024), dA_lo = umma_smem_desc(s0 + a_lo_off, a_lbo, 1024);
      const uint64_t dB_hi = umma_smem_desc(s0 + A_STAGE_BYTES, b_lbo, 1024), dB_lo = umma_smem_desc(s0 + A_STAGE_BYTES + b_lo_off, b_lbo, 1024);
      uint32_t it = 0, tile_iter = 0;
      for (int t = pair; t < total_tiles; t += num_pairs, ++tile_iter) {
        const int split = t / tiles_mn;
        const int kb_begin = split * tp.kb_per_split;
        const int kb_end = min(num_kb_total, kb_begin + tp.kb_per_split);
        const uint32_t as = tile_iter & 1;
        mbar_wait(&acc_empty[as], ((tile_iter >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc_addr = tmem_base + as * 256;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t so = s * (STAGE_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t a_hi = dA_hi + so + kk * a_kstep, a_lo = dA_lo + so + kk * a_kstep;
            const uint64_t b_hi = dB_hi + so + kk * b_kstep, b_lo = dB_lo + so + kk * b_kstep;
            umma_bf16_2cta(acc_addr, a_lo, b_hi, idesc, (kb > kb_begin || kk > 0) ? 1u : 0u);
            umma_bf16_2cta(acc_addr, a_hi, b_lo, idesc, 1u);
            umma_bf16_2cta(acc_addr, a_hi, b_hi, idesc, 1u);
          }
          umma_commit_2cta(&empty_bar[s]);   // both CTAs' producers may refill this stage
        }
        umma_commit_2cta(&acc_full[as]);     // both CTAs' epilogues may drain this accumulator
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs, each drains its own 128 accumulator rows) =====
    const int q = warp & 3;
    const int part = (warp - 4) >> 2;
    uint32_t tile_iter = 0;
    for (int t = pair; t < total_tiles; t += num_pairs, ++tile_iter) {
      const int split = t / tiles_mn, rem = t % tiles_mn;
      const int m0 = (rem / tp.n_tiles) * 256 + (int)rank * BM, n0 = (rem % tp.n_tiles) * BN;
      const int kb_begin = split * tp.kb_per_split;
      const bool has_k = kb_begin < num_kb_total;
      const uint32_t as = tile_iter & 1;
      mbar_wait(&acc_full[as], (tile_iter >> 1) & 1);
      tc_fence_after();
      drain_accumulator<BN, ACTDROP>(ep, tmem_base + as * 256, epi_stage + (warp - 4) * (32 * EPI_STAGE_LD), q, part, lane, m0, n0, split, has_k);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[as]), 0));
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// SIMT verification twin: same operands (hi+lo reconstructed to fp32), fp32 FMA, same epilogue code.
// 64x64 tile, 256 threads, 4x4 outputs per thread.
// ------------------------------------------------------------------------------------------------
struct SimtParams {
  int K, kb_per_split;
  const __nv_bfloat16* a; int64_t lda, a_ps; int a_mn;
  const __nv_bfloat16* b; int64_t ldb, b_ps; int b_mn;
};

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtParams sp, const EpiParams ep) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int split = blockIdx.z;
  const int k_begin = split * sp.kb_per_split * BK;
  const int k_end = min(sp.K, k_begin + sp.kb_per_split * BK);
  float acc[4][4] = {};
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      int mm, kk;
      if (!sp.a_mn) { kk = e % 16; mm = e / 16; } else { mm = e % 64; kk = e / 64; }
      const int m = m0 + mm, k = k0 + kk;
      float va = 0.f;
      if (m < ep.M && k < k_end) {
        const int64_t idx = sp.a_mn ? ((int64_t)k * sp.lda + m) : ((int64_t)m * sp.lda + k);
        va = plane_value(sp.a, sp.a_ps, idx);
      }
      As[kk][mm] = va;
      int nn, kb;
      if (!sp.b_mn) { kb = e % 16; nn = e / 16; } else { nn = e % 64; kb = e / 64; }
      const int n = n0 + nn, k2 = k0 + kb;
      float vb = 0.f;
      if (n < ep.N && k2 < k_end) {
        const int64_t idx = sp.b_mn ? ((int64_t)k2 * sp.ldb + n) : ((int64_t)n * sp.ldb + k2);
        vb = plane_value(sp.b, sp.b_ps, idx);
      }
      Bs[kb][nn] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int col0 = n0 + tx * 4;
    if (col0 < ep.N) epilogue_store<4>(ep, m0 + ty * 4 + i, col0, acc[i], split);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 3-D bf16 tensor map over split planes: dims (inner, outer, plane).  box = (64, box_outer, box_planes), SWIZZLE_128B.
int make_plane_tmap(CUtensorMap* out, const void* base, int64_t inner, int64_t outer, int64_t ld, int64_t plane_stride,
                    int box_outer, int box_planes) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled not available from the driver");
    return SRW_ERR_DRIVER;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) || (plane_stride % 8)) {
    set_last_error("TMA operand must be 16-byte aligned with ld and plane stride multiples of 8 elements (ld=%lld ps=%lld)",
                   (long long)ld, (long long)plane_stride);
    return SRW_ERR_ARG;
  }
  cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)outer, 2};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_outer, (cuuint32_t)box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: %d (inner=%lld outer=%lld ld=%lld ps=%lld box=%d)", (int)r, (long long)inner,
                   (long long)outer, (long long)ld, (long long)plane_stride, box_outer);
    return SRW_ERR_DRIVER;
  }
  return SRW_OK;
}

// Tile width: the candidate with the fewest (waves x tile cost) over the 148 SMs; ties go to the wider tile (less L2
// traffic per FLOP).  192 divides the ViT widths exactly, 128 / 64 give more tiles for small problems.
int gemm_pick_bn(int M, int N, int splits, int ctas) {
  if (ctas <= 0) ctas = 148;
  const int cands[3] = {192, 128, 64};
  int best = 128;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn == 192 && N % 192 != 0) continue;
    if (bn == 128 && N <= 64) continue;
    const int64_t tiles = (int64_t)cdiv(M, BM) * cdiv(N, bn) * std::max(1, splits);
    const double waves = (double)((tiles + ctas - 1) / ctas);
    const double cost = waves * (bn + 40);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// Pair-tile width for the 2-CTA kernel (0 = use the 1-CTA kernel): 256 x BN tiles, BN/2 rows of B per CTA, which must be
// a multiple of 64 for MN-major B.  Small problems (fewer pair tiles than half the SMs' worth) stay on the 1-CTA kernel.
int gemm2_pick_bn(int M, int N, int splits, bool b_mn, int ctas, int k_per_cta) {
  if (ctas <= 0) ctas = 148;
  const int npairs = std::max(1, ctas / 2);
  const int cands[3] = {256, 192, 128};
  int best = 0;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn == 192 && (b_mn || N % 192 != 0)) continue;
    if (bn == 256 && N % 256 != 0 && N < 1024) continue;
    if (N < bn) continue;
    const int64_t tiles = (int64_t)cdiv(M, 256) * cdiv(N, bn) * std::max(1, splits);
    const double waves = (double)((tiles + npairs - 1) / npairs);
    const double cost = waves * (bn + 40);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  // compare with the 1-CTA kernel's wave cost (per-SM work): prefer 2-CTA unless it quantises clearly worse
  const int bn1 = gemm_pick_bn(M, N, splits, ctas);
  const int64_t tiles1 = (int64_t)cdiv(M, BM) * cdiv(N, bn1) * std::max(1, splits);
  const double cost1 = (double)((tiles1 + ctas - 1) / ctas) * (bn1 + 40);
  // (scripts/gemm_sweep.sh: forcing the 3-stage pair kernel on the deep-K shapes wins 5 us on fc2 forward in isolation, 33.8 -> 28.5 us,
  // but the whole step did not move — 4.68 vs 4.74 ms — so the wave-quantisation rule stays as it was)
  (void)k_per_cta;
  const double slack = 1.15;
  if (best == 0 || best_cost > slack * cost1) return 0;
  return best;
}

static int fill_epi(const srw_gemm_args* a, EpiParams& ep) {
  ep.M = a->M; ep.N = a->N; ep.epilogue = a->epilogue;
  ep.bias = a->bias;
  ep.resid = a->resid; ep.ldr = a->ldr;
  ep.row_scale = a->row_scale; ep.rows_per_scale = a->rows_per_scale > 0 ? a->rows_per_scale : 1;
  ep.aux = a->aux; ep.ldaux = a->ldaux;
  ep.out_f32 = a->out_f32; ep.ldo = a->ldo;
  ep.out_planes = reinterpret_cast<__nv_bfloat16*>(a->out_planes); ep.ldp = a->ldp; ep.out_plane_stride = a->out_plane_stride;
  ep.workspace = a->workspace;
  ep.drop = make_drop(a->drop); ep.drop_rows_per_seq = a->drop_rows_per_seq > 0 ? a->drop_rows_per_seq : 1;
  SRW_REQUIRE(!ep.drop.on || a->epilogue == SRW_EPI_RESID || a->epilogue == SRW_EPI_GELU || a->epilogue == SRW_EPI_DGELU,
              "srw_gemm: dropout is part of SRW_EPI_RESID / SRW_EPI_GELU / SRW_EPI_DGELU only");
  SRW_REQUIRE(a->N % 4 == 0, "srw_gemm: N must be a multiple of 4 (N=%d)", a->N);
  switch (a->epilogue) {
    case SRW_EPI_F32: SRW_REQUIRE(a->out_f32 && a->ldo % 4 == 0, "srw_gemm: EPI_F32 needs out_f32, ldo%%4==0"); break;
    case SRW_EPI_PLANES: SRW_REQUIRE(a->out_planes && a->ldp % 4 == 0, "srw_gemm: EPI_PLANES needs out_planes"); break;
    case SRW_EPI_GELU:
      SRW_REQUIRE(a->out_planes && a->out_f32 && a->ldp % 4 == 0 && a->ldo % 4 == 0, "srw_gemm: EPI_GELU needs out_f32 and out_planes");
      break;
    case SRW_EPI_RESID: SRW_REQUIRE(a->out_f32 && a->resid && a->ldo % 4 == 0 && a->ldr % 4 == 0, "srw_gemm: EPI_RESID needs out_f32 and resid"); break;
    case SRW_EPI_DGELU: SRW_REQUIRE(a->out_planes && a->aux && a->ldp % 4 == 0 && a->ldaux % 4 == 0, "srw_gemm: EPI_DGELU needs out_planes and aux"); break;
    case SRW_EPI_SPLITK: SRW_REQUIRE(a->workspace && a->split_k >= 1, "srw_gemm: EPI_SPLITK needs workspace and split_k>=1"); break;
    default: set_last_error("srw_gemm: unknown epilogue %d", a->epilogue); return SRW_ERR_ARG;
  }
  return SRW_OK;
}

}  // namespace srw

using namespace srw;

extern "C" int srw_gemm(const srw_gemm_args* a, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  SRW_REQUIRE(a && a->M > 0 && a->N > 0 && a->K > 0, "srw_gemm: bad shape");
  SRW_REQUIRE(a->a && a->b, "srw_gemm: null operand");
  EpiParams ep;
  int rc = fill_epi(a, ep);
  if (rc) return rc;
  const int num_kb = cdiv(a->K, BK);
  int split = (a->epilogue == SRW_EPI_SPLITK) ? a->split_k : 1;
  if (split > num_kb) split = num_kb;
  const int kb_per_split = cdiv(num_kb, split);
  // note: with split < requested, the untouched workspace slices are still written (as zeros) because grid.z = split_k
  const int grid_z = (a->epilogue == SRW_EPI_SPLITK) ? a->split_k : 1;

  if (a->impl == SRW_GEMM_SIMT) {
    SRW_REQUIRE(a->a_seg_k == 0, "srw_gemm: K segments are a tcgen05-path feature");
    SimtParams sp;
    sp.K = a->K; sp.kb_per_split = kb_per_split;
    sp.a = reinterpret_cast<const __nv_bfloat16*>(a->a); sp.lda = a->lda; sp.a_ps = a->a_plane_stride; sp.a_mn = a->a_mn_major;
    sp.b = reinterpret_cast<const __nv_bfloat16*>(a->b); sp.ldb = a->ldb; sp.b_ps = a->b_plane_stride; sp.b_mn = a->b_mn_major;
    dim3 grid(cdiv(a->N, 64), cdiv(a->M, 64), grid_z);
    gemm_simt_kernel<<<grid, 256, 0, stream>>>(sp, ep);
    g_launches++;
    SRW_LAUNCH_CHECK();
    return SRW_OK;
  }
  SRW_REQUIRE(a->impl == SRW_GEMM_TCGEN05 || a->impl == SRW_GEMM_TCGEN05_1CTA, "srw_gemm: unknown impl %d", a->impl);
  SRW_REQUIRE(kb_per_split * BK <= MAX_K_PER_CTA, "srw_gemm: K per CTA is %d > %d: use SRW_EPI_SPLITK with more splits (single fp32 accumulator)",
              kb_per_split * BK, MAX_K_PER_CTA);
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  static int num_sms = 148;
  std::call_once(attr_once, [] {
    auto set = [&](const void* fn, int bytes) {
      if (attr_err == cudaSuccess) attr_err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    set((const void*)gemm_bf16x3_tcgen05_kernel<192, false>, gemm_smem_bytes(192));
    set((const void*)gemm_bf16x3_tcgen05_kernel<128, false>, gemm_smem_bytes(128));
    set((const void*)gemm_bf16x3_tcgen05_kernel<64, false>, gemm_smem_bytes(64));
    set((const void*)gemm2_bf16x3_tcgen05_kernel<256, false>, gemm2_smem_bytes(256));
    set((const void*)gemm2_bf16x3_tcgen05_kernel<192, false>, gemm2_smem_bytes(192));
    set((const void*)gemm2_bf16x3_tcgen05_kernel<128, false>, gemm2_smem_bytes(128));
    set((const void*)gemm_bf16x3_tcgen05_kernel<192, true>, gemm_smem_bytes(192));
    set((const void*)gemm_bf16x3_tcgen05_kernel<128, true>, gemm_smem_bytes(128));
    set((const void*)gemm_bf16x3_tcgen05_kernel<64, true>, gemm_smem_bytes(64));
    set((const void*)gemm2_bf16x3_tcgen05_kernel<256, true>, gemm2_smem_bytes(256));
    set((const void*)gemm2_bf16x3_tcgen05_kernel<192, true>, gemm2_smem_bytes(192));
    set((const void*)gemm2_bf16x3_tcgen05_kernel<128, true>, gemm2_smem_bytes(128));
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) num_sms = n;
  });
  SRW_CUDA(attr_err);

  TcParams tp;
  tp.trace = g_gemm_trace;
  tp.K = a->K; tp.kb_per_split = kb_per_split; tp.a_mn = a->a_mn_major ? 1 : 0; tp.b_mn = a->b_mn_major ? 1 : 0;
  tp.splits = grid_z;
  tp.a_seg_kb = 0; tp.a_seg_rows = 0;
  const bool segmented = a->a_seg_k > 0;
  const bool actdrop = ep.drop.on && (a->epilogue == SRW_EPI_GELU || a->epilogue == SRW_EPI_DGELU);   // HuBERT's activation dropout: own kernel instantiations
  int64_t a_inner = a->K, a_outer = a->M;
  if (segmented) {
    // K = nseg * a_seg_kb * 64 (every segment padded to whole k-blocks; B carries zeros there, TMA zero-fills A past a_seg_k)
    SRW_REQUIRE(!a->a_mn_major && a->a_seg_rows > 0, "srw_gemm: K segments need a K-major A and a_seg_rows > 0");
    tp.a_seg_kb = cdiv(a->a_seg_k, BK);
    SRW_REQUIRE(a->K % (tp.a_seg_kb * BK) == 0, "srw_gemm: K (%d) must be a whole number of padded segments (%d)", a->K, tp.a_seg_kb * BK);
    SRW_REQUIRE(a->a_seg_rows <= 0x7fffffff, "srw_gemm: a_seg_rows too large");
    tp.a_seg_rows = (int)a->a_seg_rows;
    a_inner = a->a_seg_k;
    a_outer = (int64_t)a->M + (int64_t)(a->K / (tp.a_seg_kb * BK) - 1) * a->a_seg_rows;
  }
  const double flops = 2.0 * a->M * a->N * a->K, bytes = 4.0 * ((double)a->M * a->K + (double)a->N * a->K + (double)a->M * a->N);

  // ---- 2-CTA path: pair tiles of 256 x BN ----
  // max_ctas > 0: this GEMM may only occupy that many SMs (it runs next to other kernels on other streams)
  const int cta_cap = (a->max_ctas > 0 && a->max_ctas < num_sms) ? std::max(2, a->max_ctas) : num_sms;
  int bn2 = srw::gemm2_pick_bn(a->M, a->N, grid_z, a->b_mn_major != 0, cta_cap, kb_per_split * BK);
  // tuning aid (scripts/gemm_bench.py --sweep): SRW_GEMM_FORCE = "1:<bn>" forces the 1-CTA kernel with that tile width, "2:<bn>" the 2-CTA kernel
  static const char* force_env = getenv("SRW_GEMM_FORCE");
  int force_bn1 = 0;
  if (force_env && force_env[0] && force_env[1] == ':') {
    const int fb = atoi(force_env + 2);
    if (force_env[0] == '1' && (fb == 64 || fb == 128 || (fb == 192 && a->N % 192 == 0))) { bn2 = 0; force_bn1 = fb; }
    if (force_env[0] == '2' && (fb == 128 || fb == 256 || (fb == 192 && a->N % 192 == 0 && !a->b_mn_major)) && a->N >= fb) bn2 = fb;
  }
  if (a->impl == SRW_GEMM_TCGEN05 && bn2 > 0 && !segmented) {
    CUtensorMap ta, tb;
    if (!a->a_mn_major) rc = make_plane_tmap(&ta, a->a, a->K, a->M, a->lda, a->a_plane_stride, 128, 2);
    else rc = make_plane_tmap(&ta, a->a, a->M, a->K, a->lda, a->a_plane_stride, 64, 2);
    if (rc) return rc;
    if (!a->b_mn_major) rc = make_plane_tmap(&tb, a->b, a->K, a->N, a->ldb, a->b_plane_stride, bn2 / 2, 2);
    else rc = make_plane_tmap(&tb, a->b, a->N, a->K, a->ldb, a->b_plane_stride, 64, 2);
    if (rc) return rc;
    tp.m_tiles = cdiv(a->M, 256); tp.n_tiles = cdiv(a->N, bn2);
    const int total_tiles = tp.m_tiles * tp.n_tiles * tp.splits;
    const int pairs = std::min(total_tiles, cta_cap / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(GEMM_THREADS); cfg.stream = stream;
    cfg.dynamicSmemBytes = gemm2_smem_bytes(bn2);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 2;
    void* prof = prof_begin(SRW_PROF_GEMM, flops, bytes, stream);
    cudaError_t le;
    if (actdrop) {
      if (bn2 == 256) le = cudaLaunchKernelEx(&cfg, gemm2_bf16x3_tcgen05_kernel<256, true>, ta, tb, tp, ep);
      else if (bn2 == 192) le = cudaLaunchKernelEx(&cfg, gemm2_bf16x3_tcgen05_kernel<192, true>, ta, tb, tp, ep);
      else le = cudaLaunchKernelEx(&cfg, gemm2_bf16x3_tcgen05_kernel<128, true>, ta, tb, tp, ep);
    } else {
      if (bn2 == 256) le = cudaLaunchKernelEx(&cfg, gemm2_bf16x3_tcgen05_kernel<256, false>, ta, tb, tp, ep);
      else if (bn2 == 192) le = cudaLaunchKernelEx(&cfg, gemm2_bf16x3_tcgen05_kernel<192, false>, ta, tb, tp, ep);
      else le = cudaLaunchKernelEx(&cfg, gemm2_bf16x3_tcgen05_kernel<128, false>, ta, tb, tp, ep);
    }
    prof_end(prof, stream);
    g_launches++;
    SRW_CUDA(le);
    SRW_LAUNCH_CHECK();
    return SRW_OK;
  }

  // ---- 1-CTA path ----
  const int bn = force_bn1 ? force_bn1 : srw::gemm_pick_bn(a->M, a->N, grid_z, cta_cap);
  CUtensorMap ta, tb;
  if (!a->a_mn_major) rc = make_plane_tmap(&ta, a->a, a_inner, a_outer, a->lda, a->a_plane_stride, 128, 2);
  else rc = make_plane_tmap(&ta, a->a, a->M, a->K, a->lda, a->a_plane_stride, 64, 2);
  if (rc) return rc;
  if (!a->b_mn_major) rc = make_plane_tmap(&tb, a->b, a->K, a->N, a->ldb, a->b_plane_stride, bn, 2);
  else rc = make_plane_tmap(&tb, a->b, a->N, a->K, a->ldb, a->b_plane_stride, 64, 2);
  if (rc) return rc;
  tp.m_tiles = cdiv(a->M, BM); tp.n_tiles = cdiv(a->N, bn);
  const int total_tiles = tp.m_tiles * tp.n_tiles * tp.splits;
  const int grid = std::min(total_tiles, cta_cap);
  void* prof = prof_begin(SRW_PROF_GEMM, flops, bytes, stream);
  cudaError_t le;
  if (actdrop) {
    if (bn == 192) le = launch_pdl(gemm_bf16x3_tcgen05_kernel<192, true>, dim3(grid), dim3(GEMM_THREADS), gemm_smem_bytes(192), stream, ta, tb, tp, ep);
    else if (bn == 128) le = launch_pdl(gemm_bf16x3_tcgen05_kernel<128, true>, dim3(grid), dim3(GEMM_THREADS), gemm_smem_bytes(128), stream, ta, tb, tp, ep);
    else le = launch_pdl(gemm_bf16x3_tcgen05_kernel<64, true>, dim3(grid), dim3(GEMM_THREADS), gemm_smem_bytes(64), stream, ta, tb, tp, ep);
  } else {
    if (bn == 192) le = launch_pdl(gemm_bf16x3_tcgen05_kernel<192, false>, dim3(grid), dim3(GEMM_THREADS), gemm_smem_bytes(192), stream, ta, tb, tp, ep);
    else if (bn == 128) le = launch_pdl(gemm_bf16x3_tcgen05_kernel<128, false>, dim3(grid), dim3(GEMM_THREADS), gemm_smem_bytes(128), stream, ta, tb, tp, ep);
    else le = launch_pdl(gemm_bf16x3_tcgen05_kernel<64, false>, dim3(grid), dim3(GEMM_THREADS), gemm_smem_bytes(64), stream, ta, tb, tp, ep);
  }
  prof_end(prof, stream);
  g_launches++;
  SRW_CUDA(le);
  SRW_LAUNCH_CHECK();
  return SRW_OK;
}

// debug: per-CTA clock64 timeline of the 1-CTA kernel into buf[grid][16] (NULL turns it off).  Not part of include/srw.h.
extern "C" int srw_gemm_set_trace(unsigned long long* buf) {
  srw::g_gemm_trace = buf;
  return SRW_OK;
}
