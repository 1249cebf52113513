// srw_engine.cuh — host-side helpers shared by the backbone engines (srw_vit.cu, srw_bert.cu): workspace carving, the split-K
// policy of the weight-gradient GEMMs, thin builders around srw_gemm / srw_colsum / srw_grad_fold, and the CUDA-graph replay cache.
#pragma once
#include <stdlib.h>

#include <atomic>
#include <functional>
#include <vector>

#include "../../include/srw.h"
#include "srw_common.cuh"

namespace srw {
extern std::atomic<int64_t> g_launches;

struct Carver {
  int64_t off = 0;
  int64_t take(int64_t bytes) {
    const int64_t o = off;
    off += (bytes + 1023) / 1024 * 1024;
    return o;
  }
};

static inline int64_t splitk_for(int M, int N, int64_t K, int* split_out) {
  // K slices so that (tiles x slices) fills the 148 SMs about twice, with at least 4 k-blocks (256 rows) per slice and
  // at most 2048 rows per slice (single fp32 accumulator per tile, see srw_gemm.cu)
  const int kb = (int)cdiv64(K, 64);
  int best = 1;
  double best_cost = 1e30;
  const int max_split = std::max(1, kb / 4), min_split = (int)cdiv64(K, 2048);
  for (int split = min_split; split <= std::max(min_split, std::min(max_split, 64)); ++split) {
    const int bn = gemm_pick_bn(M, N, split);
    const int64_t tiles = (int64_t)cdiv(M, 128) * cdiv(N, bn) * split;
    const double waves = (double)((tiles + 147) / 148);
    const double cost = waves * (cdiv(kb, split) * (bn + 40) + 0.5 * bn) + 0.02 * split;   // mainloop + epilogue + reduce traffic
    if (cost < best_cost - 1e-9) { best_cost = cost; best = split; }
  }
  if (split_out) *split_out = best;
  return (int64_t)best * M * N;
}

static inline int split_to(const float* x, int64_t ldx, int rows, int cols, void* planes, int64_t ldp, const float* row_scale, int rows_per_scale,
                    cudaStream_t s, float* colsum_out = nullptr, int colsum_accumulate = 0, float* colsum_ws = nullptr) {
  srw_split_args a = {};
  a.x = x; a.ldx = ldx; a.rows = rows; a.cols = cols; a.row_scale = row_scale; a.rows_per_scale = rows_per_scale;
  a.planes = planes; a.ldp = ldp; a.plane_stride = (int64_t)rows * ldp;
  a.colsum_out = colsum_out; a.colsum_accumulate = colsum_accumulate; a.colsum_workspace = colsum_ws;
  return srw_split_planes(&a, s);
}

struct Gemm {
  srw_gemm_args g = {};
  Gemm(int M, int N, int K, int impl) { g.M = M; g.N = N; g.K = K; g.impl = impl; g.split_k = 1; }
  // operand stored [rows, ld] as planes with `rows_total` rows per plane
  Gemm& A(const void* p, int64_t ld, int64_t rows_total, int mn) { g.a = p; g.lda = ld; g.a_plane_stride = rows_total * ld; g.a_mn_major = mn; return *this; }
  Gemm& Bm(const void* p, int64_t ld, int64_t rows_total, int mn) { g.b = p; g.ldb = ld; g.b_plane_stride = rows_total * ld; g.b_mn_major = mn; return *this; }
  int run(cudaStream_t s) { return srw_gemm(&g, s); }
};

#define SRW_TRY(expr)        \
  do {                       \
    int _rc = (expr);        \
    if (_rc) return _rc;     \
  } while (0)

static inline int wgrad(int M, int N, int64_t K, const void* a_planes, int64_t lda, int64_t a_rows, const void* b_planes, int64_t ldb, int64_t b_rows,
                 float* ws, float* out, int64_t ldo, int accumulate, int impl, cudaStream_t s, srw_grad_fold_args* fold = nullptr) {
  // out[M,N] (+)= A^T B with A stored [K, M] and B stored [K, N] (token-major activations / gradients)
  int split = 1;
  splitk_for(M, N, K, &split);
  Gemm g(M, N, (int)K, impl);
  g.A(a_planes, lda, a_rows, 1).Bm(b_planes, ldb, b_rows, 1);
  g.g.epilogue = SRW_EPI_SPLITK; g.g.split_k = split; g.g.workspace = ws;
  SRW_TRY(g.run(s));
  srw_splitk_reduce_args r = {};
  r.workspace = ws; r.split_k = split; r.M = M; r.N = N; r.out = out; r.ldo = ldo; r.accumulate = accumulate;
  if (fold) {   // folded with the rest of the block's reductions (srw_grad_fold)
    fold->splitk[fold->n_splitk++] = r;
    return SRW_OK;
  }
  return srw_splitk_reduce(&r, s);
}

static inline void fold_colsum(srw_grad_fold_args* fold, const float* partial, int nparts, int64_t stride_p, int cols, float* out, int accumulate) {
  srw_fold_colsum& c = fold->colsum[fold->n_colsum++];
  c.partial = partial; c.nparts = nparts; c.stride_p = stride_p; c.cols = cols; c.out = out; c.accumulate = accumulate;
}

static inline int colsum_planes(const void* planes, int64_t ld, int64_t rows_total, int rows, int cols, float* out, int accumulate, float* ws,
                         cudaStream_t s, srw_grad_fold_args* fold = nullptr) {
  srw_colsum_args a = {};
  a.planes = planes; a.ldp = ld; a.plane_stride = rows_total * ld; a.rows = rows; a.cols = cols; a.out = fold ? nullptr : out; a.accumulate = accumulate;
  a.workspace = ws;
  SRW_TRY(srw_colsum(&a, s));
  if (fold) fold_colsum(fold, ws, srw_colsum_nparts(rows), cols, cols, out, accumulate);
  return SRW_OK;
}

// SRW_FOLD=0: every reduction as its own launch right behind its producer (the previous behaviour; for A/B measurements)
static inline bool fold_enabled() {
  static const bool on = [] { const char* e = getenv("SRW_FOLD"); return !(e && e[0] == '0'); }();
  return on;
}

struct KeyBuilder {
  std::vector<uint8_t> k;
  template <typename T> void add(const T& v) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(&v);
    k.insert(k.end(), b, b + sizeof(T));
  }
};


// CUDA-graph replay of an engine call (defined in srw_vit.cu): the second call with an identical key is captured, later ones replay.
bool graphs_enabled(cudaStream_t s);
int run_graphed(std::vector<uint8_t>&& key, cudaStream_t s, const std::function<int(cudaStream_t)>& body);

}  // namespace srw
