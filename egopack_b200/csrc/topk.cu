// a14: cosine k-NN of nodes against a frozen prototype bank (GraphONE.__compute_edges, graphONE.py:119-141):
//      d = 1 - (F/|F|)(P/|P|)^T ;  idx[i,:] = argsort(d[i,:])[:k]
// The reference sorts the full [B,Kp] matrix twice per stage per task; here the similarity is one GEMM and the
// selection is a register-resident running top-k per row (one warp per row, shuffle merge).
//
//   exact path     : fp32 similarity (FFMA GEMM) -> k smallest d, ties -> lower prototype index.
//   tensor path    : bf16 tcgen05 similarity -> top-KK candidates (KK >= k+8) -> exact fp32 re-score of the
//                    candidates on the fp32 normalised rows -> k smallest d.  Indices therefore do not depend
//                    on bf16 rounding as long as the true top-k is inside the candidate set.
#include <float.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"

namespace egp {

int sgemm_launch(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, const void* A2,
                 int64_t lda2, const void* B2, int64_t ldb2, int64_t K2, const float* bias, const void* residual,
                 int64_t ldr, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope,
                 int in_dtype, int out_dtype, int accumulate, cudaStream_t stream);
int tc_gemm_launch(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, const void* A2,
                   int64_t lda2, const void* B2, int64_t ldb2, int64_t K2, const float* bias, const void* residual,
                   int64_t ldr, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope,
                   int out_dtype, int accumulate, cudaStream_t stream, double* rowstats, void* workspace, size_t ws_bytes);
int tc_gemm_topk_launch(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K, int keep,
                        int32_t* cand, float* cand_thr, cudaStream_t stream);
int tc_topk_groups();
bool tc_gemm_supported(const void* A, int64_t lda, const void* B, int64_t ldb, const void* A2, int64_t lda2,
                       const void* B2, int64_t ldb2);

constexpr int kTopkThreads = 256;
constexpr int64_t kTopkChunkRows = 32768;

__device__ __forceinline__ bool key_less(float ka, int ia, float kb, int ib) {
  return ka < kb || (ka == kb && ia < ib);
}

// out[row, 0..nout) = indices of the nout smallest keys of the row, ascending (ties -> lower index).
// DIST: key = 1 - s (cosine dissimilarity, as the reference computes it); otherwise key = -s.
template <int KK, bool DIST, typename IdxT>
__global__ void __launch_bounds__(kTopkThreads)
row_topk_kernel(const float* __restrict__ S, int64_t rows, int64_t cols, int nout, IdxT* __restrict__ out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float key[KK];
  int id[KK];
#pragma unroll
  for (int t = 0; t < KK; ++t) { key[t] = FLT_MAX; id[t] = INT_MAX; }
  const float* s = S + row * cols;
#pragma unroll 4
  for (int64_t j = lane; j < cols; j += 32) {
    const float v = DIST ? 1.0f - s[j] : -s[j];
    if (key_less(v, (int)j, key[KK - 1], id[KK - 1])) {
      key[KK - 1] = v;
      id[KK - 1] = (int)j;
#pragma unroll
      for (int t = KK - 1; t > 0; --t) {
        if (key_less(key[t], id[t], key[t - 1], id[t - 1])) {
          const float fk = key[t]; key[t] = key[t - 1]; key[t - 1] = fk;
          const int fi = id[t]; id[t] = id[t - 1]; id[t - 1] = fi;
        }
      }
    }
  }
  for (int o = 0; o < nout; ++o) {
    float bk = key[0];
    int bi = id[0];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ok = __shfl_xor_sync(0xffffffffu, bk, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (key_less(ok, oi, bk, bi)) { bk = ok; bi = oi; }
    }
    if (bi == id[0] && bk == key[0]) {  // the winning lane pops its head (indices are unique per row)
#pragma unroll
      for (int t = 0; t < KK - 1; ++t) { key[t] = key[t + 1]; id[t] = id[t + 1]; }
      key[KK - 1] = FLT_MAX;
      id[KK - 1] = INT_MAX;
    }
    if (lane == 0) out[row * nout + o] = (IdxT)bi;
  }
}

// exact fp32 re-score of KK candidates per row, then the k smallest d = 1 - <f, p>.
//
// Miss detector (thr != null).  The candidates were chosen by their bf16 tensor-core score s16; every column j that is
// NOT a candidate scored s16_j <= thr_g (the smallest score its column group kept).  Its exact score obeys
//   s_j = s16_j - (df.p16_j + f.dp_j)  =>  s_j <= thr_g + |df| + |dp_j| + |df||dp_j|      (Cauchy-Schwarz, |f| = |p| = 1)
// with df / dp the bf16 rounding errors of the normalised rows: |df| is measured per row by egp_row_normalize
// (f_err), max_j |dp_j| once per bank (p_err).  If the k-th best exact candidate score clears that bound for every
// group, no column outside the candidate set can belong to the top-k; otherwise the row is FLAGGED (appended to
// flagged_rows) and the host re-runs it through the exact fp32 path.
template <int KK>
__global__ void __launch_bounds__(kTopkThreads)
rerank_kernel(const float* __restrict__ fn, const float* __restrict__ pn, const int32_t* __restrict__ cand,
              int64_t rows, int64_t protos, int64_t channels, int k, int64_t* __restrict__ idx,
              const float* __restrict__ thr, int groups, const float* __restrict__ f_err, float p_err,
              int64_t row_base, int32_t* __restrict__ flagged_rows, int32_t* __restrict__ flagged_count) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* f = fn + row * channels;
  const int nv = (int)(channels / 4);
  float d[KK];
  int id[KK];
#pragma unroll
  for (int c = 0; c < KK; ++c) {
    const int j = cand[row * KK + c];
    float acc = 0.f;
    if (j >= 0 && j < protos) {
      const float* p = pn + (int64_t)j * channels;
      for (int v = lane; v < nv; v += 32) {
        const float4 a = *reinterpret_cast<const float4*>(f + 4 * v);
        const float4 b = *reinterpret_cast<const float4*>(p + 4 * v);
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      }
      acc = warp_sum(acc);
      d[c] = 1.0f - acc;
      id[c] = j;
    } else {
      d[c] = FLT_MAX;
      id[c] = INT_MAX;
    }
  }
  if (lane == 0) {
    float kth = FLT_MAX;
    for (int o = 0; o < k; ++o) {  // selection of the k best of KK (KK <= 32)
      int best = 0;
#pragma unroll
      for (int c = 1; c < KK; ++c)
        if (key_less(d[c], id[c], d[best], id[best])) best = c;
      // static-index extraction keeps d/id in registers
      float bd = FLT_MAX;
      int bi = INT_MAX;
#pragma unroll
      for (int c = 0; c < KK; ++c)
        if (c == best) { bd = d[c]; bi = id[c]; d[c] = FLT_MAX; id[c] = INT_MAX; }
      kth = bd;
      idx[row * k + o] = (int64_t)bi;
    }
    if (thr) {
      float t = -FLT_MAX;
      for (int g = 0; g < groups; ++g) t = fmaxf(t, thr[row * groups + g]);
      const float fe = f_err ? f_err[row_base + row] : 0.00390625f;   // 2^-8: the unit roundoff of bf16
      // fp32 slack: the tensor-core accumulation of s16 and the fmaf chain of the exact score (C terms of <= 1 ulp(1)
      // each, even if every rounding went the same way), and the rounding of d = 1 - s
      const float bound = fe + p_err + fe * p_err + 1e-4f;
      const float s_k = 1.0f - kth;   // k-th best exact similarity among the candidates
      if (!(s_k - t > bound)) {       // also true for NaN scores
        const int slot = atomicAdd(flagged_count, 1);
        flagged_rows[slot] = (int32_t)(row_base + row);
      }
    }
  }
}

template <typename T, typename O>
__global__ void __launch_bounds__(kTopkThreads)
row_normalize_kernel(const T* __restrict__ x, O* __restrict__ out, int64_t rows, int64_t cols,
                     float* __restrict__ round_err) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * cols;
  float q = 0.f;
  for (int64_t c = lane; c < cols; c += 32) { const float v = to_float<T>(xr[c]); q += v * v; }
  const float nrm = sqrtf(warp_sum(q));
  float e = 0.f;
  for (int64_t c = lane; c < cols; c += 32) {
    const float v = to_float<T>(xr[c]) / nrm;
    const O o = from_float<O>(v);
    out[row * cols + c] = o;
    const float dlt = to_float<O>(o) - v;
    e += dlt * dlt;
  }
  if (round_err) {   // |stored row - exact normalised row|_2, rounded up: the k-NN miss detector's per-row error
    e = warp_sum(e);
    if (lane == 0) round_err[row] = sqrtf(e) * 1.0001f + 1e-7f;
  }
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_row_normalize(const void* x, void* out, int64_t rows, int64_t cols, int in_dtype, int out_dtype, float* round_err,
                      void* stream) {
  EGP_REQUIRE(x && out, "row_normalize: null pointer");
  if (rows == 0) return EGP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (unsigned)ceil_div(rows, kTopkThreads / 32);
  if (in_dtype == EGP_F32 && out_dtype == EGP_F32)
    (void)launch_kernel(row_normalize_kernel<float, float>, grid, kTopkThreads, 0, s, (const float*)x, (float*)out, rows, cols, round_err);
  else if (in_dtype == EGP_F32 && out_dtype == EGP_BF16)
    (void)launch_kernel(row_normalize_kernel<float, __nv_bfloat16>, grid, kTopkThreads, 0, s, (const float*)x, (__nv_bfloat16*)out, rows, cols, round_err);
  else if (in_dtype == EGP_BF16 && out_dtype == EGP_F32)
    (void)launch_kernel(row_normalize_kernel<__nv_bfloat16, float>, grid, kTopkThreads, 0, s, (const __nv_bfloat16*)x, (float*)out, rows, cols, round_err);
  else if (in_dtype == EGP_BF16 && out_dtype == EGP_BF16)
    (void)launch_kernel(row_normalize_kernel<__nv_bfloat16, __nv_bfloat16>, grid, kTopkThreads, 0, s, (const __nv_bfloat16*)x, (__nv_bfloat16*)out, rows, cols, round_err);
  else {
    set_error("row_normalize: bad dtype pair %d -> %d", in_dtype, out_dtype);
    return EGP_ERR_INVALID;
  }
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

static int topk_kk(int k) { return k + 8 <= 16 ? 16 : (k + 8 <= 32 ? 32 : 0); }

size_t egp_cos_topk_workspace(int64_t num_nodes, int64_t num_protos, int64_t k) {
  const int64_t rows = num_nodes < kTopkChunkRows ? num_nodes : kTopkChunkRows;
  return (size_t)rows * (size_t)num_protos * sizeof(float) + (size_t)rows * 32 * sizeof(int32_t) +
         (size_t)rows * 4 * sizeof(float) + 1024;   // similarities (unfused paths), candidates, group thresholds
}

int egp_cos_topk(const float* fn, const float* pn, const void* fn16, const void* pn16, int64_t num_nodes,
                 int64_t num_protos, int64_t channels, int k, int64_t* idx, const float* f_round_err,
                 float p_round_err, int32_t* flagged_rows, int32_t* flagged_count, void* workspace, size_t ws_bytes,
                 void* stream) {
  EGP_REQUIRE(fn && pn && idx && workspace, "cos_topk: null pointer");
  EGP_REQUIRE((flagged_rows == nullptr) == (flagged_count == nullptr), "cos_topk: flagged_rows and flagged_count come together");
  const bool guard = flagged_rows != nullptr;
  EGP_REQUIRE(k >= 1 && k <= 32 && k <= num_protos, "cos_topk: k=%d must be in [1, min(32, num_protos)]", k);
  EGP_REQUIRE(channels % 4 == 0 && aligned16(fn) && aligned16(pn), "cos_topk: channels must be a multiple of 4");
  if (ws_bytes < egp_cos_topk_workspace(num_nodes, num_protos, k)) {
    set_error("cos_topk: workspace %zu < %zu", ws_bytes, egp_cos_topk_workspace(num_nodes, num_protos, k));
    return EGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (guard) EGP_CUDA(cudaMemsetAsync(flagged_count, 0, sizeof(int32_t), s));
  if (num_nodes == 0) return EGP_OK;
  const int kk = topk_kk(k);
  // the miss detector lives in the fused path (kk == 16, i.e. k <= 8): a guarded call with a larger k is exact fp32
  const bool tensor = fn16 && pn16 && kk > 0 && kk <= num_protos && (!guard || kk == 16) &&
                      tc_gemm_supported(fn16, channels, pn16, channels, nullptr, 0, nullptr, 0);
  const int64_t chunk = num_nodes < kTopkChunkRows ? num_nodes : kTopkChunkRows;
  float* S = (float*)workspace;
  int32_t* cand = (int32_t*)((char*)workspace + (((size_t)chunk * num_protos * sizeof(float) + 255) & ~(size_t)255));
  float* thr = (float*)((char*)cand + (((size_t)chunk * 32 * sizeof(int32_t) + 255) & ~(size_t)255));
  const int groups = tc_topk_groups();
  for (int64_t r0 = 0; r0 < num_nodes; r0 += chunk) {
    const int64_t rows = (num_nodes - r0) < chunk ? (num_nodes - r0) : chunk;
    const unsigned grid = (unsigned)ceil_div(rows, kTopkThreads / 32);
    int rc;
    static const bool unfused = [] { const char* e = getenv("EGP_TOPK_UNFUSED"); return e && e[0] == '1'; }();
    if (tensor && kk == 16 && !unfused) {
      // fused path: similarity GEMM with the best candidates selected in its epilogue (k+4 <= 8 -> keep 8, else 16),
      // then the exact fp32 re-rank of those candidates
      const __nv_bfloat16* a = (const __nv_bfloat16*)fn16 + r0 * channels;
      // every column group keeps its own set: keep 8 per group when k+4 <= 8 (k+4 covered the true top-k on every
      // row in SURVEY app. B), else 16; the re-rank sees groups*keep candidates (unused slots hold INT_MAX)
      const int keep = (k + 4 <= 8) ? 8 : 16;
      rc = tc_gemm_topk_launch(a, channels, pn16, channels, rows, num_protos, channels, keep, cand, guard ? thr : nullptr, s);
      if (rc != EGP_OK) return rc;
      if (keep * groups == 16)
        (void)launch_kernel(rerank_kernel<16>, grid, kTopkThreads, 0, s, fn + r0 * channels, pn, cand, rows, num_protos, channels, k, idx + r0 * k,
                            guard ? (const float*)thr : nullptr, groups, f_round_err, p_round_err, r0, flagged_rows, flagged_count);
      else
        (void)launch_kernel(rerank_kernel<32>, grid, kTopkThreads, 0, s, fn + r0 * channels, pn, cand, rows, num_protos, channels, k, idx + r0 * k,
                            guard ? (const float*)thr : nullptr, groups, f_round_err, p_round_err, r0, flagged_rows, flagged_count);
    } else if (tensor) {
      const __nv_bfloat16* a = (const __nv_bfloat16*)fn16 + r0 * channels;
      rc = tc_gemm_launch(a, channels, 0, pn16, channels, 0, nullptr, 0, nullptr, 0, 0, nullptr, nullptr, 0, S,
                          num_protos, rows, num_protos, channels, EGP_ACT_NONE, 0.f, EGP_F32, 0, s, nullptr, nullptr, 0);
      if (rc != EGP_OK) return rc;
      if (kk == 16) {
        (void)launch_kernel(row_topk_kernel<16, false, int32_t>, grid, kTopkThreads, 0, s, S, rows, num_protos, 16, cand);
        (void)launch_kernel(rerank_kernel<16>, grid, kTopkThreads, 0, s, fn + r0 * channels, pn, cand, rows, num_protos, channels, k, idx + r0 * k,
                            (const float*)nullptr, 0, (const float*)nullptr, 0.f, r0, (int32_t*)nullptr, (int32_t*)nullptr);
      } else {
        (void)launch_kernel(row_topk_kernel<32, false, int32_t>, grid, kTopkThreads, 0, s, S, rows, num_protos, 32, cand);
        (void)launch_kernel(rerank_kernel<32>, grid, kTopkThreads, 0, s, fn + r0 * channels, pn, cand, rows, num_protos, channels, k, idx + r0 * k,
                            (const float*)nullptr, 0, (const float*)nullptr, 0.f, r0, (int32_t*)nullptr, (int32_t*)nullptr);
      }
    } else {
      rc = sgemm_launch(fn + r0 * channels, channels, 0, pn, channels, 0, nullptr, 0, nullptr, 0, 0, nullptr, nullptr,
                        0, S, num_protos, rows, num_protos, channels, EGP_ACT_NONE, 0.f, EGP_F32, EGP_F32, 0, s);
      if (rc != EGP_OK) return rc;
      if (k <= 4) (void)launch_kernel(row_topk_kernel<4, true, int64_t>, grid, kTopkThreads, 0, s, S, rows, num_protos, k, idx + r0 * k);
      else if (k <= 8) (void)launch_kernel(row_topk_kernel<8, true, int64_t>, grid, kTopkThreads, 0, s, S, rows, num_protos, k, idx + r0 * k);
      else if (k <= 16) (void)launch_kernel(row_topk_kernel<16, true, int64_t>, grid, kTopkThreads, 0, s, S, rows, num_protos, k, idx + r0 * k);
      else (void)launch_kernel(row_topk_kernel<32, true, int64_t>, grid, kTopkThreads, 0, s, S, rows, num_protos, k, idx + r0 * k);
    }
    EGP_LAUNCH_CHECK();
  }
  return EGP_OK;
}

}  // extern "C"
