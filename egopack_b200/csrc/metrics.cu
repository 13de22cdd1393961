// Headline-metric kernels (SURVEY.md section 8 row a-M / (f)-4): the integer parts of utils/meters/ego4d.py that the
// reference runs on the CPU through torchmetrics / editdistance.  All results are integers (ranks, indices, edit
// distances), so they are bit-exact against the oracle.
#include <float.h>

#include "common.cuh"

namespace egp {

constexpr int kMetricThreads = 256;

// rank[i] = number of classes that beat the label's logit in row i: strictly larger, or equal with a lower class
// index (so rank == 0 <=> label == first arg-max, the torch.argmax rule).  One warp per row.
// rank = -1 for ignored rows (label == ignore_index) and for labels outside [0, C).
__global__ void __launch_bounds__(kMetricThreads)
label_rank_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels, int64_t label_stride,
                  int64_t n, int64_t classes, int64_t ignore_index, int32_t* __restrict__ rank) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (kMetricThreads / 32) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t t = labels[row * label_stride];
  if (t == ignore_index || t < 0 || t >= classes) {
    if (lane == 0) rank[row] = -1;
    return;
  }
  const float* r = logits + row * ld;
  const float ref = r[t];
  int cnt = 0;
  for (int64_t c = lane; c < classes; c += 32) {
    const float v = r[c];
    cnt += (v > ref || (v == ref && c < t)) ? 1 : 0;
  }
  cnt = warp_sum(cnt);
  if (lane == 0) rank[row] = cnt;
}

// out[g] = first index (relative to the graph start) of the maximum of f(values) over [ptr[g], ptr[g+1]);
// f = sigmoid in fp32 when apply_sigmoid (the meter takes argmax AFTER torch.sigmoid, whose saturation can create
// ties that the raw logits do not have); -1 for an empty graph.  One warp per graph.
__global__ void __launch_bounds__(kMetricThreads)
segment_argmax_kernel(const float* __restrict__ values, const int64_t* __restrict__ ptr, int64_t num_graphs,
                      int apply_sigmoid, int64_t* __restrict__ out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t g = (int64_t)blockIdx.x * (kMetricThreads / 32) + (threadIdx.x >> 5);
  if (g >= num_graphs) return;
  const int64_t r0 = ptr[g], r1 = ptr[g + 1];
  float best = -FLT_MAX;
  int64_t bi = INT64_MAX;
  for (int64_t i = r0 + lane; i < r1; i += 32) {
    float v = values[i];
    if (apply_sigmoid) v = 1.0f / (1.0f + expf(-v));   // torch.sigmoid's fp32 formula
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) out[g] = r1 > r0 ? bi - r0 : -1;
}

// out[n] = min over the K sampled sequences of the Levenshtein distance between preds[n, :, k] and labels[n, :]
// (editdistance.eval semantics: unit-cost insert / delete / substitute).  One thread per (n, k) pair runs the two-row
// dynamic programme in local memory; the K results of a row are combined through shared memory.
constexpr int kMaxSeq = 64;
__global__ void __launch_bounds__(kMetricThreads)
edit_distance_min_kernel(const int64_t* __restrict__ preds, const int64_t* __restrict__ labels, int64_t n, int z, int k,
                         int32_t* __restrict__ out) {
  pdl_enter();
  __shared__ int32_t res[kMetricThreads];
  const int per = kMetricThreads / k;              // rows per block
  const int local_row = threadIdx.x / k, s = threadIdx.x % k;
  const int64_t row = (int64_t)blockIdx.x * per + local_row;
  int32_t d = INT32_MAX;
  if (local_row < per && row < n) {
    int32_t prev[kMaxSeq + 1];
    int64_t b[kMaxSeq];
    for (int j = 0; j < z; ++j) b[j] = labels[row * z + j];
    for (int j = 0; j <= z; ++j) prev[j] = j;
    for (int i = 1; i <= z; ++i) {
      const int64_t ca = preds[(row * z + (i - 1)) * k + s];
      int32_t diag = prev[0];
      prev[0] = i;
      for (int j = 1; j <= z; ++j) {
        const int32_t up = prev[j];
        const int32_t sub = diag + (ca != b[j - 1] ? 1 : 0);
        const int32_t best = min(min(up + 1, prev[j - 1] + 1), sub);
        diag = up;
        prev[j] = best;
      }
    }
    d = prev[z];
  }
  res[threadIdx.x] = d;
  __syncthreads();
  if (local_row < per && row < n && s == 0) {
    int32_t m = res[threadIdx.x];
    for (int q = 1; q < k; ++q) m = min(m, res[threadIdx.x + q]);
    out[row] = m;
  }
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_label_rank(const float* logits, int64_t ld, const int64_t* labels, int64_t label_stride, int64_t n,
                   int64_t classes, int64_t ignore_index, int32_t* rank, void* stream) {
  EGP_REQUIRE(logits && labels && rank, "label_rank: null pointer");
  EGP_REQUIRE(classes >= 1 && ld >= classes && label_stride >= 1, "label_rank: bad sizes");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(label_rank_kernel, (unsigned)ceil_div(n, kMetricThreads / 32), kMetricThreads, 0, (cudaStream_t)stream, 
      logits, ld, labels, label_stride, n, classes, ignore_index, rank);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_segment_argmax(const float* values, const int64_t* ptr, int64_t num_graphs, int apply_sigmoid, int64_t* out,
                       void* stream) {
  EGP_REQUIRE(values && ptr && out, "segment_argmax: null pointer");
  if (num_graphs == 0) return EGP_OK;
  (void)launch_kernel(segment_argmax_kernel, (unsigned)ceil_div(num_graphs, kMetricThreads / 32), kMetricThreads, 0, (cudaStream_t)stream, 
      values, ptr, num_graphs, apply_sigmoid, out);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_edit_distance_min(const int64_t* preds, const int64_t* labels, int64_t n, int64_t seq_len, int64_t num_samples,
                          int32_t* out, void* stream) {
  EGP_REQUIRE(preds && labels && out, "edit_distance_min: null pointer");
  EGP_REQUIRE(seq_len >= 1 && seq_len <= kMaxSeq, "edit_distance_min: sequence length must be in [1, %d]", kMaxSeq);
  EGP_REQUIRE(num_samples >= 1 && num_samples <= kMetricThreads, "edit_distance_min: 1 <= K <= %d", kMetricThreads);
  if (n == 0) return EGP_OK;
  const int per = kMetricThreads / (int)num_samples;
  (void)launch_kernel(edit_distance_min_kernel, (unsigned)ceil_div(n, per), kMetricThreads, 0, (cudaStream_t)stream, 
      preds, labels, n, (int)seq_len, (int)num_samples, out);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

}  // extern "C"
