// Loss kernels (a11) -- HBM / latency bound, fp32 logits.
//
//  cross entropy, reduction='none', ignore_index, label smoothing:
//      nn.CrossEntropyLoss(ignore_index=-1, reduction='none') per head, summed over heads
//      (models/tasks/recognition.py:21,61-69; lta.py:21,73-74; criterion/wrapper.py:80-82; main_temporal.py:285,291)
//      F.cross_entropy(..., label_smoothing=0.1) for OSCC (models/tasks/oscc.py:88-96)
//  BCE with logits, reduction='none' (models/tasks/pnr.py:38,82-83)
//  weighted mean: `w * loss.mean()` summed over tasks (main_temporal.py:99-128)
//
// One warp per row; a row is read once in the forward (online max / sum-exp in registers for up to 32*kRowCache
// classes, a second pass through L1 beyond that) and once in the backward.
#include <math.h>

#include "common.cuh"

namespace egp {

constexpr int kLossThreads = 256;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// loss[i] (+)= (1-eps) * (lse - z_t) + eps * (lse - mean_c z_c);  0 for ignored rows (t == ignore_index or out of range)
__global__ void __launch_bounds__(kLossThreads)
ce_loss_fwd_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                   int64_t label_stride, int64_t n, int classes, int64_t ignore_index, float smoothing,
                   float* __restrict__ loss, int accumulate, float* __restrict__ lse_out) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (kLossThreads / 32) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* z = logits + row * ld;
  float m = -INFINITY, sum = 0.f;
  for (int c = lane; c < classes; c += 32) {
    const float v = z[c];
    m = fmaxf(m, v);
    sum += v;
  }
  m = warp_max(m);
  sum = warp_sum(sum);
  float se = 0.f;
  for (int c = lane; c < classes; c += 32) se += expf(z[c] - m);   // second touch: L1
  se = warp_sum(se);
  const float lse = m + logf(se);
  if (lane == 0) {
    const int64_t t = labels[row * label_stride];
    float l = 0.f;
    if (t != ignore_index && t >= 0 && t < classes)
      l = (1.f - smoothing) * (lse - z[t]) + smoothing * (lse - sum / (float)classes);
    loss[row] = accumulate ? loss[row] + l : l;
    lse_out[row] = lse;
  }
}

// dlogits[i,c] = g_i * (softmax_c - (1-eps) [c == t] - eps / C); zero rows for ignored targets.  OutT = float
// ([n, ldd], columns >= classes untouched) or bf16 (columns classes..ldd-1 zero-filled: a TMA-ready GEMM operand).
template <typename OutT>
__global__ void __launch_bounds__(kLossThreads)
ce_loss_bwd_kernel(const float* __restrict__ logits, int64_t ld, const float* __restrict__ lse,
                   const int64_t* __restrict__ labels, int64_t label_stride, const float* __restrict__ dloss,
                   int64_t dloss_stride, int64_t n, int classes, int64_t ignore_index, float smoothing,
                   OutT* __restrict__ dlogits, int64_t ldd, int pad_to) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (kLossThreads / 32) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* z = logits + row * ld;
  OutT* d = dlogits + row * ldd;
  const int64_t t = labels[row * label_stride];
  const bool live = t != ignore_index && t >= 0 && t < classes;
  const float g = live ? dloss[row * dloss_stride] : 0.f;
  const float l = lse[row];
  const float uni = smoothing / (float)classes;
  for (int c = lane; c < pad_to; c += 32) {
    float v = 0.f;
    if (c < classes && live) v = g * (expf(z[c] - l) - (c == (int)t ? 1.f - smoothing : 0.f) - uni);
    d[c] = from_float<OutT>(v);
  }
}

// BCE with logits: loss = max(z,0) - z*t + log1p(exp(-|z|))
__global__ void bce_logits_fwd_kernel(const float* __restrict__ z, const float* __restrict__ target,
                                      float* __restrict__ loss, int64_t n) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = z[i], t = target[i];
  loss[i] = fmaxf(v, 0.f) - v * t + log1pf(expf(-fabsf(v)));
}

__global__ void bce_logits_bwd_kernel(const float* __restrict__ z, const float* __restrict__ target,
                                      const float* __restrict__ dloss, int64_t dloss_stride, float* __restrict__ dz,
                                      int64_t n) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = z[i];
  const float s = 1.f / (1.f + expf(-v));
  dz[i] = dloss[i * dloss_stride] * (s - target[i]);
}

// out[0] (+)= weight * mean(x): one block, fixed reduction order (deterministic); n is a per-step loss vector
__global__ void __launch_bounds__(1024)
weighted_mean_kernel(const float* __restrict__ x, int64_t n, float weight, float* __restrict__ out, int accumulate) {
  pdl_enter();
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float v = (float)(s / (double)(n > 0 ? n : 1)) * weight;
    out[0] = accumulate ? out[0] + v : v;
  }
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_ce_loss_fwd(const float* logits, int64_t ld, const int64_t* labels, int64_t label_stride, int64_t n,
                    int64_t classes, int64_t ignore_index, float label_smoothing, float* loss, int accumulate,
                    float* lse, void* stream) {
  EGP_REQUIRE(logits && labels && loss && lse, "ce_loss_fwd: null pointer");
  EGP_REQUIRE(classes >= 1 && classes < (int64_t)INT32_MAX && ld >= classes && label_stride >= 1,
              "ce_loss_fwd: bad class count / strides");
  EGP_REQUIRE(label_smoothing >= 0.f && label_smoothing <= 1.f, "ce_loss_fwd: label_smoothing must be in [0,1]");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(ce_loss_fwd_kernel, (unsigned)ceil_div(n, kLossThreads / 32), kLossThreads, 0, (cudaStream_t)stream, logits, ld,
                      labels, label_stride, n, (int)classes, ignore_index, label_smoothing, loss, accumulate, lse);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_ce_loss_bwd(const float* logits, int64_t ld, const float* lse, const int64_t* labels, int64_t label_stride,
                    const float* dloss, int64_t dloss_stride, int64_t n, int64_t classes, int64_t ignore_index,
                    float label_smoothing, void* dlogits, int64_t ldd, int out_dtype, void* stream) {
  EGP_REQUIRE(logits && lse && labels && dloss && dlogits, "ce_loss_bwd: null pointer");
  EGP_REQUIRE(classes >= 1 && classes < (int64_t)INT32_MAX && ld >= classes && ldd >= classes && dloss_stride >= 0,
              "ce_loss_bwd: bad class count / strides");
  if (n == 0) return EGP_OK;
  const unsigned grid = (unsigned)ceil_div(n, kLossThreads / 32);
  cudaStream_t s = (cudaStream_t)stream;
  if (out_dtype == EGP_F32)
    (void)launch_kernel(ce_loss_bwd_kernel<float>, grid, kLossThreads, 0, s, logits, ld, lse, labels, label_stride, dloss, dloss_stride,
                        n, (int)classes, ignore_index, label_smoothing, (float*)dlogits, ldd, (int)classes);
  else if (out_dtype == EGP_BF16)
    (void)launch_kernel(ce_loss_bwd_kernel<__nv_bfloat16>, grid, kLossThreads, 0, s, logits, ld, lse, labels, label_stride, dloss,
                        dloss_stride, n, (int)classes, ignore_index, label_smoothing, (__nv_bfloat16*)dlogits, ldd, (int)ldd);
  else {
    set_error("ce_loss_bwd: unsupported dtype code %d", out_dtype);
    return EGP_ERR_INVALID;
  }
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_bce_logits_fwd(const float* z, const float* target, float* loss, int64_t n, void* stream) {
  EGP_REQUIRE(z && target && loss, "bce_logits_fwd: null pointer");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(bce_logits_fwd_kernel, (unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream, z, target, loss, n);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_bce_logits_bwd(const float* z, const float* target, const float* dloss, int64_t dloss_stride, float* dz,
                       int64_t n, void* stream) {
  EGP_REQUIRE(z && target && dloss && dz && dloss_stride >= 0, "bce_logits_bwd: null pointer / bad stride");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(bce_logits_bwd_kernel, (unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream, z, target, dloss, dloss_stride,
                      dz, n);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_weighted_mean(const float* x, int64_t n, float weight, float* out, int accumulate, void* stream) {
  EGP_REQUIRE(out && (x || n == 0), "weighted_mean: null pointer");
  (void)launch_kernel(weighted_mean_kernel, 1, 1024, 0, (cudaStream_t)stream, x, n, weight, out, accumulate);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

}  // extern "C"
