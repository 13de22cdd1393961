// (f)-2: the optimiser step of the training loops (torch.optim.Adam(lr, weight_decay) at main_temporal.py:265-271 /
// main_egopack.py:317-324, stepped at main_temporal.py:130) as ONE pass over flat buffers -- HBM bound:
// 16 B read + 12 B written per parameter for p / grad / exp_avg / exp_avg_sq, plus 2 B for the bf16 copy of the updated
// parameter that the tensor-core GEMMs of the bf16 compute mode read (so no separate fp32 -> bf16 weight casts run).
//
// torch.optim.Adam semantics (amsgrad off, maximize off):
//   g' = g + weight_decay * p;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),   t = completed steps + 1
// Parameters and moments of all tensors sit back to back in flat fp32 buffers (tensor starts padded to 8 elements);
// gradients stay wherever autograd (or the all-reduce buckets) left them: the kernel gets one pointer per tensor.
#include <math.h>

#include "common.cuh"

namespace egp {

constexpr int kAdamThreads = 256;
constexpr int kAdamMaxTensors = 160;   // gradient pointers passed by value per launch (1.25 KB of kernel parameters)
// work items are chunks of <= 4096 elements that never cross a tensor boundary; the CALLER builds the chunk table
// (egopack_b200/optim.py)

struct AdamGrads {
  const float* g[kAdamMaxTensors];
};

// chunk table (device, built once): chunk c covers flat elements [start[c], start[c] + len[c]) of tensor tensor[c]
__global__ void __launch_bounds__(kAdamThreads)
adam_step_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, __nv_bfloat16* __restrict__ shadow,
                 const int64_t* __restrict__ seg_off, const int32_t* __restrict__ chunk_tensor,
                 const int64_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_len, int num_chunks,
                 int tensor_base, int tensor_count, AdamGrads grads, const int64_t* __restrict__ step,
                 const float* __restrict__ lr, float beta1, float beta2, float eps, float weight_decay) {
  pdl_enter();
  const float lr0 = lr[0];
  for (int c = blockIdx.x; c < num_chunks; c += gridDim.x) {
    const int ti = chunk_tensor[c] - tensor_base;
    if (ti < 0 || ti >= tensor_count) continue;
    const float* g = grads.g[ti];
    if (!g) continue;                                   // no gradient this step: the tensor is skipped (torch semantics)
    // torch keeps one step counter PER PARAMETER: a tensor without gradient does not advance its bias correction
    const float t = (float)(step[chunk_tensor[c]] + 1);
    const float bc1 = 1.f - powf(beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
    const float step_size = lr0 / bc1;
    const int64_t start = chunk_start[c];
    const int len = chunk_len[c];
    const float* gc = g + (start - seg_off[chunk_tensor[c]]);
    const bool vec = ((reinterpret_cast<uintptr_t>(gc) & 15u) == 0) && ((start & 3) == 0);
    auto update = [&](float& pv, float& mv, float& vv, float gv) {
      gv = fmaf(weight_decay, pv, gv);
      mv = fmaf(beta1, mv, (1.f - beta1) * gv);
      vv = fmaf(beta2, vv, (1.f - beta2) * gv * gv);
      const float denom = sqrtf(vv) / bc2_sqrt + eps;
      pv -= step_size * (mv / denom);
    };
    if (vec) {
      const int nv = len / 4;
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const int64_t e = start + 4 * (int64_t)i;
        float4 pv = *reinterpret_cast<float4*>(p + e), mv = *reinterpret_cast<float4*>(m + e);
        float4 vv = *reinterpret_cast<float4*>(v + e);
        const float4 gv = *reinterpret_cast<const float4*>(gc + 4 * i);
        update(pv.x, mv.x, vv.x, gv.x);
        update(pv.y, mv.y, vv.y, gv.y);
        update(pv.z, mv.z, vv.z, gv.z);
        update(pv.w, mv.w, vv.w, gv.w);
        *reinterpret_cast<float4*>(p + e) = pv;
        *reinterpret_cast<float4*>(m + e) = mv;
        *reinterpret_cast<float4*>(v + e) = vv;
        if (shadow) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&lo);
          pk.y = *reinterpret_cast<const uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(shadow + e) = pk;
        }
      }
      for (int i = 4 * nv + threadIdx.x; i < len; i += blockDim.x) {   // tail of the tensor's last chunk
        const int64_t e = start + i;
        float pv = p[e], mv = m[e], vv = v[e];
        update(pv, mv, vv, gc[i]);
        p[e] = pv; m[e] = mv; v[e] = vv;
        if (shadow) shadow[e] = __float2bfloat16_rn(pv);
      }
    } else {
      for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const int64_t e = start + i;
        float pv = p[e], mv = m[e], vv = v[e];
        update(pv, mv, vv, gc[i]);
        p[e] = pv; m[e] = mv; v[e] = vv;
        if (shadow) shadow[e] = __float2bfloat16_rn(pv);
      }
    }
  }
}

__global__ void adam_tick_kernel(int64_t* __restrict__ step, int tensor_base, int tensor_count, AdamGrads grads) {
  pdl_enter();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < tensor_count && grads.g[i]) step[tensor_base + i] += 1;
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_adam_step(float* p, float* m, float* v, void* shadow, const int64_t* seg_off, const int32_t* chunk_tensor,
                  const int64_t* chunk_start, const int32_t* chunk_len, int64_t num_chunks, const float* const* grads,
                  int num_tensors, int64_t* step, const float* lr, float beta1, float beta2, float eps,
                  float weight_decay, void* stream) {
  EGP_REQUIRE(p && m && v && seg_off && chunk_tensor && chunk_start && chunk_len && grads && step && lr,
              "adam_step: null pointer");
  static_assert(kAdamMaxTensors <= 1024, "the tick kernel runs one thread per tensor of a launch");
  EGP_REQUIRE(num_tensors >= 0 && num_chunks >= 0 && num_chunks < (int64_t)INT32_MAX, "adam_step: bad sizes");
  EGP_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adam_step: bad hyper-parameters");
  EGP_REQUIRE(aligned16(p) && aligned16(m) && aligned16(v) && (!shadow || aligned16(shadow)), "adam_step: flat buffers must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (num_chunks > 0) {
    int64_t grid = num_chunks;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    for (int base = 0; base < num_tensors; base += kAdamMaxTensors) {
      AdamGrads tab;
      const int count = num_tensors - base < kAdamMaxTensors ? num_tensors - base : kAdamMaxTensors;
      bool any = false;
      for (int i = 0; i < kAdamMaxTensors; ++i) {
        tab.g[i] = i < count ? grads[base + i] : nullptr;
        any = any || tab.g[i] != nullptr;
      }
      if (!any) continue;
      (void)launch_kernel(adam_step_kernel, (unsigned)grid, kAdamThreads, 0, s, p, m, v, (__nv_bfloat16*)shadow, seg_off, chunk_tensor,
                          chunk_start, chunk_len, (int)num_chunks, base, count, tab, (const int64_t*)step, lr, beta1, beta2, eps,
                          weight_decay);
      EGP_LAUNCH_CHECK();
      (void)launch_kernel(adam_tick_kernel, 1, kAdamMaxTensors, 0, s, step, base, count, tab);
      EGP_LAUNCH_CHECK();
    }
  }
  return EGP_OK;
}

}  // extern "C"
