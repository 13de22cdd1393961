// Per-graph max pooling (a12) and prototype max-gather (a15) -- HBM bound gathers/segment reductions.
#include <float.h>

#include "common.cuh"

namespace egp {

constexpr int kPoolThreads = 128;

// out[g,c] = max_{i in graph g} x[i,c]; empty graph -> 0 (scatter 'amax' into zeros, include_self=False);
// ties keep the first row (torch_scatter CUDA arg-max semantics used for the gradient).
// blockDim = (kPoolThreads, kPoolRowGroups): row group `ty` scans rows r0 + ty, r0 + ty + G, ... of the graph (eight
// 16-byte loads in flight per thread), then the groups are combined through shared memory -- one CTA per graph alone is
// latency bound (256 graphs x 128 threads cannot keep 6.5 TB/s of loads in flight).
constexpr int kPoolRowGroups = 4;
template <typename T>
__global__ void __launch_bounds__(kPoolThreads * kPoolRowGroups)
segment_max_fwd_kernel(const T* __restrict__ x, const int64_t* __restrict__ ptr, T* __restrict__ out,
                       int32_t* __restrict__ arg, int64_t channels) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  constexpr int G = kPoolRowGroups;
  __shared__ float s_best[G - 1][kPoolThreads][VN];
  __shared__ int32_t s_arg[G - 1][kPoolThreads][VN];
  const int64_t g = blockIdx.x;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int64_t col = ((int64_t)blockIdx.y * kPoolThreads + tx) * VN;
  const bool live = col < channels;
  const int64_t r0 = ptr[g], r1 = ptr[g + 1];
  float best[VN];
  int32_t bi[VN];
#pragma unroll
  for (int c = 0; c < VN; ++c) { best[c] = -FLT_MAX; bi[c] = -1; }
  auto take = [&](const Vec<T>& a, int64_t i) {   // rows arrive in ascending order inside a group: strict > keeps the first
#pragma unroll
    for (int c = 0; c < VN; ++c)
      if (a.v[c] > best[c] || bi[c] < 0) { best[c] = a.v[c]; bi[c] = (int32_t)i; }
  };
  if (live) {
    int64_t i = r0 + ty;
    for (; i + 7 * G < r1; i += 8 * G) {
      Raw<T> a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = Raw<T>::load(x + (i + u * G) * channels + col);
#pragma unroll
      for (int u = 0; u < 8; ++u) take(a[u].unpack(), i + u * G);
    }
    for (; i < r1; i += G) take(Vec<T>::load(x + i * channels + col), i);
  }
  if (ty > 0) {
#pragma unroll
    for (int c = 0; c < VN; ++c) { s_best[ty - 1][tx][c] = best[c]; s_arg[ty - 1][tx][c] = bi[c]; }
  }
  __syncthreads();
  if (ty == 0 && live) {
#pragma unroll
    for (int q = 0; q < G - 1; ++q)
#pragma unroll
      for (int c = 0; c < VN; ++c) {
        const float ob = s_best[q][tx][c];
        const int32_t oi = s_arg[q][tx][c];
        // ties keep the FIRST row (torch_scatter's arg-max): larger value wins, equal values -> lower row index
        if (oi >= 0 && (bi[c] < 0 || ob > best[c] || (ob == best[c] && oi < bi[c]))) { best[c] = ob; bi[c] = oi; }
      }
    Vec<T> o;
#pragma unroll
    for (int c = 0; c < VN; ++c) {
      o.v[c] = bi[c] >= 0 ? best[c] : 0.f;
      arg[g * channels + col + c] = bi[c];
    }
    o.store(out + g * channels + col);
  }
}

template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
segment_max_bwd_kernel(const T* __restrict__ dout, const int32_t* __restrict__ arg,
                       const int64_t* __restrict__ batch, T* __restrict__ dx, int64_t nvec, int64_t channels) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  constexpr int U = 4;  // vectors per thread per iteration: U dependent chains (batch -> dout / arg) in flight
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += U * stride) {
    int64_t row[U], c0[U], g[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      const int64_t e0 = (v < nvec ? v : v0) * VN;
      row[u] = e0 / channels;
      c0[u] = e0 % channels;
      g[u] = batch[row[u]];
    }
    Raw<T> d[U];
    int4 a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      d[u] = Raw<T>::load(dout + g[u] * channels + c0[u]);
      const int4* ap = reinterpret_cast<const int4*>(arg + g[u] * channels + c0[u]);
      a0[u] = ap[0];
      if (VN == 8) a1[u] = ap[1];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v >= nvec) continue;
      const Vec<T> dv = d[u].unpack();
      const int32_t am[8] = {a0[u].x, a0[u].y, a0[u].z, a0[u].w, a1[u].x, a1[u].y, a1[u].z, a1[u].w};
      Vec<T> o;
#pragma unroll
      for (int c = 0; c < VN; ++c) o.v[c] = (am[c] == (int32_t)row[u]) ? dv.v[c] : 0.f;
      o.store(dx + v * VN);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
proto_max_gather_kernel(const T* __restrict__ protos, const int64_t* __restrict__ idx, T* __restrict__ m,
                        int64_t nvec, int64_t k, int64_t channels) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  constexpr int U = 4;  // output vectors per thread per iteration: U * k gathers in flight behind U index loads
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += U * stride) {
    int64_t row[U], c0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      const int64_t e0 = (v < nvec ? v : v0) * VN;
      row[u] = e0 / channels;
      c0[u] = e0 % channels;
    }
    if (k == 4) {  // the configured neighbour count (experiments/egopack/*.yaml): everything unrolled
      int64_t j[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) j[u][q] = idx[row[u] * 4 + q];
      Raw<T> a[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) a[u][q] = Raw<T>::load(protos + j[u][q] * channels + c0[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v >= nvec) continue;
        Vec<T> best = a[u][0].unpack();
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const Vec<T> t = a[u][q].unpack();
#pragma unroll
          for (int c = 0; c < VN; ++c) best.v[c] = fmaxf(best.v[c], t.v[c]);
        }
        best.store(m + v * VN);
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t v = v0 + u * stride;
        if (v >= nvec) continue;
        Vec<T> best = Vec<T>::load(protos + idx[row[u] * k] * channels + c0[u]);
        for (int64_t q = 1; q < k; ++q) {
          const Vec<T> t = Vec<T>::load(protos + idx[row[u] * k + q] * channels + c0[u]);
#pragma unroll
          for (int c = 0; c < VN; ++c) best.v[c] = fmaxf(best.v[c], t.v[c]);
        }
        best.store(m + v * VN);
      }
    }
  }
}

// gradient of a = max(f, max_j P[idx[:,j]]) w.r.t. the bank (GraphONE(freeze=False)): where the prototype maximum
// wins, da goes to the FIRST of the k gathered prototypes that attains it (torch_scatter arg-max semantics).
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
proto_max_scatter_bwd_kernel(const T* __restrict__ da, const T* __restrict__ f, const T* __restrict__ protos,
                             const int64_t* __restrict__ idx, float* __restrict__ dbank, int64_t nvec, int64_t k,
                             int64_t channels) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = v * VN;
    const int64_t row = e0 / channels, c0 = e0 % channels;
    const Vec<T> g = Vec<T>::load(da + e0);
    const Vec<T> fv = Vec<T>::load(f + e0);
    float best[VN];
    int64_t who[VN];
#pragma unroll
    for (int c = 0; c < VN; ++c) { best[c] = -FLT_MAX; who[c] = -1; }
    for (int64_t j = 0; j < k; ++j) {
      const int64_t pj = idx[row * k + j];
      const Vec<T> a = Vec<T>::load(protos + pj * channels + c0);
#pragma unroll
      for (int c = 0; c < VN; ++c)
        if (a.v[c] > best[c]) { best[c] = a.v[c]; who[c] = pj; }
    }
#pragma unroll
    for (int c = 0; c < VN; ++c)
      if (!(fv.v[c] >= best[c]) && who[c] >= 0 && g.v[c] != 0.f) atomicAdd(dbank + who[c] * channels + c0 + c, g.v[c]);
  }
}

// class-conditional feature sums in fp64 (prototype-bank builder, graphone.py:53): out[label[i], :] += x[i, :]
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
class_sum_kernel(const T* __restrict__ x, const int64_t* __restrict__ label, double* __restrict__ out, int64_t nvec,
                 int64_t channels, int64_t num_classes) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e0 = v * VN;
    const int64_t row = e0 / channels, c0 = e0 % channels;
    const int64_t l = label[row];
    if (l < 0 || l >= num_classes) continue;
    const Vec<T> a = Vec<T>::load(x + e0);
#pragma unroll
    for (int c = 0; c < VN; ++c) atomicAdd(out + l * channels + c0 + c, (double)a.v[c]);
  }
}

static int pool_grid(int64_t nvec) {
  int64_t g = ceil_div(nvec, (int64_t)kPoolThreads * 2);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_segment_max_pool_fwd(const void* x, const int64_t* ptr, void* out, int32_t* arg, int64_t num_graphs,
                             int64_t channels, int dtype, void* stream) {
  EGP_REQUIRE(x && ptr && out && arg, "segment_max_pool_fwd: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(x) && aligned16(out), "segment_max_pool_fwd: channels %% %d", (int)vn);
  if (num_graphs == 0 || channels == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(dtype, T, {
    const unsigned gy = (unsigned)ceil_div(channels, (int64_t)kPoolThreads * Vec<T>::N);
    (void)launch_kernel(segment_max_fwd_kernel<T>, dim3((unsigned)num_graphs, gy), dim3(kPoolThreads, kPoolRowGroups), 0, (cudaStream_t)stream, 
        (const T*)x, ptr, (T*)out, arg, channels);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_segment_max_pool_bwd(const void* dout, const int32_t* arg, const int64_t* batch, void* dx,
                             int64_t num_nodes, int64_t channels, int dtype, void* stream) {
  EGP_REQUIRE(dout && arg && batch && dx, "segment_max_pool_bwd: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(dout) && aligned16(dx), "segment_max_pool_bwd: channels %% %d", (int)vn);
  if (num_nodes == 0 || channels == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = num_nodes * channels / Vec<T>::N;
    (void)launch_kernel(segment_max_bwd_kernel<T>, pool_grid(nvec), kPoolThreads, 0, (cudaStream_t)stream, 
        (const T*)dout, arg, batch, (T*)dx, nvec, channels);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_proto_max_gather(const void* protos, const int64_t* idx, void* m, int64_t num_nodes, int64_t k,
                         int64_t channels, int proto_dtype, int out_dtype, void* stream) {
  EGP_REQUIRE(protos && idx && m, "proto_max_gather: null pointer");
  EGP_REQUIRE(proto_dtype == out_dtype, "proto_max_gather: bank and output dtypes must match");
  EGP_REQUIRE(k >= 1, "proto_max_gather: k must be >= 1");
  const int64_t vn = out_dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(protos) && aligned16(m), "proto_max_gather: channels %% %d", (int)vn);
  if (num_nodes == 0 || channels == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(out_dtype, T, {
    const int64_t nvec = num_nodes * channels / Vec<T>::N;
    (void)launch_kernel(proto_max_gather_kernel<T>, pool_grid(nvec), kPoolThreads, 0, (cudaStream_t)stream, 
        (const T*)protos, idx, (T*)m, nvec, k, channels);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_proto_max_scatter_bwd(const void* da, const void* f, const void* protos, const int64_t* idx, float* dbank,
                              int64_t num_nodes, int64_t k, int64_t channels, int dtype, void* stream) {
  EGP_REQUIRE(da && f && protos && idx && dbank, "proto_max_scatter_bwd: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && k >= 1, "proto_max_scatter_bwd: channels %% %d", (int)vn);
  if (num_nodes == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = num_nodes * channels / Vec<T>::N;
    (void)launch_kernel(proto_max_scatter_bwd_kernel<T>, pool_grid(nvec), kPoolThreads, 0, (cudaStream_t)stream, 
        (const T*)da, (const T*)f, (const T*)protos, idx, dbank, nvec, k, channels);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_class_sum_f64(const void* x, const int64_t* label, double* out, int64_t rows, int64_t channels,
                      int64_t num_classes, int dtype, void* stream) {
  EGP_REQUIRE(x && label && out, "class_sum_f64: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(x), "class_sum_f64: channels %% %d", (int)vn);
  if (rows == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = rows * channels / Vec<T>::N;
    (void)launch_kernel(class_sum_kernel<T>, pool_grid(nvec), kPoolThreads, 0, (cudaStream_t)stream, (const T*)x, label, out, nvec, channels,
                                                                                    num_classes);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

}  // extern "C"
