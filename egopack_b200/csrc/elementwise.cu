// Elementwise / small-reduction kernels -- all HBM bound, 128-bit vectorised, grid-stride.
#include "common.cuh"

namespace egp {

constexpr int kEwThreads = 256;

static int ew_grid(int64_t nvec) {
  int64_t g = ceil_div(nvec, (int64_t)kEwThreads * 2);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// a4: out = x + [sin(pos*f) | cos(pos*f)]  (gnn.PositionalEncoding, models/graph.py:37,63).  A thread owns one
// 16-byte vector of the sin half and the matching vector of the cos half, so each argument is reduced once.
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
posenc_add_kernel(const T* __restrict__ x, const int64_t* __restrict__ pos, const float* __restrict__ freq,
                  T* __restrict__ out, int64_t nwork, int64_t channels) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const int64_t half = channels / 2;
  const int64_t hv = half / VN;  // vectors per half row
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nwork; v += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = v / hv, c0 = (v % hv) * VN;
    const float p = (float)pos[row];  // int64 * float32 promotes to float32 in the reference
    const int64_t e0 = row * channels + c0;
    Vec<T> a = Vec<T>::load(x + e0);
    Vec<T> b = Vec<T>::load(x + e0 + half);
#pragma unroll
    for (int c = 0; c < VN; ++c) {
      float sn, cs;
      sincosf(p * freq[c0 + c], &sn, &cs);
      a.v[c] += sn;
      b.v[c] += cs;
    }
    a.store(out + e0);
    b.store(out + e0 + half);
  }
}

template <typename S, typename D>
__global__ void __launch_bounds__(kEwThreads)
cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  pdl_enter();
  // 8 elements per thread-iteration: 2x16B loads for fp32 sources, 1x16B for bf16
  const int64_t n8 = n / 8;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n8; v += (int64_t)gridDim.x * blockDim.x) {
    float f[8];
    if constexpr (sizeof(S) == 4) {
      const Vec<float> a = Vec<float>::load((const float*)src + v * 8);
      const Vec<float> b = Vec<float>::load((const float*)src + v * 8 + 4);
#pragma unroll
      for (int c = 0; c < 4; ++c) { f[c] = a.v[c]; f[4 + c] = b.v[c]; }
    } else {
      const Vec<__nv_bfloat16> a = Vec<__nv_bfloat16>::load((const __nv_bfloat16*)src + v * 8);
#pragma unroll
      for (int c = 0; c < 8; ++c) f[c] = a.v[c];
    }
    if constexpr (sizeof(D) == 4) {
      Vec<float> a, b;
#pragma unroll
      for (int c = 0; c < 4; ++c) { a.v[c] = f[c]; b.v[c] = f[4 + c]; }
      a.store((float*)dst + v * 8);
      b.store((float*)dst + v * 8 + 4);
    } else {
      Vec<__nv_bfloat16> a;
#pragma unroll
      for (int c = 0; c < 8; ++c) a.v[c] = f[c];
      a.store((__nv_bfloat16*)dst + v * 8);
    }
  }
  // tail
  const int64_t t = n8 * 8 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) dst[t] = from_float<D>(to_float<S>(src[t]));
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, int64_t nvec) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    Vec<T> x = Vec<T>::load(a + v * VN);
    const Vec<T> y = Vec<T>::load(b + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) x.v[c] += y.v[c];
    x.store(out + v * VN);
  }
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
axpby_kernel(const T* __restrict__ a, float alpha, const T* __restrict__ b, float beta, T* __restrict__ out,
             int64_t nvec) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    Vec<T> x = Vec<T>::load(a + v * VN);
    if (b) {
      const Vec<T> y = Vec<T>::load(b + v * VN);
#pragma unroll
      for (int c = 0; c < VN; ++c) x.v[c] = alpha * x.v[c] + beta * y.v[c];
    } else {
#pragma unroll
      for (int c = 0; c < VN; ++c) x.v[c] = alpha * x.v[c];
    }
    x.store(out + v * VN);
  }
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t nvec, int act,
               float slope) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    Vec<T> g = Vec<T>::load(dy + v * VN);
    const Vec<T> o = Vec<T>::load(y + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) {
      if (!(o.v[c] > 0.f)) g.v[c] = (act == EGP_ACT_LEAKY_RELU) ? g.v[c] * slope : 0.f;
    }
    g.store(dx + v * VN);
  }
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
mask_scale_kernel(const T* __restrict__ x, const uint8_t* __restrict__ mask, T* __restrict__ out, int64_t nvec,
                  float scale) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    Vec<T> a = Vec<T>::load(x + v * VN);
    const uint8_t* m = mask + v * VN;
#pragma unroll
    for (int c = 0; c < VN; ++c) a.v[c] = m[c] ? a.v[c] * scale : 0.f;
    a.store(out + v * VN);
  }
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
max_combine_fwd_kernel(const T* __restrict__ f, const T* __restrict__ m, T* __restrict__ a, int64_t nvec) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    Vec<T> x = Vec<T>::load(f + v * VN);
    const Vec<T> y = Vec<T>::load(m + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) x.v[c] = fmaxf(x.v[c], y.v[c]);
    x.store(a + v * VN);
  }
}

template <typename T>
__global__ void __launch_bounds__(kEwThreads)
max_combine_bwd_kernel(const T* __restrict__ da, const T* __restrict__ f, const T* __restrict__ m,
                       T* __restrict__ df, int64_t nvec) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    Vec<T> g = Vec<T>::load(da + v * VN);
    const Vec<T> x = Vec<T>::load(f + v * VN);
    const Vec<T> y = Vec<T>::load(m + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) g.v[c] = (x.v[c] >= y.v[c]) ? g.v[c] : 0.f;
    g.store(df + v * VN);
  }
}

// column sums: grid (parts, column chunks); thread owns a 16-byte column, loops over its row strip
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
colsum_partial_kernel(const T* __restrict__ x, int64_t rows, int64_t cols, int64_t ldx, int rows_per_cta,
                      int vec_ok, float* __restrict__ part) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const int64_t col = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * VN;
  if (col >= cols) return;
  float acc[VN];
#pragma unroll
  for (int c = 0; c < VN; ++c) acc[c] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(r0 + rows_per_cta, rows);
  if (vec_ok && col + VN <= cols) {
    int64_t i = r0;
    for (; i + 8 <= r1; i += 8) {  // eight independent 16-byte loads in flight per thread
      Raw<T> a[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] = Raw<T>::load(x + (i + u) * ldx + col);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const Vec<T> v = a[u].unpack();
#pragma unroll
        for (int c = 0; c < VN; ++c) acc[c] += v.v[c];
      }
    }
    for (; i < r1; ++i) {
      const Vec<T> a = Vec<T>::load(x + i * ldx + col);
#pragma unroll
      for (int c = 0; c < VN; ++c) acc[c] += a.v[c];
    }
  } else {  // ragged last chunk / unaligned rows (classifier heads): scalar path
    for (int64_t i = r0; i < r1; ++i)
#pragma unroll
      for (int c = 0; c < VN; ++c)
        if (col + c < cols) acc[c] += to_float<T>(x[i * ldx + col + c]);
  }
#pragma unroll
  for (int c = 0; c < VN; ++c)
    if (col + c < cols) part[(size_t)blockIdx.x * cols + col + c] = acc[c];
}

// out[c] = sum_g part[g][c]; a block owns kFinalCols columns and splits the part rows over 1024 / kFinalCols groups
// (latency-bound: few serial iterations per thread, ~128 CTAs for 1024 columns), combined in a fixed order
// (deterministic): shuffles inside a warp, shared memory across warps
constexpr int kFinalThreads = 1024;
constexpr int kFinalCols = 8;
constexpr int kFinalGroups = kFinalThreads / kFinalCols;
__global__ void __launch_bounds__(kFinalThreads)
colsum_final_kernel(const float* __restrict__ part, int parts, int64_t cols, float* __restrict__ out) {
  pdl_enter();
  __shared__ float red[32][kFinalCols];
  const int tx = threadIdx.x % kFinalCols, ty = threadIdx.x / kFinalCols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * kFinalCols + tx;
  float a = 0.f;
  if (c < cols) {
    int g = ty;
    for (; g + 3 * kFinalGroups < parts; g += 4 * kFinalGroups) {
      const float v0 = part[(size_t)g * cols + c], v1 = part[(size_t)(g + kFinalGroups) * cols + c];
      const float v2 = part[(size_t)(g + 2 * kFinalGroups) * cols + c], v3 = part[(size_t)(g + 3 * kFinalGroups) * cols + c];
      a += (v0 + v1) + (v2 + v3);
    }
    for (; g < parts; g += kFinalGroups) a += part[(size_t)g * cols + c];
  }
#pragma unroll
  for (int o = kFinalCols; o < 32; o <<= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane < kFinalCols) red[warp][lane] = a;
  __syncthreads();
  if (warp == 0) {
    constexpr int kPer = 32 / (32 / kFinalCols);  // warp partials summed per lane
    const int col = lane % kFinalCols, part0 = (lane / kFinalCols) * kPer;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) t += red[part0 + i][col];
#pragma unroll
    for (int o = kFinalCols; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    const int64_t cc = (int64_t)blockIdx.x * kFinalCols + col;
    if (lane < kFinalCols && cc < cols) out[cc] = t;
  }
}

// dst[m, ldd] (bf16/fp32) = cast(src[m, n]) with the columns n..ldd-1 zero-filled: gives classifier-width gradients
// (115 / 478 columns) the 16-byte row pitch TMA needs, in one pass
template <typename S, typename D>
__global__ void __launch_bounds__(kEwThreads)
cast_pad_kernel(const S* __restrict__ src, int64_t lds, D* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
  pdl_enter();
  const int64_t total = rows * ldd;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / ldd, c = e % ldd;
    dst[e] = c < cols ? from_float<D>(to_float<S>(src[r * lds + c])) : from_float<D>(0.f);
  }
}

// 1/||x_i||: one warp per row
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
row_inv_norm_kernel(const T* __restrict__ x, float* __restrict__ out, int64_t rows, int64_t cols, int64_t ldx) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ldx;
  float q = 0.f;
  for (int64_t v = lane; v < cols / VN; v += 32) {
    const Vec<T> a = Vec<T>::load(xr + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) q += a.v[c] * a.v[c];
  }
  q = warp_sum(q);
  if (lane == 0) out[row] = 1.0f / sqrtf(q);  // no epsilon, as the reference (graphONE.py:148-151)
}

// dx = dy * act'(y) fused with the column sums of dx (bias gradient of the producing Linear): threads keep a fixed
// 16-byte column (grid stride multiple of the row length), CTA partial rows -> colsum_final_kernel.
template <typename T>
__global__ void __launch_bounds__(kEwThreads)
act_bwd_colsum_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx, int64_t nvec, int period,
                      int act, float slope, float* __restrict__ part) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  __shared__ float csum[kEwThreads * VN];
  float dsum[VN];
#pragma unroll
  for (int c = 0; c < VN; ++c) dsum[c] = 0.f;
#pragma unroll 2
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    Vec<T> g = Vec<T>::load(dy + v * VN);
    const Vec<T> o = Vec<T>::load(y + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) {
      if (!(o.v[c] > 0.f)) g.v[c] = (act == EGP_ACT_LEAKY_RELU) ? g.v[c] * slope : 0.f;
      dsum[c] += to_float<T>(from_float<T>(g.v[c]));
    }
    g.store(dx + v * VN);
  }
  block_column_partial<VN>(dsum, period, part + (size_t)blockIdx.x * period * VN, csum);
}

static int colsum_parts(int64_t rows) {
  int64_t p = ceil_div(rows, 32);
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(p < 1 ? 1 : (p > cap ? cap : p));
}

size_t colsum_workspace_bytes(int64_t rows, int64_t cols) {
  return sizeof(float) * (size_t)colsum_parts(rows) * (size_t)cols + 64;
}

int colsum_launch(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx, int dtype, void* workspace,
                  size_t ws_bytes, cudaStream_t s) {
  EGP_REQUIRE(x && out && workspace, "colsum: null pointer");
  if (ws_bytes < colsum_workspace_bytes(rows, cols)) {
    set_error("colsum: workspace too small");
    return EGP_ERR_WORKSPACE;
  }
  if (cols == 0) return EGP_OK;
  if (rows == 0) {
    EGP_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * cols, s));
    return EGP_OK;
  }
  const int parts = colsum_parts(rows);
  const int rows_per = (int)ceil_div(rows, parts);
  float* part = (float*)workspace;
  EGP_DISPATCH_DTYPE(dtype, T, {
    constexpr int VN = Vec<T>::N;
    const int vec_ok = aligned16(x) && (ldx % VN) == 0;
    const int64_t nvec = ceil_div(cols, (int64_t)VN);
    const int threads = (int)(nvec >= kEwThreads ? kEwThreads : (nvec + 31) / 32 * 32);   // no idle half-blocks
    const int gy = (int)ceil_div(nvec, (int64_t)threads);
    (void)launch_kernel(colsum_partial_kernel<T>, dim3(parts, gy), threads, 0, s, (const T*)x, rows, cols, ldx, rows_per, vec_ok, part);
    EGP_LAUNCH_CHECK();
  });
  (void)launch_kernel(colsum_final_kernel, (unsigned)ceil_div(cols, kFinalCols), kFinalThreads, 0, s, part, parts, cols, out);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}


// ---------------------------------------------------------------------------------------------------------
// fp32 operand -> bf16 split terms for the fp32-on-tensor-cores GEMM (bf16x3 / bf16x6)
// x = x1 + x2 + x3 (+ <= 2^-24 |x|) with x1 = bf16(x), x2 = bf16(x - x1), x3 = bf16(x - x1 - x2): three bf16 numbers
// carry the 24-bit significand of an fp32 value.  The GEMM  sum_k a_k b_k  is then evaluated as the sum of bf16 x bf16
// products (a_i, b_j) on the tcgen05 pipe with fp32 accumulation -- the products are concatenated along K, so ONE launch
// of the ordinary bf16 kernel (same epilogue: bias, activation, residual) computes it.
// Output: T blocks; block t holds term sel[t] of every element, element (r, c) of block t at
//   dst[t * block_stride + r * ldd + c];  rows R..Rp-1 and columns C..Cp-1 of every block are zero-filled.
// ---------------------------------------------------------------------------------------------------------
struct SplitSel { int t[6]; };

__global__ void __launch_bounds__(kEwThreads)
split_bf16_kernel(const float* __restrict__ src, int64_t lds, int64_t R, int64_t C, int64_t Rp, int64_t Cp,
                  __nv_bfloat16* __restrict__ dst, int64_t block_stride, int64_t ldd, int T, SplitSel sel) {
  pdl_enter();
  const int64_t cvec = Cp / 4;                      // Cp is a multiple of 8
  const int64_t total = Rp * cvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cvec, c0 = (i - r * cvec) * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (r < R && c0 + j < C) ? src[r * lds + c0 + j] : 0.f;
    __nv_bfloat16 term[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat16 x1 = __float2bfloat16_rn(v[j]);
      const float r1 = v[j] - __bfloat162float(x1);
      const __nv_bfloat16 x2 = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(x2);
      term[0][j] = x1; term[1][j] = x2; term[2][j] = __float2bfloat16_rn(r2);
    }
    for (int t = 0; t < T; ++t) {
      const int w = sel.t[t];
      __nv_bfloat16* d = dst + (int64_t)t * block_stride + r * ldd + c0;
      // 8-byte store: c0 % 4 == 0, ldd % 8 == 0, block_stride % 8 == 0, dst 16-byte aligned
      uint2 pk;
      const __nv_bfloat16* tw = w == 0 ? term[0] : (w == 1 ? term[1] : term[2]);
      pk.x = (uint32_t)__bfloat16_as_ushort(tw[0]) | ((uint32_t)__bfloat16_as_ushort(tw[1]) << 16);
      pk.y = (uint32_t)__bfloat16_as_ushort(tw[2]) | ((uint32_t)__bfloat16_as_ushort(tw[3]) << 16);
      *reinterpret_cast<uint2*>(d) = pk;
    }
  }
}

}  // namespace egp

using namespace egp;

#define EGP_EW_CHECK(name, n, dtype, ...)                                                          \
  const int64_t _vn = (dtype) == EGP_BF16 ? 8 : 4;                                                 \
  EGP_REQUIRE((n) % _vn == 0, name ": element count %lld not a multiple of %d", (long long)(n), (int)_vn); \
  if ((n) == 0) return EGP_OK;

extern "C" {

int egp_posenc_add(const void* x, const int64_t* pos, const float* frequency, void* out, int64_t n,
                   int64_t channels, int dtype, void* stream) {
  EGP_REQUIRE(x && pos && frequency && out, "posenc_add: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % (2 * vn) == 0 && aligned16(x) && aligned16(out),
              "posenc_add: channels must be a multiple of %d", (int)(2 * vn));
  if (n == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nwork = n * channels / (2 * Vec<T>::N);
    (void)launch_kernel(posenc_add_kernel<T>, ew_grid(nwork), kEwThreads, 0, (cudaStream_t)stream, (const T*)x, pos, frequency,
                                                                                  (T*)out, nwork, channels);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream) {
  EGP_REQUIRE(src && dst, "cast: null pointer");
  EGP_REQUIRE(aligned16(src) && aligned16(dst), "cast: pointers must be 16-byte aligned");
  if (n == 0) return EGP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = ew_grid(ceil_div(n, 8));
  if (src_dtype == EGP_F32 && dst_dtype == EGP_BF16)
    (void)launch_kernel(cast_kernel<float, __nv_bfloat16>, grid, kEwThreads, 0, s, (const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == EGP_BF16 && dst_dtype == EGP_F32)
    (void)launch_kernel(cast_kernel<__nv_bfloat16, float>, grid, kEwThreads, 0, s, (const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == dst_dtype && (src_dtype == EGP_F32 || src_dtype == EGP_BF16)) {
    EGP_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * (src_dtype == EGP_F32 ? 4 : 2), cudaMemcpyDeviceToDevice, s));
    return EGP_OK;
  } else {
    set_error("cast: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
    return EGP_ERR_INVALID;
  }
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_cast_pad(const void* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols, int src_dtype,
                 int dst_dtype, void* stream) {
  EGP_REQUIRE(src && dst, "cast_pad: null pointer");
  EGP_REQUIRE(ldd >= cols && lds >= cols, "cast_pad: row pitches must cover the columns");
  if (rows == 0 || ldd == 0) return EGP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = ew_grid(ceil_div(rows * ldd, 4));
  if (src_dtype == EGP_F32 && dst_dtype == EGP_BF16)
    (void)launch_kernel(cast_pad_kernel<float, __nv_bfloat16>, grid, kEwThreads, 0, s, (const float*)src, lds, (__nv_bfloat16*)dst, ldd, rows, cols);
  else if (src_dtype == EGP_BF16 && dst_dtype == EGP_BF16)
    (void)launch_kernel(cast_pad_kernel<__nv_bfloat16, __nv_bfloat16>, grid, kEwThreads, 0, s, (const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, rows, cols);
  else if (src_dtype == EGP_F32 && dst_dtype == EGP_F32)
    (void)launch_kernel(cast_pad_kernel<float, float>, grid, kEwThreads, 0, s, (const float*)src, lds, (float*)dst, ldd, rows, cols);
  else if (src_dtype == EGP_BF16 && dst_dtype == EGP_F32)
    (void)launch_kernel(cast_pad_kernel<__nv_bfloat16, float>, grid, kEwThreads, 0, s, (const __nv_bfloat16*)src, lds, (float*)dst, ldd, rows, cols);
  else {
    set_error("cast_pad: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
    return EGP_ERR_INVALID;
  }
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_split_bf16(const float* src, int64_t lds, int64_t rows, int64_t cols, int64_t rows_padded, int64_t cols_padded,
                   void* dst, int64_t block_stride, int64_t ldd, int num_blocks, const int* term_of_block, void* stream) {
  EGP_REQUIRE(src && dst && term_of_block, "split_bf16: null pointer");
  EGP_REQUIRE(num_blocks >= 1 && num_blocks <= 6, "split_bf16: 1..6 blocks");
  EGP_REQUIRE(rows_padded >= rows && cols_padded >= cols && cols_padded % 8 == 0 && ldd % 8 == 0 && block_stride % 8 == 0 &&
              aligned16(dst) && lds >= cols, "split_bf16: padded sizes / pitches must keep 16-byte bf16 rows");
  SplitSel sel;
  for (int t = 0; t < 6; ++t) {
    sel.t[t] = t < num_blocks ? term_of_block[t] : 0;
    EGP_REQUIRE(sel.t[t] >= 0 && sel.t[t] <= 2, "split_bf16: terms are 0, 1, 2");
  }
  if (rows_padded == 0 || cols_padded == 0) return EGP_OK;
  (void)launch_kernel(split_bf16_kernel, ew_grid(rows_padded * (cols_padded / 4)), kEwThreads, 0, (cudaStream_t)stream, src, lds, rows,
                      cols, rows_padded, cols_padded, (__nv_bfloat16*)dst, block_stride, ldd, num_blocks, sel);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream) {
  EGP_REQUIRE(a && b && out, "add: null pointer");
  EGP_EW_CHECK("add", n, dtype);
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = n / Vec<T>::N;
    (void)launch_kernel(add_kernel<T>, ew_grid(nvec), kEwThreads, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, (T*)out, nvec);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_axpby(const void* a, float alpha, const void* b, float beta, void* out, int64_t n, int dtype, void* stream) {
  EGP_REQUIRE(a && out, "axpby: null pointer");
  EGP_EW_CHECK("axpby", n, dtype);
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = n / Vec<T>::N;
    (void)launch_kernel(axpby_kernel<T>, ew_grid(nvec), kEwThreads, 0, (cudaStream_t)stream, (const T*)a, alpha, (const T*)b, beta, (T*)out, nvec);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_act_bwd(const void* dy, const void* y, void* dx, int64_t n, int act, float slope, int dtype, void* stream) {
  EGP_REQUIRE(dy && y && dx, "act_bwd: null pointer");
  EGP_REQUIRE(act == EGP_ACT_RELU || act == EGP_ACT_LEAKY_RELU, "act_bwd: act must be relu or leaky relu");
  EGP_EW_CHECK("act_bwd", n, dtype);
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = n / Vec<T>::N;
    (void)launch_kernel(act_bwd_kernel<T>, ew_grid(nvec), kEwThreads, 0, (cudaStream_t)stream, (const T*)dy, (const T*)y, (T*)dx,
                                                                              nvec, act, slope);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_mask_scale(const void* x, const uint8_t* mask, void* out, int64_t n, float scale, int dtype, void* stream) {
  EGP_REQUIRE(x && mask && out, "mask_scale: null pointer");
  EGP_EW_CHECK("mask_scale", n, dtype);
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = n / Vec<T>::N;
    (void)launch_kernel(mask_scale_kernel<T>, ew_grid(nvec), kEwThreads, 0, (cudaStream_t)stream, (const T*)x, mask, (T*)out, nvec, scale);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_max_combine_fwd(const void* f, const void* m, void* a, int64_t n, int dtype, void* stream) {
  EGP_REQUIRE(f && m && a, "max_combine_fwd: null pointer");
  EGP_EW_CHECK("max_combine_fwd", n, dtype);
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = n / Vec<T>::N;
    (void)launch_kernel(max_combine_fwd_kernel<T>, ew_grid(nvec), kEwThreads, 0, (cudaStream_t)stream, (const T*)f, (const T*)m, (T*)a, nvec);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_max_combine_bwd(const void* da, const void* f, const void* m, void* df, int64_t n, int dtype, void* stream) {
  EGP_REQUIRE(da && f && m && df, "max_combine_bwd: null pointer");
  EGP_EW_CHECK("max_combine_bwd", n, dtype);
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = n / Vec<T>::N;
    (void)launch_kernel(max_combine_bwd_kernel<T>, ew_grid(nvec), kEwThreads, 0, (cudaStream_t)stream, (const T*)da, (const T*)f,
                                                                                      (const T*)m, (T*)df, nvec);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

size_t egp_colsum_workspace(int64_t rows, int64_t cols) { return colsum_workspace_bytes(rows, cols); }

int egp_colsum(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx, int dtype, void* workspace,
               size_t ws_bytes, void* stream) {
  return colsum_launch(x, out, rows, cols, ldx, dtype, workspace, ws_bytes, (cudaStream_t)stream);
}

size_t egp_act_bwd_colsum_workspace(int64_t rows, int64_t cols) {
  return sizeof(float) * (size_t)(sm_count() * 8) * (size_t)cols + colsum_workspace_bytes(rows, cols) + 128;
}

int egp_act_bwd_colsum(const void* dy, const void* y, void* dx, float* dx_colsum, int64_t rows, int64_t cols, int act,
                       float slope, int dtype, void* workspace, size_t ws_bytes, void* stream) {
  EGP_REQUIRE(dy && y && dx && dx_colsum && workspace, "act_bwd_colsum: null pointer");
  EGP_REQUIRE(act == EGP_ACT_RELU || act == EGP_ACT_LEAKY_RELU, "act_bwd_colsum: act must be relu or leaky relu");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(cols % vn == 0, "act_bwd_colsum: cols must be a multiple of %d", (int)vn);
  if (ws_bytes < egp_act_bwd_colsum_workspace(rows, cols)) {
    set_error("act_bwd_colsum: workspace too small");
    return EGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (rows == 0) {
    EGP_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * cols, s));
    return EGP_OK;
  }
  float* part = (float*)workspace;
  void* cs_ws = part + (size_t)(sm_count() * 8) * cols;
  EGP_DISPATCH_DTYPE(dtype, T, {
    constexpr int VN = Vec<T>::N;
    const int64_t nvec = rows * cols / VN, period = cols / VN;
    const int grid = ew_grid(nvec);
    const bool fuse = ((int64_t)grid * kEwThreads) % period == 0 && period <= kEwThreads && kEwThreads % period == 0;
    if (fuse) {
      (void)launch_kernel(act_bwd_colsum_kernel<T>, grid, kEwThreads, 0, s, (const T*)dy, (const T*)y, (T*)dx, nvec, (int)period, act, slope, part);
      EGP_LAUNCH_CHECK();
      (void)launch_kernel(colsum_final_kernel, (unsigned)ceil_div(cols, kFinalCols), kFinalThreads, 0, s, part, grid, cols, dx_colsum);
      EGP_LAUNCH_CHECK();
    } else {
      (void)launch_kernel(act_bwd_kernel<T>, grid, kEwThreads, 0, s, (const T*)dy, (const T*)y, (T*)dx, nvec, act, slope);
      EGP_LAUNCH_CHECK();
      const int rc = colsum_launch(dx, dx_colsum, rows, cols, cols, dtype, cs_ws, colsum_workspace_bytes(rows, cols), s);
      if (rc != EGP_OK) return rc;
    }
  });
  return EGP_OK;
}

int egp_row_inv_norm(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx, int dtype, void* stream) {
  EGP_REQUIRE(x && out, "row_inv_norm: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(cols % vn == 0 && ldx % vn == 0 && aligned16(x), "row_inv_norm: cols/ld must be multiples of %d", (int)vn);
  if (rows == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(dtype, T, {
    (void)launch_kernel(row_inv_norm_kernel<T>, (unsigned)ceil_div(rows, kEwThreads / 32), kEwThreads, 0, (cudaStream_t)stream, 
        (const T*)x, out, rows, cols, ldx);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

}  // extern "C"
