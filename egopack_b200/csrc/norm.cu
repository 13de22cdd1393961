// Normalisation kernels -- HBM bound.
//
//  graph-mode LayerNorm (+LeakyReLU): gnn.LayerNorm(H) called without `batch` (models/graph.py:43-44), i.e.
//      mu = mean over ALL N*C elements, sigma = sqrt(mean((x-mu)^2)), y = act((x-mu)/(sigma+eps)*w_c + b_c).
//      forward : stats pass (read C*b per node) + apply pass (read+write 2*C*b)          = 3*C*b per node
//      backward: reduce pass (read dy,x) + apply pass (read dy,x, write dx)               = 5*C*b per node
//  row LayerNorm (+ReLU): nn.LayerNorm in TRNPooling / task heads / GraphONE stages.
//      forward : one warp per row, row cached in registers                                = 2*C*b per node
//      backward: dx by one warp per row; dweight/dbias by per-lane column accumulators
#include <type_traits>

#include "common.cuh"

namespace egp {

// elementwise.cu
size_t colsum_workspace_bytes(int64_t rows, int64_t cols);
int colsum_launch(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx, int dtype, void* workspace,
                  size_t ws_bytes, cudaStream_t stream);

constexpr int kNormThreads = 256;

// ---------------------------------------------------------------------------------------------------------
// graph LayerNorm forward
// ---------------------------------------------------------------------------------------------------------
// Segments: the statistics of gnn.LayerNorm(batch=None) span ONE forward call.  Graph.forward_many runs several task
// batches through the shared weights as one tensor, so the rows are split into up to kMaxGlnSegs consecutive
// segments (one per original forward call), each with its own mu / sigma; dweight / dbias sum over all of them.
constexpr int kMaxGlnSegs = 8;
struct GlnSegs {
  int count;
  int64_t row[kMaxGlnSegs + 1];  // segment s = rows [row[s], row[s+1])
};

struct GlnWorkspace {            // layout of the caller-provided workspace
  unsigned int ticket[kMaxGlnSegs];   // last-block election, per segment
  double scal[2 * kMaxGlnSegs];  // backward, per segment: S1 = sum(g_hat), S2 = sum(g_hat * (x-mu))
};                               // followed by: double partial[2*G*S]; float colpart[G*S][2][C]

template <typename T>
__global__ void __launch_bounds__(kNormThreads)
gln_stats_kernel(const T* __restrict__ x, GlnSegs segs, int64_t channels, double* __restrict__ partial,
                 unsigned int* __restrict__ ticket, double* __restrict__ stats) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  __shared__ double red[32];
  __shared__ bool is_last;
  const int seg = blockIdx.y;
  const int64_t rows = segs.row[seg + 1] - segs.row[seg];
  const int64_t nvec = rows * channels / VN;
  x += segs.row[seg] * channels;
  partial += (size_t)seg * 2 * gridDim.x;
  double s = 0.0, q = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; v + 3 * stride < nvec; v += 4 * stride) {  // four independent 16-byte loads in flight per thread
    Raw<T> r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) r[u] = Raw<T>::load(x + (v + u * stride) * VN);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const Vec<T> a = r[u].unpack();
      float ls = 0.f, lq = 0.f;
#pragma unroll
      for (int c = 0; c < VN; ++c) { ls += a.v[c]; lq += a.v[c] * a.v[c]; }
      s += (double)ls;
      q += (double)lq;
    }
  }
  for (; v < nvec; v += stride) {
    const Vec<T> a = Vec<T>::load(x + v * VN);
    float ls = 0.f, lq = 0.f;
#pragma unroll
    for (int c = 0; c < VN; ++c) { ls += a.v[c]; lq += a.v[c] * a.v[c]; }
    s += (double)ls;
    q += (double)lq;
  }
  s = block_sum(s, red);
  q = block_sum(q, red);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = s;
    partial[2 * blockIdx.x + 1] = q;
    __threadfence();
    const unsigned int t = atomicAdd(ticket + seg, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {  // deterministic final reduction: fixed order over the per-block partials
    __threadfence();
    double ts = 0.0, tq = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
      ts += partial[2 * b];
      tq += partial[2 * b + 1];
    }
    ts = block_sum(ts, red);
    tq = block_sum(tq, red);
    if (threadIdx.x == 0) {
      const double inv_count = rows > 0 ? 1.0 / ((double)rows * (double)channels) : 0.0;
      const double mu = ts * inv_count;
      double var = tq * inv_count - mu * mu;
      var = var > 0.0 ? var : 0.0;
      stats[2 * seg] = mu;
      stats[2 * seg + 1] = sqrt(var);
      ticket[seg] = 0u;
    }
  }
}

// Statistics of a segment from the {sum, sum of squares} pairs that the producing GEMM's epilogue left per (128-row
// block, TMEM quarter, n tile) (egp_gemm_rowstats): one block per segment, fixed summation order (deterministic).
// Replaces gln_stats_kernel -- a full read of the tensor -- when the input of the LayerNorm comes straight out of a GEMM.
constexpr int kRowstatThreads = 1024;
__global__ void __launch_bounds__(kRowstatThreads)
gln_stats_from_rowstats_kernel(const double* __restrict__ rowstats, int pairs_per_rowblock, GlnSegs segs, int64_t channels,
                               double* __restrict__ stats) {
  pdl_enter();
  __shared__ double red[32];
  const int seg = blockIdx.x;
  const int64_t rb0 = segs.row[seg] / 128, rb1 = (segs.row[seg + 1] + 127) / 128;
  const double2* pr = reinterpret_cast<const double2*>(rowstats) + rb0 * pairs_per_rowblock;
  const int64_t count = (rb1 - rb0) * pairs_per_rowblock;
  double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
  int64_t i = threadIdx.x;
  for (; i + 3 * kRowstatThreads < count; i += 4 * kRowstatThreads) {   // four independent 16-byte loads in flight
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const double2 v = pr[i + u * kRowstatThreads];
      s[u] += v.x;
      q[u] += v.y;
    }
  }
  for (; i < count; i += kRowstatThreads) {
    const double2 v = pr[i];
    s[0] += v.x;
    q[0] += v.y;
  }
  double ts = block_sum((s[0] + s[1]) + (s[2] + s[3]), red);
  double tq = block_sum((q[0] + q[1]) + (q[2] + q[3]), red);
  if (threadIdx.x == 0) {
    const int64_t rows = segs.row[seg + 1] - segs.row[seg];
    const double inv_count = rows > 0 ? 1.0 / ((double)rows * (double)channels) : 0.0;
    const double mu = ts * inv_count;
    double var = tq * inv_count - mu * mu;
    var = var > 0.0 ? var : 0.0;
    stats[2 * seg] = mu;
    stats[2 * seg + 1] = sqrt(var);
  }
}

// FIXED: the grid stride is a multiple of the row length, so a thread always lands on the same 16-byte column and
// keeps that column's weight/bias in registers (otherwise they are re-read, vectorised, every iteration).
template <typename T, bool FIXED>
__global__ void __launch_bounds__(kNormThreads)
gln_apply_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 T* __restrict__ y, const double* __restrict__ stats, GlnSegs segs, int64_t channels, float eps,
                 int act, float slope) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const int seg = blockIdx.y;
  const int64_t nvec = (segs.row[seg + 1] - segs.row[seg]) * channels / VN;
  x += segs.row[seg] * channels;
  y += segs.row[seg] * channels;
  const float mu = (float)stats[2 * seg];
  const float rs = (float)(1.0 / (stats[2 * seg + 1] + (double)eps));
  const int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float wv[VN], bv[VN];
  if (FIXED) {
    const int64_t c0 = (v0 * VN) % channels;
#pragma unroll
    for (int c = 0; c < VN; c += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + c0 + c);
      const float4 b4 = *reinterpret_cast<const float4*>(b + c0 + c);
      wv[c] = w4.x; wv[c + 1] = w4.y; wv[c + 2] = w4.z; wv[c + 3] = w4.w;
      bv[c] = b4.x; bv[c + 1] = b4.y; bv[c + 2] = b4.z; bv[c + 3] = b4.w;
    }
  }
#pragma unroll 2
  for (int64_t v = v0; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    if (!FIXED) {
      const int64_t c0 = (v * VN) % channels;
#pragma unroll
      for (int c = 0; c < VN; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + c0 + c);
        const float4 b4 = *reinterpret_cast<const float4*>(b + c0 + c);
        wv[c] = w4.x; wv[c + 1] = w4.y; wv[c + 2] = w4.z; wv[c + 3] = w4.w;
        bv[c] = b4.x; bv[c + 1] = b4.y; bv[c + 2] = b4.z; bv[c + 3] = b4.w;
      }
    }
    Vec<T> a = Vec<T>::load(x + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) a.v[c] = apply_act((a.v[c] - mu) * rs * wv[c] + bv[c], act, slope);
    a.store(y + v * VN);
  }
}

// ---------------------------------------------------------------------------------------------------------
// graph LayerNorm backward
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float pre, int act, float slope) {
  if (act == EGP_ACT_RELU) return pre > 0.f ? 1.f : 0.f;
  if (act == EGP_ACT_LEAKY_RELU) return pre > 0.f ? 1.f : slope;
  return 1.f;
}

// grid (G row strips, column chunks); thread owns one 16-byte column
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
gln_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ b, const double* __restrict__ stats, GlnSegs segs, int64_t channels,
                      int rows_per_cta, float eps, int act, float slope, float* __restrict__ colpart,
                      double* __restrict__ scalpart) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  __shared__ double red[32];
  const int seg = blockIdx.z;
  const int64_t col = ((int64_t)blockIdx.y * blockDim.x + threadIdx.x) * VN;
  const bool live = col < channels;
  const float mu = (float)stats[2 * seg];
  const float rs = (float)(1.0 / (stats[2 * seg + 1] + (double)eps));
  float wv[VN], bv[VN], dw[VN], db[VN];
#pragma unroll
  for (int c = 0; c < VN; ++c) {
    wv[c] = live ? w[col + c] : 0.f;
    bv[c] = live ? b[col + c] : 0.f;
    dw[c] = 0.f;
    db[c] = 0.f;
  }
  double s1 = 0.0, s2 = 0.0;
  const int64_t r0 = min(segs.row[seg] + (int64_t)blockIdx.x * rows_per_cta, segs.row[seg + 1]);
  const int64_t r1 = min(r0 + rows_per_cta, segs.row[seg + 1]);
  if (live) {
    auto row_update = [&](const Vec<T>& g, const Vec<T>& a) {
      float l1 = 0.f, l2 = 0.f;
#pragma unroll
      for (int c = 0; c < VN; ++c) {
        const float d = a.v[c] - mu;
        const float xh = d * rs;
        const float gp = g.v[c] * act_grad(xh * wv[c] + bv[c], act, slope);
        dw[c] += gp * xh;
        db[c] += gp;
        const float gh = gp * wv[c];
        l1 += gh;
        l2 += gh * d;
      }
      s1 += (double)l1;
      s2 += (double)l2;
    };
    int64_t i = r0;
    for (; i + 4 <= r1; i += 4) {  // four rows = eight independent 16-byte loads in flight per thread
      Raw<T> gr[4], ar[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        gr[u] = Raw<T>::load(dy + (i + u) * channels + col);
        ar[u] = Raw<T>::load(x + (i + u) * channels + col);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) row_update(gr[u].unpack(), ar[u].unpack());
    }
    for (; i < r1; ++i) row_update(Vec<T>::load(dy + i * channels + col), Vec<T>::load(x + i * channels + col));
    float* cp = colpart + ((size_t)seg * gridDim.x + blockIdx.x) * 2 * channels;
#pragma unroll
    for (int c = 0; c < VN; ++c) {
      cp[col + c] = dw[c];
      cp[channels + col + c] = db[c];
    }
  }
  s1 = block_sum(s1, red);
  s2 = block_sum(s2, red);
  if (threadIdx.x == 0) {
    const size_t p = (((size_t)seg * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 2;
    scalpart[p] = s1;
    scalpart[p + 1] = s2;
  }
}

// Deterministic reduction of per-CTA column partials.  colpart is [parts][nout][channels]; a block owns kFinCols
// columns and splits the part rows over 1024 / kFinCols groups (few serial iterations per thread, ~128 CTAs for 1024
// channels: the kernel is pure latency), then combines the groups in a fixed order: shuffles inside a warp, shared
// memory across warps.  out_k[c] = sum_g colpart[g][k][c].
// Blocks colblocks.. (one per graph-LN segment, if scal != null) reduce the scalar pairs: scal[2s..2s+1] = sum over
// that segment's scal_parts entries of scalpart.
constexpr int kFinThreads = 1024;
constexpr int kFinCols = 8;                       // columns per block: 32-byte segments of a partial row
constexpr int kFinGroups = kFinThreads / kFinCols;
__global__ void __launch_bounds__(kFinThreads)
col_finalize_kernel(const float* __restrict__ colpart, int parts, int64_t channels, int nout, float* __restrict__ out0,
                    float* __restrict__ out1, float* __restrict__ out2, const double* __restrict__ scalpart,
                    int scal_parts, double* __restrict__ scal, int colblocks) {
  pdl_enter();
  __shared__ double red[32];
  __shared__ float rs[3][32][kFinCols];
  if ((int)blockIdx.x < colblocks) {
    const int tx = threadIdx.x % kFinCols, ty = threadIdx.x / kFinCols;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * kFinCols + tx;
    float a[3] = {0.f, 0.f, 0.f};
    if (c < channels) {
      const size_t stride = (size_t)nout * channels;
      int g = ty;
      for (; g + 3 * kFinGroups < parts; g += 4 * kFinGroups) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          if (k < nout) {
            const float* p = colpart + (size_t)k * channels + c;
            const float v0 = p[(size_t)g * stride], v1 = p[(size_t)(g + kFinGroups) * stride];
            const float v2 = p[(size_t)(g + 2 * kFinGroups) * stride], v3 = p[(size_t)(g + 3 * kFinGroups) * stride];
            a[k] += (v0 + v1) + (v2 + v3);
          }
        }
      }
      for (; g < parts; g += kFinGroups)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          if (k < nout) a[k] += colpart[(size_t)g * stride + (size_t)k * channels + c];
    }
    // a warp holds 32 / kFinCols consecutive groups of the same kFinCols columns
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int o = kFinCols; o < 32; o <<= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
      if (lane < kFinCols) rs[k][warp][lane] = a[k];
    }
    __syncthreads();
    if (warp < 3 && warp < nout) {  // warp k finishes output k: lane = (quarter of the 32 warp partials, column)
      const int col = lane % kFinCols, part0 = (lane / kFinCols) * (32 / (32 / kFinCols));
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 32 / (32 / kFinCols); ++i) t += rs[warp][part0 + i][col];
#pragma unroll
      for (int o = kFinCols; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      const int64_t cc = (int64_t)blockIdx.x * kFinCols + col;
      float* o = warp == 0 ? out0 : (warp == 1 ? out1 : out2);
      if (lane < kFinCols && cc < channels && o) o[cc] = t;
    }
  } else {  // one extra block per segment reduces that segment's scalar pairs
    const int seg = (int)blockIdx.x - colblocks;
    scalpart += (size_t)seg * 2 * scal_parts;
    double s1 = 0.0, s2 = 0.0;
    for (int p = threadIdx.x; p < scal_parts; p += blockDim.x) {
      s1 += scalpart[2 * p];
      s2 += scalpart[2 * p + 1];
    }
    s1 = block_sum(s1, red);
    s2 = block_sum(s2, red);
    if (threadIdx.x == 0) { scal[2 * seg] = s1; scal[2 * seg + 1] = s2; }
  }
}

template <typename T, bool FIXED>
__global__ void __launch_bounds__(kNormThreads)
gln_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ b, const double* __restrict__ stats, const double* __restrict__ scal,
                     T* __restrict__ dx, GlnSegs segs, int64_t channels, float eps, int act,
                     float slope, float* __restrict__ dxpart /* [gridDim.y * gridDim.x, C] or null; FIXED only */) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  __shared__ float csum[kNormThreads * VN];
  float dsum[VN];
#pragma unroll
  for (int c = 0; c < VN; ++c) dsum[c] = 0.f;
  const int seg = blockIdx.y;
  const int64_t seg_rows = segs.row[seg + 1] - segs.row[seg];
  const int64_t nvec = seg_rows * channels / VN;
  dy += segs.row[seg] * channels;
  x += segs.row[seg] * channels;
  dx += segs.row[seg] * channels;
  const double inv_count = seg_rows > 0 ? 1.0 / ((double)seg_rows * (double)channels) : 0.0;
  const double sigma = stats[2 * seg + 1];
  const float mu = (float)stats[2 * seg];
  const float rs = (float)(1.0 / (sigma + (double)eps));
  const float m1 = (float)(scal[2 * seg] * inv_count);  // mean(g_hat)
  const double den = sigma * (sigma + (double)eps) * (sigma + (double)eps);
  const float k2 = den > 0.0 ? (float)(scal[2 * seg + 1] * inv_count / den) : 0.f;
  const int64_t v0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float wv[VN], bv[VN];
  auto load_params = [&](int64_t c0) {
#pragma unroll
    for (int c = 0; c < VN; c += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + c0 + c);
      const float4 b4 = *reinterpret_cast<const float4*>(b + c0 + c);
      wv[c] = w4.x; wv[c + 1] = w4.y; wv[c + 2] = w4.z; wv[c + 3] = w4.w;
      bv[c] = b4.x; bv[c + 1] = b4.y; bv[c + 2] = b4.z; bv[c + 3] = b4.w;
    }
  };
  if (FIXED) load_params((v0 * VN) % channels);
#pragma unroll 2
  for (int64_t v = v0; v < nvec; v += (int64_t)gridDim.x * blockDim.x) {
    if (!FIXED) load_params((v * VN) % channels);
    const Vec<T> g = Vec<T>::load(dy + v * VN);
    Vec<T> a = Vec<T>::load(x + v * VN);
#pragma unroll
    for (int c = 0; c < VN; ++c) {
      const float d = a.v[c] - mu;
      const float xh = d * rs;
      const float gh = g.v[c] * act_grad(xh * wv[c] + bv[c], act, slope) * wv[c];
      a.v[c] = (gh - m1) * rs - d * k2;
      if (FIXED) dsum[c] += to_float<T>(from_float<T>(a.v[c]));  // column sum of dx as stored
    }
    a.store(dx + v * VN);
  }
  if (FIXED && dxpart)
    block_column_partial<VN>(dsum, (int)(channels / VN), dxpart + ((size_t)seg * gridDim.x + blockIdx.x) * channels, csum);
}

// ---------------------------------------------------------------------------------------------------------
// row LayerNorm: one warp per row, lane owns vectors lane, lane+32, ...  (NVL of them, cached in registers)
// ---------------------------------------------------------------------------------------------------------
// weight / bias live in shared memory as float4, laid out [NVL][VN/4][lane] so a warp's reads are conflict-free;
// the NEXT row of the warp is already in flight (packed registers) while the current one is reduced and written.
template <typename T, int NVL>
__global__ void __launch_bounds__(kNormThreads, NVL * (16 / (int)sizeof(T)) <= 16 ? 4 : (NVL * (16 / (int)sizeof(T)) <= 32 ? 3 : 2))
rln_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
               T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t n, int64_t channels,
               float eps, int act, uint32_t drop_thr16, float keep_scale, uint64_t seed, uint64_t offset,
               const uint64_t* __restrict__ rng_state) {
  pdl_enter();
  // CUDA-graph replays cannot change kernel arguments, so the per-step part of the Philox stream may come from device
  // memory: rng_state = {seed, step}; the call-site index stays in `offset`
  if (rng_state) { seed = rng_state[0]; offset += rng_state[1] << 20; }
  const uint32_t drop_key = drop_thr16 ? dropout_key(seed, offset) : 0u;   // once per thread, not per vector
  constexpr int VN = Vec<T>::N;
  constexpr int Q = VN / 4;  // float4 pieces per 16-byte vector of T
  __shared__ float4 sw[NVL * Q * 32], sb[NVL * Q * 32];
  const int lane = threadIdx.x & 31;
  const int nvec = (int)(channels / VN);
  for (int i = threadIdx.x; i < NVL * Q * 32; i += blockDim.x) {
    const int ln = i & 31, h = (i >> 5) % Q, it = (i >> 5) / Q;
    const int v = ln + 32 * it;
    float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = w4;
    if (v < nvec) {
      w4 = *reinterpret_cast<const float4*>(w + v * VN + 4 * h);
      b4 = *reinterpret_cast<const float4*>(b + v * VN + 4 * h);
    }
    sw[i] = w4;
    sb[i] = b4;
  }
  __syncthreads();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float inv_c = 1.f / (float)channels;
  Raw<T> cur[NVL], nxt[NVL];
  auto load_row = [&](Raw<T>* dst, int64_t r) {
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      const int v = lane + 32 * it;
      dst[it] = v < nvec ? Raw<T>::load(x + r * channels + (int64_t)v * VN) : Raw<T>::zero();
    }
  };
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < n) load_row(cur, row);
  for (; row < n; row += warps) {
    if (row + warps < n) load_row(nxt, row + warps);
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      const Vec<T> a = cur[it].unpack();   // lanes past the row end hold zeros
#pragma unroll
      for (int c = 0; c < VN; ++c) s += a.v[c];
    }
    const float mu = warp_sum(s) * inv_c;
    float q = 0.f;
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      if (lane + 32 * it < nvec) {
        const Vec<T> a = cur[it].unpack();
#pragma unroll
        for (int c = 0; c < VN; ++c) { const float d = a.v[c] - mu; q += d * d; }
      }
    }
    const float rs = rsqrtf(warp_sum(q) * inv_c + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    T* yr = y + row * channels;
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      const int v = lane + 32 * it;
      if (v < nvec) {
        const Vec<T> a = cur[it].unpack();
        Vec<T> o;
#pragma unroll
        for (int h = 0; h < Q; ++h) {
          const float4 w4 = sw[(it * Q + h) * 32 + lane], b4 = sb[(it * Q + h) * 32 + lane];
          o.v[4 * h + 0] = apply_act((a.v[4 * h + 0] - mu) * rs * w4.x + b4.x, act, 0.f);
          o.v[4 * h + 1] = apply_act((a.v[4 * h + 1] - mu) * rs * w4.y + b4.y, act, 0.f);
          o.v[4 * h + 2] = apply_act((a.v[4 * h + 2] - mu) * rs * w4.z + b4.z, act, 0.f);
          o.v[4 * h + 3] = apply_act((a.v[4 * h + 3] - mu) * rs * w4.w + b4.w, act, 0.f);
        }
        if (drop_thr16) {  // fused inverted dropout (nn.Dropout after the ReLU, trn_pooling.py:31,36)
          const uint32_t keep = dropout_keep_bits_keyed((uint64_t)row * nvec + v, drop_key, drop_thr16);
#pragma unroll
          for (int c = 0; c < VN; ++c) o.v[c] = ((keep >> c) & 1u) ? o.v[c] * keep_scale : 0.f;
        }
        o.store(yr + (int64_t)v * VN);
      }
    }
#pragma unroll
    for (int it = 0; it < NVL; ++it) cur[it] = nxt[it];
  }
}

// The common shape -- every lane owns exactly NVL full vectors (channels == 32 * NVL * VN), activation known at compile
// time -- without the per-vector guards and runtime activation branches of rln_fwd_kernel, and with the arithmetic in
// packed fp32x2 instructions: 43 % fewer instructions executed (ncu, 1024 bf16 channels, ReLU + dropout: 68.9 M -> 39.5 M
// warp instructions, issue slots 77 % -> 43 %).  Same operation order per element ((x - mu) * rstd * w + b, ReLU,
// * keep_scale), so results are bit-identical.  Measured same-box at 98 304 rows: 88.1 -> 75.3 us with dropout (4.57 ->
// 5.35 TB/s); without dropout both kernels sit at 73.7 us (5.46 TB/s), already bound by memory.
// Dropout zeroes the dropped elements of the PACKED output with byte masks (dropout_mask_packed).  A shared-memory table
// of multiplier pairs indexed by the keep bits measured no faster than the guarded kernel: the randomly indexed 16-byte
// loads conflict (ncu: 9.4 short-scoreboard stalls per issue).
template <typename T, int NVL, bool RELU, bool DROP>
__global__ void __launch_bounds__(kNormThreads, NVL * (16 / (int)sizeof(T)) <= 16 ? 4 : (NVL * (16 / (int)sizeof(T)) <= 32 ? 3 : 2))
rln_fwd_full_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                    T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t n, float eps,
                    uint32_t drop_thr16, float keep_scale, uint64_t seed, uint64_t offset,
                    const uint64_t* __restrict__ rng_state) {
  pdl_enter();
  if (rng_state) { seed = rng_state[0]; offset += rng_state[1] << 20; }
  const uint32_t drop_key = DROP ? dropout_key(seed, offset) : 0u;
  constexpr int VN = Vec<T>::N;
  constexpr int NP = Pairs<T>::NP;
  constexpr int Q = VN / 4;  // float4 pieces per 16-byte vector of T
  constexpr int nvec = 32 * NVL;
  constexpr int64_t channels = (int64_t)nvec * VN;
  __shared__ float4 sw[NVL * Q * 32], sb[NVL * Q * 32];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NVL * Q * 32; i += blockDim.x) {
    const int ln = i & 31, h = (i >> 5) % Q, it = (i >> 5) / Q;
    const int v = ln + 32 * it;
    sw[i] = *reinterpret_cast<const float4*>(w + v * VN + 4 * h);
    sb[i] = *reinterpret_cast<const float4*>(b + v * VN + 4 * h);
  }
  __syncthreads();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  constexpr float inv_c = 1.f / (float)channels;
  Raw<T> cur[NVL], nxt[NVL];
  auto load_row = [&](Raw<T>* dst, int64_t r) {
    const T* xr = x + r * channels + (int64_t)lane * VN;
#pragma unroll
    for (int it = 0; it < NVL; ++it) dst[it] = Raw<T>::load(xr + it * 32 * VN);
  };
  int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row < n) load_row(cur, row);
  for (; row < n; row += warps) {
    if (row + warps < n) load_row(nxt, row + warps);
    Pairs<T> a[NVL];
    float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      a[it] = Pairs<T>::from(cur[it]);
#pragma unroll
      for (int c = 0; c < NP; ++c) s2 = __fadd2_rn(s2, a[it].p[c]);
    }
    const float mu = warp_sum(s2.x + s2.y) * inv_c;
    const float2 nmu = splat2(-mu);
    float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
#pragma unroll
      for (int c = 0; c < NP; ++c) {
        a[it].p[c] = __fadd2_rn(a[it].p[c], nmu);
        q2 = __ffma2_rn(a[it].p[c], a[it].p[c], q2);
      }
    }
    const float rs = rsqrtf(warp_sum(q2.x + q2.y) * inv_c + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    const float2 rs2 = splat2(rs), ks2 = splat2(keep_scale);
    T* yr = y + row * channels + (int64_t)lane * VN;
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      Pairs<T> o;
#pragma unroll
      for (int h = 0; h < Q; ++h) {
        const float4 w4 = sw[(it * Q + h) * 32 + lane], b4 = sb[(it * Q + h) * 32 + lane];
        o.p[2 * h] = __ffma2_rn(__fmul2_rn(a[it].p[2 * h], rs2), make_float2(w4.x, w4.y), make_float2(b4.x, b4.y));
        o.p[2 * h + 1] = __ffma2_rn(__fmul2_rn(a[it].p[2 * h + 1], rs2), make_float2(w4.z, w4.w), make_float2(b4.z, b4.w));
      }
      if constexpr (RELU) {
#pragma unroll
        for (int c = 0; c < NP; ++c) o.p[c] = make_float2(fmaxf(o.p[c].x, 0.f), fmaxf(o.p[c].y, 0.f));
      }
      uint4 pk;
      if constexpr (DROP) {  // fused inverted dropout (nn.Dropout after the ReLU, trn_pooling.py:31,36)
#pragma unroll
        for (int c = 0; c < NP; ++c) o.p[c] = __fmul2_rn(o.p[c], ks2);
        pk = o.pack();
        dropout_mask_packed<T>(pk, dropout_keep_bits_keyed((uint64_t)row * nvec + (lane + 32 * it), drop_key, drop_thr16));
      } else {
        pk = o.pack();
      }
      *reinterpret_cast<uint4*>(yr + it * 32 * VN) = pk;
    }
#pragma unroll
    for (int it = 0; it < NVL; ++it) cur[it] = nxt[it];
  }
}

template <typename T, int NVL>
__global__ void __launch_bounds__(kNormThreads)
rln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
               const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd,
               T* __restrict__ dx, int64_t n, int64_t channels, int act, float out_scale,
               float* __restrict__ colpart) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const bool use_y = act == EGP_ACT_RELU || out_scale != 1.f;  // mask = output != 0 (ReLU and/or dropout)
  extern __shared__ float cacc[];  // [2][channels] block accumulators for dweight / dbias
  const int lane = threadIdx.x & 31;
  const int nvec = (int)(channels / VN);
  for (int64_t c = threadIdx.x; c < 2 * channels; c += blockDim.x) cacc[c] = 0.f;
  __syncthreads();
  float dw[NVL][VN], db[NVL][VN], wv[NVL][VN];
#pragma unroll
  for (int it = 0; it < NVL; ++it) {
    const int v = lane + 32 * it;
#pragma unroll
    for (int c = 0; c < VN; ++c) { dw[it][c] = 0.f; db[it][c] = 0.f; wv[it][c] = 0.f; }
    if (v < nvec) {
#pragma unroll
      for (int c = 0; c < VN; c += 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(w + v * VN + c);
        wv[it][c] = w4.x; wv[it][c + 1] = w4.y; wv[it][c + 2] = w4.z; wv[it][c + 3] = w4.w;
      }
    }
  }

  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    const float mu = mean[row], rs = rstd[row];
    Vec<T> gh[NVL], xh[NVL];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      const int v = lane + 32 * it;
      if (v < nvec) {
        const int64_t o = row * channels + (int64_t)v * VN;
        const Vec<T> g = Vec<T>::load(dy + o);
        const Vec<T> a = Vec<T>::load(x + o);
        Vec<T> yo;
        if (use_y) yo = Vec<T>::load(y + o);
#pragma unroll
        for (int c = 0; c < VN; ++c) {
          const float h = (a.v[c] - mu) * rs;
          float gp = g.v[c] * out_scale;
          if (use_y && yo.v[c] == 0.f) gp = 0.f;
          dw[it][c] += gp * h;
          db[it][c] += gp;
          const float t = gp * wv[it][c];
          gh[it].v[c] = t;
          xh[it].v[c] = h;
          s1 += t;
          s2 += t * h;
        }
      }
    }
    const float c1 = warp_sum(s1) / (float)channels;
    const float c2 = warp_sum(s2) / (float)channels;
#pragma unroll
    for (int it = 0; it < NVL; ++it) {
      const int v = lane + 32 * it;
      if (v < nvec) {
        Vec<T> o;
#pragma unroll
        for (int c = 0; c < VN; ++c) o.v[c] = rs * (gh[it].v[c] - c1 - xh[it].v[c] * c2);
        o.store(dx + row * channels + (int64_t)v * VN);
      }
    }
  }
#pragma unroll
  for (int it = 0; it < NVL; ++it) {
    const int v = lane + 32 * it;
    if (v < nvec) {
#pragma unroll
      for (int c = 0; c < VN; ++c) {
        atomicAdd(&cacc[v * VN + c], dw[it][c]);
        atomicAdd(&cacc[channels + v * VN + c], db[it][c]);
      }
    }
  }
  __syncthreads();
  float* cp = colpart + (size_t)blockIdx.x * 2 * channels;
  for (int64_t c = threadIdx.x; c < 2 * channels; c += blockDim.x) cp[c] = cacc[c];
}

// Row-LN backward, one CTA per row (rows strided over the grid): thread t owns the 16-byte column t of every row its
// CTA visits, so dweight/dbias accumulate in registers with no atomics, the two row statistics need one
// __syncthreads per row (double-buffered scratch), and the register footprint stays small enough for full occupancy.
template <typename T, int MAXT>
__global__ void __launch_bounds__(MAXT)
rln_bwd_block_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                     const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd,
                     T* __restrict__ dx, int64_t n, int64_t channels, int act, float out_scale, int nout,
                     float* __restrict__ colpart) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const bool use_y = act == EGP_ACT_RELU || out_scale != 1.f;
  __shared__ float red[2][32][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int64_t col = (int64_t)threadIdx.x * VN;
  const bool live = col < channels;
  constexpr int NP = Pairs<T>::NP;
  float2 wv2[NP], dw2[NP], db2[NP], ds2[NP];
#pragma unroll
  for (int c = 0; c < NP; ++c) { wv2[c] = make_float2(0.f, 0.f); dw2[c] = wv2[c]; db2[c] = wv2[c]; ds2[c] = wv2[c]; }
  if (live) {
#pragma unroll
    for (int c = 0; c < VN; c += 4) {
      const float4 w4 = *reinterpret_cast<const float4*>(w + col + c);
      wv2[c / 2] = make_float2(w4.x, w4.y);
      wv2[c / 2 + 1] = make_float2(w4.z, w4.w);
    }
  }
  const float inv_c = 1.f / (float)channels;
  int buf = 0;
  // the next row's operands (and statistics) are requested before the current row is reduced
  Raw<T> ng = Raw<T>::zero(), na = ng, ny = ng;
  float nmu = 0.f, nrs = 0.f;
  auto prefetch = [&](int64_t r) {
    nrs = rstd[r];
    if (live) {
      nmu = mean[r];
      const int64_t o = r * channels + col;
      ng = Raw<T>::load(dy + o);
      na = Raw<T>::load(x + o);
      if (use_y) ny = Raw<T>::load(y + o);
    }
  };
  if ((int64_t)blockIdx.x < n) prefetch(blockIdx.x);
  // packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2): 128 M -> 102 M warp instructions at 98 304 x 1024 (ncu), issue slots
  // 72 % -> 40 %; same box 183 -> 177 us for the whole backward (the kernel is bound by load latency at 5 CTAs per SM)
  for (int64_t row = blockIdx.x; row < n; row += gridDim.x, buf ^= 1) {
    float2 gh[NP], xh[NP];
    float2 s1 = make_float2(0.f, 0.f), s2 = s1;
    const float rs = nrs, mu = nmu;
    const Pairs<T> g = Pairs<T>::from(ng), a = Pairs<T>::from(na), yo = Pairs<T>::from(ny);
    if (row + gridDim.x < n) prefetch(row + gridDim.x);
    // keep the next row's loads HERE: without the fence ptxas sinks them to their first use (ncu: 9.9 long-scoreboard
    // stalls per issue, 166 -> 230 us) to save registers
    asm volatile("" ::: "memory");
    if (live) {
      const float2 nmu2 = splat2(-mu), rs2 = splat2(rs), sc2 = splat2(out_scale);
#pragma unroll
      for (int c = 0; c < NP; ++c) {
        const float2 h = __fmul2_rn(__fadd2_rn(a.p[c], nmu2), rs2);
        float2 gp = __fmul2_rn(g.p[c], sc2);
        if (use_y) {
          gp.x = yo.p[c].x == 0.f ? 0.f : gp.x;
          gp.y = yo.p[c].y == 0.f ? 0.f : gp.y;
        }
        dw2[c] = __ffma2_rn(gp, h, dw2[c]);
        db2[c] = __fadd2_rn(db2[c], gp);
        const float2 t = __fmul2_rn(gp, wv2[c]);
        gh[c] = t;
        xh[c] = h;
        s1 = __fadd2_rn(s1, t);
        s2 = __ffma2_rn(t, h, s2);
      }
    }
    const float w1 = warp_sum(s1.x + s1.y);
    const float w2 = warp_sum(s2.x + s2.y);
    if (lane == 0) { red[buf][warp][0] = w1; red[buf][warp][1] = w2; }
    __syncthreads();
    float c1 = 0.f, c2 = 0.f;
    for (int i = 0; i < nw; ++i) { c1 += red[buf][i][0]; c2 += red[buf][i][1]; }
    c1 *= inv_c;
    c2 *= inv_c;
    if (live) {
      const float2 nc1 = splat2(-c1), nc2 = splat2(-c2), rs2 = splat2(rs);
      Pairs<T> o;
#pragma unroll
      for (int c = 0; c < NP; ++c) o.p[c] = __fmul2_rn(__ffma2_rn(xh[c], nc2, __fadd2_rn(gh[c], nc1)), rs2);
      const uint4 pk = o.pack();
      if (nout == 3) {   // column sum of dx AS STORED (bias gradient upstream)
        Raw<T> stored;
        stored.u = pk;
        const Pairs<T> r = Pairs<T>::from(stored);
#pragma unroll
        for (int c = 0; c < NP; ++c) ds2[c] = __fadd2_rn(ds2[c], r.p[c]);
      }
      *reinterpret_cast<uint4*>(dx + row * channels + col) = pk;
    }
  }
  if (live) {
    float* cp = colpart + (size_t)blockIdx.x * nout * channels;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
      *reinterpret_cast<float2*>(cp + col + 2 * c) = dw2[c];
      *reinterpret_cast<float2*>(cp + channels + col + 2 * c) = db2[c];
      if (nout == 3) *reinterpret_cast<float2*>(cp + 2 * channels + col + 2 * c) = ds2[c];
    }
  }
}

// generic fallbacks for very wide rows (no register cache; rows are re-read through L1/L2)
template <typename T>
__global__ void __launch_bounds__(kNormThreads)
rln_fwd_wide_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                    T* __restrict__ y, float* __restrict__ mean, float* __restrict__ rstd, int64_t n,
                    int64_t channels, float eps, int act, uint32_t drop_thr16, float keep_scale, uint64_t seed,
                    uint64_t offset, const uint64_t* __restrict__ rng_state) {
  pdl_enter();
  if (rng_state) { seed = rng_state[0]; offset += rng_state[1] << 20; }
  const uint32_t drop_key = drop_thr16 ? dropout_key(seed, offset) : 0u;
  constexpr int VN = Vec<T>::N;
  const int lane = threadIdx.x & 31;
  const int nvec = (int)(channels / VN);
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    const T* xr = x + row * channels;
    float s = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      const Vec<T> a = Vec<T>::load(xr + (int64_t)v * VN);
#pragma unroll
      for (int c = 0; c < VN; ++c) s += a.v[c];
    }
    const float mu = warp_sum(s) / (float)channels;
    float q = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      const Vec<T> a = Vec<T>::load(xr + (int64_t)v * VN);
#pragma unroll
      for (int c = 0; c < VN; ++c) { const float d = a.v[c] - mu; q += d * d; }
    }
    const float rs = rsqrtf(warp_sum(q) / (float)channels + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    for (int v = lane; v < nvec; v += 32) {
      Vec<T> a = Vec<T>::load(xr + (int64_t)v * VN);
#pragma unroll
      for (int c = 0; c < VN; ++c) a.v[c] = apply_act((a.v[c] - mu) * rs * w[v * VN + c] + b[v * VN + c], act, 0.f);
      if (drop_thr16) {
        const uint32_t keep = dropout_keep_bits_keyed((uint64_t)row * nvec + v, drop_key, drop_thr16);
#pragma unroll
        for (int c = 0; c < VN; ++c) a.v[c] = ((keep >> c) & 1u) ? a.v[c] * keep_scale : 0.f;
      }
      a.store(y + row * channels + (int64_t)v * VN);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kNormThreads)
rln_bwd_wide_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                    const float* __restrict__ w, const float* __restrict__ mean, const float* __restrict__ rstd,
                    T* __restrict__ dx, int64_t n, int64_t channels, int act, float out_scale,
                    float* __restrict__ colpart) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const bool use_y = act == EGP_ACT_RELU || out_scale != 1.f;
  extern __shared__ float cacc[];
  const int lane = threadIdx.x & 31;
  const int nvec = (int)(channels / VN);
  for (int64_t c = threadIdx.x; c < 2 * channels; c += blockDim.x) cacc[c] = 0.f;
  __syncthreads();
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      const int64_t o = row * channels + (int64_t)v * VN;
      const Vec<T> g = Vec<T>::load(dy + o);
      const Vec<T> a = Vec<T>::load(x + o);
      Vec<T> yo;
      if (use_y) yo = Vec<T>::load(y + o);
#pragma unroll
      for (int c = 0; c < VN; ++c) {
        const float h = (a.v[c] - mu) * rs;
        float gp = g.v[c] * out_scale;
        if (use_y && yo.v[c] == 0.f) gp = 0.f;
        atomicAdd(&cacc[v * VN + c], gp * h);
        atomicAdd(&cacc[channels + v * VN + c], gp);
        const float t = gp * w[v * VN + c];
        s1 += t;
        s2 += t * h;
      }
    }
    const float c1 = warp_sum(s1) / (float)channels;
    const float c2 = warp_sum(s2) / (float)channels;
    for (int v = lane; v < nvec; v += 32) {
      const int64_t o = row * channels + (int64_t)v * VN;
      const Vec<T> g = Vec<T>::load(dy + o);
      Vec<T> a = Vec<T>::load(x + o);
      Vec<T> yo;
      if (use_y) yo = Vec<T>::load(y + o);
#pragma unroll
      for (int c = 0; c < VN; ++c) {
        const float h = (a.v[c] - mu) * rs;
        float gp = g.v[c] * out_scale;
        if (use_y && yo.v[c] == 0.f) gp = 0.f;
        a.v[c] = rs * (gp * w[v * VN + c] - c1 - h * c2);
      }
      a.store(dx + o);
    }
  }
  __syncthreads();
  float* cp = colpart + (size_t)blockIdx.x * 2 * channels;
  for (int64_t c = threadIdx.x; c < 2 * channels; c += blockDim.x) cp[c] = cacc[c];
}

static int norm_grid(int64_t work_items, int per_block) {
  int64_t g = ceil_div(work_items, per_block);
  const int64_t cap = (int64_t)sm_count() * 4;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int gln_parts(int64_t n) {  // row strips of >= 16 rows, <= 8 per SM
  const int64_t g = ceil_div(n, 16), cap = (int64_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace egp

using namespace egp;

extern "C" {

static size_t gln_workspace_bytes(int64_t n, int64_t channels, int nseg) {
  const int parts = gln_parts(n);
  const int64_t gy = ceil_div(channels, (int64_t)kNormThreads * 4);
  const int g_stats = sm_count() * 4;
  size_t bytes = sizeof(GlnWorkspace);
  const size_t np = (size_t)(parts * gy > g_stats ? parts * gy : g_stats) * (size_t)nseg;
  bytes += sizeof(double) * 2 * np;
  bytes += sizeof(float) * 2 * (size_t)parts * (size_t)nseg * (size_t)channels;      // dweight/dbias partials
  bytes += sizeof(float) * (size_t)(sm_count() * 8) * (size_t)nseg * (size_t)channels;  // dx column-sum partials
  const size_t cs = colsum_workspace_bytes(n, channels);                              // fallback column sum of dx
  return bytes + cs + 128;
}

static int gln_segments(int64_t n, int nseg, const int64_t* seg_rows, GlnSegs* out, const char* who) {
  if (nseg < 1 || nseg > kMaxGlnSegs) {
    set_error("%s: %d segments (1..%d supported)", who, nseg, kMaxGlnSegs);
    return EGP_ERR_INVALID;
  }
  out->count = nseg;
  if (!seg_rows) {
    if (nseg != 1) { set_error("%s: seg_rows is required for more than one segment", who); return EGP_ERR_INVALID; }
    out->row[0] = 0;
    out->row[1] = n;
    return EGP_OK;
  }
  for (int i = 0; i <= nseg; ++i) out->row[i] = seg_rows[i];
  for (int i = 0; i < nseg; ++i)
    if (out->row[i + 1] < out->row[i]) { set_error("%s: seg_rows must be non-decreasing", who); return EGP_ERR_INVALID; }
  if (out->row[0] != 0 || out->row[nseg] != n) {
    set_error("%s: seg_rows must run from 0 to num_nodes", who);
    return EGP_ERR_INVALID;
  }
  return EGP_OK;
}

static int64_t gln_max_seg_rows(const GlnSegs& sg) {
  int64_t m = 0;
  for (int i = 0; i < sg.count; ++i) m = m > sg.row[i + 1] - sg.row[i] ? m : sg.row[i + 1] - sg.row[i];
  return m;
}

size_t egp_graph_layernorm_workspace(int64_t n, int64_t channels) { return gln_workspace_bytes(n, channels, 1); }
size_t egp_graph_layernorm_seg_workspace(int64_t n, int64_t channels, int nseg) {
  return gln_workspace_bytes(n, channels, nseg < 1 ? 1 : (nseg > kMaxGlnSegs ? kMaxGlnSegs : nseg));
}

int egp_graph_layernorm_seg_fwd(const void* x, const float* weight, const float* bias, void* y, double* stats,
                                int64_t n, int64_t channels, int nseg, const int64_t* seg_rows, float eps, int act,
                                float slope, int dtype, void* workspace, size_t ws_bytes, void* stream) {
  EGP_REQUIRE(x && weight && bias && y && stats && workspace, "graph_layernorm_fwd: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(x) && aligned16(y), "graph_layernorm_fwd: channels %% %d != 0 or unaligned", (int)vn);
  GlnSegs sg;
  int rc = gln_segments(n, nseg, seg_rows, &sg, "graph_layernorm_fwd");
  if (rc != EGP_OK) return rc;
  if (ws_bytes < gln_workspace_bytes(n, channels, nseg)) {
    set_error("graph_layernorm_fwd: workspace %zu < %zu", ws_bytes, gln_workspace_bytes(n, channels, nseg));
    return EGP_ERR_WORKSPACE;
  }
  if (n == 0) return EGP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  GlnWorkspace* ws = (GlnWorkspace*)workspace;
  double* partial = (double*)(ws + 1);
  EGP_CUDA(cudaMemsetAsync(ws->ticket, 0, sizeof(ws->ticket), s));
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = gln_max_seg_rows(sg) * channels / Vec<T>::N;     // grid sized for the largest segment
    int g1 = norm_grid(nvec, kNormThreads * 4);
    if (nseg > 1) {                                           // keep the chip-wide CTA count of a one-segment launch
      const int cap = sm_count() * 4 / nseg;
      g1 = g1 > cap ? (cap > 1 ? cap : 1) : g1;
    }
    (void)launch_kernel(gln_stats_kernel<T>, dim3(g1, nseg), kNormThreads, 0, s, (const T*)x, sg, channels, partial, ws->ticket, stats);
    EGP_LAUNCH_CHECK();
    const int g2 = norm_grid(nvec, kNormThreads * 2) * 2;
    const bool fixed = ((int64_t)g2 * kNormThreads) % (channels / Vec<T>::N) == 0;
    if (fixed) (void)launch_kernel(gln_apply_kernel<T, true>, dim3(g2, nseg), kNormThreads, 0, s, (const T*)x, weight, bias, (T*)y, stats, sg, channels, eps, act, slope);
    else (void)launch_kernel(gln_apply_kernel<T, false>, dim3(g2, nseg), kNormThreads, 0, s, (const T*)x, weight, bias, (T*)y, stats, sg, channels, eps, act, slope);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_graph_layernorm_seg_fwd_rowstats(const void* x, const float* weight, const float* bias, void* y, double* stats,
                                         int64_t n, int64_t channels, int nseg, const int64_t* seg_rows,
                                         const double* rowstats, float eps, int act, float slope, int dtype,
                                         void* stream) {
  EGP_REQUIRE(x && weight && bias && y && stats && rowstats, "graph_layernorm_fwd_rowstats: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && channels % 64 == 0 && aligned16(x) && aligned16(y),
              "graph_layernorm_fwd_rowstats: channels must be a multiple of 64 and the tensors 16-byte aligned");
  GlnSegs sg;
  int rc = gln_segments(n, nseg, seg_rows, &sg, "graph_layernorm_fwd_rowstats");
  if (rc != EGP_OK) return rc;
  for (int i = 1; i < nseg; ++i)
    EGP_REQUIRE(sg.row[i] % 128 == 0, "graph_layernorm_fwd_rowstats: inner segment boundaries must be multiples of 128 rows");
  if (n == 0) return EGP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int pairs = 4 * (int)ceil_div(channels, 256);   // (quarter, n tile) pairs per 128-row block: egp_gemm_rowstats layout
  (void)launch_kernel(gln_stats_from_rowstats_kernel, nseg, kRowstatThreads, 0, s, rowstats, pairs, sg, channels, stats);
  EGP_LAUNCH_CHECK();
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = gln_max_seg_rows(sg) * channels / Vec<T>::N;
    const int g2 = norm_grid(nvec, kNormThreads * 2) * 2;
    const bool fixed = ((int64_t)g2 * kNormThreads) % (channels / Vec<T>::N) == 0;
    if (fixed) (void)launch_kernel(gln_apply_kernel<T, true>, dim3(g2, nseg), kNormThreads, 0, s, (const T*)x, weight, bias, (T*)y, (const double*)stats, sg, channels, eps, act, slope);
    else (void)launch_kernel(gln_apply_kernel<T, false>, dim3(g2, nseg), kNormThreads, 0, s, (const T*)x, weight, bias, (T*)y, (const double*)stats, sg, channels, eps, act, slope);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_graph_layernorm_fwd(const void* x, const float* weight, const float* bias, void* y, double* stats,
                            int64_t n, int64_t channels, float eps, int act, float slope, int dtype,
                            void* workspace, size_t ws_bytes, void* stream) {
  return egp_graph_layernorm_seg_fwd(x, weight, bias, y, stats, n, channels, 1, nullptr, eps, act, slope, dtype, workspace,
                                     ws_bytes, stream);
}

int egp_graph_layernorm_seg_bwd(const void* dy, const void* x, const float* weight, const float* bias,
                                const double* stats, void* dx, float* dweight, float* dbias, float* dx_colsum, int64_t n,
                                int64_t channels, int nseg, const int64_t* seg_rows, float eps, int act, float slope,
                                int dtype, void* workspace, size_t ws_bytes, void* stream) {
  EGP_REQUIRE(dy && x && weight && bias && stats && dx && workspace, "graph_layernorm_bwd: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(x) && aligned16(dy) && aligned16(dx),
              "graph_layernorm_bwd: channels %% %d != 0 or unaligned", (int)vn);
  GlnSegs sg;
  int rcs = gln_segments(n, nseg, seg_rows, &sg, "graph_layernorm_bwd");
  if (rcs != EGP_OK) return rcs;
  if (ws_bytes < gln_workspace_bytes(n, channels, nseg)) {
    set_error("graph_layernorm_bwd: workspace too small");
    return EGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    if (dx_colsum) EGP_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * channels, s));
    if (dweight) EGP_CUDA(cudaMemsetAsync(dweight, 0, sizeof(float) * channels, s));
    if (dbias) EGP_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * channels, s));
    return EGP_OK;
  }
  GlnWorkspace* ws = (GlnWorkspace*)workspace;
  const int64_t max_rows = gln_max_seg_rows(sg);
  int parts = gln_parts(max_rows);
  if (nseg > 1) {                                             // keep the chip-wide CTA count of the one-segment launch
    const int cap = gln_parts(n) / nseg > 1 ? gln_parts(n) / nseg : 1;
    if (parts > cap) parts = cap;
  }
  const int rows_per = (int)ceil_div(max_rows, parts);
  EGP_DISPATCH_DTYPE(dtype, T, {
    constexpr int VN = Vec<T>::N;
    const int64_t nvec_row = channels / VN;
    const int rthreads = (int)(nvec_row >= kNormThreads ? kNormThreads : (nvec_row + 31) / 32 * 32);  // no idle lanes
    const int gy = (int)ceil_div(nvec_row, (int64_t)rthreads);
    const int gy_ws = (int)ceil_div(channels, (int64_t)kNormThreads * 4);
    const int g_stats = sm_count() * 4;
    const int parts_ws = gln_parts(n);
    const size_t np = (size_t)(parts_ws * gy_ws > g_stats ? parts_ws * gy_ws : g_stats) * (size_t)nseg;
    double* scalpart = (double*)(ws + 1);
    float* colpart = (float*)(scalpart + 2 * np);
    float* dxpart = colpart + 2 * (size_t)parts_ws * nseg * channels;
    void* cs_ws = dxpart + (size_t)(sm_count() * 8) * nseg * channels;
    EGP_REQUIRE((size_t)parts * gy * nseg <= np && parts <= parts_ws, "graph_layernorm_bwd: internal partial sizing");
    (void)launch_kernel(gln_bwd_reduce_kernel<T>, dim3(parts, gy, nseg), rthreads, 0, s,
        (const T*)dy, (const T*)x, weight, bias, stats, sg, channels, rows_per, eps, act, slope, colpart, scalpart);
    EGP_LAUNCH_CHECK();
    const int colblocks = (int)ceil_div(channels, kFinCols);
    (void)launch_kernel(col_finalize_kernel, colblocks + nseg, kFinThreads, 0, s, colpart, parts * nseg, channels, 2, dweight, dbias, nullptr,
                                                              scalpart, parts * gy, ws->scal, colblocks);
    EGP_LAUNCH_CHECK();
    const int64_t nvec = max_rows * channels / VN;
    const int g2 = norm_grid(nvec, kNormThreads * 2) * 2;     // <= 8 per SM: dxpart holds that many rows per segment
    const bool fixed = ((int64_t)g2 * kNormThreads) % nvec_row == 0;
    const bool fuse = dx_colsum && fixed && nvec_row <= kNormThreads && kNormThreads % nvec_row == 0;
    if (fixed)
      (void)launch_kernel(gln_bwd_apply_kernel<T, true>, dim3(g2, nseg), kNormThreads, 0, s, (const T*)dy, (const T*)x, weight, bias, stats, ws->scal,
                                                                (T*)dx, sg, channels, eps, act, slope,
                                                                fuse ? dxpart : nullptr);
    else
      (void)launch_kernel(gln_bwd_apply_kernel<T, false>, dim3(g2, nseg), kNormThreads, 0, s, (const T*)dy, (const T*)x, weight, bias, stats, ws->scal,
                                                                 (T*)dx, sg, channels, eps, act, slope, nullptr);
    EGP_LAUNCH_CHECK();
    if (fuse) {
      (void)launch_kernel(col_finalize_kernel, colblocks, kFinThreads, 0, s, dxpart, g2 * nseg, channels, 1, dx_colsum, nullptr, nullptr, nullptr, 0,
                                                            nullptr, colblocks);
      EGP_LAUNCH_CHECK();
    } else if (dx_colsum) {
      const int rc = colsum_launch(dx, dx_colsum, n, channels, channels, dtype, cs_ws, colsum_workspace_bytes(n, channels), s);
      if (rc != EGP_OK) return rc;
    }
  });
  return EGP_OK;
}

int egp_graph_layernorm_bwd(const void* dy, const void* x, const float* weight, const float* bias,
                            const double* stats, void* dx, float* dweight, float* dbias, float* dx_colsum, int64_t n,
                            int64_t channels, float eps, int act, float slope, int dtype, void* workspace,
                            size_t ws_bytes, void* stream) {
  return egp_graph_layernorm_seg_bwd(dy, x, weight, bias, stats, dx, dweight, dbias, dx_colsum, n, channels, 1, nullptr, eps,
                                     act, slope, dtype, workspace, ws_bytes, stream);
}

size_t egp_row_layernorm_workspace(int64_t n, int64_t channels) {
  return sizeof(float) * 3 * (size_t)(sm_count() * 8) * (size_t)channels + colsum_workspace_bytes(n, channels) + 128;
}

#define EGP_RLN_DISPATCH_NVL(nvec, ...)                                               \
  do {                                                                                \
    const int _per = (int)(((nvec) + 31) / 32);                                       \
    if (_per <= 1) { constexpr int NVL = 1; __VA_ARGS__ }                             \
    else if (_per <= 2) { constexpr int NVL = 2; __VA_ARGS__ }                        \
    else if (_per <= 4) { constexpr int NVL = 4; __VA_ARGS__ }                        \
    else if (_per <= 8) { constexpr int NVL = 8; __VA_ARGS__ }                        \
    else { constexpr int NVL = 0; __VA_ARGS__ }                                       \
  } while (0)

int egp_row_layernorm_fwd(const void* x, const float* weight, const float* bias, void* y, float* mean,
                          float* rstd, int64_t n, int64_t channels, float eps, int act, float dropout_p,
                          uint64_t seed, uint64_t offset, const uint64_t* rng_state, int dtype, void* stream) {
  EGP_REQUIRE(x && weight && bias && y && mean && rstd, "row_layernorm_fwd: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(x) && aligned16(y), "row_layernorm_fwd: channels %% %d != 0 or unaligned", (int)vn);
  EGP_REQUIRE(act == EGP_ACT_NONE || act == EGP_ACT_RELU, "row_layernorm_fwd: act must be none or relu");
  EGP_REQUIRE(dropout_p >= 0.f && dropout_p < 1.f, "row_layernorm_fwd: dropout_p must be in [0,1)");
  if (n == 0) return EGP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = norm_grid(n, kNormThreads / 32);
  uint32_t thr = (uint32_t)(dropout_p * 65536.0f + 0.5f);
  if (dropout_p > 0.f && thr == 0) thr = 1;
  const float keep_scale = dropout_p > 0.f ? 1.0f / (1.0f - dropout_p) : 1.0f;
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = channels / Vec<T>::N;
    EGP_RLN_DISPATCH_NVL(nvec, {
      if constexpr (NVL == 0)
        (void)launch_kernel(rln_fwd_wide_kernel<T>, grid, kNormThreads, 0, s, (const T*)x, weight, bias, (T*)y, mean, rstd, n, channels, eps,
                                                             act, thr, keep_scale, seed, offset, rng_state);
      else
      {
        static int resident = 0;   // CTAs per SM of this instantiation (rows are strided over whatever grid runs)
        if (!resident) {
          int guarded = 0, full_rows = 0;   // one grid size for both kernels: the smaller residency
          EGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&guarded, rln_fwd_kernel<T, (NVL ? NVL : 1)>, kNormThreads, 0));
          EGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&full_rows, rln_fwd_full_kernel<T, (NVL ? NVL : 1), true, true>, kNormThreads, 0));
          resident = guarded < full_rows ? guarded : full_rows;
          if (resident < 1) resident = 1;
        }
        const int64_t want = ceil_div(n, kNormThreads / 32), cap = (int64_t)sm_count() * resident;
        const int g = (int)(want < cap ? want : cap);
        constexpr int NV = NVL ? NVL : 1;
        auto full = [&](auto relu, auto drop) {
          (void)launch_kernel(rln_fwd_full_kernel<T, NV, decltype(relu)::value, decltype(drop)::value>, g, kNormThreads, 0, s,
                              (const T*)x, weight, bias, (T*)y, mean, rstd, n, eps, thr, keep_scale, seed, offset, rng_state);
        };
        using Yes = std::true_type;
        using No = std::false_type;
        if (nvec != 32 * NV)   // ragged rows: guarded kernel
          (void)launch_kernel(rln_fwd_kernel<T, NV>, g, kNormThreads, 0, s, (const T*)x, weight, bias, (T*)y, mean, rstd, n,
                              channels, eps, act, thr, keep_scale, seed, offset, rng_state);
        else if (act == EGP_ACT_RELU && thr) full(Yes{}, Yes{});
        else if (act == EGP_ACT_RELU) full(Yes{}, No{});
        else if (thr) full(No{}, Yes{});
        else full(No{}, No{});
      }
    });
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

int egp_row_layernorm_bwd(const void* dy, const void* x, const void* y, const float* weight, const float* mean,
                          const float* rstd, void* dx, float* dweight, float* dbias, float* dx_colsum, int64_t n,
                          int64_t channels, int act, float out_scale, int dtype, void* workspace, size_t ws_bytes,
                          void* stream) {
  EGP_REQUIRE(dy && x && weight && mean && rstd && dx && workspace, "row_layernorm_bwd: null pointer");
  EGP_REQUIRE(act == EGP_ACT_NONE || act == EGP_ACT_RELU, "row_layernorm_bwd: act must be none or relu");
  EGP_REQUIRE(y || (act == EGP_ACT_NONE && out_scale == 1.f), "row_layernorm_bwd: relu/dropout need the saved output");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && aligned16(x) && aligned16(dy) && aligned16(dx), "row_layernorm_bwd: channels/alignment");
  if (ws_bytes < egp_row_layernorm_workspace(n, channels)) {
    set_error("row_layernorm_bwd: workspace too small");
    return EGP_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    if (dweight) EGP_CUDA(cudaMemsetAsync(dweight, 0, sizeof(float) * channels, s));
    if (dbias) EGP_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * channels, s));
    if (dx_colsum) EGP_CUDA(cudaMemsetAsync(dx_colsum, 0, sizeof(float) * channels, s));
    return EGP_OK;
  }
  int grid = norm_grid(n, (kNormThreads / 32) * 16);   // warp-per-row kernels: >= 16 rows per warp
  const size_t smem = sizeof(float) * 2 * channels;
  float* colpart = (float*)workspace;
  void* cs_ws = colpart + 3 * (size_t)(sm_count() * 8) * channels;
  int nout = 2;
  EGP_DISPATCH_DTYPE(dtype, T, {
    const int64_t nvec = channels / Vec<T>::N;
    if (nvec > 32 && nvec <= 1024) {                     // one CTA per row, thread-owned columns
      const int threads = (int)((nvec + 31) / 32 * 32);
      // one wave of resident CTAs (rows are strided over the grid); the partial workspace holds 8 per SM
      static int resident[2] = {0, 0};
      int& res = resident[threads <= 256 ? 0 : 1];
      if (!res) {
        if (threads <= 256) EGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, rln_bwd_block_kernel<T, 256>, 128, 0));
        else EGP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, rln_bwd_block_kernel<T, 1024>, 1024, 0));
        res = res < 1 ? 1 : (res > 8 ? 8 : res);
      }
      int per_sm = res;
      if (threads > 128 && threads <= 256) per_sm = res / 2 > 0 ? res / 2 : 1;
      const int64_t cap = (int64_t)sm_count() * per_sm;
      grid = (int)(n < cap ? n : cap);
      nout = dx_colsum ? 3 : 2;
      if (threads <= 256)
        (void)launch_kernel(rln_bwd_block_kernel<T, 256>, grid, threads, 0, s, (const T*)dy, (const T*)x, (const T*)y, weight, mean, rstd,
                                                              (T*)dx, n, channels, act, out_scale, nout, colpart);
      else
        (void)launch_kernel(rln_bwd_block_kernel<T, 1024>, grid, threads, 0, s, (const T*)dy, (const T*)x, (const T*)y, weight, mean, rstd,
                                                               (T*)dx, n, channels, act, out_scale, nout, colpart);
    } else {
      EGP_RLN_DISPATCH_NVL(nvec, {
        auto kern = rln_bwd_wide_kernel<T>;
        if constexpr (NVL != 0) kern = rln_bwd_kernel<T, NVL>;
        if (smem > 48 * 1024) EGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        (void)launch_kernel(kern, grid, kNormThreads, smem, s, (const T*)dy, (const T*)x, (const T*)y, weight, mean, rstd, (T*)dx, n,
                                              channels, act, out_scale, colpart);
      });
    }
    EGP_LAUNCH_CHECK();
  });
  const int colblocks = (int)ceil_div(channels, kFinCols);
  (void)launch_kernel(col_finalize_kernel, colblocks, kFinThreads, 0, s, colpart, grid, channels, nout, dweight, dbias,
                                                        nout == 3 ? dx_colsum : nullptr, nullptr, 0, nullptr, colblocks);
  EGP_LAUNCH_CHECK();
  if (dx_colsum && nout != 3) {
    const int rc = colsum_launch(dx, dx_colsum, n, channels, channels, dtype, cs_ws, colsum_workspace_bytes(n, channels), s);
    if (rc != EGP_OK) return rc;
  }
  return EGP_OK;
}

}  // extern "C"
