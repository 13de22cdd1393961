// fp32 FFMA GEMM -- the fp32 "parity mode" of every Linear on the path (1e-4 relative to the fp32 oracle).
// The bf16 product path is gemm_tcgen05.cu; this kernel exists because TF32/bf16 tensor-core rounding cannot
// meet the fp32 tolerance.  128x128x16 tiles, 256 threads, 8x8 register micro-tiles, operands staged in shared
// memory k-major so the inner product reads are 128-bit and conflict-free.
//
//   C[M,N] = act( sum_k A(m,k) B(n,k) + sum_k A2(m,k) B2(n,k) + bias[n] ) + residual[m,n]   (+ C if accumulate)
#include "common.cuh"

namespace egp {

constexpr int SBM = 128, SBN = 128, SBK = 16, STHREADS = 256;

// load a [ROWS x SBK] operand tile into smem[k][row] (k-major), zero-filling out-of-range elements
template <typename InT>
__device__ __forceinline__ void load8(const InT* p, float* v);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* v) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* v) {
  const Vec<__nv_bfloat16> a = Vec<__nv_bfloat16>::load(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = a.v[i];
}

template <bool TRANS, int ROWS, typename InT>
__device__ __forceinline__ void load_tile(const InT* __restrict__ p, int64_t ld, int64_t row0, int64_t k0,
                                          int64_t rows, int64_t kdim, bool vec_ok, float (*sm)[ROWS + 4]) {
  const int t = threadIdx.x;
  if (!TRANS) {  // element (row, k) at p[row*ld + k]: thread reads 8 consecutive k of one row
    const int r = t >> 1, kk = (t & 1) * 8;
    const int64_t gr = row0 + r, gk = k0 + kk;
    float v[8];
    if (vec_ok && gr < rows && gk + 8 <= kdim) {
      load8<InT>(p + gr * ld + gk, v);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (gr < rows && gk + i < kdim) ? to_float<InT>(p[gr * ld + gk + i]) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[kk + i][r] = v[i];
  } else {  // element (row, k) at p[k*ld + row]: thread reads 8 consecutive rows of one k
    const int kk = t >> 4, r = (t & 15) * 8;
    const int64_t gk = k0 + kk, gr = row0 + r;
    float v[8];
    if (vec_ok && gk < kdim && gr + 8 <= rows) {
      load8<InT>(p + gk * ld + gr, v);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (gk < kdim && gr + i < rows) ? to_float<InT>(p[gk * ld + gr + i]) : 0.f;
    }
    *reinterpret_cast<float4*>(&sm[kk][r]) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(&sm[kk][r + 4]) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <bool AT, bool BT, typename InT, typename OutT>
__global__ void __launch_bounds__(STHREADS)
sgemm_kernel(const InT* __restrict__ A, int64_t lda, const InT* __restrict__ B, int64_t ldb,
             const InT* __restrict__ A2, int64_t lda2, const InT* __restrict__ B2, int64_t ldb2, int64_t K2,
             const float* __restrict__ bias, const OutT* __restrict__ residual, int64_t ldr, OutT* __restrict__ C,
             int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope, int accumulate, int vec_a,
             int vec_b, int vec_a2, int vec_b2) {
  pdl_enter();
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Bs[SBK][SBN + 4];
  const int64_t m0 = (int64_t)blockIdx.y * SBM, n0 = (int64_t)blockIdx.x * SBN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int pass = 0; pass < 2; ++pass) {
    const InT* Ap = pass == 0 ? A : A2;
    const InT* Bp = pass == 0 ? B : B2;
    if (!Ap || !Bp) continue;
    const int64_t la = pass == 0 ? lda : lda2, lb = pass == 0 ? ldb : ldb2, kd = pass == 0 ? K : K2;
    const bool va = pass == 0 ? vec_a : vec_a2, vb = pass == 0 ? vec_b : vec_b2;
    for (int64_t k0 = 0; k0 < kd; k0 += SBK) {
      load_tile<AT, SBM, InT>(Ap, la, m0, k0, M, kd, va, As);
      load_tile<BT, SBN, InT>(Bp, lb, n0, k0, N, kd, vb, Bs);
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SBK; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 8 + 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + tx * 8 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      v = apply_act(v, act, slope);
      if (residual) v += to_float<OutT>(residual[m * ldr + n]);
      if (accumulate) v += to_float<OutT>(C[m * ldc + n]);
      C[m * ldc + n] = from_float<OutT>(v);
    }
  }
}

static inline int vec_ok(const void* p, int64_t ld, int n) { return p && aligned16(p) && (ld % n == 0); }

template <typename InT>
static int sgemm_launch_t(const InT* A, int64_t lda, int a_trans, const InT* B, int64_t ldb, int b_trans, const InT* A2,
                          int64_t lda2, const InT* B2, int64_t ldb2, int64_t K2, const float* bias,
                          const void* residual, int64_t ldr, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                          int act, float slope, int out_dtype, int accumulate, cudaStream_t stream) {
  if (M == 0 || N == 0) return EGP_OK;
  dim3 grid((unsigned)ceil_div(N, SBN), (unsigned)ceil_div(M, SBM));
  constexpr int VE = 16 / (int)sizeof(InT);  // elements per 16-byte load
  const int va = vec_ok(A, lda, VE), vb = vec_ok(B, ldb, VE), va2 = vec_ok(A2, lda2, VE), vb2 = vec_ok(B2, ldb2, VE);
#define EGP_SGEMM(AT, BT, OT)                                                                                       \
  (void)launch_kernel(sgemm_kernel<AT, BT, InT, OT>, grid, STHREADS, 0, stream, A, lda, B, ldb, A2, lda2, B2, ldb2, K2, bias,       \
                                                               (const OT*)residual, ldr, (OT*)C, ldc, M, N, K, act, \
                                                               slope, accumulate, va, vb, va2, vb2)
#define EGP_SGEMM_T(OT)                                       \
  do {                                                        \
    if (!a_trans && !b_trans) EGP_SGEMM(false, false, OT);    \
    else if (!a_trans && b_trans) EGP_SGEMM(false, true, OT); \
    else if (a_trans && !b_trans) EGP_SGEMM(true, false, OT); \
    else EGP_SGEMM(true, true, OT);                           \
  } while (0)
  if (out_dtype == EGP_F32) EGP_SGEMM_T(float);
  else EGP_SGEMM_T(__nv_bfloat16);
#undef EGP_SGEMM_T
#undef EGP_SGEMM
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int sgemm_launch(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, const void* A2,
                 int64_t lda2, const void* B2, int64_t ldb2, int64_t K2, const float* bias, const void* residual,
                 int64_t ldr, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope,
                 int in_dtype, int out_dtype, int accumulate, cudaStream_t stream) {
  if (in_dtype == EGP_F32)
    return sgemm_launch_t<float>((const float*)A, lda, a_trans, (const float*)B, ldb, b_trans, (const float*)A2, lda2,
                                 (const float*)B2, ldb2, K2, bias, residual, ldr, C, ldc, M, N, K, act, slope,
                                 out_dtype, accumulate, stream);
  return sgemm_launch_t<__nv_bfloat16>((const __nv_bfloat16*)A, lda, a_trans, (const __nv_bfloat16*)B, ldb, b_trans,
                                       (const __nv_bfloat16*)A2, lda2, (const __nv_bfloat16*)B2, ldb2, K2, bias,
                                       residual, ldr, C, ldc, M, N, K, act, slope, out_dtype, accumulate, stream);
}

}  // namespace egp
