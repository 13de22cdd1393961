// egp_gemm: the single Linear-layer entry point.  bf16 inputs -> tcgen05/TMEM/TMA kernel (gemm_tcgen05.cu);
// fp32 inputs (parity mode) and bf16 operands that TMA cannot address (row stride not a multiple of 16 bytes,
// e.g. a 2-class head) -> FFMA kernel (gemm_simt.cu).  Both are CUDA; there is no host fallback.
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
int64_t tc_gemm_rowstats_slots(int64_t N);
void tc_set_deterministic(int on);
int tc_get_deterministic();
size_t tc_gemm_workspace_bytes();
bool tc_gemm_supported(const void* A, int64_t lda, const void* B, int64_t ldb, const void* A2, int64_t lda2,
                       const void* B2, int64_t ldb2);
}  // namespace egp

using namespace egp;

extern "C" {

size_t egp_gemm_workspace(int64_t M, int64_t N, int64_t K) {
  (void)M; (void)N; (void)K;
  // default: split-K reduce-adds straight into C, no workspace.  Deterministic mode: slabs for an ordered reduction.
  return tc_gemm_workspace_bytes();
}

int egp_set_deterministic(int on) {
  tc_set_deterministic(on);
  return EGP_OK;
}

int egp_get_deterministic(void) { return tc_get_deterministic(); }

int egp_gemm(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, const void* A2,
             int64_t lda2, const void* B2, int64_t ldb2, int64_t K2, const float* bias, const void* residual,
             int64_t ldr, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope, int in_dtype,
             int out_dtype, int accumulate, void* workspace, size_t ws_bytes, void* stream) {
  EGP_REQUIRE(A && B && C, "gemm: null operand");
  EGP_REQUIRE(M >= 0 && N >= 0 && K >= 0 && K2 >= 0, "gemm: negative size");
  EGP_REQUIRE((A2 == nullptr) == (B2 == nullptr), "gemm: A2 and B2 must be given together");
  EGP_REQUIRE(in_dtype == EGP_F32 || in_dtype == EGP_BF16, "gemm: bad in_dtype %d", in_dtype);
  EGP_REQUIRE(out_dtype == EGP_F32 || out_dtype == EGP_BF16, "gemm: bad out_dtype %d", out_dtype);
  EGP_REQUIRE(act >= EGP_ACT_NONE && act <= EGP_ACT_LEAKY_RELU, "gemm: bad activation %d", act);
  EGP_REQUIRE(!accumulate || out_dtype == EGP_F32, "gemm: accumulate needs an fp32 output");
  cudaStream_t s = (cudaStream_t)stream;
  if (!A2) { K2 = 0; lda2 = 0; ldb2 = 0; }
  // EGP_FORCE_SIMT=1 routes bf16 GEMMs to the FFMA kernel (debugging aid for the tensor-core path)
  static const bool force_simt = [] { const char* e = getenv("EGP_FORCE_SIMT"); return e && e[0] == '1'; }();
  if (!force_simt && in_dtype == EGP_BF16 && tc_gemm_supported(A, lda, B, ldb, A2, lda2, B2, ldb2) && K > 0)
    return tc_gemm_launch(A, lda, a_trans, B, ldb, b_trans, A2, lda2, B2, ldb2, K2, bias, residual, ldr, C, ldc, M, N,
                          K, act, slope, out_dtype, accumulate, s, nullptr, workspace, ws_bytes);
  return sgemm_launch(A, lda, a_trans, B, ldb, b_trans, A2, lda2, B2, ldb2, K2, bias, residual, ldr, C, ldc, M, N, K,
                      act, slope, in_dtype, out_dtype, accumulate, s);
}

size_t egp_gemm_rowstats_bytes(int64_t M, int64_t N) {
  return sizeof(double) * 2 * (size_t)ceil_div(M, 128) * 4 * (size_t)tc_gemm_rowstats_slots(N);
}

int egp_gemm_rowstats(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, const void* A2,
                      int64_t lda2, const void* B2, int64_t ldb2, int64_t K2, const float* bias, const void* residual,
                      int64_t ldr, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope,
                      int out_dtype, double* rowstats, void* stream) {
  EGP_REQUIRE(A && B && C && rowstats, "gemm_rowstats: null pointer");
  EGP_REQUIRE(M >= 0 && N >= 0 && K > 0 && K2 >= 0, "gemm_rowstats: bad size");
  EGP_REQUIRE((A2 == nullptr) == (B2 == nullptr), "gemm_rowstats: A2 and B2 must be given together");
  EGP_REQUIRE(out_dtype == EGP_F32 || out_dtype == EGP_BF16, "gemm_rowstats: bad out_dtype %d", out_dtype);
  EGP_REQUIRE(act >= EGP_ACT_NONE && act <= EGP_ACT_LEAKY_RELU, "gemm_rowstats: bad activation %d", act);
  if (!A2) { K2 = 0; lda2 = 0; ldb2 = 0; }
  if (!tc_gemm_supported(A, lda, B, ldb, A2, lda2, B2, ldb2)) {
    set_error("gemm_rowstats: operands must be 16-byte aligned bf16 with 16-byte row pitches (tensor-core path only)");
    return EGP_ERR_UNSUPPORTED;
  }
  return tc_gemm_launch(A, lda, a_trans, B, ldb, b_trans, A2, lda2, B2, ldb2, K2, bias, residual, ldr, C, ldc, M, N, K, act,
                        slope, out_dtype, 0, (cudaStream_t)stream, rowstats, nullptr, 0);
}

}  // extern "C"
