// Shared helpers for the egopack_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/egopack_b200.h"

namespace egp {

void set_error(const char* fmt, ...);

#define EGP_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      egp::set_error(__VA_ARGS__);             \
      return EGP_ERR_INVALID;                  \
    }                                          \
  } while (0)

#define EGP_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      egp::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return EGP_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

// launch-configuration errors only; never synchronises
#define EGP_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    cudaError_t _e = cudaPeekAtLastError();                                                  \
    if (_e != cudaSuccess) {                                                                 \
      egp::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      (void)cudaGetLastError();                                                              \
      return EGP_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

constexpr int kMaxDevices = 64;
int current_device();  // cudaGetDevice(), clamped to [0, kMaxDevices): index for per-device one-time state
int sm_count();
bool pdl_enabled();   // EGP_PDL=0 turns programmatic dependent launch off (abi.cu)

// ------------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the library starts with pdl_enter(): it lets the NEXT kernel in the
// stream be scheduled early (its CTAs become resident and run up to their own pdl_enter as SMs free up) and then waits
// until the PREVIOUS kernel has completed and flushed its memory -- so the data dependencies are exactly those of
// plain stream order, but the 1-3 us between the end of one kernel and the first CTA of the next are hidden.
// Kernels launched without the attribute (or after a non-kernel stream operation) see both instructions as no-ops.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                        Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------------
// 16-byte vectors of the storage type, computed on as fp32
// ------------------------------------------------------------------------------------------------------
template <typename T>
struct Vec;  // 16 bytes

template <>
struct Vec<float> {
  static constexpr int N = 4;
  float v[4];
  __device__ __forceinline__ static Vec load(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    Vec r;
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
    return r;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  float v[8];
  __device__ __forceinline__ static Vec load(const __nv_bfloat16* p) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    Vec r;
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r.v[2 * i] = __uint_as_float(w[i] << 16);
      r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// a 16-byte vector kept PACKED in registers (4 regs) and unpacked only where it is consumed
template <typename T>
struct Raw {
  uint4 u;
  __device__ __forceinline__ static Raw load(const T* p) { Raw r; r.u = *reinterpret_cast<const uint4*>(p); return r; }
  __device__ __forceinline__ static Raw zero() { Raw r; r.u = make_uint4(0u, 0u, 0u, 0u); return r; }
  __device__ __forceinline__ Vec<T> unpack() const;
};
template <>
__device__ __forceinline__ Vec<float> Raw<float>::unpack() const {
  Vec<float> r;
  r.v[0] = __uint_as_float(u.x); r.v[1] = __uint_as_float(u.y); r.v[2] = __uint_as_float(u.z); r.v[3] = __uint_as_float(u.w);
  return r;
}
template <>
__device__ __forceinline__ Vec<__nv_bfloat16> Raw<__nv_bfloat16>::unpack() const {
  Vec<__nv_bfloat16> r;
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}

// A 16-byte vector as fp32 PAIRS, for the packed fp32x2 instructions of sm_100 (FADD2 / FMUL2 / FFMA2: one issue slot
// per two lanes of arithmetic -- what the instruction-issue-bound streaming kernels need).
template <typename T>
struct Pairs;
template <>
struct Pairs<float> {
  static constexpr int NP = 2;
  float2 p[2];
  __device__ __forceinline__ static Pairs from(const Raw<float>& r) {
    Pairs q;
    q.p[0] = make_float2(__uint_as_float(r.u.x), __uint_as_float(r.u.y));
    q.p[1] = make_float2(__uint_as_float(r.u.z), __uint_as_float(r.u.w));
    return q;
  }
  __device__ __forceinline__ uint4 pack() const {
    return make_uint4(__float_as_uint(p[0].x), __float_as_uint(p[0].y), __float_as_uint(p[1].x), __float_as_uint(p[1].y));
  }
};
template <>
struct Pairs<__nv_bfloat16> {
  static constexpr int NP = 4;
  float2 p[4];
  __device__ __forceinline__ static Pairs from(const Raw<__nv_bfloat16>& r) {
    Pairs q;
    const uint32_t w[4] = {r.u.x, r.u.y, r.u.z, r.u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) q.p[i] = make_float2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xffff0000u));
    return q;
  }
  __device__ __forceinline__ uint4 pack() const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(p[i].x, p[i].y);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float2 splat2(float s) { return make_float2(s, s); }

template <typename T>
__device__ __forceinline__ float to_float(T x);
template <>
__device__ __forceinline__ float to_float<float>(float x) { return x; }
template <>
__device__ __forceinline__ float to_float<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T>
__device__ __forceinline__ T from_float(float x);
template <>
__device__ __forceinline__ float from_float<float>(float x) { return x; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  if (act == EGP_ACT_RELU) return x > 0.f ? x : 0.f;
  if (act == EGP_ACT_LEAKY_RELU) return x > 0.f ? x : x * slope;
  return x;
}

// ------------------------------------------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in every thread; `smem` holds >= 32 T; blockDim.x multiple of 32
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = (lane < nw) ? smem[lane] : T(0);
  r = warp_sum(r);
  return r;
}

// ------------------------------------------------------------------------------------------------------
// counter-based RNG for fused dropout: Philox4x32-10 keyed by (seed), counted by (element-vector index, offset)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// keep-decisions for the (up to 8) elements of the 16-byte vector with global index `vec`: bit c = keep element c.
// Drop probability is quantised to 1/65536.
//
// Counter-based and stateless like Philox (a pure function of (vec, seed, offset), so forward and any recomputation
// agree and CUDA-graph replays only need a new `offset`), but an order of magnitude cheaper: the 64-bit seed / offset
// are folded into one 32-bit key with Philox-grade mixing ONCE per thread (dropout_key), and each vector then draws its
// 8 x 16 random bits from four rounds of a 32-bit multiply-xorshift hash (the PCG-RXS-M-XS output permutation applied
// to key + 4*vec + i).  With Philox4x32-10 per vector the fused LayerNorm+ReLU+Dropout kernel was issue-bound
// (ncu: 68 % issue slots, 0.50 of HBM peak against 0.69 without dropout); a dropout mask needs independence between
// elements, not cryptographic strength.
__device__ __forceinline__ uint32_t pcg_hash(uint32_t v) {
  const uint32_t state = v * 747796405u + 2891336453u;
  const uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
  return (word >> 22u) ^ word;
}

__device__ __forceinline__ uint32_t dropout_key(uint64_t seed, uint64_t offset) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)offset, (uint32_t)(offset >> 32), 0x9E3779B9u, 0xBB67AE85u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  return r.x ^ r.z;
}

__device__ __forceinline__ uint32_t dropout_keep_bits_keyed(uint64_t vec, uint32_t key, uint32_t thr16) {
  // vectors beyond 2^30 fold their high bits into the key (a tensor that large is > 8.6e9 elements)
  const uint32_t k = key ^ pcg_hash((uint32_t)(vec >> 30));
  const uint32_t base = k + ((uint32_t)vec << 2);
  // p = 0.5 (the configured dropout, experiments/mtl.yaml): one fair bit per element, so ONE hash serves the vector
  if (thr16 == 0x8000u) return pcg_hash(base) >> 24;
  uint32_t bits = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t w = pcg_hash(base + (uint32_t)i);
    bits |= ((w & 0xffffu) >= thr16 ? 1u : 0u) << (2 * i);
    bits |= ((w >> 16) >= thr16 ? 1u : 0u) << (2 * i + 1);
  }
  return bits;
}

// Zero the dropped elements of a PACKED 16-byte output vector: the keep bits are spread so that each lands in the top
// bit of its own byte (one multiply per four bits), and PRMT's sign-replicate mode turns those into byte masks --
// 1.5 integer instructions per bf16 element instead of a bit test + select each.
__device__ __forceinline__ uint32_t prmt_sx(uint32_t a, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
  return d;
}
template <typename T>
__device__ __forceinline__ void dropout_mask_packed(uint4& v, uint32_t keep);
template <>
__device__ __forceinline__ void dropout_mask_packed<__nv_bfloat16>(uint4& v, uint32_t keep) {   // 8 bits, 2 per word
  const uint32_t r0 = (keep & 0xFu) * 0x10204080u, r1 = ((keep >> 4) & 0xFu) * 0x10204080u;
  v.x &= prmt_sx(r0, 0x9988u);
  v.y &= prmt_sx(r0, 0xBBAAu);
  v.z &= prmt_sx(r1, 0x9988u);
  v.w &= prmt_sx(r1, 0xBBAAu);
}
template <>
__device__ __forceinline__ void dropout_mask_packed<float>(uint4& v, uint32_t keep) {            // 4 bits, 1 per word
  const uint32_t r0 = (keep & 0xFu) * 0x10204080u;
  v.x &= prmt_sx(r0, 0x8888u);
  v.y &= prmt_sx(r0, 0x9999u);
  v.z &= prmt_sx(r0, 0xAAAAu);
  v.w &= prmt_sx(r0, 0xBBBBu);
}

__device__ __forceinline__ uint32_t dropout_keep_bits(uint64_t vec, uint64_t seed, uint64_t offset, uint32_t thr16) {
  return dropout_keep_bits_keyed(vec, dropout_key(seed, offset), thr16);
}

// Per-CTA column partial for flat grid-stride kernels whose threads keep a FIXED 16-byte column (the grid stride is
// a multiple of the row length).  `period` = vectors per row; requires period <= blockDim.x and blockDim.x % period
// == 0.  Threads sharing a column are combined through shared memory and row `blockIdx.x` of the [gridDim.x, C]
// partial matrix is written; a col-finalize kernel then reduces the partial rows in a fixed order (deterministic).
template <int VN>
__device__ __forceinline__ void block_column_partial(const float (&dsum)[VN], int period, float* __restrict__ part_row,
                                                     float* smem /* blockDim.x * VN floats */) {
#pragma unroll
  for (int c = 0; c < VN; ++c) smem[threadIdx.x * VN + c] = dsum[c];
  __syncthreads();
  if ((int)threadIdx.x < period) {
#pragma unroll
    for (int c = 0; c < VN; ++c) {
      float t = 0.f;
      for (int j = threadIdx.x; j < (int)blockDim.x; j += period) t += smem[j * VN + c];
      part_row[threadIdx.x * VN + c] = t;
    }
  }
}

// dispatch on the dtype code
#define EGP_DISPATCH_DTYPE(dtype, T, ...)                         \
  do {                                                            \
    if ((dtype) == EGP_F32) {                                     \
      using T = float;                                            \
      __VA_ARGS__                                                 \
    } else if ((dtype) == EGP_BF16) {                             \
      using T = __nv_bfloat16;                                    \
      __VA_ARGS__                                                 \
    } else {                                                      \
      egp::set_error("unsupported dtype code %d", (int)(dtype));  \
      return EGP_ERR_INVALID;                                     \
    }                                                             \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace egp
