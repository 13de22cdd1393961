// bf16 tensor-core GEMM for sm_100a: tcgen05.mma with fp32 accumulators in TMEM, operands staged by TMA.
//
//   C[M,N] = act( sum_k A(m,k) B(n,k) + sum_k A2(m,k) B2(n,k) + bias[n] ) + residual[m,n]
//
// Serves every Linear on EgoPack's path (TRNPooling, SAGE lin/lin_l/lin_r, task nets, classifiers, GraphONE
// stages) in forward (A K-major, B K-major), dgrad (B MN-major) and wgrad (A and B MN-major, split-K), plus the
// node x prototype similarity.  The second operand pair accumulates into the SAME TMEM tile, which fuses
// SAGEConv's  lin_l(agg) + lin_r(x)  (and its two dgrads / wgrads) into one pass.
//
// Structure (one persistent CTA per SM, 192 threads):
//   warp 0      TMA producer   : cp.async.bulk.tensor -> 128B-swizzled smem ring (full/empty mbarriers)
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (128 x BN x 16), tcgen05.commit frees slots
//   warps 2..5  epilogue       : tcgen05.ld TMEM -> registers -> bias/act/residual -> global
// The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of tile t overlaps the mainloop
// of tile t+1.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace egp {

constexpr int TBM = 128;     // tile M (UMMA_M)
constexpr int TBK = 64;      // k-block: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;   // bf16
constexpr int TC_THREADS = 192;
// fused top-k kernels run kTopkGroups epilogue warps per TMEM lane quarter (each takes every kTopkGroups-th 32-column
// chunk and keeps its own candidate set), so every scheduler has several epilogue warps to interleave
constexpr int kTopkGroups = 2;
constexpr int TC_THREADS_TOPK = 64 + 128 * kTopkGroups;

// ------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (~2 s) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xfffu) == 0 && clock64() - t0 > 4000000000LL) {
      printf("egopack_b200: mbarrier timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// `bar_addr` is a shared-window address; with CG == 2 it may name the mbarrier of the pair's leader CTA
template <int CG>
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar_addr, void* smem, int c0, int c1) {
  if constexpr (CG == 2)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(bar_addr), "r"(c0), "r"(c1)
        : "memory");
}

// thread-block-cluster plumbing for the CTA-pair (cta_group::2) kernels
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
// smem -> global tile store (clips rows/columns outside the tensor); `reduce` adds into global memory instead
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem, int c0, int c1, bool reduce) {
  if (reduce)
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(smem_u32(smem)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// L2 prefetch of a tile that a later TMA load will fetch (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"((uint64_t)map), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// CG == 2: executed by the same warp of BOTH CTAs of the pair (each gets the address in its own shared memory)
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  if constexpr (CG == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 2)
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// CG == 2: issued by the leader CTA only; multiplies the pair's 256 x BN tile (A rows and B columns split over the
// two CTAs' shared memories, accumulator rows split over their TMEMs)
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if constexpr (CG == 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once all previously issued tcgen05.mma of this thread have completed; CG == 2 arrives on
// the barrier at the same shared-memory offset in BOTH CTAs of the pair
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  if constexpr (CG == 2)
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=2 (SWIZZLE_128B)
// K-major  : rows of 128 B (64 bf16 of K); 8-row groups 1024 B apart (SBO); LBO unused.
// MN-major : 64-element MN atoms; inside an atom K rows are 128 B apart, 8-row groups 1024 B apart (SBO);
//            atoms are TBK*128 B apart (LBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Shared-memory budget (227 KB): a ring of TMA stages for A/B, then the epilogue's staging ring (each epilogue warp
// owns kEpiBufs buffers of 32 rows x 128 B so that a tile's boxes are written back-to-back without waiting for the
// previous TMA store to drain), the per-warp bias slices and the mbarriers.
//
// CG == 2 (CTA pair, cta_group::2): the pair computes a 256 x BN tile; each CTA stages its own 128 rows of A and
// HALF of the B tile, so a k-block costs 32 KB per SM instead of 48 KB (BN = 256) -- 6 stages instead of 4 and a
// third less L2 -> shared-memory traffic per flop.
constexpr int kEpiBufs = 2;
template <int BN, int CG = 1>
struct TcCfg {
  static constexpr int kABytes = TBM * TBK * 2;
  static constexpr int kBBytes = (BN / CG) * TBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiStageBytes = 4 * kEpiBufs * 32 * 128;   // 32 KB
  static constexpr int kEpiBiasBytes = 2 * 256 * 4;                // one bias slice per accumulator stage
  // the dynamic shared-memory window is declared 1024-byte aligned (checked at kernel entry), no slack needed
  static constexpr int kFixedBytes = kEpiStageBytes + kEpiBiasBytes + 256 /*barriers*/;
  static constexpr int kStagesFit = (232448 - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;  // 4 for BN=256 (6 as a CTA pair), 6 for BN=128
  static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;      // power of two for BN in {16,...,256}
  static constexpr int kSmemBytes = kStages * kStageBytes + kFixedBytes;
};

struct TcParams {   // exactly 128 bytes (static_assert below): see the note at cand_thr
  int64_t M, N;
  int kb1, kb2;        // k-blocks of the first / second operand pair
  int splits;          // split-K factor (atomic fp32 accumulation when > 1)
  int m_tiles, n_tiles;
  int slab_rows;       // deterministic split-K: split s writes rows [s*slab_rows, ...) of a [splits*slab_rows, N] workspace
                       // with plain stores (summed afterwards in a fixed order); 0 = reduce-add straight into C.
                       // (sits in what was alignment padding: the struct stays at 128 bytes)
  const float* bias;
  const void* residual;
  int64_t ldr;
  void* C;
  int64_t ldc;
  int act;
  float slope;
  int accumulate;      // C += result (fp32 atomics)
  int tma_store;       // epilogue stages 32x128B boxes in smem and stores them with TMA (coalesced, clipped)
  int topk;            // TOPK kernels: candidates kept per row (8 or 16); 0 otherwise
  int32_t* cand;       // TOPK kernels: [M, TOPK] column indices of the largest entries of each row (unordered)
  float* cand_thr;     // TOPK kernels (optional): [M, kTopkGroups] smallest score each group KEPT -- every column of the
                       // group that is not a candidate scored at most this (the k-NN miss detector's threshold)
                       // STATS kernels reuse the two TOPK fields (the variants are exclusive, and the struct must stay at
                       // 128 bytes: with 144 the 896-byte kernel parameter block grows and ptxas stops keeping the
                       // parameters in uniform registers -- every instantiation got ~25 % more instructions and the
                       // forward GEMMs lost 5-19 %): cand_thr = double* rowstats, per (128-row block, TMEM quarter, n tile)
                       // {sum, sum of squares} of the values stored; topk = n-tile slots per (row block, quarter)
  int debug;           // EGP_TC_DEBUG (timing experiments) bit 0: skip the stores, bit 1: skip the TMEM loads too, bit 2: all CTAs load tile (0,0), bit 3: skip the B loads of odd k-blocks
  uint32_t idesc;
};
static_assert(sizeof(TcParams) == 128, "TcParams must stay at 128 bytes (kernel parameter block of 896 bytes)");

// CG = 2 (launched as clusters of two CTAs): the pair owns a 256 x BN tile.  CTA `rank` stages rows
// [rank*128, rank*128+128) of A and columns [rank*BN/2, (rank+1)*BN/2) of B; the leader (rank 0) issues
// tcgen05.mma.cta_group::2, which reads both shared memories and writes each CTA's 128 accumulator rows into its
// own TMEM; each CTA runs its own epilogue.  Barriers: `full` lives in the leader (both producers' TMA bytes land
// on it), `empty` / `tfull` are signalled in both CTAs by a multicast tcgen05.commit, `tempty` lives in the leader
// and collects the (remote) arrivals of both CTAs' epilogue warps.
// STATS: the epilogue also accumulates {sum, sum of squares} of what it stores (a separate instantiation: the extra
// registers and instructions cost every GEMM ~5 % when they were a run-time branch of the common kernel).
template <int BN, bool A_MN, bool B_MN, typename OutT, int TOPK = 0, int CG = 1, bool STATS = false>
__global__ void __launch_bounds__(TOPK > 0 ? TC_THREADS_TOPK : TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapA2, const __grid_constant__ CUtensorMap mapB2,
               const __grid_constant__ CUtensorMap mapC, const __grid_constant__ CUtensorMap mapR,
               const TcParams p) {
  static_assert(CG == 1 || (TOPK == 0 && BN >= 128), "CTA pairs: plain GEMM with BN >= 128 only");
  using Cfg = TcCfg<BN, CG>;
  constexpr int S = Cfg::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;   // 128-byte-swizzled TMA boxes and UMMA descriptors need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("egopack_b200: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  uint8_t* epi_stage = smem + S * Cfg::kStageBytes;                      // 1024-aligned (stage bytes are)
  float* epi_bias = reinterpret_cast<float*>(epi_stage + Cfg::kEpiStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_stage + Cfg::kEpiStageBytes + Cfg::kEpiBiasBytes);
  uint64_t* full = bars;            // [S]  TMA -> MMA
  uint64_t* empty = bars + S;       // [S]  MMA -> TMA
  uint64_t* tfull = bars + 2 * S;   // [2]  MMA -> epilogue
  uint64_t* tempty = tfull + 2;     // [2]  epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint64_t* rbar = bars + 24;       // [4 warps][kEpiBufs]  residual box landed in the warp's staging buffer
  static_assert((2 * S + 4) * 8 + 4 <= 192 && (24 + 4 * kEpiBufs) * 8 <= 256, "barrier area layout");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
    if (p.kb2 > 0) { tma_prefetch_desc(&mapA2); tma_prefetch_desc(&mapB2); }
    if (p.tma_store) tma_prefetch_desc(&mapC);
    if (p.tma_store && p.residual) tma_prefetch_desc(&mapR);
    for (int s = 0; s < 4 * kEpiBufs; ++s) mbar_init(&rbar[s], 1);
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], (TOPK > 0 ? 4 * kTopkGroups : 4) * CG); }
    fence_barrier_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc<CG>(tmem_slot, Cfg::kTmemCols);
  // everything above touches only this CTA's shared memory / TMEM and the kernel parameters, so it may overlap the
  // tail of the previous kernel; global memory is first read below
  pdl_enter();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // CTA pairs tile M in units of 256 rows: m_units pair-rows, this CTA taking the `cta_rank`-th half of each
  const int m_units = CG == 2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int tiles_mn = m_units * p.n_tiles;
  const int total = tiles_mn * p.splits;
  const int worker = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int workers = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int kb_all = p.kb1 + p.kb2;

  // k-block range of split s: contiguous chunks of the concatenated [pair1 | pair2] k-block list
  auto split_range = [&](int s, int& b, int& e) {
    const int per = (kb_all + p.splits - 1) / p.splits;
    b = s * per;
    e = min(b + per, kb_all);
  };

  // The i-th tile of this CTA.  Normal GEMMs stride the linear tile index by the grid (n fastest, so CTAs running
  // side by side share an A row block).  TOPK kernels give a CTA whole row blocks and walk all of its n tiles in
  // order, so a row's running top-k stays in the registers of one epilogue thread.
  auto tile_at = [&](int iter, int& split, int& m_blk, int& n_blk) -> bool {
    if (TOPK) {
      const int mb = blockIdx.x + (iter / p.n_tiles) * gridDim.x;
      if (mb >= p.m_tiles) return false;
      split = 0; m_blk = mb; n_blk = iter % p.n_tiles;
      return true;
    }
    const int t = worker + iter * workers;
    if (t >= total) return false;
    split = t / tiles_mn;
    const int mn = t % tiles_mn;
    m_blk = (mn / p.n_tiles) * CG + (int)cta_rank; n_blk = mn % p.n_tiles;
    return true;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int split, m_blk, n_blk;
      for (int it = 0; tile_at(it, split, m_blk, n_blk); ++it) {
        int m0 = m_blk * TBM, n0 = n_blk * BN + (int)cta_rank * (BN / CG);
        if (p.debug & 4) { m0 = 0; n0 = 0; }   // timing experiment: every CTA streams the SAME operand tiles
        int kb_b, kb_e;
        split_range(split, kb_b, kb_e);
        for (int kb = kb_b; kb < kb_e; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          const bool skip_b = (p.debug & 8) && (kb & 1);   // timing experiment: 25 % less operand traffic (stale B)
          if (leader) mbar_expect_tx(&full[stage], CG * (skip_b ? Cfg::kABytes : Cfg::kStageBytes));
          const uint32_t full_bar = CG == 2 ? mapa_u32(smem_u32(&full[stage]), 0u) : smem_u32(&full[stage]);
          const bool second = kb >= p.kb1;
          const CUtensorMap* ma = second ? &mapA2 : &mapA;
          const CUtensorMap* mb = second ? &mapB2 : &mapB;
          const int k0 = (second ? kb - p.kb1 : kb) * TBK;
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if (!A_MN) {
            tma_load_2d<CG>(ma, full_bar, sa, k0, m0);               // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int a = 0; a < TBM / 64; ++a)                       // box {64 m, 64 k} per MN atom
              tma_load_2d<CG>(ma, full_bar, sa + a * (TBK * 128), m0 + a * 64, k0);
          }
          if (skip_b) {
          } else if (!B_MN) {
            tma_load_2d<CG>(mb, full_bar, sb, k0, n0);               // box {64 k, BN / CG rows}
          } else {
#pragma unroll
            for (int a = 0; a < BN / CG / 64; ++a)
              tma_load_2d<CG>(mb, full_bar, sb + a * (TBK * 128), n0 + a * 64, k0);
          }
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int split, m_blk, n_blk;
      for (int it = 0; tile_at(it, split, m_blk, n_blk); ++it) {
        int kb_b, kb_e;
        split_range(split, kb_b, kb_e);
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&tempty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = kb_b; kb < kb_e; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
          const uint64_t adesc = A_MN ? make_smem_desc(sa, TBK * 128, 1024) : make_smem_desc(sa, 0, 1024);
          const uint64_t bdesc = B_MN ? make_smem_desc(sb, TBK * 128, 1024) : make_smem_desc(sb, 0, 1024);
#pragma unroll
          for (int k = 0; k < TBK / UMMA_K; ++k) {
            // advance the 14-bit start-address field: 32 B per UMMA_K step (K-major) or 16 rows x 128 B (MN-major)
            const uint64_t ao = (uint64_t)((A_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
            const uint64_t bo = (uint64_t)((B_MN ? k * UMMA_K * 128 : k * UMMA_K * 2) >> 4);
            umma_bf16<CG>(d_tmem, adesc + ao, bdesc + bo, p.idesc, (kb > kb_b || k > 0) ? 1u : 0u);
          }
          umma_commit<CG>(&empty[stage]);
          if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        umma_commit<CG>(&tfull[as]);
      }
    }
  } else {
    // epilogue: TMEM lane quarter is fixed by warp id % 4
    const int quarter = warp & 3;
    OutT* C = reinterpret_cast<OutT*>(p.C);
    const OutT* R = reinterpret_cast<const OutT*>(p.residual);
    int ebuf = 0;  // next staging buffer of this warp's ring
    uint32_t rphase = 0;  // parity bit per staging buffer of this warp's residual barriers
    constexpr int KK = TOPK > 0 ? TOPK : 1;
    float tv[KK];   // fused top-k: the KK largest entries of this thread's row so far (unsorted)
    int ti[KK];
    float tmin = -3.0e38f;
    int smin = 0;
    // release accumulator stage `as` to the MMA issuer (CTA pairs: the leader's barrier, from both CTAs)
    auto tempty_arrive = [&](int as) {
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty[as]), 0u));
      else mbar_arrive(&tempty[as]);
    };
    int split, m_blk, n_blk;
    for (int it = 0; tile_at(it, split, m_blk, n_blk); ++it) {
      const int64_t m0 = (int64_t)m_blk * TBM, n0 = (int64_t)n_blk * BN;
      int kb_b, kb_e;
      split_range(split, kb_b, kb_e);
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const int64_t m = m0 + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      if constexpr (TOPK > 0) {
        // similarity tile -> running set of the KK largest entries of the row; nothing is written to C.
        // The set is UNSORTED (the exact re-rank orders it): an insertion overwrites the slot holding the current
        // minimum and re-derives (minimum, slot) with a tournament -- ~40 short-dependency selects instead of a
        // sorted shift, which matters because an epilogue warp has its scheduler to itself (no latency hiding).
        // Ties keep the earlier (lower) column: only a strictly larger value evicts.
        if (n_blk == 0) {
#pragma unroll
          for (int q = 0; q < KK; ++q) { tv[q] = -3.0e38f; ti[q] = 0x7fffffff; }
          tmin = -3.0e38f;
          smin = 0;
        }
        const int group = (warp - 2) >> 2;  // which of the kTopkGroups column groups this warp scans
#pragma unroll 1
        for (int c = group; c < BN / 32; c += kTopkGroups) {
          const int64_t nb = n0 + c * 32;
          if (nb >= p.N) break;
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]);
            if (v > tmin && nb + j < p.N) {
#pragma unroll
              for (int q = 0; q < KK; ++q) {
                const bool here = q == smin;
                tv[q] = here ? v : tv[q];
                ti[q] = here ? (int)(nb + j) : ti[q];
              }
              // tournament for the new (minimum, slot)
              float mv[KK];
              int ms[KK];
#pragma unroll
              for (int q = 0; q < KK; ++q) { mv[q] = tv[q]; ms[q] = q; }
#pragma unroll
              for (int w = KK / 2; w >= 1; w /= 2) {
#pragma unroll
                for (int q = 0; q < w; ++q) {
                  const bool lt = mv[q + w] < mv[q];
                  mv[q] = lt ? mv[q + w] : mv[q];
                  ms[q] = lt ? ms[q + w] : ms[q];
                }
              }
              tmin = mv[0];
              smin = ms[0];
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tempty_arrive(as);
        if (n_blk == p.n_tiles - 1 && row_ok) {
#pragma unroll
          for (int q = 0; q < KK; ++q) p.cand[(m * kTopkGroups + group) * KK + q] = ti[q];
          if (p.cand_thr) p.cand_thr[m * kTopkGroups + group] = tmin;
        }
        continue;
      }
      const bool atomic = (p.splits > 1 && p.slab_rows == 0) || p.accumulate;
      const int64_t slab_off = (int64_t)split * p.slab_rows;   // 0 unless deterministic split-K
      const bool has_k = kb_e > kb_b;
      constexpr int CHT = 128 / (int)sizeof(OutT);  // columns per 128-byte staged row: 64 (bf16) / 32 (fp32)
      if constexpr (BN >= CHT) {
        if (p.tma_store) {
          uint8_t* stage_w0 = epi_stage + quarter * (kEpiBufs * 32 * 128);
          float* bias_w = epi_bias + as * 256;
          const bool use_bias = p.bias != nullptr;   // every tile stages (zeros for split > 0): barrier stays uniform
          // residual: the matching 32 x 128 B box of R is TMA-loaded INTO the staging buffer one chunk ahead, the
          // accumulator is added onto it in place and the same buffer is stored
          const bool has_res = p.residual != nullptr;
          uint64_t* rbar_w = rbar + quarter * kEpiBufs;
          auto load_residual = [&](int buf, int64_t col) {
            mbar_expect_tx(&rbar_w[buf], 32 * 128);
            tma_load_2d<1>(&mapR, smem_u32(&rbar_w[buf]), stage_w0 + buf * (32 * 128), (int)col, (int)(m0 + quarter * 32));
          };
          if (has_res && lane == 0) {
            bulk_wait_read<kEpiBufs - 1>();
            load_residual(ebuf, n0);
          }
          if (use_bias) {
            // the tile's bias slice, staged once by the 128 epilogue threads.  Double-buffered by accumulator
            // stage: a warp can only reach the tile that reuses this buffer after every warp passed the barrier
            // of the tile in between, i.e. finished reading it.
            const int et = (int)threadIdx.x - 64;
            for (int j = et; j < BN; j += 128)
              bias_w[j] = (split == 0 && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
            asm volatile("bar.sync 1, 128;" ::: "memory");
          }
          [[maybe_unused]] float st_s = 0.f, st_q = 0.f;   // STATS: this thread's row, all columns of the tile
#pragma unroll 1
          for (int c = 0; c < BN / CHT; ++c) {
            const int64_t nb = n0 + c * CHT;
            if (nb >= p.N) break;  // warp-uniform
            uint32_t r[CHT];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c * CHT);
            if (!(p.debug & 2)) {
              tmem_ld32(taddr, r);
              if constexpr (CHT == 64) tmem_ld32(taddr + 32, r + 32);
              tmem_ld_wait();
            } else {
#pragma unroll
              for (int j = 0; j < CHT; ++j) r[j] = 0u;
            }
            if (p.debug & 1) continue;
            float v[CHT];
#pragma unroll
            for (int j = 0; j < CHT; ++j) v[j] = has_k ? __uint_as_float(r[j]) : 0.f;
            if (use_bias) {
#pragma unroll
              for (int j = 0; j < CHT; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_w + c * CHT + j);  // broadcast read
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
            if (p.act != EGP_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < CHT; ++j) v[j] = apply_act(v[j], p.act, p.slope);
            }
            // staging ring: buffer `ebuf` is free once all but the newest kEpiBufs-1 stores have been read out
            uint8_t* stage_w = stage_w0 + ebuf * (32 * 128);
            const int cur = ebuf;
            ebuf = (ebuf + 1) % kEpiBufs;
            if (!has_res) {
              if (lane == 0) bulk_wait_read<kEpiBufs - 1>();
              __syncwarp();
            } else {
              if (lane == 0 && c + 1 < BN / CHT && nb + CHT < p.N) {
                bulk_wait_read<0>();       // the other buffer's store (previous chunk) has been read out
                load_residual(ebuf, nb + CHT);
              }
              mbar_wait(&rbar_w[cur], (rphase >> cur) & 1u);
              rphase ^= 1u << cur;
            }
            uint8_t* row = stage_w + lane * 128;
#pragma unroll
            for (int q = 0; q < 8; ++q) {  // 16-byte pieces, XOR-swizzled with the row index (SWIZZLE_128B)
              if (has_res) {
                const Vec<OutT> rr = Vec<OutT>::load(reinterpret_cast<const OutT*>(row + ((q ^ (lane & 7)) << 4)));
#pragma unroll
                for (int e = 0; e < Vec<OutT>::N; ++e) v[q * Vec<OutT>::N + e] += rr.v[e];
              }
              if constexpr (STATS) {   // statistics of the values as computed (fp32, before the output rounding)
#pragma unroll
                for (int e = 0; e < Vec<OutT>::N; ++e) {
                  const float t = v[q * Vec<OutT>::N + e];
                  st_s += t;
                  st_q = fmaf(t, t, st_q);
                }
              }
              uint4 pk;
              if constexpr (sizeof(OutT) == 2) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
                pk = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
              } else {
                pk = make_uint4(__float_as_uint(v[4 * q + 0]), __float_as_uint(v[4 * q + 1]),
                                __float_as_uint(v[4 * q + 2]), __float_as_uint(v[4 * q + 3]));
              }
              *reinterpret_cast<uint4*>(row + ((q ^ (lane & 7)) << 4)) = pk;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&mapC, stage_w, (int)nb, (int)(slab_off + m0 + quarter * 32), atomic);
              bulk_commit();
            }
          }
          if constexpr (STATS) {
            // rows past M hold bias-only garbage: excluded.  One {sum, sumsq} pair per (row block, quarter, n tile), written
            // exactly once, so a later reduction in a fixed order is deterministic
            double ds = row_ok ? (double)st_s : 0.0, dq = row_ok ? (double)st_q : 0.0;
            ds = warp_sum(ds);
            dq = warp_sum(dq);
            if (lane == 0) {
              double* sp = reinterpret_cast<double*>(p.cand_thr) + (((int64_t)m_blk * 4 + quarter) * p.topk + n_blk) * 2;
              sp[0] = ds;
              sp[1] = dq;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) tempty_arrive(as);
          continue;
        }
      }
      constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
      for (int c = 0; c < BN / CH; ++c) {
        uint32_t r[CH];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + c * CH);
        if (CH == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
        tmem_ld_wait();
        const int64_t nb = n0 + c * CH;
        if (row_ok && nb < p.N) {
          float v[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j) v[j] = has_k ? __uint_as_float(r[j]) : 0.f;
          const bool full_chunk = nb + CH <= p.N;
          if (p.bias && split == 0) {
#pragma unroll
            for (int j = 0; j < CH; ++j)
              if (full_chunk || nb + j < p.N) v[j] += __ldg(p.bias + nb + j);
          }
          if (p.act != EGP_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < CH; ++j) v[j] = apply_act(v[j], p.act, p.slope);
          }
          OutT* crow = C + (slab_off + m) * p.ldc + nb;
          if (atomic) {
            if constexpr (sizeof(OutT) == 4) {
#pragma unroll
              for (int j = 0; j < CH; ++j)
                if (full_chunk || nb + j < p.N) atomicAdd(reinterpret_cast<float*>(crow) + j, v[j]);
            }
          } else {
            constexpr int VN = 16 / (int)sizeof(OutT);
            const bool vec = full_chunk && ((reinterpret_cast<uintptr_t>(crow) & 15u) == 0) &&
                             (!R || ((reinterpret_cast<uintptr_t>(R + m * p.ldr + nb) & 15u) == 0));
            if (vec) {
#pragma unroll
              for (int j = 0; j < CH; j += VN) {
                Vec<OutT> o;
                if (R) {
                  const Vec<OutT> rr = Vec<OutT>::load(R + m * p.ldr + nb + j);
#pragma unroll
                  for (int q = 0; q < VN; ++q) o.v[q] = v[j + q] + rr.v[q];
                } else {
#pragma unroll
                  for (int q = 0; q < VN; ++q) o.v[q] = v[j + q];
                }
                o.store(crow + j);
              }
            } else {
#pragma unroll
              for (int j = 0; j < CH; ++j) {
                if (full_chunk || nb + j < p.N) {
                  float o = v[j];
                  if (R) o += to_float<OutT>(R[m * p.ldr + nb + j]);
                  crow[j] = from_float<OutT>(o);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) tempty_arrive(as);
    }
    if (p.tma_store && lane == 0) bulk_wait0();  // all staged boxes have landed in global memory
  }

  // teardown.  CTA pairs: neither CTA may exit (or free TMEM) while its partner can still touch its shared memory
  // or barriers, so the whole cluster meets here first.
  __syncwarp();
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------------
// host side: tensor maps (driver entry point resolved at run time: the library has no link-time libcuda
// dependency, so it loads on machines without a driver) and launch
// ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

struct MapKey {
  const void* ptr;
  int64_t inner, outer, ld;
  int box_inner, box_outer, esize;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer && esize == o.esize;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&](int64_t v) { h ^= std::hash<int64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box_inner); mix(k.box_outer); mix(k.esize);
    return h;
  }
};

// 2-D tensor [outer, inner] (inner contiguous, row stride ld elements of esize bytes: 2 = bf16, 4 = fp32),
// box {box_inner, box_outer} with box_inner * esize == 128, 128B swizzle
static int make_map(const void* ptr, int64_t inner, int64_t outer, int64_t ld, int box_inner, int box_outer, int esize,
                    CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{ptr, inner, outer, ld, box_inner, box_outer, esize};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return EGP_OK; }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
    return EGP_ERR_UNSUPPORTED;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * (cuuint64_t)esize};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUresult r = enc(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                         const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%lld outer=%lld ld=%lld box=%dx%d esize=%d", (int)r, ptr,
              (long long)inner, (long long)outer, (long long)ld, box_inner, box_outer, esize);
    return EGP_ERR_CUDA;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return EGP_OK;
}

// operand with `rows` (M or N) and `k`: trans=0 -> [rows,k] row-major (K-major); trans=1 -> [k,rows] (MN-major)
static int operand_map(const void* ptr, int64_t rows, int64_t k, int64_t ld, int trans, int tile_rows, CUtensorMap* out) {
  if (!trans) return make_map(ptr, k, rows, ld, 64, tile_rows, 2, out);
  return make_map(ptr, rows, k, ld, 64, 64, 2, out);
}

bool tc_gemm_supported(const void* A, int64_t lda, const void* B, int64_t ldb, const void* A2, int64_t lda2,
                       const void* B2, int64_t ldb2) {
  auto ok = [](const void* p, int64_t ld) { return !p || (aligned16(p) && ld % 8 == 0); };
  return ok(A, lda) && ok(B, ldb) && ok(A2, lda2) && ok(B2, ldb2);
}

// `work` = tiles (CG == 1) or pair tiles (CG == 2) to distribute; the grid is min(work, resident CTAs / clusters)
template <int BN, bool A_MN, bool B_MN, typename OutT, int CG, bool STATS = false>
static int tc_launch_inst(const CUtensorMap* maps, const TcParams& p, int work, cudaStream_t stream) {
  auto kern = tc_gemm_kernel<BN, A_MN, B_MN, OutT, 0, CG, STATS>;
  constexpr int kSmem = TcCfg<BN, CG>::kSmemBytes;
  // cudaFuncSetAttribute is per DEVICE: a process that drives several GPUs (or a feeder thread racing the training
  // thread) must set it once on each of them
  static std::mutex attr_mu;
  static bool attr_done[kMaxDevices] = {};
  static int max_clusters_dev[kMaxDevices] = {};
  const int dev = current_device();
  std::unique_lock<std::mutex> attr_lock(attr_mu);
  bool& attr_set = attr_done[dev];
  int& max_clusters = max_clusters_dev[dev];
  if (!attr_set) {
    EGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    if (CG == 2) {
      cudaLaunchConfig_t q = {};
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.gridDim = dim3(2 * (unsigned)sm_count()); q.blockDim = dim3(TC_THREADS); q.dynamicSmemBytes = kSmem;
      q.attrs = qa; q.numAttrs = 1;
      EGP_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kern, &q));
      if (max_clusters < 1) {
        set_error("tc_gemm: no CTA pair of the cta_group::2 kernel fits on this device");
        return EGP_ERR_UNSUPPORTED;
      }
      // timing experiment: run on fewer SMs (is the kernel bound per SM, or by a chip-wide L2 limit?)
      if (const char* e = getenv("EGP_TC_MAX_CLUSTERS")) { const int lim = atoi(e); if (lim >= 1 && lim < max_clusters) max_clusters = lim; }
    }
    attr_set = true;
  }
  const int max_clusters_now = max_clusters;
  attr_lock.unlock();
  if constexpr (CG == 2) {
    const int clusters = work < max_clusters_now ? work : max_clusters_now;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(2 * (unsigned)clusters); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = kSmem;
    cfg.stream = stream; cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    EGP_CUDA(cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p));
  } else {
    const int sms = sm_count();
    (void)launch_kernel(kern, work < sms ? work : sms, TC_THREADS, kSmem, stream, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
  }
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

template <int BN, typename OutT, int CG>
static int tc_launch_major(int a_trans, int b_trans, const CUtensorMap* maps, const TcParams& p, int work,
                           cudaStream_t stream) {
  if (!a_trans && !b_trans) return tc_launch_inst<BN, false, false, OutT, CG>(maps, p, work, stream);
  if (!a_trans && b_trans) {
    if constexpr (BN >= 64) return tc_launch_inst<BN, false, true, OutT, CG>(maps, p, work, stream);
  }
  if (a_trans && !b_trans) return tc_launch_inst<BN, true, false, OutT, CG>(maps, p, work, stream);
  if (a_trans && b_trans) {
    if constexpr (BN >= 64) return tc_launch_inst<BN, true, true, OutT, CG>(maps, p, work, stream);
  }
  set_error("tc_gemm: MN-major B needs a tile N of at least 64");
  return EGP_ERR_INVALID;
}

// Deterministic split-K (egp_set_deterministic): the splits write their partial tiles into slabs of a workspace and this
// kernel sums the slabs in split order -- bit-reproducible weight gradients, at the price of one extra pass over
// splits x [M,N] fp32 (the default reduce-adds the splits into C with TMA, whose arrival order varies run to run).
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t slab_stride, float* __restrict__ C, int64_t ldc,
                     int64_t M, int64_t N, int accumulate) {
  pdl_enter();
  const int64_t total = M * N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N, n = i - m * N;
    float acc = accumulate ? C[m * ldc + n] : 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[(int64_t)s * slab_stride + i];
    C[m * ldc + n] = acc;
  }
}

static std::atomic<int> g_deterministic{0};
void tc_set_deterministic(int on) { g_deterministic.store(on ? 1 : 0); }
int tc_get_deterministic() { return g_deterministic.load(); }
// upper bound of the split-K workspace: at most one tile per concurrently running CTA (pair) is in flight per wave
size_t tc_gemm_workspace_bytes() {
  return g_deterministic.load() ? sizeof(float) * (size_t)sm_count() * 128 * 256 + 256 : 0;
}

// one slot per 256-wide n tile: the statistics epilogue only exists in the CTA-pair kernel (BN = 256)
int64_t tc_gemm_rowstats_slots(int64_t N) { return ceil_div(N, 256); }

int tc_gemm_launch(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans, const void* A2,
                   int64_t lda2, const void* B2, int64_t ldb2, int64_t K2, const float* bias, const void* residual,
                   int64_t ldr, void* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope,
                   int out_dtype, int accumulate, cudaStream_t stream, double* rowstats, void* workspace, size_t ws_bytes) {
  if (M == 0 || N == 0) return EGP_OK;
  const int sms = sm_count();
  const int m_tiles = (int)ceil_div(M, TBM);
  // tile N: no wider than N needs.  When the epilogue is linear and fp32 (wgrad) SMs are filled by split-K with
  // the widest tile; otherwise the tile is halved (not below 64: narrower tiles are smem-bandwidth bound) while
  // the tile count leaves SMs idle.  MN-major B needs >= 64 (one swizzle atom).
  const int bn_min = b_trans ? 64 : 16;
  const bool can_split = out_dtype == EGP_F32 && act == EGP_ACT_NONE && !residual;
  const int kb_total = (int)ceil_div(K, TBK) + ((A2 && B2 && K2 > 0) ? (int)ceil_div(K2, TBK) : 0);
  int bn = 256;
  while (bn > bn_min && bn / 2 >= N) bn /= 2;
  if (!(can_split && kb_total >= 16))
    while (bn > 64 && (int64_t)m_tiles * ceil_div(N, bn) < sms) bn /= 2;
  const int n_tiles = (int)ceil_div(N, bn);
  const bool has2 = A2 && B2 && K2 > 0;
  // CTA pairs (cta_group::2) for the widest tile whenever there are at least two row blocks; EGP_TC_CG=1 disables
  static const int tc_cg = [] { const char* e = getenv("EGP_TC_CG"); return e ? atoi(e) : 2; }();
  const int cg = (tc_cg == 2 && bn == 256 && m_tiles >= 2) ? 2 : 1;
  TcParams p;
  p.M = M; p.N = N;
  p.kb1 = (int)ceil_div(K, TBK);
  p.kb2 = has2 ? (int)ceil_div(K2, TBK) : 0;
  p.m_tiles = m_tiles; p.n_tiles = n_tiles;
  p.bias = bias; p.residual = residual; p.ldr = ldr; p.C = C; p.ldc = ldc;
  p.act = act; p.slope = slope; p.accumulate = accumulate;
  p.topk = 0; p.cand = nullptr; p.cand_thr = nullptr;
  const int stat_slots = (int)tc_gemm_rowstats_slots(N);
  if (rowstats) { p.cand_thr = reinterpret_cast<float*>(rowstats); p.topk = stat_slots; }
  static const int tc_debug = [] { const char* e = getenv("EGP_TC_DEBUG"); return e ? atoi(e) : 0; }();
  p.debug = tc_debug;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(a_trans ? 1 : 0) << 15) |
            ((uint32_t)(b_trans ? 1 : 0) << 16) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)((TBM * cg) >> 4) << 24);
  // split-K: only for fp32 outputs without a non-linear epilogue (wgrad); keeps >= 4 k-blocks per split
  int splits = 1;
  const int tiles = (cg == 2 ? (m_tiles + 1) / 2 : m_tiles) * n_tiles, kb_all = p.kb1 + p.kb2;
  const int slots = sms / cg;  // concurrently running tiles (CTAs, or CTA pairs)
  if (can_split && !rowstats && tiles * 2 <= slots && kb_all >= 8) {
    splits = slots / tiles;
    if (splits > kb_all / 4) splits = kb_all / 4;
    if (splits < 1) splits = 1;
    const int per = (kb_all + splits - 1) / splits;
    splits = (kb_all + per - 1) / per;  // no empty split
  }
  p.splits = splits;
  p.slab_rows = 0;
  if (accumulate && out_dtype != EGP_F32) {
    set_error("tc_gemm: accumulate needs an fp32 output");
    return EGP_ERR_INVALID;
  }
  // deterministic split-K: slabs in the caller's workspace + an ordered reduction (N contiguous in the slabs)
  void* const c_user = C;
  const int64_t ldc_user = ldc;
  const int acc_user = accumulate;
  const int64_t slab_rows = (int64_t)m_tiles * TBM;
  const bool slab = splits > 1 && g_deterministic.load() != 0;
  if (slab) {
    const size_t need = sizeof(float) * (size_t)splits * (size_t)slab_rows * (size_t)N;
    if (!workspace || ws_bytes < need + 128) {
      set_error("tc_gemm: deterministic split-K needs a workspace of %zu bytes (egp_gemm_workspace)", need + 128);
      return EGP_ERR_WORKSPACE;
    }
    C = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 127) & ~(uintptr_t)127);
    ldc = N;
    accumulate = 0;
    p.C = C; p.ldc = ldc; p.accumulate = 0;
    p.slab_rows = (int)slab_rows;
  }
  if (splits > 1 && !accumulate && !slab) {  // atomics need a zeroed destination
    if (ldc == N) EGP_CUDA(cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * (size_t)N, stream));
    else EGP_CUDA(cudaMemset2DAsync(C, sizeof(float) * ldc, 0, sizeof(float) * N, (size_t)M, stream));
  }
  CUtensorMap maps[6];
  int rc;
  if ((rc = operand_map(A, M, K, lda, a_trans, TBM, &maps[0])) != EGP_OK) return rc;
  if ((rc = operand_map(B, N, K, ldb, b_trans, bn / cg, &maps[1])) != EGP_OK) return rc;
  if (has2) {
    if ((rc = operand_map(A2, M, K2, lda2, a_trans, TBM, &maps[2])) != EGP_OK) return rc;
    if ((rc = operand_map(B2, N, K2, ldb2, b_trans, bn / cg, &maps[3])) != EGP_OK) return rc;
  } else {
    maps[2] = maps[0];
    maps[3] = maps[1];
  }
  // TMA-store epilogue: needs a 16-byte-aligned C (and residual) with 16-byte row pitch and a tile at least one
  // 128-byte staged row wide; anything else keeps the register-direct epilogue
  const int esz = out_dtype == EGP_F32 ? 4 : 2;
  p.tma_store = (aligned16(C) && (ldc * esz) % 16 == 0 && bn >= 128 / esz &&
                 (!residual || (aligned16(residual) && (ldr * esz) % 16 == 0))) ? 1 : 0;
  if (rowstats) {
    // the statistics ride on the TMA-store epilogue of the CTA-pair kernel (256-wide tiles, K-major operands): the
    // shape of the SAGE layer's dual GEMM.  Anything else: the caller runs the plain GEMM and a statistics pass.
    if (!p.tma_store || cg != 2 || bn != 256 || a_trans || b_trans || N % 64 != 0 || splits != 1 || accumulate) {
      set_error("tc_gemm: row statistics need K-major operands, 256-wide CTA-pair tiles, N %% 64 == 0, no split-K");
      return EGP_ERR_UNSUPPORTED;
    }
    EGP_CUDA(cudaMemsetAsync(rowstats, 0, sizeof(double) * 2 * (size_t)m_tiles * 4 * (size_t)stat_slots, stream));
  }
  maps[4] = maps[0];
  maps[5] = maps[0];
  if (p.tma_store) {
    if ((rc = make_map(C, N, slab ? (int64_t)splits * slab_rows : M, ldc, 128 / esz, 32, esz, &maps[4])) != EGP_OK) return rc;
    if (residual && (rc = make_map(residual, N, M, ldr, 128 / esz, 32, esz, &maps[5])) != EGP_OK) return rc;
  }
  const int total = tiles * splits;
  auto launch = [&]() -> int {
    if (rowstats)
      return out_dtype == EGP_F32 ? tc_launch_inst<256, false, false, float, 2, true>(maps, p, total, stream)
                                  : tc_launch_inst<256, false, false, __nv_bfloat16, 2, true>(maps, p, total, stream);
    if (cg == 2)
      return out_dtype == EGP_F32 ? tc_launch_major<256, float, 2>(a_trans, b_trans, maps, p, total, stream)
                                  : tc_launch_major<256, __nv_bfloat16, 2>(a_trans, b_trans, maps, p, total, stream);
#define EGP_TC_BN(BNV)                                                                                          \
  case BNV:                                                                                                     \
    return out_dtype == EGP_F32 ? tc_launch_major<BNV, float, 1>(a_trans, b_trans, maps, p, total, stream)      \
                                : tc_launch_major<BNV, __nv_bfloat16, 1>(a_trans, b_trans, maps, p, total, stream);
    switch (bn) {
      EGP_TC_BN(256)
      EGP_TC_BN(128)
      EGP_TC_BN(64)
      EGP_TC_BN(32)
      EGP_TC_BN(16)
    }
#undef EGP_TC_BN
    set_error("tc_gemm: no kernel for tile N %d", bn);
    return EGP_ERR_INVALID;
  };
  rc = launch();
  if (rc != EGP_OK || !slab) return rc;
  const int64_t elems = M * N;
  int64_t grid = ceil_div(elems, 256 * 4);
  const int64_t cap = (int64_t)sms * 8;
  grid = grid > cap ? cap : (grid < 1 ? 1 : grid);
  (void)launch_kernel(splitk_reduce_kernel, (unsigned)grid, 256, 0, stream, (const float*)C, splits, slab_rows * N,
                      reinterpret_cast<float*>(c_user), ldc_user, M, N, acc_user);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

// Fused similarity + top-k: S = A[M,K] B[N,K]^T on the tensor cores, the `keep` (8 or 16) largest entries of every row
// are selected in the epilogue (registers of the thread that owns the row) and only their column indices leave the
// chip: cand int32 [M, kTopkGroups * keep], unordered (the re-rank scores and sorts them).  Replaces the [M,N] fp32 similarity round trip through HBM.
template <int KEEP>
static int tc_gemm_topk_inst(const CUtensorMap* maps, const TcParams& p, int grid, cudaStream_t stream) {
  constexpr int BN = 256;
  auto kern = tc_gemm_kernel<BN, false, false, float, KEEP>;
  static std::mutex attr_mu;
  static bool attr_done[kMaxDevices] = {};
  {
    std::lock_guard<std::mutex> lock(attr_mu);
    bool& attr_set = attr_done[current_device()];
    if (!attr_set) {
      EGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<BN>::kSmemBytes));
      attr_set = true;
    }
  }
  (void)launch_kernel(kern, grid, TC_THREADS_TOPK, TcCfg<BN>::kSmemBytes, stream, maps[0], maps[1], maps[2], maps[3], maps[4], maps[4], p);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int tc_topk_groups() { return kTopkGroups; }

int tc_gemm_topk_launch(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K, int keep,
                        int32_t* cand, float* cand_thr, cudaStream_t stream) {
  if (M == 0) return EGP_OK;
  if (keep != 8 && keep != 16) {
    set_error("tc_gemm_topk: keep must be 8 or 16");
    return EGP_ERR_INVALID;
  }
  constexpr int BN = 256;
  const int sms = sm_count();
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N;
  p.kb1 = (int)ceil_div(K, TBK);
  p.kb2 = 0;
  p.splits = 1;
  p.m_tiles = (int)ceil_div(M, TBM);
  p.n_tiles = (int)ceil_div(N, BN);
  p.act = EGP_ACT_NONE;
  p.topk = keep;
  p.cand = cand;
  p.cand_thr = cand_thr;
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);
  CUtensorMap maps[5];
  int rc;
  if ((rc = operand_map(A, M, K, lda, 0, TBM, &maps[0])) != EGP_OK) return rc;
  if ((rc = operand_map(B, N, K, ldb, 0, BN, &maps[1])) != EGP_OK) return rc;
  maps[2] = maps[0]; maps[3] = maps[1]; maps[4] = maps[0];
  const int grid = p.m_tiles < sms ? p.m_tiles : sms;
  return keep == 8 ? tc_gemm_topk_inst<8>(maps, p, grid, stream) : tc_gemm_topk_inst<16>(maps, p, grid, stream);
}

}  // namespace egp
