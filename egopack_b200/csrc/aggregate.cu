// SAGE mean aggregation (a5): out[i] = s_out(i) * sum_{j in N(i)} s_in(j) * x[j]      -- HBM bound.
//
// Replaces the reference's index_select (materialises [E,C]) + atomic scatter_add_ + count + divide inside
// gnn.SAGEConv (models/graph.py:42).  Algorithmic traffic is one read + one write of the [N,C] activations
// (2*C*b bytes per node), independent of the window radius, because the adjacency of a temporal graph is a
// band: consecutive output rows share all but one of their neighbour rows.
//
// band kernels: a CTA owns a strip of consecutive rows x a 128-vector (16 B each) column chunk and a thread only
// ever touches its own 16-byte column, so no block barrier is needed.
//   radius <= 4 : the 2k+1-row window is held in registers (sage_mean_band_reg_kernel), summed in ascending
//                 neighbour order (bit-identical to a sequential scatter_add);
//   radius  > 4 : a running window sum adds the entering row (from HBM) and subtracts the leaving row (re-read from
//                 L2), software-pipelined over few resident CTAs per SM (sage_mean_band_run_kernel).  Variants that
//                 measured slower and were removed: a cp.async shared-memory ring (occupancy 1-2 CTAs per SM, 3.5 ms
//                 at 524 288 nodes against 0.56 ms) and one thread per output vector with L1 re-reads.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace egp {

constexpr int kAggThreads = 128;

// Small radii (K <= 4): the sliding window lives in REGISTERS.  A thread owns one 16-byte column of a strip of
// consecutive rows; every row is loaded exactly once (U independent loads in flight per thread), unpacked,
// optionally pre-scaled, and then reused by the 2K neighbouring outputs straight from registers.  All window
// indices are compile-time constants (the loops are fully unrolled), so nothing spills to local memory.
//
// MODE selects the star extension used by LTA graphs (band + "last input clips -> every forecast clip", see
// band_star_windows_kernel in edges.cu):
//   kBandPlain : pure band.
//   kBandExt   : forward direction.  A star target also sums the few source rows outside its band, [ext_lo, ext_hi)
//                (at most K rows, the same ones for every target of a graph: L1/L2 hits).
//   kBandHub   : backward direction.  A star source receives from ~every row of its graph; instead of a 125-row gather
//                in one thread, every thread adds the (scaled) rows it streams anyway to a per-graph hub sum when they
//                are star targets, and flushes that partial once per (strip, graph) to hub_part[strip + graph]
//                (a unique, monotone index).  sage_hub_fixup_kernel then rewrites the <= K source rows of each graph
//                from the partials in a fixed order -- deterministic, and the tensor is still read exactly once.
enum { kBandPlain = 0, kBandExt = 1, kBandHub = 2 };

template <typename T, int K, int U, int MODE>
__global__ void __launch_bounds__(kAggThreads, MODE == kBandHub ? 6 : 8)
sage_mean_band_reg_kernel(const T* __restrict__ x, T* __restrict__ out, int n, int64_t channels, int64_t ldx,
                          int64_t ldo, int rows_per_cta, const int32_t* __restrict__ win_lo,
                          const int32_t* __restrict__ win_hi, const float* __restrict__ scale_out,
                          const float* __restrict__ scale_in, const int32_t* __restrict__ ext_lo,
                          const int32_t* __restrict__ ext_hi, const int32_t* __restrict__ hub_slot,
                          float* __restrict__ hub_part) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  constexpr int W = 2 * K + U;
  float hub[MODE == kBandHub ? VN : 1];
  int cur_slot = -1;
  auto hub_flush = [&](int64_t col_) {
    if constexpr (MODE == kBandHub) {
      if (cur_slot >= 0) {
        float* hp = hub_part + ((int64_t)blockIdx.x + cur_slot) * channels + col_;
#pragma unroll
        for (int c = 0; c < VN; c += 4) *reinterpret_cast<float4*>(hp + c) = make_float4(hub[c], hub[c + 1], hub[c + 2], hub[c + 3]);
      }
    }
  };
  const int64_t col = ((int64_t)blockIdx.y * kAggThreads + threadIdx.x) * VN;
  if (col >= channels) return;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(r0 + rows_per_cta, n);
  if (r0 >= r1) return;
  const T* xc = x + col;
  Raw<T> w[W];  // packed window: rows [g-K, g+U+K)
#pragma unroll
  for (int d = 0; d < 2 * K; ++d) {
    const int j = r0 - K + d;
    w[d] = (j >= 0 && j < n) ? Raw<T>::load(xc + (int64_t)j * ldx) : Raw<T>::zero();
  }
#pragma unroll 1
  for (int g = r0; g < r1; g += U) {
#pragma unroll
    for (int u = 0; u < U; ++u) {  // U independent 16-byte loads in flight
      const int j = g + K + u;
      w[2 * K + u] = (j < n) ? Raw<T>::load(xc + (int64_t)j * ldx) : Raw<T>::zero();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = g + u;
      if (i < r1) {
        const int lo = win_lo[i], hi = win_hi[i];
        Vec<T> acc;
#pragma unroll
        for (int c = 0; c < VN; ++c) acc.v[c] = 0.f;
#pragma unroll
        for (int d = 0; d <= 2 * K; ++d) {
          if (d == K) continue;
          const int j = i - K + d;
          if (j >= lo && j <= hi) {
            const Vec<T> v = w[u + d].unpack();
            const float s = scale_in ? scale_in[j] : 1.f;
#pragma unroll
            for (int c = 0; c < VN; ++c) acc.v[c] += s * v.v[c];
          }
        }
        if constexpr (MODE == kBandExt) {
          const int el = ext_lo[i], eh = ext_hi[i];
          for (int j = el; j < eh; ++j) {
            const Vec<T> v = Vec<T>::load(xc + (int64_t)j * ldx);
            const float s = scale_in ? scale_in[j] : 1.f;
#pragma unroll
            for (int c = 0; c < VN; ++c) acc.v[c] += s * v.v[c];
          }
        }
        if constexpr (MODE == kBandHub) {
          const int slot = hub_slot[i];
          if (slot != cur_slot) {
            hub_flush(col);
            cur_slot = slot;
#pragma unroll
            for (int c = 0; c < VN; ++c) hub[c] = 0.f;
          }
          if (slot >= 0) {
            const Vec<T> v = w[u + K].unpack();   // the row itself
            const float s = scale_in ? scale_in[i] : 1.f;
#pragma unroll
            for (int c = 0; c < VN; ++c) hub[c] += s * v.v[c];
          }
        }
        const float so = scale_out ? scale_out[i] : 1.f;
#pragma unroll
        for (int c = 0; c < VN; ++c) acc.v[c] *= so;
        acc.store(out + (int64_t)i * ldo + col);
      }
    }
#pragma unroll
    for (int d = 0; d < 2 * K; ++d) w[d] = w[d + U];
  }
  hub_flush(col);
}

// Rewrites the star-source rows of every graph after a kBandHub pass: out[s] = s_out(s) * (sum of the band neighbours
// of s that are NOT star targets + the graph's hub sum).  grid (graphs, column chunks).
template <typename T>
__global__ void __launch_bounds__(kAggThreads)
sage_hub_fixup_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t channels, int64_t ldx, int64_t ldo,
                      int rows_per_cta, const int32_t* __restrict__ win_lo, const int32_t* __restrict__ win_hi,
                      const float* __restrict__ scale_out, const float* __restrict__ scale_in,
                      const int32_t* __restrict__ graph_meta, const float* __restrict__ hub_part) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const int64_t col = ((int64_t)blockIdx.y * kAggThreads + threadIdx.x) * VN;
  if (col >= channels) return;
  const int g = blockIdx.x;
  const int src_lo = graph_meta[4 * g], src_hi = graph_meta[4 * g + 1];
  const int tgt_lo = graph_meta[4 * g + 2], tgt_hi = graph_meta[4 * g + 3];
  if (src_lo >= src_hi || tgt_lo >= tgt_hi) return;
  float h[VN];
#pragma unroll
  for (int c = 0; c < VN; ++c) h[c] = 0.f;
  for (int strip = tgt_lo / rows_per_cta; strip <= (tgt_hi - 1) / rows_per_cta; ++strip) {   // fixed order
    const float* hp = hub_part + ((int64_t)strip + g) * channels + col;
#pragma unroll
    for (int c = 0; c < VN; c += 4) {
      const float4 t = *reinterpret_cast<const float4*>(hp + c);
      h[c] += t.x; h[c + 1] += t.y; h[c + 2] += t.z; h[c + 3] += t.w;
    }
  }
  for (int s = src_lo; s < src_hi; ++s) {
    Vec<T> acc;
#pragma unroll
    for (int c = 0; c < VN; ++c) acc.v[c] = 0.f;
    const int lo = win_lo[s], hi = win_hi[s];
    for (int j = lo; j <= hi; ++j) {
      if (j == s || (j >= tgt_lo && j < tgt_hi)) continue;
      const Vec<T> v = Vec<T>::load(x + (int64_t)j * ldx + col);
      const float sj = scale_in ? scale_in[j] : 1.f;
#pragma unroll
      for (int c = 0; c < VN; ++c) acc.v[c] += sj * v.v[c];
    }
    const float so = scale_out ? scale_out[s] : 1.f;
#pragma unroll
    for (int c = 0; c < VN; ++c) acc.v[c] = (acc.v[c] + h[c]) * so;
    acc.store(out + (int64_t)s * ldo + col);
  }
}

// Large radii (K > 4): running window sum held in REGISTERS.  A thread owns one 16-byte column of a strip of rows.
// For every output row it needs three rows: the one entering the window (first touch -> HBM), the one leaving it
// and the row itself (both touched <= 2K+1 rows ago by the same CTA -> L2 hits), so DRAM traffic stays at one read + one
// write of the tensor while no shared-memory ring limits occupancy.  At a graph boundary (or strip start) the sum restarts
// from scratch, so no cancellation residue survives a graph.
//   * A group of U rows whose windows all slide by exactly one row on both ends (every interior row of a graph) takes a
//     branch-free path: 3*U unclamped 16-byte loads, issued one group ahead into a ping-pong pair of register buffers
//     (no copies between groups), the window bounds of the group checked with two 16-byte loads, the arithmetic in
//     packed fp32x2 instructions (FFMA2 / FADD2 / FMUL2 halve the FP32 instruction count), scale vectors only in the
//     instantiations that have them (SI: backward, SO: forward).
//   * Every other row (strip start, the k rows either side of a graph boundary, array ends) goes through one small
//     rolled slow path that loads what it needs directly.
// The combined windows of all RESIDENT threads ((2k+1) rows x 16 B each) must stay in L2 for the leaving / own rows to be
// re-read from there instead of DRAM: strips are strided over a grid of kBandRunCtasPerSm CTAs per SM (148 x 3 x 128
// threads x 33 rows x 16 B = 30 MB at radius 16; with every strip resident it was 78 MB, L2 hit rate 2 %, every row read
// from DRAM three times).
// History (profiles/README.md): the first version of this kernel spent ~165 instructions per output vector on clamps,
// per-row window branches and scalar FP32 (ncu: 56 % issue slots, 3869 GB/s at radius 16); this one ~82 (39 % issue,
// 5066 GB/s, 1.03x algorithmic DRAM traffic) and is bound by load latency at 12 resident warps per SM.  A TMA bulk
// prefetch of the entering rows into L2 (cp.async.bulk.prefetch.L2, 0..16 groups ahead) measured 5-50 % SLOWER and was
// removed.
constexpr int kBandRunCtasPerSm = 3;
template <typename T, int U, bool SI, bool SO>
__global__ void __launch_bounds__(kAggThreads, kBandRunCtasPerSm)
sage_mean_band_run_kernel(const T* __restrict__ x, T* __restrict__ out, int n, int64_t channels, int64_t ldx,
                           int64_t ldo, int rows_per_cta, int k, const int32_t* __restrict__ win_lo,
                           const int32_t* __restrict__ win_hi, const float* __restrict__ scale_out,
                           const float* __restrict__ scale_in, bool win_vec) {
  static_assert(U == 4, "the window bounds of a group are fetched as one int4");
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  constexpr int NP = Pairs<T>::NP;
  const int64_t col = ((int64_t)blockIdx.y * kAggThreads + threadIdx.x) * VN;
  if (col >= channels) return;
  const T* xc = x + col;
  T* oc = out + col;
  struct Group {      // prefetched rows of one group: entering (i+k), leaving (i-k-1), self (i)
    Raw<T> en[U], lv[U], sf[U];
    float s_en[SI ? U : 1], s_lv[SI ? U : 1], s_sf[SI ? U : 1], s_out[SO ? U : 1];
    int4 lo, hi;
    bool ok;          // false: nothing was loaded (group touches an array end or starts a strip)
  };
  // g must be a multiple of U.  Loads are unclamped: only issued when every address is inside the arrays.
  auto load_group = [&](Group& q, int g, int r1) {
    q.ok = g + U <= r1 && g - k - 1 >= 0 && g + U - 1 + k < n;
    if (q.ok) {
      if (win_vec) {   // 16-byte aligned window arrays (always, for tensors the host side allocates)
        q.lo = *reinterpret_cast<const int4*>(win_lo + g);
        q.hi = *reinterpret_cast<const int4*>(win_hi + g);
      } else {
        q.lo = make_int4(win_lo[g], win_lo[g + 1], win_lo[g + 2], win_lo[g + 3]);
        q.hi = make_int4(win_hi[g], win_hi[g + 1], win_hi[g + 2], win_hi[g + 3]);
      }
      const T* pe = xc + (int64_t)(g + k) * ldx;
      const T* pl = xc + (int64_t)(g - k - 1) * ldx;
      const T* ps = xc + (int64_t)g * ldx;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        q.en[u] = Raw<T>::load(pe + (int64_t)u * ldx);
        q.lv[u] = Raw<T>::load(pl + (int64_t)u * ldx);
        q.sf[u] = Raw<T>::load(ps + (int64_t)u * ldx);
        if constexpr (SI) {
          q.s_en[u] = scale_in[g + k + u];
          q.s_lv[u] = -scale_in[g - k - 1 + u];
          q.s_sf[u] = -scale_in[g + u];
        }
        if constexpr (SO) q.s_out[u] = scale_out[g + u];
      }
    }
  };
  float2 acc[NP];
  auto clear = [&]() {
#pragma unroll
    for (int c = 0; c < NP; ++c) acc[c] = make_float2(0.f, 0.f);
  };
  auto add_row = [&](int j, float sign) {   // slow path: acc += sign * s_in(j) * x[j]
    const Pairs<T> v = Pairs<T>::from(Raw<T>::load(xc + (int64_t)j * ldx));
    const float s = SI ? sign * scale_in[j] : sign;
    const float2 s2 = make_float2(s, s);
#pragma unroll
    for (int c = 0; c < NP; ++c) acc[c] = __ffma2_rn(v.p[c], s2, acc[c]);
  };
  const float2 neg1 = make_float2(-1.f, -1.f);
  const int stride = (int)gridDim.x * rows_per_cta;
  int r0 = (int)blockIdx.x * rows_per_cta;
  int plo = 0, phi = -1;  // window currently summed in acc: [plo, phi] (empty)
  // one group of U rows: `cur` holds its prefetched rows, the next group's are fetched into `nxt` first
  auto step = [&](int g, int r1, Group& cur, Group& nxt) {
    if (g + U < r1) load_group(nxt, g + U, r1);
    else nxt.ok = false;
    bool steady = cur.ok && plo == g - k - 1 && phi == g + k - 1;
    if (steady) {
      const int a = g - k, b = g + k;
      steady = cur.lo.x == a && cur.lo.y == a + 1 && cur.lo.z == a + 2 && cur.lo.w == a + 3 &&
               cur.hi.x == b && cur.hi.y == b + 1 && cur.hi.z == b + 2 && cur.hi.w == b + 3;
    }
    if (steady) {
      T* po = oc + (int64_t)g * ldo;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const Pairs<T> e = Pairs<T>::from(cur.en[u]), l = Pairs<T>::from(cur.lv[u]), s = Pairs<T>::from(cur.sf[u]);
        Pairs<T> res;
        if constexpr (SI) {
          const float2 se = make_float2(cur.s_en[u], cur.s_en[u]), sl = make_float2(cur.s_lv[u], cur.s_lv[u]);
          const float2 ss = make_float2(cur.s_sf[u], cur.s_sf[u]);
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            acc[c] = __ffma2_rn(e.p[c], se, acc[c]);
            acc[c] = __ffma2_rn(l.p[c], sl, acc[c]);
            res.p[c] = __ffma2_rn(s.p[c], ss, acc[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < NP; ++c) {
            acc[c] = __fadd2_rn(acc[c], e.p[c]);
            acc[c] = __ffma2_rn(l.p[c], neg1, acc[c]);
            res.p[c] = __ffma2_rn(s.p[c], neg1, acc[c]);
          }
        }
        if constexpr (SO) {
          const float2 so = make_float2(cur.s_out[u], cur.s_out[u]);
#pragma unroll
          for (int c = 0; c < NP; ++c) res.p[c] = __fmul2_rn(res.p[c], so);
        }
        // streaming store: the output must not push the window rows out of L2
        __stcs(reinterpret_cast<uint4*>(po + (int64_t)u * ldo), res.pack());
      }
      plo += U;
      phi += U;
    } else {
      const int ge = min(g + U, r1);
#pragma unroll 1
      for (int i = g; i < ge; ++i) {
        const int lo = win_lo[i], hi = win_hi[i];
        if (lo > phi || phi < plo) {  // new graph / strip start: rebuild the window sum
          clear();
          for (int j = lo; j <= hi; ++j) add_row(j, 1.f);
        } else {
          for (int j = phi + 1; j <= hi; ++j) add_row(j, 1.f);
          for (int j = plo; j < lo; ++j) add_row(j, -1.f);
        }
        plo = lo;
        phi = hi;
        const Pairs<T> s = Pairs<T>::from(Raw<T>::load(xc + (int64_t)i * ldx));
        const float nss = SI ? -scale_in[i] : -1.f;
        const float so = SO ? scale_out[i] : 1.f;
        const float2 nss2 = make_float2(nss, nss), so2 = make_float2(so, so);
        Pairs<T> res;
#pragma unroll
        for (int c = 0; c < NP; ++c) res.p[c] = __fmul2_rn(__ffma2_rn(s.p[c], nss2, acc[c]), so2);
        __stcs(reinterpret_cast<uint4*>(oc + (int64_t)i * ldo), res.pack());
      }
    }
  };
  Group ga, gb;   // ping-pong: no register copies between groups
#pragma unroll 1
  for (; r0 < n; r0 += stride) {
    const int r1 = min(r0 + rows_per_cta, n);
    clear();
    plo = 0;
    phi = -1;
    ga.ok = false;          // a strip starts with a rebuild
#pragma unroll 1
    for (int g = r0; g < r1; g += 2 * U) {
      step(g, r1, ga, gb);
      if (g + U < r1) step(g + U, r1, gb, ga);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kAggThreads)
sage_mean_csr_kernel(const T* __restrict__ x, T* __restrict__ out, int64_t n, int64_t channels, int64_t ldx,
                     int64_t ldo, int rows_per_cta, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ colidx, const float* __restrict__ scale_out,
                     const float* __restrict__ scale_in) {
  pdl_enter();
  constexpr int VN = Vec<T>::N;
  const int64_t col = ((int64_t)blockIdx.y * kAggThreads + threadIdx.x) * VN;
  if (col >= channels) return;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, n);
  const T* xc = x + col;
  for (int64_t i = r0; i < r1; ++i) {
    const int b = rowptr[i], e = rowptr[i + 1];
    Vec<T> acc;
#pragma unroll
    for (int c = 0; c < VN; ++c) acc.v[c] = 0.f;
    int p = b;
    for (; p + 4 <= e; p += 4) {  // 4 independent 16-byte gathers in flight per thread
      int j[4];
      Vec<T> v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) j[u] = colidx[p + u];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = Vec<T>::load(xc + (int64_t)j[u] * ldx);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float s = scale_in ? scale_in[j[u]] : 1.f;
#pragma unroll
        for (int c = 0; c < VN; ++c) acc.v[c] += s * v[u].v[c];
      }
    }
    for (; p < e; ++p) {
      const int j = colidx[p];
      const Vec<T> v = Vec<T>::load(xc + (int64_t)j * ldx);
      const float s = scale_in ? scale_in[j] : 1.f;
#pragma unroll
      for (int c = 0; c < VN; ++c) acc.v[c] += s * v.v[c];
    }
    const float so = scale_out ? scale_out[i] : 1.f;
#pragma unroll
    for (int c = 0; c < VN; ++c) acc.v[c] *= so;
    acc.store(out + i * ldo + col);
  }
}

static int64_t band_reg_rows(int64_t n, int64_t gy, int U) {
  int64_t rows = (n * gy) / ((int64_t)sm_count() * 16);      // ~2 waves of 8 CTAs per SM
  rows = rows < 4 * U ? 4 * U : (rows > 512 ? 512 : rows);   // halo re-read <= 2K/(4U)
  return (rows + U - 1) / U * U;
}

struct StarArgs {                 // all null for a plain band
  const int32_t* ext_lo = nullptr;
  const int32_t* ext_hi = nullptr;
  const int32_t* hub_slot = nullptr;
  const int32_t* graph_meta = nullptr;
  int64_t num_graphs = 0;
  float* hub_part = nullptr;      // [strips + num_graphs, channels]
};

template <typename T, int K, int U>
static int launch_band_reg(const void* x, void* out, int64_t n, int64_t channels, int64_t ldx, int64_t ldo,
                           const int32_t* win_lo, const int32_t* win_hi, const float* scale_out, const float* scale_in,
                           cudaStream_t stream, const StarArgs& st = StarArgs()) {
  constexpr int VN = Vec<T>::N;
  const unsigned gy = (unsigned)ceil_div(channels, (int64_t)kAggThreads * VN);
  const int64_t rows = band_reg_rows(n, gy, U);
  dim3 grid((unsigned)ceil_div(n, rows), gy);
#define EGP_BAND_ARGS (const T*)x, (T*)out, (int)n, channels, ldx, ldo, (int)rows, win_lo, win_hi, scale_out, scale_in, \
                      st.ext_lo, st.ext_hi, st.hub_slot, st.hub_part
  if (st.hub_slot) {
    (void)launch_kernel(sage_mean_band_reg_kernel<T, K, U, kBandHub>, grid, kAggThreads, 0, stream, EGP_BAND_ARGS);
    EGP_LAUNCH_CHECK();
    (void)launch_kernel(sage_hub_fixup_kernel<T>, dim3((unsigned)st.num_graphs, gy), kAggThreads, 0, stream, (const T*)x, (T*)out,
                        channels, ldx, ldo, (int)rows, win_lo, win_hi, scale_out, scale_in, st.graph_meta,
                        (const float*)st.hub_part);
  } else if (st.ext_lo) {
    (void)launch_kernel(sage_mean_band_reg_kernel<T, K, U, kBandExt>, grid, kAggThreads, 0, stream, EGP_BAND_ARGS);
  } else {
    (void)launch_kernel(sage_mean_band_reg_kernel<T, K, U, kBandPlain>, grid, kAggThreads, 0, stream, EGP_BAND_ARGS);
  }
#undef EGP_BAND_ARGS
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

template <typename T>
static int launch_band_run(const void* x, void* out, int64_t n, int64_t channels, int64_t ldx, int64_t ldo, int k,
                           const int32_t* win_lo, const int32_t* win_hi, const float* scale_out, const float* scale_in,
                           cudaStream_t stream) {
  constexpr int VN = Vec<T>::N, U = 4;
  const unsigned gy = (unsigned)ceil_div(channels, (int64_t)kAggThreads * VN);
  static const int per_sm_env = [] { const char* e = getenv("EGP_BAND_RUN_CTAS"); return e ? atoi(e) : 0; }();
  const int per_sm = per_sm_env > 0 && per_sm_env < kBandRunCtasPerSm ? per_sm_env : kBandRunCtasPerSm;
  const int64_t cap = ceil_div((int64_t)sm_count() * per_sm, (int64_t)gy);   // resident CTAs along the row dimension
  // strips: as long as possible (every strip start rebuilds a 2k+1-row window sum) while every resident CTA gets the
  // same number of them -- one strip each unless that would exceed 2048 rows
  const int64_t min_rows = 2 * (2 * k + 1);                    // restart (2k+1 rows, L2 hits) <= half a strip's loads
  const int64_t waves = ceil_div(n, cap * 2048);
  int64_t rows = ceil_div(n, cap * waves);
  rows = rows < min_rows ? min_rows : rows;
  rows = (rows + U - 1) / U * U;
  const int64_t strips = ceil_div(n, rows);
  dim3 grid((unsigned)(strips < cap ? strips : cap), gy);
  auto run2 = [&](auto si, auto so) {
    (void)launch_kernel(sage_mean_band_run_kernel<T, U, decltype(si)::value, decltype(so)::value>, grid, kAggThreads, 0, stream,
                        (const T*)x, (T*)out, (int)n, channels, ldx, ldo, (int)rows, k, win_lo, win_hi, scale_out, scale_in,
                        aligned16(win_lo) && aligned16(win_hi));
  };
  using Yes = std::true_type;
  using No = std::false_type;
  if (scale_in && scale_out) run2(Yes{}, Yes{});
  else if (scale_in) run2(Yes{}, No{});
  else if (scale_out) run2(No{}, Yes{});
  else run2(No{}, No{});
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_sage_mean_band(const void* x, void* out, int64_t n, int64_t channels, int64_t ldx, int64_t ldo, int k,
                       const int32_t* win_lo, const int32_t* win_hi, const float* scale_out,
                       const float* scale_in, int dtype, void* stream) {
  EGP_REQUIRE(x && out && win_lo && win_hi, "sage_mean_band: null pointer");
  EGP_REQUIRE(k >= 0 && k <= 32, "sage_mean_band: radius %d out of range [0,32]", k);
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && ldx % vn == 0 && ldo % vn == 0 && aligned16(x) && aligned16(out),
              "sage_mean_band: channels/strides must keep rows 16-byte aligned");
  if (n == 0 || channels == 0) return EGP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  EGP_DISPATCH_DTYPE(dtype, T, {
    if (k <= 1) return launch_band_reg<T, 1, 8>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s);
    if (k == 2) return launch_band_reg<T, 2, 8>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s);
    if (k == 3) return launch_band_reg<T, 3, 4>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s);
    if (k == 4) return launch_band_reg<T, 4, 4>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s);
    return launch_band_run<T>(x, out, n, channels, ldx, ldo, k, win_lo, win_hi, scale_out, scale_in, s);
  });
  return EGP_OK;
}

size_t egp_sage_mean_band_star_workspace(int64_t n, int64_t channels, int64_t num_graphs) {
  // hub partials: one fp32 row per (strip, graph) pair; strips are at least 16 rows long
  return sizeof(float) * (size_t)(ceil_div(n, 16) + num_graphs + 1) * (size_t)channels + 128;
}

int egp_sage_mean_band_star(const void* x, void* out, int64_t n, int64_t channels, int64_t ldx, int64_t ldo, int k,
                            const int32_t* win_lo, const int32_t* win_hi, const float* scale_out,
                            const float* scale_in, const int32_t* ext_lo, const int32_t* ext_hi,
                            const int32_t* hub_slot, const int32_t* graph_meta, int64_t num_graphs, int dtype,
                            void* workspace, size_t ws_bytes, void* stream) {
  EGP_REQUIRE(x && out && win_lo && win_hi, "sage_mean_band_star: null pointer");
  EGP_REQUIRE(k >= 0 && k <= 4, "sage_mean_band_star: radius %d out of range [0,4] (wider stars use the CSR path)", k);
  EGP_REQUIRE((ext_lo && ext_hi && !hub_slot) || (hub_slot && graph_meta && !ext_lo && !ext_hi),
              "sage_mean_band_star: pass either the forward extension (ext_lo/ext_hi) or the backward hub (hub_slot/graph_meta)");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && ldx % vn == 0 && ldo % vn == 0 && aligned16(x) && aligned16(out),
              "sage_mean_band_star: channels/strides must keep rows 16-byte aligned");
  if (n == 0 || channels == 0) return EGP_OK;
  StarArgs st;
  st.ext_lo = ext_lo; st.ext_hi = ext_hi; st.hub_slot = hub_slot; st.graph_meta = graph_meta; st.num_graphs = num_graphs;
  if (hub_slot) {
    if (!workspace || ws_bytes < egp_sage_mean_band_star_workspace(n, channels, num_graphs)) {
      set_error("sage_mean_band_star: workspace too small");
      return EGP_ERR_WORKSPACE;
    }
    st.hub_part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 15) & ~(uintptr_t)15);
  }
  cudaStream_t s = (cudaStream_t)stream;
  EGP_DISPATCH_DTYPE(dtype, T, {
    if (k <= 1) return launch_band_reg<T, 1, 8>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s, st);
    if (k == 2) return launch_band_reg<T, 2, 8>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s, st);
    if (k == 3) return launch_band_reg<T, 3, 4>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s, st);
    return launch_band_reg<T, 4, 4>(x, out, n, channels, ldx, ldo, win_lo, win_hi, scale_out, scale_in, s, st);
  });
  return EGP_OK;
}

int egp_sage_mean_csr(const void* x, void* out, int64_t n, int64_t channels, int64_t ldx, int64_t ldo,
                      const int32_t* rowptr, const int32_t* col, const float* scale_out, const float* scale_in,
                      int dtype, void* stream) {
  EGP_REQUIRE(x && out && rowptr, "sage_mean_csr: null pointer");
  const int64_t vn = dtype == EGP_BF16 ? 8 : 4;
  EGP_REQUIRE(channels % vn == 0 && ldx % vn == 0 && ldo % vn == 0 && aligned16(x) && aligned16(out),
              "sage_mean_csr: channels/strides must keep rows 16-byte aligned");
  if (n == 0 || channels == 0) return EGP_OK;
  EGP_DISPATCH_DTYPE(dtype, T, {
    constexpr int VN = Vec<T>::N;
    const unsigned gy = (unsigned)ceil_div(channels, (int64_t)kAggThreads * VN);
    int64_t rows = (n * gy) / ((int64_t)sm_count() * 16);
    rows = rows < 4 ? 4 : (rows > 256 ? 256 : rows);
    dim3 grid((unsigned)ceil_div(n, rows), gy);
    (void)launch_kernel(sage_mean_csr_kernel<T>, grid, kAggThreads, 0, (cudaStream_t)stream, 
        (const T*)x, (T*)out, n, channels, ldx, ldo, (int)rows, rowptr, col, scale_out, scale_in);
    EGP_LAUNCH_CHECK();
  });
  return EGP_OK;
}

}  // extern "C"
