// Temporal edge construction and graph-structure helpers (integer work, HBM/latency bound).
//
//  band edges   : replaces torch_cluster.radius_graph behind RadiusGraph(r=k+0.5)  (main_temporal.py:168)
//  LTA edges    : models/transforms/lta_temp_connectivity.py:30-56
//  band windows : per-node [lo,hi] neighbour window + 1/deg for the sliding-window aggregation kernel
//  CSR build    : arbitrary edge_index -> deterministic CSR (rows sorted) for the generic aggregation kernel
#include "common.cuh"

namespace egp {

// ---------------------------------------------------------------------------------------------------------
// per-node match enumeration shared by count and fill.  Calls emit(j) for each kept neighbour, ascending j.
// torch_cluster keeps at most `cap` matches per centre INCLUDING the self match, then drops the self loop.
// ---------------------------------------------------------------------------------------------------------
template <typename Emit>
__device__ __forceinline__ int enumerate_band(const int64_t* __restrict__ pos, int64_t i, int64_t lo, int64_t hi,
                                              float r2, int cap, bool monotone, Emit emit) {
  const int64_t pi = pos[i];
  int kept = 0, out = 0;
  if (monotone) {
    int64_t jl = i;
    while (jl - 1 >= lo) {
      const float d = (float)(pi - pos[jl - 1]);
      if (d * d < r2) --jl; else break;
    }
    for (int64_t j = jl; j < hi && kept < cap; ++j) {
      const float d = (float)(pi - pos[j]);
      if (!(d * d < r2)) break;
      ++kept;
      if (j != i) { emit(j); ++out; }
    }
  } else {
    for (int64_t j = lo; j < hi && kept < cap; ++j) {
      const float d = (float)(pi - pos[j]);
      if (d * d < r2) {
        ++kept;
        if (j != i) { emit(j); ++out; }
      }
    }
  }
  return out;
}

__global__ void band_edge_count_kernel(const int64_t* __restrict__ pos, const int64_t* __restrict__ batch,
                                       const int64_t* __restrict__ ptr, int64_t n, float r2, int cap,
                                       int monotone, int32_t* __restrict__ deg) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = batch[i];
  deg[i] = enumerate_band(pos, i, ptr[g], ptr[g + 1], r2, cap, monotone != 0, [](int64_t) {});
}

__global__ void band_edge_fill_kernel(const int64_t* __restrict__ pos, const int64_t* __restrict__ batch,
                                      const int64_t* __restrict__ ptr, int64_t n, float r2, int cap,
                                      int monotone, const int64_t* __restrict__ rowptr, int64_t num_edges,
                                      int64_t* __restrict__ edge_index) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = batch[i];
  int64_t e = rowptr[i];
  int64_t* src = edge_index;
  int64_t* dst = edge_index + num_edges;
  enumerate_band(pos, i, ptr[g], ptr[g + 1], r2, cap, monotone != 0, [&](int64_t j) {
    src[e] = j;
    dst[e] = i;
    ++e;
  });
}

// ---------------------------------------------------------------------------------------------------------
// LTA connectivity: out-edges of node s = band(s) U star(s), ascending target, de-duplicated.
// ---------------------------------------------------------------------------------------------------------
template <typename Emit>
__device__ __forceinline__ int enumerate_lta(const int64_t* __restrict__ pos, const int64_t* __restrict__ y,
                                             int64_t y_cols, int64_t s, int64_t lo, int64_t hi, float r,
                                             Emit emit) {
  // per-graph counts (lta_temp_connectivity.py:48,50): inputs carry verb label -1, forecasts verb label > 0
  int64_t n_in = 0, n_fc = 0;
  for (int64_t t = lo; t < hi; ++t) {
    const int64_t v = y[t * y_cols];
    n_in += (v == -1);
    n_fc += (v > 0);
  }
  const float r2 = r * r;
  int64_t first_src = (int64_t)ceilf((float)n_in - r);
  if (first_src < 0) first_src = 0;
  const int64_t sl = s - lo;
  const bool star = (sl >= first_src) && (sl < n_in);
  const int64_t ps = pos[s];
  int out = 0;
  for (int64_t t = lo; t < hi; ++t) {
    const float d = (float)(ps - pos[t]);
    const bool band = (t != s) && (d * d < r2);
    const int64_t tl = t - lo;
    const bool st = star && (tl >= n_in) && (tl < n_in + n_fc);
    if (band || st) { emit(t); ++out; }
  }
  return out;
}

__global__ void lta_edge_count_kernel(const int64_t* __restrict__ pos, const int64_t* __restrict__ y,
                                      int64_t y_cols, const int64_t* __restrict__ batch,
                                      const int64_t* __restrict__ ptr, int64_t n, float r,
                                      int32_t* __restrict__ deg) {
  pdl_enter();
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int64_t g = batch[s];
  deg[s] = enumerate_lta(pos, y, y_cols, s, ptr[g], ptr[g + 1], r, [](int64_t) {});
}

__global__ void lta_edge_fill_kernel(const int64_t* __restrict__ pos, const int64_t* __restrict__ y,
                                     int64_t y_cols, const int64_t* __restrict__ batch,
                                     const int64_t* __restrict__ ptr, int64_t n, float r,
                                     const int64_t* __restrict__ rowptr, int64_t num_edges,
                                     int64_t* __restrict__ edge_index) {
  pdl_enter();
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const int64_t g = batch[s];
  int64_t e = rowptr[s];
  int64_t* src = edge_index;
  int64_t* dst = edge_index + num_edges;
  enumerate_lta(pos, y, y_cols, s, ptr[g], ptr[g + 1], r, [&](int64_t t) {
    src[e] = s;
    dst[e] = t;
    ++e;
  });
}

// ---------------------------------------------------------------------------------------------------------
// single-block exclusive scan (n up to a few million: data-pipeline op, not on the training critical path)
// ---------------------------------------------------------------------------------------------------------
template <typename Out>
__global__ void exclusive_scan_kernel(const int32_t* __restrict__ in, int64_t n, Out* __restrict__ out) {
  pdl_enter();
  __shared__ long long warp_tot[32];
  __shared__ long long carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) { carry_s = 0; out[0] = 0; }
  __syncthreads();
  for (int64_t base = 0; base < n; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    long long v = (i < n) ? (long long)in[i] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long w = (lane < nw) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += t;
      }
      warp_tot[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const long long carry = carry_s;
    const long long prefix = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + incl;
    if (i < n) out[i + 1] = (Out)prefix;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + warp_tot[nw - 1];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// band windows
// ---------------------------------------------------------------------------------------------------------
__global__ void band_windows_kernel(const int64_t* __restrict__ batch, const int64_t* __restrict__ ptr,
                                    int64_t n, int k, int32_t* __restrict__ win_lo,
                                    int32_t* __restrict__ win_hi, float* __restrict__ inv_deg) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = batch[i];
  const int64_t lo = max(i - k, ptr[g]);
  const int64_t hi = min(i + k, ptr[g + 1] - 1);
  win_lo[i] = (int32_t)lo;
  win_hi[i] = (int32_t)hi;
  const int d = (int)(hi - lo);
  inv_deg[i] = 1.0f / (float)(d > 1 ? d : 1);
}


// ---------------------------------------------------------------------------------------------------------
// band + star structure (LTA graphs without materialising edge_index)
// ---------------------------------------------------------------------------------------------------------
// one warp per graph: n_in = #(y[:,0] == -1), n_fc = #(y[:,0] > 0), first_src = max(ceil(n_in - r), 0)
// (lta_temp_connectivity.py:48-52; the `> 0` -- verb label 0 is not counted -- is the reference's)
__global__ void lta_star_counts_kernel(const int64_t* __restrict__ y, int64_t y_cols, const int64_t* __restrict__ ptr,
                                       int64_t num_graphs, float r, int32_t* __restrict__ star) {
  pdl_enter();
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (g >= num_graphs) return;
  const int64_t lo = ptr[g], hi = ptr[g + 1];
  int n_in = 0, n_fc = 0;
  for (int64_t t = lo + lane; t < hi; t += 32) {
    const int64_t v = y[t * y_cols];
    n_in += (v == -1);
    n_fc += (v > 0);
  }
  n_in = warp_sum(n_in);
  n_fc = warp_sum(n_fc);
  if (lane == 0) {
    int first = (int)ceilf((float)n_in - r);
    if (first < 0) first = 0;
    star[3 * g + 0] = n_in;
    star[3 * g + 1] = n_fc;
    star[3 * g + 2] = first;
  }
}

// Per node: band window [lo,hi] clipped to the graph, the forward extension (for a star TARGET t: the star sources that
// are not already inside its band, a contiguous range [ext_lo, ext_hi)), the backward hub slot (for a star target of a
// graph that has at least one source: the graph's global index; -1 otherwise) and 1/max(in-degree, 1).
// Per graph: graph_meta = {src_lo, src_hi, tgt_lo, tgt_hi} (absolute rows; empty ranges without a star).
// All row / graph indices written are GLOBAL (row_offset / graph_offset added) so that several task batches can be
// laid out back to back in one structure (Graph.forward_many).
__global__ void band_star_windows_kernel(const int64_t* __restrict__ batch, const int64_t* __restrict__ ptr, int64_t n,
                                         int k, const int32_t* __restrict__ star, int64_t row_offset,
                                         int64_t graph_offset, int32_t* __restrict__ win_lo,
                                         int32_t* __restrict__ win_hi, float* __restrict__ inv_deg,
                                         int32_t* __restrict__ ext_lo, int32_t* __restrict__ ext_hi,
                                         int32_t* __restrict__ hub_slot, int32_t* __restrict__ graph_meta) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t g = batch[i];
  const int64_t g_lo = ptr[g], g_hi = ptr[g + 1];
  const int64_t lo = max(i - k, g_lo);
  const int64_t hi = min(i + k, g_hi - 1);
  int deg = (int)(hi - lo);
  int64_t src_lo = 0, src_hi = 0, tgt_lo = 0, tgt_hi = 0;
  if (star) {
    const int64_t n_in = star[3 * g], n_fc = star[3 * g + 1], first = star[3 * g + 2];
    if (n_fc > 0 && n_in > first) {
      src_lo = g_lo + first;
      src_hi = g_lo + n_in;
      tgt_lo = g_lo + n_in;
      tgt_hi = min(g_lo + n_in + n_fc, g_hi);
    }
  }
  int64_t el = 0, eh = 0;
  int slot = -1;
  if (i >= tgt_lo && i < tgt_hi) {          // star target: sources outside its band
    el = src_lo;
    eh = min(src_hi, i - k);
    if (eh < el) eh = el;
    deg += (int)(eh - el);
    slot = (int)(g + graph_offset);
  }
  win_lo[i] = (int32_t)(lo + row_offset);
  win_hi[i] = (int32_t)(hi + row_offset);
  inv_deg[i] = 1.0f / (float)(deg > 1 ? deg : 1);
  if (ext_lo) {
    ext_lo[i] = (int32_t)(el + row_offset);
    ext_hi[i] = (int32_t)(eh + row_offset);
    hub_slot[i] = slot;
    if (i == g_lo) {
      int32_t* m = graph_meta + 4 * g;
      m[0] = (int32_t)(src_lo + row_offset);
      m[1] = (int32_t)(src_hi + row_offset);
      m[2] = (int32_t)(tgt_lo + row_offset);
      m[3] = (int32_t)(tgt_hi + row_offset);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// CSR build
// ---------------------------------------------------------------------------------------------------------
// Edges whose endpoints fall outside [0, n) are dropped (count and fill apply the same test), so a malformed
// edge_index can never index the CSR arrays out of bounds.
__global__ void csr_count_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ val, int64_t e_count,
                                 int64_t n, int32_t* __restrict__ deg) {
  pdl_enter();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e_count) return;
  const int64_t r = key[e], c = val[e];
  if (r < 0 || r >= n || c < 0 || c >= n) return;
  atomicAdd(&deg[r], 1);
}

__global__ void csr_fill_kernel(const int64_t* __restrict__ key, const int64_t* __restrict__ val, int64_t e_count,
                                int64_t n, const int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ col) {
  pdl_enter();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= e_count) return;
  const int64_t r = key[e], c = val[e];
  if (r < 0 || r >= n || c < 0 || c >= n) return;
  const int slot = atomicAdd(&cursor[r], 1);
  col[rowptr[r] + slot] = (int32_t)c;
}

__global__ void csr_sort_rows_kernel(const int32_t* __restrict__ rowptr, int64_t n, int32_t* __restrict__ col) {
  pdl_enter();
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int b = rowptr[r], e = rowptr[r + 1];
  for (int i = b + 1; i < e; ++i) {  // insertion sort: rows are short (temporal graphs)
    const int32_t v = col[i];
    int j = i - 1;
    while (j >= b && col[j] > v) { col[j + 1] = col[j]; --j; }
    col[j + 1] = v;
  }
}

__global__ void csr_inv_degree_kernel(const int32_t* __restrict__ rowptr, int64_t n, float* __restrict__ inv_deg) {
  pdl_enter();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int d = rowptr[i + 1] - rowptr[i];
  inv_deg[i] = 1.0f / (float)(d > 1 ? d : 1);
}

}  // namespace egp

using namespace egp;

extern "C" {

int egp_band_edge_count(const int64_t* pos, const int64_t* batch, const int64_t* ptr, int64_t n, float r,
                        int max_num_neighbors, int monotone, int32_t* deg, void* stream) {
  EGP_REQUIRE(pos && batch && ptr && deg, "band_edge_count: null pointer");
  EGP_REQUIRE(n >= 0 && r > 0.f && max_num_neighbors > 0, "band_edge_count: bad size/radius");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(band_edge_count_kernel, (unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream, 
      pos, batch, ptr, n, r * r, max_num_neighbors + 1, monotone, deg);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_band_edge_fill(const int64_t* pos, const int64_t* batch, const int64_t* ptr, int64_t n, float r,
                       int max_num_neighbors, int monotone, const int64_t* rowptr, int64_t num_edges,
                       int64_t* edge_index, void* stream) {
  EGP_REQUIRE(pos && batch && ptr && rowptr, "band_edge_fill: null pointer");
  EGP_REQUIRE(edge_index || num_edges == 0, "band_edge_fill: null edge_index");
  if (n == 0 || num_edges == 0) return EGP_OK;
  (void)launch_kernel(band_edge_fill_kernel, (unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream, 
      pos, batch, ptr, n, r * r, max_num_neighbors + 1, monotone, rowptr, num_edges, edge_index);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, void* stream) {
  EGP_REQUIRE(out && (in || n == 0), "exclusive_scan: null pointer");
  (void)launch_kernel(exclusive_scan_kernel<int64_t>, 1, 1024, 0, (cudaStream_t)stream, in, n, out);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_lta_edge_count(const int64_t* pos, const int64_t* y, int64_t y_cols, const int64_t* batch,
                       const int64_t* ptr, int64_t n, float r, int max_num_neighbors, int32_t* deg,
                       void* stream) {
  EGP_REQUIRE(pos && y && batch && ptr && deg, "lta_edge_count: null pointer");
  EGP_REQUIRE(y_cols >= 1 && r > 0.f, "lta_edge_count: bad arguments");
  if (2 * (int)floorf(r) + 1 > max_num_neighbors + 1) {
    set_error("lta_edge_count: radius %.2f exceeds the neighbour cap %d (truncated bands are not symmetric)",
              r, max_num_neighbors);
    return EGP_ERR_UNSUPPORTED;
  }
  if (n == 0) return EGP_OK;
  (void)launch_kernel(lta_edge_count_kernel, (unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream, pos, y, y_cols, batch,
                                                                                      ptr, n, r, deg);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_lta_edge_fill(const int64_t* pos, const int64_t* y, int64_t y_cols, const int64_t* batch,
                      const int64_t* ptr, int64_t n, float r, int max_num_neighbors, const int64_t* rowptr,
                      int64_t num_edges, int64_t* edge_index, void* stream) {
  (void)max_num_neighbors;
  EGP_REQUIRE(pos && y && batch && ptr && rowptr, "lta_edge_fill: null pointer");
  EGP_REQUIRE(edge_index || num_edges == 0, "lta_edge_fill: null edge_index");
  if (n == 0 || num_edges == 0) return EGP_OK;
  (void)launch_kernel(lta_edge_fill_kernel, (unsigned)ceil_div(n, 128), 128, 0, (cudaStream_t)stream, 
      pos, y, y_cols, batch, ptr, n, r, rowptr, num_edges, edge_index);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_band_windows(const int64_t* batch, const int64_t* ptr, int64_t n, int k, int32_t* win_lo,
                     int32_t* win_hi, float* inv_deg, void* stream) {
  EGP_REQUIRE(batch && ptr && win_lo && win_hi && inv_deg, "band_windows: null pointer");
  EGP_REQUIRE(k >= 0 && n < (int64_t)INT32_MAX, "band_windows: bad k or too many nodes");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(band_windows_kernel, (unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream, batch, ptr, n, k, win_lo,
                                                                                    win_hi, inv_deg);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_lta_star_counts(const int64_t* y, int64_t y_cols, const int64_t* ptr, int64_t num_graphs, float r,
                        int32_t* star, void* stream) {
  EGP_REQUIRE(y && ptr && star, "lta_star_counts: null pointer");
  EGP_REQUIRE(y_cols >= 1 && r > 0.f, "lta_star_counts: bad arguments");
  if (num_graphs == 0) return EGP_OK;
  (void)launch_kernel(lta_star_counts_kernel, (unsigned)ceil_div(num_graphs * 32, 128), 128, 0, (cudaStream_t)stream, y, y_cols, ptr,
                      num_graphs, r, star);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_band_star_windows(const int64_t* batch, const int64_t* ptr, int64_t n, int64_t num_graphs, int k,
                          const int32_t* star, int64_t row_offset, int64_t graph_offset, int32_t* win_lo,
                          int32_t* win_hi, float* inv_deg, int32_t* ext_lo, int32_t* ext_hi, int32_t* hub_slot,
                          int32_t* graph_meta, void* stream) {
  EGP_REQUIRE(batch && ptr && win_lo && win_hi && inv_deg, "band_star_windows: null pointer");
  EGP_REQUIRE((ext_lo && ext_hi && hub_slot && graph_meta) || (!ext_lo && !ext_hi && !hub_slot && !graph_meta && !star),
              "band_star_windows: the star outputs come together (and are needed when a star descriptor is given)");
  EGP_REQUIRE(k >= 0 && n + row_offset < (int64_t)INT32_MAX && num_graphs + graph_offset < (int64_t)INT32_MAX,
              "band_star_windows: bad k or too many nodes");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(band_star_windows_kernel, (unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream, batch, ptr, n, k, star,
                      row_offset, graph_offset, win_lo, win_hi, inv_deg, ext_lo, ext_hi, hub_slot, graph_meta);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

int egp_csr_build(const int64_t* edge_index, int64_t num_edges, int64_t n, int group_by_dst, int32_t* rowptr,
                  int32_t* col, int32_t* cursor, void* stream) {
  EGP_REQUIRE(rowptr && cursor && (col || num_edges == 0) && (edge_index || num_edges == 0),
              "csr_build: null pointer");
  EGP_REQUIRE(n < (int64_t)INT32_MAX && num_edges < (int64_t)INT32_MAX, "csr_build: graph too large for int32");
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t* key = group_by_dst ? edge_index + num_edges : edge_index;
  const int64_t* val = group_by_dst ? edge_index : edge_index + num_edges;
  if (n > 0) EGP_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n, s));
  if (num_edges > 0) {
    (void)launch_kernel(csr_count_kernel, (unsigned)ceil_div(num_edges, 256), 256, 0, s, key, val, num_edges, n, cursor);
    EGP_LAUNCH_CHECK();
  }
  (void)launch_kernel(exclusive_scan_kernel<int32_t>, 1, 1024, 0, s, cursor, n, rowptr);
  EGP_LAUNCH_CHECK();
  if (num_edges > 0) {
    EGP_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n, s));
    (void)launch_kernel(csr_fill_kernel, (unsigned)ceil_div(num_edges, 256), 256, 0, s, key, val, num_edges, n, rowptr, cursor, col);
    EGP_LAUNCH_CHECK();
    (void)launch_kernel(csr_sort_rows_kernel, (unsigned)ceil_div(n, 128), 128, 0, s, rowptr, n, col);
    EGP_LAUNCH_CHECK();
  }
  return EGP_OK;
}

int egp_csr_inv_degree(const int32_t* rowptr, int64_t n, float* inv_deg, void* stream) {
  EGP_REQUIRE(rowptr && inv_deg, "csr_inv_degree: null pointer");
  if (n == 0) return EGP_OK;
  (void)launch_kernel(csr_inv_degree_kernel, (unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream, rowptr, n, inv_deg);
  EGP_LAUNCH_CHECK();
  return EGP_OK;
}

}  // extern "C"
