/* egopack_b200 -- C ABI of the B200 (sm_100a) kernels behind EgoPack's temporal-graph hot path.
 *
 * This is the drop-in boundary.  The reference (sapeirone/EgoPack) is pure Python and has
 * no FFI of its own: its hot path dispatches into torch / torch_geometric 2.3.0 / torch_cluster 1.6.1 /
 * torch_scatter 2.1.1 kernels.  Each entry point below replaces one of those dispatches; the reference call
 * site it serves is cited as (file:line, relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is DEVICE memory owned by the caller (the library never
 *     allocates, frees or retains device memory; `workspace` is caller-provided scratch).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no device synchronisation.
 *   - return value: EGP_OK (0) or a negative egp_status; the message is kept per thread (egp_last_error).
 *   - dtype codes: EGP_F32 = 0 (float), EGP_BF16 = 1 (__nv_bfloat16).  Index tensors are int64 where the
 *     reference hands int64 around (pos, batch, ptr, edge_index, knn indices); library-internal graph
 *     structure (window bounds, CSR) is int32.
 *   - row-major everywhere; `ld*` arguments are row strides in ELEMENTS.
 */
#ifndef EGOPACK_B200_H
#define EGOPACK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  EGP_OK = 0,
  EGP_ERR_INVALID = -1,     /* bad argument (null pointer, unsupported size/alignment) */
  EGP_ERR_CUDA = -2,        /* a CUDA runtime/driver call or a kernel launch failed */
  EGP_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve (e.g. no sm_100 device) */
  EGP_ERR_WORKSPACE = -4    /* workspace too small */
} egp_status;

enum { EGP_F32 = 0, EGP_BF16 = 1 };
enum { EGP_ACT_NONE = 0, EGP_ACT_RELU = 1, EGP_ACT_LEAKY_RELU = 2 };

/* ---- library ------------------------------------------------------------------------------------------ */
int egp_version(void);                               /* ABI version (2), bumped on any signature change */
int egp_last_error(char* buf, size_t len);           /* copies the calling thread's last error message */
int egp_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- a1: band edge_index (replaces torch_cluster.radius_graph behind RadiusGraph(r=k+0.5, loop=False);
 *      main_temporal.py:168-169,189-190,225-226; main_egopack.py:196-197,213-214,248-249) -------------------
 * Graph g owns nodes [ptr[g], ptr[g+1]).  For node i the matches are the nodes j of the same graph with
 * (pos_i-pos_j)^2 < r^2, visited in ascending j, truncated to the first `max_num_neighbors + 1` (self
 * included), self then removed.  `monotone` != 0 promises pos is non-decreasing inside every graph
 * (window scan); 0 scans the whole graph.
 *   count: deg[i]   = number of edges into i                              (int32 [N])
 *   fill : edge_index[0][e] = src j, edge_index[1][e] = dst i, dst-major, src ascending, e = rowptr[i]..
 *          (rowptr = exclusive scan of deg, int64 [N+1], see egp_exclusive_scan_i32)                       */
int egp_band_edge_count(const int64_t* pos, const int64_t* batch, const int64_t* ptr, int64_t num_nodes,
                        float r, int max_num_neighbors, int monotone, int32_t* deg, void* stream);
int egp_band_edge_fill(const int64_t* pos, const int64_t* batch, const int64_t* ptr, int64_t num_nodes,
                       float r, int max_num_neighbors, int monotone, const int64_t* rowptr,
                       int64_t num_edges, int64_t* edge_index, void* stream);
/* out[0]=0, out[i+1]=sum_{j<=i} in[j]; one block, workspace-free; n <= 2^31 */
int egp_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, void* stream);

/* ---- a2: LTA connectivity (models/transforms/lta_temp_connectivity.py:30-56), batched over graphs --------
 * Per graph: n_in = #(y[:,0]==-1), n_fc = #(y[:,0]>0); edges = band(r) U {s->t : s in [max(ceil(n_in-r),0), n_in),
 * t in [n_in, n_in+n_fc)}, de-duplicated, sorted by (src,dst) (RemoveDuplicatedEdges == coalesce).
 *   count: deg_out[s]; fill: edge_index sorted by (src,dst) using rowptr = scan(deg_out).
 * `y` is int64 [N, y_cols], column 0 = verb label.                                                         */
int egp_lta_edge_count(const int64_t* pos, const int64_t* y, int64_t y_cols, const int64_t* batch,
                       const int64_t* ptr, int64_t num_nodes, float r, int max_num_neighbors,
                       int32_t* deg_out, void* stream);
int egp_lta_edge_fill(const int64_t* pos, const int64_t* y, int64_t y_cols, const int64_t* batch,
                      const int64_t* ptr, int64_t num_nodes, float r, int max_num_neighbors,
                      const int64_t* rowptr, int64_t num_edges, int64_t* edge_index, void* stream);

/* ---- graph structure used by the aggregation kernels ------------------------------------------------------
 * band windows: win_lo[i], win_hi[i] (inclusive, self included) for unit-spaced pos and radius k, and
 * inv_deg[i] = 1/max(win_hi-win_lo, 1).                                                                      */
int egp_band_windows(const int64_t* batch, const int64_t* ptr, int64_t num_nodes, int k,
                     int32_t* win_lo, int32_t* win_hi, float* inv_deg, void* stream);
/* LTA star descriptor per graph (lta_temp_connectivity.py:48-53), no host round trip:
 * star[3g] = n_in = #(y[:,0]==-1), star[3g+1] = n_fc = #(y[:,0]>0), star[3g+2] = first_src = max(ceil(n_in-r),0).
 * The star edges are {s -> t : s in [first_src, n_in), t in [n_in, n_in+n_fc)} (graph-local indices).            */
int egp_lta_star_counts(const int64_t* y, int64_t y_cols, const int64_t* ptr, int64_t num_graphs, float r,
                        int32_t* star, void* stream);
/* band windows + the star extension of egp_sage_mean_band_star for one collated batch, written with GLOBAL row / graph
 * indices (row_offset / graph_offset added) so several task batches can share one structure (Graph.forward_many).
 *   win_lo/win_hi/inv_deg [n]  as egp_band_windows, inv_deg counting the star in-edges too
 *   ext_lo/ext_hi [n]          forward: for a star target, the star sources outside its band (a contiguous range)
 *   hub_slot [n]               backward: global graph index for star targets of a graph with >= 1 source, else -1
 *   graph_meta [num_graphs,4]  {src_lo, src_hi, tgt_lo, tgt_hi} absolute rows (caller-zeroed; empty without a star)
 * star == NULL (then the four star outputs may be NULL too) gives the plain band.                                  */
int egp_band_star_windows(const int64_t* batch, const int64_t* ptr, int64_t num_nodes, int64_t num_graphs, int k,
                          const int32_t* star, int64_t row_offset, int64_t graph_offset, int32_t* win_lo,
                          int32_t* win_hi, float* inv_deg, int32_t* ext_lo, int32_t* ext_hi, int32_t* hub_slot,
                          int32_t* graph_meta, void* stream);
/* CSR from an arbitrary int64 edge_index [2,E]: group_by = 1 groups by dst (rows = dst, cols = src; forward),
 * 0 groups by src (rows = src, cols = dst; backward).  Columns inside a row are sorted ascending
 * (deterministic).  rowptr int32 [N+1], col int32 [E]; `cursor` is int32 [N] scratch.                       */
int egp_csr_build(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, int group_by_dst,
                  int32_t* rowptr, int32_t* col, int32_t* cursor, void* stream);
/* inv_deg[i] = 1 / max(rowptr[i+1]-rowptr[i], 1) */
int egp_csr_inv_degree(const int32_t* rowptr, int64_t num_nodes, float* inv_deg, void* stream);

/* ---- a5: SAGE mean aggregation (replaces index_select + scatter_add_/count/div inside gnn.SAGEConv,
 *      models/graph.py:42) --------------------------------------------------------------------------------
 * band:  out[i] = s_out(i) * sum_{j in [win_lo[i],win_hi[i]], j != i} s_in(j) * x[j]
 *        forward : scale_out = inv_deg, scale_in = NULL      (mean over in-neighbours)
 *        backward: scale_out = NULL,    scale_in = inv_deg   (band is symmetric)
 * csr :  out[i] = s_out(i) * sum_{e in row i} s_in(col[e]) * x[col[e]]                                     */
int egp_sage_mean_band(const void* x, void* out, int64_t num_nodes, int64_t channels, int64_t ldx,
                       int64_t ldo, int k, const int32_t* win_lo, const int32_t* win_hi,
                       const float* scale_out, const float* scale_in, int dtype, void* stream);
/* band + star (LTATemporalConnectivity graphs, k <= 4) without an edge list and with the tensor read once:
 *   forward  (ext_lo/ext_hi given, hub_slot NULL): a star target also sums its out-of-band source rows;
 *   backward (hub_slot/graph_meta given, ext NULL): star targets are accumulated into per-graph hub sums inside the
 *            band pass (per-(strip, graph) partials in `workspace`, combined in a fixed order -> deterministic) and a
 *            second small kernel rewrites the star-source rows.  scale_in / scale_out as in egp_sage_mean_band.      */
size_t egp_sage_mean_band_star_workspace(int64_t num_nodes, int64_t channels, int64_t num_graphs);
int egp_sage_mean_band_star(const void* x, void* out, int64_t num_nodes, int64_t channels, int64_t ldx, int64_t ldo,
                            int k, const int32_t* win_lo, const int32_t* win_hi, const float* scale_out,
                            const float* scale_in, const int32_t* ext_lo, const int32_t* ext_hi,
                            const int32_t* hub_slot, const int32_t* graph_meta, int64_t num_graphs, int dtype,
                            void* workspace, size_t ws_bytes, void* stream);
int egp_sage_mean_csr(const void* x, void* out, int64_t num_nodes, int64_t channels, int64_t ldx,
                      int64_t ldo, const int32_t* rowptr, const int32_t* col, const float* scale_out,
                      const float* scale_in, int dtype, void* stream);

/* ---- a6+a7: graph-mode LayerNorm over the WHOLE [N,C] tensor + LeakyReLU (gnn.LayerNorm batch=None,
 *      models/graph.py:43-44): mu = mean(x), sigma = sqrt(mean((x-mu)^2)), y = act((x-mu)/(sigma+eps)*w+b) ---
 * stats: double[2] = {mu, sigma};  workspace: see egp_graph_layernorm_workspace.                           */
size_t egp_graph_layernorm_workspace(int64_t num_nodes, int64_t channels);
int egp_graph_layernorm_fwd(const void* x, const float* weight, const float* bias, void* y, double* stats,
                            int64_t num_nodes, int64_t channels, float eps, int act, float slope, int dtype,
                            void* workspace, size_t ws_bytes, void* stream);
/* dx, dweight[C], dbias[C] from dy (gradient w.r.t. the activated output), the saved input x and stats.
 * dx_colsum (optional, float[C]) receives the column sums of dx -- the bias gradient of the Linear that produced x,
 * computed in the same pass instead of by a separate egp_colsum. */
int egp_graph_layernorm_bwd(const void* dy, const void* x, const float* weight, const float* bias,
                            const double* stats, void* dx, float* dweight, float* dbias, float* dx_colsum,
                            int64_t num_nodes, int64_t channels, float eps, int act, float slope, int dtype,
                            void* workspace, size_t ws_bytes, void* stream);

/* Segmented form: rows are split into `nseg` (<= 8) consecutive segments [seg_rows[s], seg_rows[s+1]) -- HOST array of
 * nseg+1 offsets, seg_rows[0] = 0, seg_rows[nseg] = num_nodes -- each normalised with its OWN global statistics
 * (stats: double[2*nseg]); dweight / dbias / dx_colsum sum over all segments.  One segment per original forward call
 * when several task batches share the weights in one pass (Graph.forward_many): the reference normalises per call. */
size_t egp_graph_layernorm_seg_workspace(int64_t num_nodes, int64_t channels, int nseg);
int egp_graph_layernorm_seg_fwd(const void* x, const float* weight, const float* bias, void* y, double* stats,
                                int64_t num_nodes, int64_t channels, int nseg, const int64_t* seg_rows, float eps,
                                int act, float slope, int dtype, void* workspace, size_t ws_bytes, void* stream);
int egp_graph_layernorm_seg_bwd(const void* dy, const void* x, const float* weight, const float* bias,
                                const double* stats, void* dx, float* dweight, float* dbias, float* dx_colsum,
                                int64_t num_nodes, int64_t channels, int nseg, const int64_t* seg_rows, float eps,
                                int act, float slope, int dtype, void* workspace, size_t ws_bytes, void* stream);

/* Forward from the row-block statistics of the producing GEMM (egp_gemm_rowstats) instead of a stats pass over x:
 * a one-block-per-segment reduction of the pairs + the normalise pass (2*C*b bytes per node instead of 3*C*b).
 * Inner segment boundaries must be multiples of 128 rows; channels a multiple of 64. */
int egp_graph_layernorm_seg_fwd_rowstats(const void* x, const float* weight, const float* bias, void* y, double* stats,
                                         int64_t num_nodes, int64_t channels, int nseg, const int64_t* seg_rows,
                                         const double* rowstats, float eps, int act, float slope, int dtype,
                                         void* stream);

/* ---- row LayerNorm (+ReLU) (+Dropout) (nn.LayerNorm -> ReLU -> Dropout in TRNPooling trn_pooling.py:30-37; tasks
 *      task.py:20; GraphONE graphONE.py:61); mean/rstd float [N] are saved for the backward ---------------------
 * fwd: y = dropout_p(act(LN(x))); the keep decisions are a pure function of (seed, offset, element-vector index) -- Philox4x32-10 folds seed/offset into a key, a 32-bit multiply-xorshift hash draws the bits per vector -- so
 *      no mask is stored; rng_state (optional, device uint64[2] = {seed, step}) overrides the seed and adds step<<20 to
 *      the offset so that CUDA-graph replays draw fresh masks.  bwd: a zero in the saved output y means "ReLU inactive or dropped"; the incoming gradient
 *      is scaled by out_scale = 1/(1-p).  dx_colsum (optional) as in egp_graph_layernorm_bwd. */
size_t egp_row_layernorm_workspace(int64_t num_nodes, int64_t channels);
int egp_row_layernorm_fwd(const void* x, const float* weight, const float* bias, void* y, float* mean,
                          float* rstd, int64_t num_nodes, int64_t channels, float eps, int act, float dropout_p,
                          uint64_t seed, uint64_t offset, const uint64_t* rng_state, int dtype, void* stream);
int egp_row_layernorm_bwd(const void* dy, const void* x, const void* y, const float* weight, const float* mean,
                          const float* rstd, void* dx, float* dweight, float* dbias, float* dx_colsum,
                          int64_t num_nodes, int64_t channels, int act, float out_scale, int dtype, void* workspace,
                          size_t ws_bytes, void* stream);

/* ---- a4: out = x + [sin(pos*f) | cos(pos*f)]  (gnn.PositionalEncoding, models/graph.py:37,63) ---------- */
int egp_posenc_add(const void* x, const int64_t* pos, const float* frequency, void* out, int64_t num_nodes,
                   int64_t channels, int dtype, void* stream);

/* ---- elementwise helpers ------------------------------------------------------------------------------- */
int egp_cast(const void* src, void* dst, int64_t n, int src_dtype, int dst_dtype, void* stream);
/* dst[rows, ldd] = cast(src[rows, cols] with pitch lds), columns cols..ldd-1 zero-filled (16-byte pitch for TMA) */
int egp_cast_pad(const void* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols, int src_dtype,
                 int dst_dtype, void* stream);
/* fp32 -> bf16 split terms for the fp32-on-tensor-cores GEMM: x = x0 + x1 + x2 with x0 = bf16(x), x1 = bf16(x - x0),
 * x2 = bf16(x - x0 - x1) (24 significand bits).  src fp32 [rows, cols] (pitch lds); dst receives num_blocks (<= 6) bf16
 * blocks, block t = term term_of_block[t] (HOST array, values 0..2) of every element, element (r,c) at
 * dst[t*block_stride + r*ldd + c]; rows rows..rows_padded-1 and columns cols..cols_padded-1 of each block are zeroed.
 * Blocks side by side (block_stride = cols_padded, ldd = num_blocks*cols_padded) extend a K-major GEMM operand along K;
 * blocks stacked (block_stride = rows_padded*ldd) extend an MN-major one.  The fp32 Linear of the parity mode
 * (nn.Linear in fp32, models/graph.py:46 and every other Linear of the path) is ONE egp_gemm over such operands. */
int egp_split_bf16(const float* src, int64_t lds, int64_t rows, int64_t cols, int64_t rows_padded, int64_t cols_padded,
                   void* dst, int64_t block_stride, int64_t ldd, int num_blocks, const int* term_of_block, void* stream);
/* out = a + b (same dtype) */
int egp_add(const void* a, const void* b, void* out, int64_t n, int dtype, void* stream);
/* out = alpha*a + beta*b (b may be NULL: out = alpha*a) */
int egp_axpby(const void* a, float alpha, const void* b, float beta, void* out, int64_t n, int dtype, void* stream);
/* dx = dy * act'(y) with y the activated output (ReLU / LeakyReLU) */
int egp_act_bwd(const void* dy, const void* y, void* dx, int64_t n, int act, float slope, int dtype, void* stream);
/* act_bwd fused with the column sums of dx (bias gradient of the Linear whose epilogue applied the activation) */
size_t egp_act_bwd_colsum_workspace(int64_t rows, int64_t cols);
int egp_act_bwd_colsum(const void* dy, const void* y, void* dx, float* dx_colsum, int64_t rows, int64_t cols, int act,
                       float slope, int dtype, void* workspace, size_t ws_bytes, void* stream);
/* out[c] = sum_i x[i,c] (bias gradients); workspace float [blocks*C], see egp_colsum_workspace */
size_t egp_colsum_workspace(int64_t rows, int64_t cols);
int egp_colsum(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx, int dtype, void* workspace,
               size_t ws_bytes, void* stream);
/* dropout apply with a caller-generated keep mask (uint8): out = x * mask * scale */
int egp_mask_scale(const void* x, const uint8_t* mask, void* out, int64_t n, float scale, int dtype, void* stream);

/* ---- Linear layers (nn.Linear / gnn.Linear everywhere on the path) ---------------------------------------
 * C[M,N] = act( A[M,K] * B^T + A2[M,K2] * B2^T + bias[N] ) + residual[M,N]
 *   a_trans = 0: A is [M,K] row-major (lda);   a_trans = 1: A is [K,M] row-major (lda)      (same for A2)
 *   b_trans = 0: B is [N,K] row-major (ldb);   b_trans = 1: B is [K,N] row-major (ldb)      (same for B2)
 *   forward  y = x W^T      : A=x  (a_trans 0), B=W  (b_trans 0)
 *   dgrad    dx = dy W      : A=dy (a_trans 0), B=W  (b_trans 1)
 *   wgrad    dW = dy^T x    : A=dy (a_trans 1), B=x  (b_trans 1)
 * in_dtype EGP_BF16 -> tcgen05/TMEM tensor-core kernel fed by TMA (fp32 accumulate);
 * in_dtype EGP_F32  -> fp32 FFMA kernel (parity mode).  out_dtype may differ from in_dtype.
 * accumulate != 0: C += (fp32 C only; used by split-K wgrad).  bias/residual/A2/B2 may be NULL.            */
size_t egp_gemm_workspace(int64_t M, int64_t N, int64_t K);
/* Determinism switch.  Default (0): split-K weight gradients reduce-add their splits into C with TMA, whose arrival order
 * varies from run to run (last-bit differences).  1: the splits write slabs of `workspace` (egp_gemm_workspace bytes) and
 * are summed in split order by a second kernel -- bit-reproducible.  Everything else on the training path (aggregation,
 * normalisation statistics, column sums, hub sums, losses) is deterministic in both modes. */
int egp_set_deterministic(int on);
int egp_get_deterministic(void);
int egp_gemm(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans,
             const void* A2, int64_t lda2, const void* B2, int64_t ldb2, int64_t K2,
             const float* bias, const void* residual, int64_t ldr, void* C, int64_t ldc,
             int64_t M, int64_t N, int64_t K, int act, float slope, int in_dtype, int out_dtype,
             int accumulate, void* workspace, size_t ws_bytes, void* stream);

/* ---- a11: losses (fp32 logits) ---------------------------------------------------------------------------------
 * Cross entropy, reduction='none', with ignore_index and label smoothing -- nn.CrossEntropyLoss(ignore_index=-1,
 * reduction='none') per label head, summed over heads (models/tasks/recognition.py:21,61-69; lta.py:21,73-74;
 * criterion/wrapper.py:80-82; main_temporal.py:285,291) and F.cross_entropy(label_smoothing=0.1) for OSCC (oscc.py:88-96):
 *   loss[i] = (1-eps) * (lse_i - z[i,t_i]) + eps * (lse_i - mean_c z[i,c]);   0 when t_i == ignore_index (or outside
 *   [0, classes)).  labels int64, read at labels[i*label_stride] (a column of y [N,2]).
 *   fwd: accumulate != 0 ADDS to loss (the sum over heads); lse float [n] is saved for the backward.
 *   bwd: dlogits[i,c] = dloss[i*dloss_stride] * (softmax(z_i)[c] - (1-eps)[c==t_i] - eps/classes) (dloss_stride 0
 *        broadcasts a scalar); out_dtype EGP_F32 writes [n, ldd]; EGP_BF16 writes bf16 and zero-fills columns
 *        classes..ldd-1 (a 16-byte-pitch operand for the classifier's dgrad / wgrad GEMMs).                        */
int egp_ce_loss_fwd(const float* logits, int64_t ld, const int64_t* labels, int64_t label_stride, int64_t n,
                    int64_t classes, int64_t ignore_index, float label_smoothing, float* loss, int accumulate,
                    float* lse, void* stream);
int egp_ce_loss_bwd(const float* logits, int64_t ld, const float* lse, const int64_t* labels, int64_t label_stride,
                    const float* dloss, int64_t dloss_stride, int64_t n, int64_t classes, int64_t ignore_index,
                    float label_smoothing, void* dlogits, int64_t ldd, int out_dtype, void* stream);
/* nn.BCEWithLogitsLoss(reduction='none') (models/tasks/pnr.py:38,82-83): loss = max(z,0) - z*t + log1p(exp(-|z|));
 * dz = dloss * (sigmoid(z) - t).  z, target, loss float [n]. */
int egp_bce_logits_fwd(const float* z, const float* target, float* loss, int64_t n, void* stream);
int egp_bce_logits_bwd(const float* z, const float* target, const float* dloss, int64_t dloss_stride, float* dz,
                       int64_t n, void* stream);
/* out[0] (+)= weight * mean(x[0..n)) -- `w * loss.mean()` summed over tasks (main_temporal.py:99-128); one block, fixed
 * summation order (fp64 partials). */
int egp_weighted_mean(const float* x, int64_t n, float weight, float* out, int accumulate, void* stream);

/* ---- (f)-2: optimiser step (torch.optim.Adam(lr, weight_decay) of main_temporal.py:265-271 / main_egopack.py:317-324,
 *      stepped at main_temporal.py:130) over FLAT buffers, one pass -----------------------------------------------------
 * g' = g + weight_decay*p; m = b1 m + (1-b1) g'; v = b2 v + (1-b2) g'^2; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 *   p, m, v      float [total]: parameters / exp_avg / exp_avg_sq of all tensors back to back (starts padded to 8 elements)
 *   shadow       bf16 [total] or NULL: receives bf16(p) of every updated element (the bf16 GEMM operands)
 *   seg_off      int64 device [num_tensors+1]: element offset of every tensor in the flat buffers
 *   chunk_*      device work table, num_chunks entries: tensor index / flat start / length (<= 4096, inside one tensor)
 *   grads        HOST array of num_tensors DEVICE pointers (float, contiguous); NULL = tensor without gradient: skipped
 *   step         device int64[num_tensors]: completed steps PER TENSOR (torch keeps one counter per parameter); a tensor
 *                with a gradient uses t = step+1 for its bias corrections and has its counter incremented
 *   lr           device float[1] (LR schedulers and CUDA-graph replays share one code path)                           */
int egp_adam_step(float* p, float* m, float* v, void* shadow, const int64_t* seg_off, const int32_t* chunk_tensor,
                  const int64_t* chunk_start, const int32_t* chunk_len, int64_t num_chunks, const float* const* grads,
                  int num_tensors, int64_t* step, const float* lr, float beta1, float beta2, float eps,
                  float weight_decay, void* stream);

/* egp_gemm (bf16 operands, tensor-core path only) that ALSO leaves, per (128-row block, 32-row quarter, 256-column tile),
 * the {sum, sum of squares} of the values it stores -- computed in the epilogue from the fp32 accumulators, so a
 * whole-tensor statistic of C (graph-mode LayerNorm right after SAGEConv, models/graph.py:42-43) needs no extra pass.
 * rowstats: double [ceil(M/128)][4][ceil(N/256)][2] (egp_gemm_rowstats_bytes), zeroed by the call.  Served by the CTA-pair
 * kernel only: K-major operands, N % 64 == 0, enough rows for 256-wide tiles, 16-byte aligned and pitched C;
 * EGP_ERR_UNSUPPORTED otherwise (callers fall back to egp_gemm + a statistics pass). */
size_t egp_gemm_rowstats_bytes(int64_t M, int64_t N);
int egp_gemm_rowstats(const void* A, int64_t lda, int a_trans, const void* B, int64_t ldb, int b_trans,
                      const void* A2, int64_t lda2, const void* B2, int64_t ldb2, int64_t K2,
                      const float* bias, const void* residual, int64_t ldr, void* C, int64_t ldc,
                      int64_t M, int64_t N, int64_t K, int act, float slope, int out_dtype, double* rowstats,
                      void* stream);

/* ---- a14: cosine k-NN of nodes against a prototype bank (GraphONE.__compute_edges, graphONE.py:119-141) ---
 * d = 1 - (F/|F|)(P/|P|)^T ; idx[i,:] = the k smallest d, ascending, ties -> lower prototype index.
 *   egp_row_normalize: out = x / ||x||_2 per row (fp32 math, no epsilon -- cos_dissimilarity, graphONE.py:148-151);
 *                      round_err (optional, float [rows]) receives |stored row - exact normalised row|_2, i.e. the
 *                      rounding error of a bf16 output row (the miss detector's per-row error bound).
 *   egp_cos_topk     : fn/pn = fp32 NORMALISED rows [B,C] / [Kp,C].  With fn16/pn16 == NULL the similarity is an
 *                      fp32 GEMM and the selection is exact.  With bf16 normalised copies the similarity runs on the
 *                      tensor cores with the best candidates kept in the GEMM epilogue, and the candidates are
 *                      re-scored exactly in fp32 from fn/pn.
 *                      Miss detector (flagged_rows/flagged_count != NULL): a row whose k-th exact candidate score does
 *                      not clear [largest bf16 score outside the candidate set] + f_round_err[row] + p_round_err (+
 *                      their product + fp32 slack) -- i.e. a row where bf16 rounding COULD have hidden a true
 *                      neighbour -- is appended to flagged_rows (int32 [B], unordered) and counted in flagged_count
 *                      (int32 [1], zeroed by the call); the caller re-runs those rows with fn16 == NULL.
 *                      f_round_err NULL assumes the worst case 2^-8 per row; p_round_err = max over prototypes.
 *   workspace        : egp_cos_topk_workspace(B, Kp, k) bytes (row-chunked [<=32768, Kp] fp32 similarities).    */
int egp_row_normalize(const void* x, void* out, int64_t rows, int64_t cols, int in_dtype, int out_dtype,
                      float* round_err, void* stream);
int egp_row_inv_norm(const void* x, float* out, int64_t rows, int64_t cols, int64_t ldx, int dtype, void* stream);
size_t egp_cos_topk_workspace(int64_t num_nodes, int64_t num_protos, int64_t k);
int egp_cos_topk(const float* fn, const float* pn, const void* fn16, const void* pn16, int64_t num_nodes,
                 int64_t num_protos, int64_t channels, int k, int64_t* idx, const float* f_round_err,
                 float p_round_err, int32_t* flagged_rows, int32_t* flagged_count, void* workspace, size_t ws_bytes,
                 void* stream);

/* ---- a15: prototype max-gather and max-combine (reduced GraphONE stage) ------------
 * m[i,c] = max_j P[idx[i,j], c];  a = max(f, m);  backward of the combine: df = da * (f >= m).              */
int egp_proto_max_gather(const void* protos, const int64_t* idx, void* m, int64_t num_nodes, int64_t k,
                         int64_t channels, int proto_dtype, int out_dtype, void* stream);
int egp_max_combine_fwd(const void* f, const void* m, void* a, int64_t n, int dtype, void* stream);
int egp_max_combine_bwd(const void* da, const void* f, const void* m, void* df, int64_t n, int dtype, void* stream);

/* gradient of the combine w.r.t. a TRAINABLE bank (GraphONE(freeze=False), graphONE.py:47-49): where the prototype
 * maximum beats f, da[i,c] is added (fp32 atomics) to dbank[idx[i,j*],c], j* = first of the k that attains the max. */
int egp_proto_max_scatter_bwd(const void* da, const void* f, const void* protos, const int64_t* idx, float* dbank,
                              int64_t num_nodes, int64_t k, int64_t channels, int dtype, void* stream);

/* ---- (f)-1: prototype-bank builder (graphone.py:16-63): out[label[i], :] += x[i, :] accumulated in fp64;
 *      rows with a label outside [0, num_classes) are skipped.  out is double [num_classes, C], caller-zeroed. */
int egp_class_sum_f64(const void* x, const int64_t* label, double* out, int64_t rows, int64_t channels,
                      int64_t num_classes, int dtype, void* stream);

/* ---- a12: per-graph channel-wise max pooling (gnn.pool.global_max_pool, models/tasks/oscc.py:68,85) ------
 * out[g,c] = max_{i in [ptr[g],ptr[g+1])} x[i,c] (0 for an empty graph); arg int32 [G,C] (-1 if empty).     */
int egp_segment_max_pool_fwd(const void* x, const int64_t* ptr, void* out, int32_t* arg, int64_t num_graphs,
                             int64_t channels, int dtype, void* stream);
int egp_segment_max_pool_bwd(const void* dout, const int32_t* arg, const int64_t* batch, void* dx,
                             int64_t num_nodes, int64_t channels, int dtype, void* stream);

/* ---- a-M / (f)-4: headline metrics (utils/meters/ego4d.py) -- integer results, bit-exact ----------------------
 * The reference computes these on the host through torchmetrics 1.0.1 / editdistance 0.6.2 (environment.yml:308,247).
 *
 * egp_label_rank: rank[i] = #{c : logits[i,c] > logits[i,t] or (== and c < t)}, t = labels[i*label_stride]; -1 when
 * t == ignore_index or t outside [0, classes).  top-k hit <=> 0 <= rank < k; k = 1 is the first-arg-max rule of
 * MulticlassAccuracy(top_k=1) (utils/meters/ego4d.py:46-49,60-63,93-96,107-110,306,432-433).  logits fp32 [n, ld]. */
int egp_label_rank(const float* logits, int64_t ld, const int64_t* labels, int64_t label_stride, int64_t n,
                   int64_t classes, int64_t ignore_index, int32_t* rank, void* stream);
/* PNR key-frame localisation (utils/meters/ego4d.py:356-358): out[g] = first arg-max, relative to ptr[g], of
 * sigmoid(values) (apply_sigmoid != 0; fp32 1/(1+exp(-v))) or of values over nodes [ptr[g], ptr[g+1]); -1 if empty. */
int egp_segment_argmax(const float* values, const int64_t* ptr, int64_t num_graphs, int apply_sigmoid, int64_t* out,
                       void* stream);
/* LTA edit distance (utils/meters/ego4d.py:410-422): out[i] = min_k Levenshtein(preds[i,:,k], labels[i,:]) with
 * editdistance.eval's unit costs; preds int64 [n, seq_len, num_samples], labels int64 [n, seq_len], seq_len <= 64. */
int egp_edit_distance_min(const int64_t* preds, const int64_t* labels, int64_t n, int64_t seq_len, int64_t num_samples,
                          int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EGOPACK_B200_H */
