"""Host-side operators: thin wrappers over the C ABI plus the ``torch.autograd.Function``s the modules use.

PyTorch supplies device memory, streams and autograd bookkeeping; every arithmetic step is a kernel of
``libegopack_b200.so``.  Nothing here runs on the CPU and nothing falls back to aten math.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L

ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
F32, BF16 = 0, 1


# EGP_ROWSTATS=0 disables the GEMM-epilogue statistics (A/B switch: graph-LN then runs its own stats pass)
import os as _os
ROWSTATS = _os.environ.get("EGP_ROWSTATS", "1") not in ("", "0")

# Optional per-launch tracing for bench.py's roofline numbers: when TRACE is a list, traced ops append
# (name, work, unit, start_event, end_event, detail) with CUDA events recorded on the launching (current) stream.
TRACE = None


class _Traced:
    def __init__(self, name: str, work: float, unit: str, detail: str = ""):
        self.on = TRACE is not None
        if self.on:
            self.name, self.work, self.unit, self.detail = name, work, unit, detail
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def __enter__(self):
        if self.on:
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.e1.record()
            TRACE.append((self.name, self.work, self.unit, self.e0, self.e1, self.detail))
        return False


def _code(t: Tensor) -> int:
    try:
        return L.DTYPE_CODE[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}; egopack_b200 computes in float32 or bfloat16") from None


def _c(t: Optional[Tensor]) -> Optional[Tensor]:
    return None if t is None else (t if t.is_contiguous() else t.contiguous())


def _i64(t: Tensor, name: str) -> Tensor:
    """Index tensors cross the C ABI as raw ``int64_t*`` (the reference's layout for pos / batch / ptr / edge_index /
    labels): anything else would be reinterpreted, so it is refused here instead of read out of bounds there."""
    if t.dtype != torch.int64:
        raise TypeError(f"{name} must be int64 (torch.long), got {t.dtype}; convert with .long() "
                        "(float positions as in plain PyG are not accepted by the band kernels)")
    return _c(t)


# =====================================================================================================
# graph structure
# =====================================================================================================
def band_edge_index(pos: Tensor, batch: Tensor, ptr: Tensor, r: float, max_num_neighbors: int = 32,
                    monotone: Optional[bool] = None) -> Tensor:
    """``RadiusGraph(r, loop=False)`` over a batch of graphs: int64 [2,E], dst-major, src ascending."""
    pos, batch, ptr = _i64(pos.view(-1), "pos"), _i64(batch, "batch"), _i64(ptr, "ptr")
    n = pos.numel()
    if n == 0:
        return torch.empty((2, 0), dtype=torch.int64, device=pos.device)
    if monotone is None:
        monotone = True if n < 2 else bool(((pos[1:] >= pos[:-1]) | (batch[1:] != batch[:-1])).all().item())
    deg = torch.empty(n, dtype=torch.int32, device=pos.device)
    L.call("egp_band_edge_count", L.ptr(pos), L.ptr(batch), L.ptr(ptr), n, float(r), int(max_num_neighbors),
           int(monotone), L.ptr(deg), L.stream())
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=pos.device)
    L.call("egp_exclusive_scan_i32", L.ptr(deg), n, L.ptr(rowptr), L.stream())
    e = int(rowptr[-1].item())
    edge_index = torch.empty((2, e), dtype=torch.int64, device=pos.device)
    L.call("egp_band_edge_fill", L.ptr(pos), L.ptr(batch), L.ptr(ptr), n, float(r), int(max_num_neighbors),
           int(monotone), L.ptr(rowptr), e, L.ptr(edge_index), L.stream())
    return edge_index


def lta_edge_index(pos: Tensor, y: Tensor, batch: Tensor, ptr: Tensor, r: float, max_num_neighbors: int = 32) -> Tensor:
    """``LTATemporalConnectivity(r)`` applied per graph of a batch: int64 [2,E] sorted by (src,dst)."""
    pos, y, batch, ptr = _i64(pos.view(-1), "pos"), _i64(y, "y"), _i64(batch, "batch"), _i64(ptr, "ptr")
    n = pos.numel()
    if n == 0:
        return torch.empty((2, 0), dtype=torch.int64, device=pos.device)
    ycols = y.shape[1] if y.dim() > 1 else 1
    deg = torch.empty(n, dtype=torch.int32, device=pos.device)
    L.call("egp_lta_edge_count", L.ptr(pos), L.ptr(y), ycols, L.ptr(batch), L.ptr(ptr), n, float(r),
           int(max_num_neighbors), L.ptr(deg), L.stream())
    rowptr = torch.empty(n + 1, dtype=torch.int64, device=pos.device)
    L.call("egp_exclusive_scan_i32", L.ptr(deg), n, L.ptr(rowptr), L.stream())
    e = int(rowptr[-1].item())
    edge_index = torch.empty((2, e), dtype=torch.int64, device=pos.device)
    L.call("egp_lta_edge_fill", L.ptr(pos), L.ptr(y), ycols, L.ptr(batch), L.ptr(ptr), n, float(r),
           int(max_num_neighbors), L.ptr(rowptr), e, L.ptr(edge_index), L.stream())
    return edge_index


@dataclass
class GraphStructure:
    """What the aggregation kernels need; built once per batch and cached on the batch object."""
    n: int
    band_k: Optional[int] = None
    win_lo: Optional[Tensor] = None
    win_hi: Optional[Tensor] = None
    inv_deg: Optional[Tensor] = None          # 1/max(in-degree,1)
    # band + star (LTATemporalConnectivity without an edge list; see egp_band_star_windows)
    ext_lo: Optional[Tensor] = None
    ext_hi: Optional[Tensor] = None
    hub_slot: Optional[Tensor] = None
    graph_meta: Optional[Tensor] = None
    num_graphs: int = 0
    rowptr_in: Optional[Tensor] = None        # CSR grouped by dst (forward)
    col_in: Optional[Tensor] = None
    rowptr_out: Optional[Tensor] = None       # CSR grouped by src (backward)
    col_out: Optional[Tensor] = None


def lta_star_counts(y: Tensor, ptr: Tensor, r: float) -> Tensor:
    """int32 [G,3] = (n_in, n_fc, first_src) per graph (lta_temp_connectivity.py:48-52), computed on the device."""
    y, ptr = _i64(y, "y"), _i64(ptr, "ptr")
    g = ptr.numel() - 1
    ycols = y.shape[1] if y.dim() > 1 else 1
    star = torch.empty((g, 3), dtype=torch.int32, device=y.device)
    L.call("egp_lta_star_counts", L.ptr(y), ycols, L.ptr(ptr), g, float(r), L.ptr(star), L.stream())
    return star


def band_structure(batch: Tensor, ptr: Tensor, k: int, star: Optional[Tensor] = None) -> GraphStructure:
    """Band (radius k, unit-spaced positions) of ONE collated batch, optionally with the LTA star of every graph."""
    return band_structure_many([(batch, ptr, star)], k)


def band_structure_many(parts, k: int) -> GraphStructure:
    """One structure over several collated batches laid out back to back (``Graph.forward_many``): ``parts`` is a list
    of ``(batch, ptr, star_or_None)``; rows / graphs of part p are offset by the sizes of the parts before it."""
    dev = parts[0][0].device
    n = sum(b.numel() for b, _, _ in parts)
    g = sum(p.numel() - 1 for _, p, _ in parts)
    with_star = any(st is not None for _, _, st in parts)
    lo = torch.empty(n, dtype=torch.int32, device=dev)
    hi = torch.empty(n, dtype=torch.int32, device=dev)
    inv = torch.empty(n, dtype=torch.float32, device=dev)
    elo = ehi = slot = meta = None
    if with_star:
        elo = torch.empty(n, dtype=torch.int32, device=dev)
        ehi = torch.empty(n, dtype=torch.int32, device=dev)
        slot = torch.empty(n, dtype=torch.int32, device=dev)
        meta = torch.zeros((g, 4), dtype=torch.int32, device=dev)
    ro = go = 0
    for batch, ptr, star in parts:
        nn, gg = batch.numel(), ptr.numel() - 1
        sl = slice(ro, ro + nn)
        L.call("egp_band_star_windows", L.ptr(_i64(batch, "batch")), L.ptr(_i64(ptr, "ptr")), nn, gg, int(k),
               L.ptr(_c(star)), ro, go, L.ptr(lo[sl]), L.ptr(hi[sl]), L.ptr(inv[sl]),
               L.ptr(elo[sl]) if with_star else None, L.ptr(ehi[sl]) if with_star else None,
               L.ptr(slot[sl]) if with_star else None, L.ptr(meta[go:go + gg]) if with_star else None, L.stream())
        ro += nn
        go += gg
    return GraphStructure(n=n, band_k=int(k), win_lo=lo, win_hi=hi, inv_deg=inv, ext_lo=elo, ext_hi=ehi, hub_slot=slot,
                          graph_meta=meta, num_graphs=g)


def csr_structure(edge_index: Tensor, n: int) -> GraphStructure:
    edge_index = _i64(edge_index, "edge_index")
    if edge_index.dim() != 2 or edge_index.shape[0] != 2:
        raise ValueError(f"edge_index must be [2, E], got {tuple(edge_index.shape)}")
    e = edge_index.shape[1]
    dev = edge_index.device
    out = GraphStructure(n=n)
    cursor = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    for by_dst in (1, 0):
        rowptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        L.call("egp_csr_build", L.ptr(edge_index), e, n, by_dst, L.ptr(rowptr), L.ptr(col), L.ptr(cursor), L.stream())
        if by_dst:
            out.rowptr_in, out.col_in = rowptr, col
        else:
            out.rowptr_out, out.col_out = rowptr, col
    out.inv_deg = torch.empty(n, dtype=torch.float32, device=dev)
    L.call("egp_csr_inv_degree", L.ptr(out.rowptr_in), n, L.ptr(out.inv_deg), L.stream())
    return out


def _aggregate(x: Tensor, gs: GraphStructure, backward: bool) -> Tensor:
    x = _c(x)
    n, c = x.shape
    out = torch.empty_like(x)
    so, si = (None, gs.inv_deg) if backward else (gs.inv_deg, None)
    kind = "sage_mean_csr" if gs.band_k is None else ("sage_mean_band_star" if gs.ext_lo is not None else "sage_mean_band")
    with _Traced(kind, 2.0 * n * c * x.element_size(), "B", f"k={gs.band_k}" if gs.band_k is not None else ""):
        _aggregate_launch(x, out, gs, backward, so, si, n, c)
    return out


def _aggregate_launch(x, out, gs, backward, so, si, n, c):
    if gs.band_k is not None and gs.ext_lo is not None and gs.band_k <= 4:
        nb, ws = 0, None
        if backward:
            nb = L.size("egp_sage_mean_band_star_workspace", n, c, gs.num_graphs)
            ws = L.workspace(nb, x.device, "hub")
            L.CALL_COUNTS["egp_sage_hub_fixup(kernel)"] = L.CALL_COUNTS.get("egp_sage_hub_fixup(kernel)", 0) + 1
        L.call("egp_sage_mean_band_star", L.ptr(x), L.ptr(out), n, c, c, c, gs.band_k, L.ptr(gs.win_lo), L.ptr(gs.win_hi),
               L.ptr(so), L.ptr(si), None if backward else L.ptr(gs.ext_lo), None if backward else L.ptr(gs.ext_hi),
               L.ptr(gs.hub_slot) if backward else None, L.ptr(gs.graph_meta) if backward else None, gs.num_graphs,
               _code(x), L.ptr(ws), nb, L.stream())
    elif gs.band_k is not None and gs.ext_lo is None:
        L.call("egp_sage_mean_band", L.ptr(x), L.ptr(out), n, c, c, c, gs.band_k, L.ptr(gs.win_lo), L.ptr(gs.win_hi),
               L.ptr(so), L.ptr(si), _code(x), L.stream())
    else:
        if gs.rowptr_in is None:
            raise RuntimeError("this graph structure needs a CSR (star with radius > 4): build it with csr_structure")
        rp, col = (gs.rowptr_out, gs.col_out) if backward else (gs.rowptr_in, gs.col_in)
        L.call("egp_sage_mean_csr", L.ptr(x), L.ptr(out), n, c, c, c, L.ptr(rp), L.ptr(col), L.ptr(so), L.ptr(si),
               _code(x), L.stream())


class SageMean(torch.autograd.Function):
    """agg_i = mean_{j -> i} x_j (0 for isolated nodes) -- the aggregation inside gnn.SAGEConv (models/graph.py:42)."""

    @staticmethod
    def forward(ctx, x: Tensor, gs: GraphStructure) -> Tensor:
        ctx.gs = gs
        return _aggregate(x, gs, backward=False)

    @staticmethod
    def backward(ctx, g: Tensor):
        return _aggregate(g, ctx.gs, backward=True), None


# =====================================================================================================
# GEMM
# =====================================================================================================
def gemm(a: Tensor, a_trans: bool, b: Tensor, b_trans: bool, m: int, n: int, k: int, *, a2: Optional[Tensor] = None,
         b2: Optional[Tensor] = None, k2: int = 0, bias: Optional[Tensor] = None, residual: Optional[Tensor] = None,
         act: int = ACT_NONE, slope: float = 0.0, out_dtype: Optional[torch.dtype] = None,
         out: Optional[Tensor] = None, accumulate: bool = False, rowstats: bool = False) -> Tensor:
    """C[m,n] = act(A B^T + A2 B2^T + bias) + residual, see ``egp_gemm``.  Operands must be 2-D contiguous.
    ``rowstats=True`` tags the result with the row-block {sum, sum of squares} pairs of ``egp_gemm_rowstats`` when the
    shape allows it (``_take_rowstats``), for a whole-tensor statistic downstream without another pass."""
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty((m, n), dtype=out_dtype, device=a.device)
    a, b, a2, b2, residual = _c(a), _c(b), _c(a2), _c(b2), _c(residual)
    if bias is not None:
        assert bias.dtype == torch.float32
    if residual is not None:
        assert residual.dtype == out.dtype and residual.shape == out.shape
    name = "gemm_tcgen05" if a.dtype == torch.bfloat16 else "gemm_fp32"
    detail = ""
    if TRACE is not None:
        detail = (f"{m}x{n}x{k}" + (f"+{k2}" if k2 else "") + f" {'T' if a_trans else 'N'}{'T' if b_trans else 'N'}"
                  f" -> {'f32' if out.dtype == torch.float32 else 'bf16'}" + (" +bias" if bias is not None else "")
                  + (f" act{act}" if act else "") + (" +res" if residual is not None else "") + (" acc" if accumulate else ""))
    stats = None
    if rowstats and ROWSTATS and not accumulate and not a_trans and not b_trans and n % 256 == 0 and m >= 128 * 148 \
            and out.is_contiguous() and (
            a.dtype == torch.bfloat16 or _fp32_on_tensor_cores()):
        stats = torch.empty(L.size("egp_gemm_rowstats_bytes", m, n) // 8, dtype=torch.float64, device=a.device)
    with _Traced(name, 2.0 * m * n * (k + k2), "FLOP", detail):
        if stats is not None:
            try:
                _gemm_launch(a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope, accumulate, stats)
            except RuntimeError as ex:       # shape without the statistics epilogue (status -3): plain GEMM, no tag
                if "status -3" not in str(ex):
                    raise
                stats = None
                _gemm_launch(a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope, accumulate)
        else:
            _gemm_launch(a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope, accumulate)
    if stats is not None:
        out._egp_rowstats = (out._version, stats)
    return out


def _fp32_on_tensor_cores() -> bool:
    from . import config
    return config.get_fp32_gemm() != "ffma"


def _take_rowstats(x: Tensor) -> Optional[Tensor]:
    tag = getattr(x, "_egp_rowstats", None)
    if tag is not None and tag[0] == x._version:
        return tag[1]
    return None


def _gemm_ws(device):
    """(pointer, bytes) of the split-K slab workspace in the deterministic mode, (None, 0) otherwise."""
    from . import config
    if not config.is_deterministic():
        return None, 0
    nb = L.size("egp_gemm_workspace", 0, 0, 0)
    return L.ptr(L.workspace(nb, device, "gemm")), nb


def _tc_call(a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope, accumulate, stats):
    """bf16 operands -> egp_gemm, or egp_gemm_rowstats when the epilogue statistics are wanted."""
    if stats is not None:
        L.call("egp_gemm_rowstats", L.ptr(a), a.stride(0), int(a_trans), L.ptr(b), b.stride(0), int(b_trans),
               L.ptr(a2), a2.stride(0) if a2 is not None else 0, L.ptr(b2), b2.stride(0) if b2 is not None else 0, int(k2),
               L.ptr(_c(bias)), L.ptr(residual), residual.stride(0) if residual is not None else 0,
               L.ptr(out), out.stride(0), m, n, k, int(act), float(slope), L.DTYPE_CODE[out.dtype], L.ptr(stats), L.stream())
        return
    wsp, wsb = _gemm_ws(a.device)
    L.call("egp_gemm", L.ptr(a), a.stride(0), int(a_trans), L.ptr(b), b.stride(0), int(b_trans),
           L.ptr(a2), a2.stride(0) if a2 is not None else 0, L.ptr(b2), b2.stride(0) if b2 is not None else 0, int(k2),
           L.ptr(_c(bias)), L.ptr(residual), residual.stride(0) if residual is not None else 0,
           L.ptr(out), out.stride(0), m, n, k, int(act), float(slope), BF16, L.DTYPE_CODE[out.dtype],
           int(accumulate), wsp, wsb, L.stream())


def _gemm_launch(a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope, accumulate, stats=None):
    if a.dtype == torch.float32:
        from . import config
        kind = config.get_fp32_gemm()
        if kind != "ffma" and m > 0 and n > 0 and k > 0:
            return _gemm_fp32_tensor(kind, a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope,
                                     accumulate, stats)
    elif stats is not None:
        return _tc_call(a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope, accumulate, stats)
    wsp, wsb = _gemm_ws(a.device)
    L.call("egp_gemm", L.ptr(a), a.stride(0), int(a_trans), L.ptr(b), b.stride(0), int(b_trans),
           L.ptr(a2), a2.stride(0) if a2 is not None else 0, L.ptr(b2), b2.stride(0) if b2 is not None else 0, int(k2),
           L.ptr(_c(bias)), L.ptr(residual), residual.stride(0) if residual is not None else 0,
           L.ptr(out), out.stride(0), m, n, k, int(act), float(slope), _code(a), L.DTYPE_CODE[out.dtype],
           int(accumulate), wsp, wsb, L.stream())


# term products (A term, B term) of the split GEMM, smallest first so the fp32 accumulator adds them in increasing
# magnitude: bf16x3 keeps the products down to 2^-16 |a||b|, bf16x6 down to 2^-24
_SPLIT_PRODUCTS = {"bf16x3": ((1, 0), (0, 1), (0, 0)),
                   "bf16x6": ((1, 1), (2, 0), (0, 2), (1, 0), (0, 1), (0, 0))}


def _split_operand(x: Tensor, rows: int, k: int, trans: bool, terms) -> Tuple[Tensor, int]:
    """bf16 operand [rows, T*Kseg] (K-major) or [T*Kseg, rows_padded] (MN-major) holding the chosen split terms of the
    fp32 operand ``x`` side by side along K.  Returns (tensor, K') with K' = T * Kseg, Kseg = K rounded up to 8."""
    import ctypes
    T = len(terms)
    kseg = (k + 7) // 8 * 8
    sel = (ctypes.c_int * T)(*terms)
    if not trans:                                     # x is [rows, k]
        out = torch.empty((rows, T * kseg), dtype=torch.bfloat16, device=x.device)
        L.call("egp_split_bf16", L.ptr(x), x.stride(0), rows, k, rows, kseg, L.ptr(out), kseg, T * kseg, T, sel, L.stream())
    else:                                             # x is [k, rows]
        rp = (rows + 7) // 8 * 8
        out = torch.empty((T * kseg, rp), dtype=torch.bfloat16, device=x.device)
        L.call("egp_split_bf16", L.ptr(x), x.stride(0), k, rows, kseg, rp, L.ptr(out), kseg * rp, rp, T, sel, L.stream())
    return out, T * kseg


def _gemm_fp32_tensor(kind, a, a_trans, b, b_trans, a2, b2, k2, bias, residual, out, m, n, k, act, slope, accumulate,
                      stats=None):
    """fp32 GEMM on the tcgen05 pipe: both operands are split into bf16 terms and the term products are laid out along
    K, so the ordinary bf16 kernel (TMA, TMEM fp32 accumulation, fused epilogue) computes the fp32 result in one launch."""
    prods = _SPLIT_PRODUCTS[kind]
    ta, tb = [p[0] for p in prods], [p[1] for p in prods]
    a_s, kk = _split_operand(a, m, k, a_trans, ta)
    b_s, _ = _split_operand(b, n, k, b_trans, tb)
    a2_s = b2_s = None
    kk2 = 0
    if a2 is not None and k2 > 0:
        a2_s, kk2 = _split_operand(a2, m, k2, a_trans, ta)
        b2_s, _ = _split_operand(b2, n, k2, b_trans, tb)
    _tc_call(a_s, a_trans, b_s, b_trans, a2_s, b2_s, kk2, bias, residual, out, m, n, kk, act, slope, accumulate, stats)


def colsum(x: Tensor) -> Tensor:
    x = _c(x)
    rows, cols = x.shape
    out = torch.empty(cols, dtype=torch.float32, device=x.device)
    nb = L.size("egp_colsum_workspace", rows, cols)
    ws = L.workspace(nb, x.device)
    with _Traced("colsum", 1.0 * rows * cols * x.element_size(), "B", f"c={cols}"):
        L.call("egp_colsum", L.ptr(x), L.ptr(out), rows, cols, x.stride(0), _code(x), L.ptr(ws), nb, L.stream())
    return out


def cast(x: Tensor, dtype: torch.dtype) -> Tensor:
    if x.dtype == dtype:
        return x
    x = _c(x)
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    with _Traced("cast", 1.0 * x.numel() * (x.element_size() + out.element_size()), "B"):
        L.call("egp_cast", L.ptr(x), L.ptr(out), x.numel(), _code(x), L.DTYPE_CODE[dtype], L.stream())
    return out


def act_bwd(dy: Tensor, y: Tensor, act: int, slope: float) -> Tensor:
    dy, y = _c(dy), _c(y)
    dx = torch.empty_like(dy)
    L.call("egp_act_bwd", L.ptr(dy), L.ptr(y), L.ptr(dx), dy.numel(), act, float(slope), _code(dy), L.stream())
    return dx


def add(a: Tensor, b: Tensor) -> Tensor:
    a, b = _c(a), _c(b)
    out = torch.empty_like(a)
    L.call("egp_add", L.ptr(a), L.ptr(b), L.ptr(out), a.numel(), _code(a), L.stream())
    return out


def axpby(a: Tensor, alpha: float, b: Optional[Tensor] = None, beta: float = 0.0) -> Tensor:
    a, b = _c(a), _c(b)
    out = torch.empty_like(a)
    n = a.numel()
    vn = 8 if a.dtype == torch.bfloat16 else 4
    if n % vn:                                           # ragged tail (tiny logits tensors): pad to a vector multiple
        pad = vn - n % vn
        af = torch.cat([a.reshape(-1), a.new_zeros(pad)])
        bf = torch.cat([b.reshape(-1), b.new_zeros(pad)]) if b is not None else None
        of = torch.empty_like(af)
        L.call("egp_axpby", L.ptr(af), float(alpha), L.ptr(bf), float(beta), L.ptr(of), af.numel(), _code(a), L.stream())
        return of[:n].reshape(a.shape)
    L.call("egp_axpby", L.ptr(a), float(alpha), L.ptr(b), float(beta), L.ptr(out), n, _code(a), L.stream())
    return out


class Scale(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha: float):
        ctx.alpha = alpha
        return axpby(x, alpha)

    @staticmethod
    def backward(ctx, g):
        return axpby(g, ctx.alpha), None


def _attach_colsum(dx: Tensor, cs: Tensor) -> Tensor:
    """Remember the column sums of a gradient that a kernel produced as a by-product.  The tag is bound to the
    tensor's version counter, so an in-place accumulation by autograd invalidates it."""
    dx._egp_colsum = (dx._version, cs)
    return dx


def _take_colsum(dy: Tensor) -> Optional[Tensor]:
    tag = getattr(dy, "_egp_colsum", None)
    if tag is not None and tag[0] == dy._version and tag[1].shape[0] == dy.shape[1]:
        return tag[1]
    return None


def act_bwd_colsum(dy: Tensor, y: Tensor, act: int, slope: float) -> Tuple[Tensor, Tensor]:
    """dx = dy * act'(y) and the column sums of dx in one pass."""
    dy, y = _c(dy), _c(y)
    rows, cols = dy.shape
    dx = torch.empty_like(dy)
    cs = torch.empty(cols, dtype=torch.float32, device=dy.device)
    nb = L.size("egp_act_bwd_colsum_workspace", rows, cols)
    ws = L.workspace(nb, dy.device)
    with _Traced("act_bwd_colsum", 3.0 * rows * cols * dy.element_size(), "B"):
        L.call("egp_act_bwd_colsum", L.ptr(dy), L.ptr(y), L.ptr(dx), L.ptr(cs), rows, cols, act, float(slope), _code(dy),
               L.ptr(ws), nb, L.stream())
    return dx, cs


_dropout_calls = 0
# Set by egopack_b200.graphs while a training step is captured into a CUDA graph: a device uint64[2] = {seed, step}
# that the replayed kernels read, because their scalar arguments are frozen at capture time.
RNG_STATE: Optional[Tensor] = None


def _dropout_stream() -> Tuple[int, int]:
    """(seed, offset) of the next fused-dropout call: reproducible under torch.manual_seed and call order."""
    global _dropout_calls
    _dropout_calls += 1
    return int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF, _dropout_calls


# Bumped whenever parameters may have changed WITHOUT their Python version counters moving: a CUDA-graph replay
# runs the captured optimizer kernels on the device only (egopack_b200.graphs.GraphedStep.__call__).  Every cache of
# values derived from parameters (bf16 weight copies here, normalised prototype banks in GraphONE) keys on it.
_param_generation = 0


def param_generation() -> int:
    return _param_generation


def bump_param_generation() -> None:
    global _param_generation
    _param_generation += 1


class _WeightCache:
    """bf16 copies of fp32 parameters, refreshed when the parameter is updated in place (optimizer step: the tensor's
    version counter moves) or by a CUDA-graph replay (``param_generation`` moves).

    Entries are keyed by storage address but validated through a weak reference to the source tensor, so a new
    parameter that happens to reuse a freed address can never pick up a stale copy."""

    def __init__(self):
        self._store = {}
        self._shadow = {}                # id(param) -> [weakref(param), bf16 view kept current by FlatAdam, version]

    def register_shadow(self, w: Tensor, view: Tensor) -> None:
        """``view`` is a low-precision copy of ``w`` that an optimizer rewrites in the same kernel that updates ``w``
        (egopack_b200.optim.FlatAdam).  It is handed out instead of a cast while ``w`` has not been modified by
        anything else since the optimizer last synced it (version counter)."""
        self._shadow[id(w)] = [weakref.ref(w), view, w._version]

    def shadows_synced(self, params) -> None:
        for w in params:
            ent = self._shadow.get(id(w))
            if ent is not None and ent[0]() is w:
                ent[2] = w._version

    def get(self, w: Tensor, dtype: torch.dtype, pad_rows: int = 0) -> Tensor:
        """`pad_rows` > rows appends zero rows (dgrad of a classifier whose gradient had its class dim padded)."""
        pad_rows = pad_rows if pad_rows > w.shape[0] else 0
        if w.dtype == dtype and not pad_rows:
            return w
        if not pad_rows:
            ent = self._shadow.get(id(w))
            if ent is not None and ent[0]() is w and ent[1].dtype == dtype and ent[2] == w._version \
                    and ent[1].shape == w.shape:
                return ent[1]
        key = (w.data_ptr(), tuple(w.shape), dtype, pad_rows)
        hit = self._store.get(key)
        if hit is not None:
            src = hit[0]()
            if src is not None and src.data_ptr() == w.data_ptr() and hit[1] == (w._version, _param_generation):
                return hit[2]
        c = cast(w.detach(), dtype)
        if pad_rows:
            c = torch.cat([c, c.new_zeros((pad_rows - w.shape[0], w.shape[1]))], 0)
        if len(self._store) > 4096:
            self._store = {k: v for k, v in self._store.items() if v[0]() is not None}
        self._store[key] = (weakref.ref(w), (w._version, _param_generation), c)
        return c


weight_cache = _WeightCache()


def cast_pad(x: Tensor, dtype: torch.dtype, mult: int) -> Tensor:
    """cast + zero-pad the last dim up to a multiple of `mult` in one kernel (bf16 operands need 16-byte row
    pitches for TMA; classifier gradients have 115 / 478 columns)."""
    r, c = x.shape
    cp = (c + mult - 1) // mult * mult
    if cp == c:
        return cast(x, dtype)
    if x.stride(1) != 1:
        x = x.contiguous()
    out = torch.empty((r, cp), dtype=dtype, device=x.device)
    L.call("egp_cast_pad", L.ptr(x), x.stride(0), L.ptr(out), cp, r, c, _code(x), L.DTYPE_CODE[dtype], L.stream())
    return out


class Linear(torch.autograd.Function):
    """y = act(x W^T [+ x2 W2^T] + b) [+ residual].

    Forward / dgrad / wgrad are all ``egp_gemm`` calls (tcgen05 for bf16 activations, FFMA for fp32); the
    two-operand form is SAGEConv's ``lin_l(agg) + lin_r(x)`` accumulated in one TMEM tile.
    Weights stay fp32 parameters; bf16 copies come from ``weight_cache``.  ``out_dtype`` lets classifier heads
    emit fp32 logits from bf16 features.
    """

    @staticmethod
    def forward(ctx, x, w, b, x2, w2, residual, act: int, slope: float, out_dtype):
        cd = x.dtype                                     # compute dtype follows the activations
        m, k = x.shape
        n = w.shape[0]
        wc = weight_cache.get(w, cd)
        w2c = weight_cache.get(w2, cd) if w2 is not None else None
        out_dtype = out_dtype or cd
        out = None
        per16 = 16 // (4 if out_dtype == torch.float32 else 2)
        if n % per16 and residual is None and n > per16:
            # classifier heads (115 / 478 classes): pad the row pitch to 16 bytes so the TMA-store epilogue applies;
            # the caller sees the [m, n] view
            out = torch.empty((m, (n + per16 - 1) // per16 * per16), dtype=out_dtype, device=x.device)[:, :n]
        y = gemm(x, False, wc, False, m, n, k, a2=x2, b2=w2c, k2=(x2.shape[1] if x2 is not None else 0), bias=b,
                 residual=residual, act=act, slope=slope, out_dtype=out_dtype, out=out)
        ctx.save_for_backward(x, w, x2, w2, y if act != ACT_NONE else None)
        ctx.act, ctx.slope, ctx.has_bias, ctx.has_res = act, slope, b is not None, residual is not None
        ctx.res_dtype = residual.dtype if residual is not None else None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, x2, w2, y = ctx.saved_tensors
        cd = x.dtype
        m, k = x.shape
        n = w.shape[0]
        dy = _c(dy)
        dres = dy if ctx.has_res else None
        g = dy
        want_db = ctx.has_bias and ctx.needs_input_grad[2]
        db = None
        if ctx.act != ACT_NONE:
            if ctx.has_res:
                raise RuntimeError("Linear: activation together with a residual is not differentiable here")
            vn = 8 if dy.dtype == torch.bfloat16 else 4
            if want_db and n % vn == 0:
                g, db = act_bwd_colsum(dy, y, ctx.act, ctx.slope)      # bias gradient in the same pass
            else:
                g = act_bwd(dy, y, ctx.act, ctx.slope)
        elif want_db:
            db = _take_colsum(dy)                                     # by-product of the LN backward upstream
        if want_db and db is None:
            db = colsum(g)
        # fp32 logits gradients -> bf16 operand; TMA needs 16-byte row pitches, so the class dim is zero-padded
        gc = cast_pad(g, cd, 8) if cd == torch.bfloat16 else cast(g, cd)
        npad = gc.shape[1]
        dx = dx2 = dw = dw2 = None
        wc = weight_cache.get(w, cd, pad_rows=npad)
        if ctx.needs_input_grad[0]:
            dx = gemm(gc, False, wc, True, m, k, npad)                           # dx = g W
        if ctx.needs_input_grad[1]:
            dw = gemm(gc, True, x, True, n, k, m, out_dtype=torch.float32)       # dW = g^T x
        if x2 is not None:
            k2 = x2.shape[1]
            w2c = weight_cache.get(w2, cd, pad_rows=npad)
            if ctx.needs_input_grad[3]:
                dx2 = gemm(gc, False, w2c, True, m, k2, npad)
            if ctx.needs_input_grad[4]:
                dw2 = gemm(gc, True, x2, True, n, k2, m, out_dtype=torch.float32)
        if dres is not None and dres.dtype != ctx.res_dtype:
            dres = cast(dres, ctx.res_dtype)
        return dx, dw, db, dx2, dw2, dres, None, None, None


def _posenc_add(x: Tensor, pos: Tensor, freq: Tensor) -> Tensor:
    n, c = x.shape
    out = torch.empty_like(x)
    with _Traced("posenc_add", 2.0 * n * c * x.element_size(), "B"):
        L.call("egp_posenc_add", L.ptr(x), L.ptr(_i64(pos.view(-1), "pos")), L.ptr(_c(freq)), L.ptr(out), n, c, _code(x),
               L.stream())
    return out


def _sage_forward(ctx, z, wp, bp, wl, bl, wr, gs):
    cd = z.dtype
    m, h = z.shape
    ho = wl.shape[0]
    wpc, wlc, wrc = weight_cache.get(wp, cd), weight_cache.get(wl, cd), weight_cache.get(wr, cd)
    xs = gemm(z, False, wpc, False, m, h, h, bias=bp, act=ACT_RELU)
    agg = _aggregate(xs, gs, backward=False)
    # the statistics of the graph-mode LayerNorm that follows (models/graph.py:43) ride on this GEMM's epilogue
    u = gemm(agg, False, wlc, False, m, ho, h, a2=z, b2=wrc, k2=h, bias=bl, rowstats=True)
    ctx.save_for_backward(z, xs, agg, wp, wl, wr)
    ctx.gs, ctx.has_bl = gs, bl is not None
    return u


def _sage_backward(ctx, du, need, extra_dz=None):
    """need = (dz, dwp, dbp, dwl, dbl, dwr) flags.  ``extra_dz`` is added to dz in the dgrad GEMM's epilogue."""
    z, xs, agg, wp, wl, wr = ctx.saved_tensors
    cd = z.dtype
    m, h = z.shape
    ho = wl.shape[0]
    du = _c(du)
    dbl = None
    if ctx.has_bl and need[4]:
        dbl = _take_colsum(du)                                        # by-product of the graph-LN backward upstream
        if dbl is None:
            dbl = colsum(du)
    duc = cast(du, cd)
    wpc, wlc, wrc = weight_cache.get(wp, cd), weight_cache.get(wl, cd), weight_cache.get(wr, cd)
    dagg = gemm(duc, False, wlc, True, m, h, ho)                      # through lin_l
    dxs = _aggregate(dagg, ctx.gs, backward=True)                     # transposed mean aggregation
    g, dbp = act_bwd_colsum(dxs, xs, ACT_RELU, 0.0)                   # through the projection's ReLU (+ its bias grad)
    dz = None
    if need[0]:
        if extra_dz is not None and extra_dz.dtype != cd:
            extra_dz = cast(_c(extra_dz), cd)
        # g Wp + du Wr in one TMEM tile (+ the gradient that reaches the layer input along another path)
        dz = gemm(g, False, wpc, True, m, h, h, a2=duc, b2=wrc, k2=ho, residual=_c(extra_dz))
    dwp = gemm(g, True, z, True, h, h, m, out_dtype=torch.float32) if need[1] else None
    dwl = gemm(duc, True, agg, True, ho, h, m, out_dtype=torch.float32) if need[3] else None
    dwr = gemm(duc, True, z, True, ho, h, m, out_dtype=torch.float32) if need[5] else None
    return dz, dwp, (dbp if need[2] else None), dwl, dbl, dwr


class SageLayer(torch.autograd.Function):
    """u = lin_l(mean_{j -> i} relu(lin(z))_j) + lin_r(z): gnn.SAGEConv(H, H, project=True, aggr='mean') as ONE autograd
    node (models/graph.py:42).  Forward: ReLU-epilogue GEMM -> band / band+star / CSR mean -> dual-operand GEMM.
    Backward: the two gradient contributions of the layer input -- through the projection and through lin_r -- come
    out of ONE dual-operand dgrad GEMM, ``dz = g Wp + du Wr`` accumulated in the same TMEM tile, instead of two GEMMs
    and an elementwise add of two [N, H] tensors by autograd."""

    @staticmethod
    def forward(ctx, z, wp, bp, wl, bl, wr, gs):
        return _sage_forward(ctx, _c(z), wp, bp, wl, bl, wr, gs)

    @staticmethod
    def backward(ctx, du):
        return (*_sage_backward(ctx, du, ctx.needs_input_grad[:6]), None)


class SageLayerPE(torch.autograd.Function):
    """The FIRST layer of ``Graph``'s stack together with what surrounds it in models/graph.py:63,
    ``x + net(x + PE(pos))``: takes x, adds the positional encoding itself, and hands x back as a second output for the
    outer residual.  The residual's gradient therefore arrives HERE, and is folded into the layer's dgrad GEMM epilogue
    (``dx = g Wp + du Wr + d_residual``) -- the [N, H] add of the two paths into x never runs as a separate kernel."""

    @staticmethod
    def forward(ctx, x, wp, bp, wl, bl, wr, gs, pos, freq):
        x = _c(x)
        z = _posenc_add(x, pos, freq)
        u = _sage_forward(ctx, z, wp, bp, wl, bl, wr, gs)
        return u, x                                                   # x comes back as an alias carrying this node's grad_fn

    @staticmethod
    def backward(ctx, du, dres):
        return (*_sage_backward(ctx, du, ctx.needs_input_grad[:6], extra_dz=dres), None, None, None)


class LinearCat(torch.autograd.Function):
    """``cat([x_1, ..., x_T]) W^T + b`` without materialising the concatenation: every part's GEMM writes its rows of
    ONE output tensor (a single GEMM when the parts already sit back to back in memory).  Used by
    ``Graph.forward_many`` for the first TRN Linear (K = S*D = 4608): the task batches arrive as separate feature
    tensors, everything after this layer runs on the stacked rows.  The inputs are data (no input gradient); the weight
    gradient is accumulated part by part into one fp32 buffer."""

    @staticmethod
    def forward(ctx, w, b, *xs):
        cd = xs[0].dtype
        n_out, k = w.shape
        rows = [x.shape[0] for x in xs]
        xs = [_c(x) for x in xs]
        wc = weight_cache.get(w, cd)
        out = torch.empty((sum(rows), n_out), dtype=cd, device=xs[0].device)
        esz = xs[0].element_size()
        # parts that are consecutive views of ONE allocation (DeviceFeeder(fuse_features=True)) are a single operand
        adjacent = all(a.untyped_storage().data_ptr() == b_.untyped_storage().data_ptr()
                       and a.data_ptr() + a.numel() * esz == b_.data_ptr() for a, b_ in zip(xs, xs[1:]))
        if adjacent:
            whole = torch.as_strided(xs[0], (sum(rows), k), (k, 1))
            gemm(whole, False, wc, False, sum(rows), n_out, k, bias=b, out=out)
        else:
            off = 0
            for x, m in zip(xs, rows):
                if m:
                    gemm(x, False, wc, False, m, n_out, k, bias=b, out=out[off:off + m])
                off += m
        ctx.save_for_backward(w, *xs)
        ctx.rows, ctx.has_bias, ctx.adjacent = rows, b is not None, adjacent
        return out

    @staticmethod
    def backward(ctx, dy):
        w, *xs = ctx.saved_tensors
        cd = xs[0].dtype
        n_out, k = w.shape
        dy = _c(dy)
        db = None
        if ctx.has_bias and ctx.needs_input_grad[1]:
            db = _take_colsum(dy)
            if db is None:
                db = colsum(dy)
        dw = None
        if ctx.needs_input_grad[0]:
            gc = cast(dy, cd)
            dw = torch.empty((n_out, k), dtype=torch.float32, device=dy.device)
            if ctx.adjacent:
                total = sum(ctx.rows)
                whole = torch.as_strided(xs[0], (total, k), (k, 1))
                gemm(gc, True, whole, True, n_out, k, total, out_dtype=torch.float32, out=dw)
            else:
                off, first = 0, True
                for x, m in zip(xs, ctx.rows):
                    if m:
                        gemm(gc[off:off + m], True, x, True, n_out, k, m, out_dtype=torch.float32, out=dw, accumulate=not first)
                        first = False
                    off += m
                if first:
                    dw.zero_()
        return (dw, db, *([None] * len(xs)))


class SplitRows(torch.autograd.Function):
    """Row blocks of a stacked activation as separate tensors (the per-task features coming out of
    ``Graph.forward_many``); the backward stacks the blocks' gradients again (missing ones are zero)."""

    @staticmethod
    def forward(ctx, x, *rows):
        ctx.rows = rows
        ctx.meta = (x.shape[1:], x.dtype, x.device)
        outs, off = [], 0
        for m in rows:
            outs.append(x[off:off + m])
            off += m
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        tail, dtype, dev = ctx.meta
        parts = [g if g is not None else torch.zeros((m, *tail), dtype=dtype, device=dev) for g, m in zip(grads, ctx.rows)]
        return (torch.cat(parts, 0), *([None] * len(ctx.rows)))


def linear(x, w, b=None, *, x2=None, w2=None, residual=None, act=ACT_NONE, slope=0.0, out_dtype=None):
    return Linear.apply(x, w, b, x2, w2, residual, act, slope, out_dtype)


# =====================================================================================================
# normalisation / elementwise
# =====================================================================================================
class RowLayerNorm(torch.autograd.Function):
    """nn.LayerNorm over the last dim, optionally fused with ReLU and with the Dropout that follows it
    (TRNPooling: Linear -> LayerNorm -> ReLU -> Dropout; task nets; GraphONE stages).  No dropout mask is stored:
    a zero in the saved output means "ReLU inactive or dropped".  The backward also returns the column sums of dx
    (the bias gradient of the Linear that produced x) as a by-product."""

    @staticmethod
    def forward(ctx, x, w, b, eps: float, act: int, dropout_p: float = 0.0):
        x = _c(x)
        n, c = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(n, dtype=torch.float32, device=x.device)
        rstd = torch.empty(n, dtype=torch.float32, device=x.device)
        seed, offset = _dropout_stream() if dropout_p > 0 else (0, 0)
        rng = RNG_STATE if dropout_p > 0 else None
        # eager: the full 64-bit call counter (never repeats); under CUDA-graph capture the per-step part comes from the
        # device-side state (added as step << 20), so only the call-site index inside a step is passed here
        offset = (offset & 0xFFFFF) if rng is not None else (offset & 0xFFFFFFFFFFFFFFFF)
        with _Traced("row_layernorm_fwd", 2.0 * n * c * x.element_size(), "B"):
            L.call("egp_row_layernorm_fwd", L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(mean), L.ptr(rstd), n, c,
                   float(eps), act, float(dropout_p), seed, offset, L.ptr(rng), _code(x), L.stream())
        need_y = act == ACT_RELU or dropout_p > 0
        ctx.save_for_backward(x, y if need_y else None, w, mean, rstd)
        ctx.act, ctx.out_scale = act, (1.0 / (1.0 - dropout_p) if dropout_p > 0 else 1.0)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, w, mean, rstd = ctx.saved_tensors
        dy = _c(dy)
        n, c = x.shape
        dx = torch.empty_like(x)
        dw = torch.empty(c, dtype=torch.float32, device=x.device)
        db = torch.empty(c, dtype=torch.float32, device=x.device)
        dxs = torch.empty(c, dtype=torch.float32, device=x.device)
        nb = L.size("egp_row_layernorm_workspace", n, c)
        ws = L.workspace(nb, x.device)
        # dy, x (and y when a ReLU / dropout mask is read back) in, dx out
        with _Traced("row_layernorm_bwd", (4.0 if y is not None else 3.0) * n * c * x.element_size(), "B"):
            L.call("egp_row_layernorm_bwd", L.ptr(dy), L.ptr(x), L.ptr(y), L.ptr(w), L.ptr(mean), L.ptr(rstd), L.ptr(dx),
                   L.ptr(dw), L.ptr(db), L.ptr(dxs), n, c, ctx.act, float(ctx.out_scale), _code(x), L.ptr(ws), nb, L.stream())
        return _attach_colsum(dx, dxs), dw, db, None, None, None


class GraphLayerNorm(torch.autograd.Function):
    """gnn.LayerNorm in graph mode WITHOUT a batch vector (models/graph.py:43): statistics over the whole
    [N,C] tensor, ``x / (std + eps)``, per-channel affine; fused with LeakyReLU (models/graph.py:44).

    ``seg_rows`` (a tuple of cumulative row offsets ``(0, n1, n1+n2, ..., N)``) splits the rows into consecutive
    segments that are normalised independently -- one per original forward call when ``Graph.forward_many`` runs
    several task batches through the shared weights at once (the statistics of the reference are per call)."""

    @staticmethod
    def forward(ctx, x, w, b, eps: float, act: int, slope: float, seg_rows=None):
        import ctypes
        x = _c(x)
        n, c = x.shape
        segs = tuple(int(v) for v in seg_rows) if seg_rows is not None else (0, n)
        nseg = len(segs) - 1
        seg_arr = (ctypes.c_int64 * (nseg + 1))(*segs)
        y = torch.empty_like(x)
        stats = torch.empty(2 * nseg, dtype=torch.float64, device=x.device)
        pre = _take_rowstats(x) if ROWSTATS else None
        if pre is not None and c % 64 == 0 and all(v % 128 == 0 for v in segs[1:-1]):
            # the producing GEMM left row-block {sum, sumsq} pairs: no stats pass over x
            with _Traced("graph_layernorm_fwd", 2.0 * n * c * x.element_size(), "B", "stats from the GEMM epilogue"):
                L.call("egp_graph_layernorm_seg_fwd_rowstats", L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(stats), n, c,
                       nseg, seg_arr, L.ptr(pre), float(eps), act, float(slope), _code(x), L.stream())
        else:
            nb = L.size("egp_graph_layernorm_seg_workspace", n, c, nseg)
            ws = L.workspace(nb, x.device)
            with _Traced("graph_layernorm_fwd", 3.0 * n * c * x.element_size(), "B"):      # stats pass + apply pass
                L.call("egp_graph_layernorm_seg_fwd", L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(stats), n, c, nseg,
                       seg_arr, float(eps), act, float(slope), _code(x), L.ptr(ws), nb, L.stream())
        ctx.save_for_backward(x, w, b, stats)
        ctx.cfg = (float(eps), act, float(slope), segs)
        return y

    @staticmethod
    def backward(ctx, dy):
        import ctypes
        x, w, b, stats = ctx.saved_tensors
        eps, act, slope, segs = ctx.cfg
        nseg = len(segs) - 1
        seg_arr = (ctypes.c_int64 * (nseg + 1))(*segs)
        dy = _c(dy)
        n, c = x.shape
        dx = torch.empty_like(x)
        dw = torch.empty(c, dtype=torch.float32, device=x.device)
        db = torch.empty(c, dtype=torch.float32, device=x.device)
        nb = L.size("egp_graph_layernorm_seg_workspace", n, c, nseg)
        ws = L.workspace(nb, x.device)
        dxs = torch.empty(c, dtype=torch.float32, device=x.device)
        with _Traced("graph_layernorm_bwd", 5.0 * n * c * x.element_size(), "B"):      # reduce (dy,x) + apply (dy,x -> dx)
            L.call("egp_graph_layernorm_seg_bwd", L.ptr(dy), L.ptr(x), L.ptr(w), L.ptr(b), L.ptr(stats), L.ptr(dx), L.ptr(dw),
                   L.ptr(db), L.ptr(dxs), n, c, nseg, seg_arr, eps, act, slope, _code(x), L.ptr(ws), nb, L.stream())
        return _attach_colsum(dx, dxs), dw, db, None, None, None, None


class PosEncAdd(torch.autograd.Function):
    """x + gnn.PositionalEncoding(pos) (models/graph.py:63); the encoding has no parameters."""

    @staticmethod
    def forward(ctx, x, pos, freq):
        return _posenc_add(_c(x), pos, freq)

    @staticmethod
    def backward(ctx, g):
        return g, None, None


class Cast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return cast(x, dtype)

    @staticmethod
    def backward(ctx, g):
        return cast(g, ctx.src), None


class Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return add(a, b)

    @staticmethod
    def backward(ctx, g):
        return g, g


class Dropout(torch.autograd.Function):
    """Inverted dropout.  The keep mask is drawn by torch's generator (plumbing); apply/backward are kernels."""

    @staticmethod
    def forward(ctx, x, p: float):
        x = _c(x)
        mask = torch.empty(x.shape, dtype=torch.uint8, device=x.device).bernoulli_(1.0 - p)
        ctx.save_for_backward(mask)
        ctx.scale = 1.0 / (1.0 - p)
        out = torch.empty_like(x)
        L.call("egp_mask_scale", L.ptr(x), L.ptr(mask), L.ptr(out), x.numel(), ctx.scale, _code(x), L.stream())
        return out

    @staticmethod
    def backward(ctx, g):
        (mask,) = ctx.saved_tensors
        g = _c(g)
        out = torch.empty_like(g)
        L.call("egp_mask_scale", L.ptr(g), L.ptr(mask), L.ptr(out), g.numel(), ctx.scale, _code(g), L.stream())
        return out, None


def dropout(x: Tensor, p: float, training: bool) -> Tensor:
    if not training or p <= 0.0:
        return x
    if p >= 1.0:
        return x * 0
    return Dropout.apply(x, float(p))


# =====================================================================================================
# losses (a11)
# =====================================================================================================
def _logits_2d(l: Tensor) -> Tensor:
    if l.dtype != torch.float32:
        raise TypeError(f"loss kernels take fp32 logits (the classifier heads emit fp32), got {l.dtype}")
    if l.dim() != 2 or l.stride(1) != 1:
        l = l.reshape(l.shape[0], -1).contiguous()
    return l


class MultiHeadCrossEntropy(torch.autograd.Function):
    """sum_h CrossEntropy(logits_h, targets[:, h], ignore_index, label_smoothing, reduction='none') -> [N]
    (criterion/wrapper.py:80-82 around nn.CrossEntropyLoss, models/tasks/recognition.py:61-69, lta.py:73-74; one head
    with smoothing 0.1 is OSCC's loss, oscc.py:88-96).  ``targets`` is int64 [N, H] (or [N] for one head)."""

    @staticmethod
    def forward(ctx, targets, ignore_index: int, label_smoothing: float, *logits):
        targets = _i64(targets, "targets")
        n = logits[0].shape[0]
        heads = len(logits)
        tcols = targets.shape[1] if targets.dim() > 1 else 1
        if tcols < heads or targets.shape[0] != n:
            raise ValueError(f"targets {tuple(targets.shape)} do not match {heads} heads of {n} rows")
        dev = logits[0].device
        loss = torch.empty(n, dtype=torch.float32, device=dev)
        lses = torch.empty((heads, n), dtype=torch.float32, device=dev)
        ls = [_logits_2d(l) for l in logits]
        for h, l in enumerate(ls):
            L.call("egp_ce_loss_fwd", L.ptr(l), l.stride(0), L.ptr(targets) + 8 * h, tcols, n, l.shape[1], int(ignore_index),
                   float(label_smoothing), L.ptr(loss), int(h > 0), L.ptr(lses[h]), L.stream())
        ctx.save_for_backward(targets, lses, *ls)
        ctx.cfg = (int(ignore_index), float(label_smoothing), tcols)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        targets, lses, *ls = ctx.saved_tensors
        ignore_index, smoothing, tcols = ctx.cfg
        n = ls[0].shape[0]
        if dloss.dtype != torch.float32:
            dloss = dloss.float()
        gstride = dloss.stride(0) if dloss.dim() else 0        # an expanded scalar (stride 0) is read in place
        grads = []
        for h, l in enumerate(ls):
            c = l.shape[1]
            d = torch.empty((n, c), dtype=torch.float32, device=l.device)
            L.call("egp_ce_loss_bwd", L.ptr(l), l.stride(0), L.ptr(lses[h]), L.ptr(targets) + 8 * h, tcols, L.ptr(dloss),
                   gstride, n, c, ignore_index, smoothing, L.ptr(d), d.stride(0), F32, L.stream())
            grads.append(d)
        return (None, None, None, *grads)


class LinearCrossEntropy(torch.autograd.Function):
    """sum_h CrossEntropy(f W_h^T + b_h, targets[:, h]) -> per-sample loss [N]: the classifier heads of a multi-head task
    (recognition.py:28-37,39-49: ``Sequential(Dropout, Linear)`` per label head) fused with their loss
    (recognition.py:61-69) into one autograd node, for the training step.

    What the fusion buys in the backward: the loss kernel writes d(logits) directly as the zero-padded bf16 operand of
    the head's dgrad / wgrad GEMMs (no fp32 gradient round trip, no cast+pad kernel), the bias gradient is a column sum
    of that operand, and the heads' contributions to d(features) chain through the GEMM epilogue (the second head's
    dgrad takes the first one's result as its residual) instead of an elementwise add.  Logits are still materialised
    in fp32 (they are what the loss reads), so values equal the unfused path."""

    @staticmethod
    def forward(ctx, f, targets, ignore_index: int, label_smoothing: float, *wb):
        f = _c(f)
        cd = f.dtype
        n, k = f.shape
        ws, bs = wb[0::2], wb[1::2]
        heads = len(ws)
        targets = _i64(targets, "targets")
        tcols = targets.shape[1] if targets.dim() > 1 else 1
        if tcols < heads or targets.shape[0] != n:
            raise ValueError(f"targets {tuple(targets.shape)} do not match {heads} heads of {n} rows")
        loss = torch.empty(n, dtype=torch.float32, device=f.device)
        lses = torch.empty((heads, n), dtype=torch.float32, device=f.device)
        logits = []
        for h, (w, b) in enumerate(zip(ws, bs)):
            c = w.shape[0]
            cp = (c + 3) // 4 * 4                                    # 16-byte fp32 row pitch: TMA-store epilogue
            lg = torch.empty((n, cp), dtype=torch.float32, device=f.device)[:, :c]
            gemm(f, False, weight_cache.get(w, cd), False, n, c, k, bias=b, out_dtype=torch.float32, out=lg)
            L.call("egp_ce_loss_fwd", L.ptr(lg), lg.stride(0), L.ptr(targets) + 8 * h, tcols, n, c, int(ignore_index),
                   float(label_smoothing), L.ptr(loss), int(h > 0), L.ptr(lses[h]), L.stream())
            logits.append(lg)
        ctx.save_for_backward(f, targets, lses, *ws, *logits)
        ctx.cfg = (int(ignore_index), float(label_smoothing), tcols, heads, tuple(b is not None for b in bs))
        return loss

    @staticmethod
    def backward(ctx, dloss):
        ignore_index, smoothing, tcols, heads, has_b = ctx.cfg
        f, targets, lses, *rest = ctx.saved_tensors
        ws, logits = rest[:heads], rest[heads:]
        cd = f.dtype
        n, k = f.shape
        if dloss.dtype != torch.float32:
            dloss = dloss.float()
        gstride = dloss.stride(0) if dloss.dim() else 0
        df = None
        grads = []
        for h, (w, lg) in enumerate(zip(ws, logits)):
            c = w.shape[0]
            if cd == torch.bfloat16:
                cp = (c + 7) // 8 * 8
                g = torch.empty((n, cp), dtype=torch.bfloat16, device=f.device)
                L.call("egp_ce_loss_bwd", L.ptr(lg), lg.stride(0), L.ptr(lses[h]), L.ptr(targets) + 8 * h, tcols, L.ptr(dloss),
                       gstride, n, c, ignore_index, smoothing, L.ptr(g), cp, BF16, L.stream())
            else:
                cp = c
                g = torch.empty((n, c), dtype=torch.float32, device=f.device)
                L.call("egp_ce_loss_bwd", L.ptr(lg), lg.stride(0), L.ptr(lses[h]), L.ptr(targets) + 8 * h, tcols, L.ptr(dloss),
                       gstride, n, c, ignore_index, smoothing, L.ptr(g), c, F32, L.stream())
            db = None
            if has_b[h] and ctx.needs_input_grad[5 + 2 * h]:
                db = colsum(g)[:c]
            if ctx.needs_input_grad[0]:
                wc = weight_cache.get(w, cd, pad_rows=cp)
                df = gemm(g, False, wc, True, n, k, cp, residual=df)       # heads chain through the epilogue residual
            dw = gemm(g, True, f, True, c, k, n, out_dtype=torch.float32) if ctx.needs_input_grad[4 + 2 * h] else None
            grads += [dw, db]
        return (df, None, None, None, *grads)


def cross_entropy(logits, targets: Tensor, ignore_index: int = -100, label_smoothing: float = 0.0) -> Tensor:
    """Per-sample cross entropy (reduction='none'); ``logits`` is one fp32 [N,C] tensor or a tuple of heads."""
    if torch.is_tensor(logits):
        logits = (logits,)
    return MultiHeadCrossEntropy.apply(targets, ignore_index, label_smoothing, *logits)


class BCEWithLogits(torch.autograd.Function):
    """nn.BCEWithLogitsLoss(reduction='none') (models/tasks/pnr.py:38,82-83)."""

    @staticmethod
    def forward(ctx, z, target):
        if z.dtype != torch.float32:
            raise TypeError(f"loss kernels take fp32 logits, got {z.dtype}")
        z, target = _c(z), _c(target if target.dtype == torch.float32 else target.float())
        if z.shape != target.shape:
            raise ValueError(f"BCE: logits {tuple(z.shape)} vs targets {tuple(target.shape)}")
        loss = torch.empty_like(z)
        L.call("egp_bce_logits_fwd", L.ptr(z), L.ptr(target), L.ptr(loss), z.numel(), L.stream())
        ctx.save_for_backward(z, target)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        z, target = ctx.saved_tensors
        dz = torch.empty_like(z)
        flat_stride = 0 if (dloss.dim() == 0 or all(s == 0 for s in dloss.stride())) else 1
        g = dloss if flat_stride == 0 else _c(dloss)
        L.call("egp_bce_logits_bwd", L.ptr(z), L.ptr(target), L.ptr(g), flat_stride, L.ptr(dz), z.numel(), L.stream())
        return dz, None


def bce_with_logits(z: Tensor, target: Tensor) -> Tensor:
    return BCEWithLogits.apply(z, target)


class WeightedMeanSum(torch.autograd.Function):
    """total = sum_t w_t * mean(loss_t) (main_temporal.py:99-128: ``losses.append(w * loss.mean())`` ...
    ``torch.stack(losses).sum()``) as one accumulation chain; the backward hands every loss_t an EXPANDED scalar
    (stride 0), which the loss kernels read in place -- no [N] gradient is materialised."""

    @staticmethod
    def forward(ctx, weights, *losses):
        dev = losses[0].device
        out = torch.empty((), dtype=torch.float32, device=dev)
        flat = []
        for i, (w, l) in enumerate(zip(weights, losses)):
            if l.dtype != torch.float32:
                raise TypeError(f"per-sample losses are fp32, got {l.dtype}")
            l = _c(l).view(-1)
            flat.append(l.shape[0])
            L.call("egp_weighted_mean", L.ptr(l), l.shape[0], float(w), L.ptr(out), int(i > 0), L.stream())
        ctx.cfg = (tuple(float(w) for w in weights), tuple(flat), tuple(l.shape for l in losses))
        return out

    @staticmethod
    def backward(ctx, g):
        weights, counts, shapes = ctx.cfg
        outs = []
        for w, n, shp in zip(weights, counts, shapes):
            outs.append(axpby_scalar(g, w / max(n, 1)).expand(shp))
        return (None, *outs)


def axpby_scalar(g: Tensor, alpha: float) -> Tensor:
    """alpha * g for a 0-dim fp32 tensor (one tiny kernel; the result is expanded, never materialised per row)."""
    out = torch.empty((), dtype=torch.float32, device=g.device)
    gg = g if g.dtype == torch.float32 else g.float()
    L.call("egp_weighted_mean", L.ptr(gg), 1, float(alpha), L.ptr(out), 0, L.stream())
    return out


def weighted_mean_sum(losses, weights) -> Tensor:
    return WeightedMeanSum.apply(tuple(weights), *losses)


# =====================================================================================================
# pooling / GraphONE pieces
# =====================================================================================================
class SegmentMaxPool(torch.autograd.Function):
    """gnn.pool.global_max_pool (models/tasks/oscc.py:68,85): per-graph channel-wise max."""

    @staticmethod
    def forward(ctx, x, ptr, batch):
        x = _c(x)
        g, c = ptr.numel() - 1, x.shape[1]
        out = torch.empty((g, c), dtype=x.dtype, device=x.device)
        arg = torch.empty((g, c), dtype=torch.int32, device=x.device)
        with _Traced("segment_max_pool_fwd", 1.0 * x.shape[0] * c * x.element_size(), "B"):
            L.call("egp_segment_max_pool_fwd", L.ptr(x), L.ptr(_i64(ptr, "ptr")), L.ptr(out), L.ptr(arg), g, c, _code(x),
                   L.stream())
        ctx.save_for_backward(arg, _i64(batch, "batch"))
        ctx.n = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        arg, batch = ctx.saved_tensors
        dout = _c(dout)
        c = dout.shape[1]
        dx = torch.empty((ctx.n, c), dtype=dout.dtype, device=dout.device)
        L.call("egp_segment_max_pool_bwd", L.ptr(dout), L.ptr(arg), L.ptr(batch), L.ptr(dx), ctx.n, c, _code(dout),
               L.stream())
        return dx, None, None


class MaxCombine(torch.autograd.Function):
    """a = max(f, m) with m constant: the max-aggregation over {self} U {k nearest prototypes}."""

    @staticmethod
    def forward(ctx, f, m):
        f, m = _c(f), _c(m)
        a = torch.empty_like(f)
        with _Traced("max_combine_fwd", 3.0 * f.numel() * f.element_size(), "B"):
            L.call("egp_max_combine_fwd", L.ptr(f), L.ptr(m), L.ptr(a), f.numel(), _code(f), L.stream())
        ctx.save_for_backward(f, m)
        return a

    @staticmethod
    def backward(ctx, da):
        f, m = ctx.saved_tensors
        da = _c(da)
        df = torch.empty_like(f)
        L.call("egp_max_combine_bwd", L.ptr(da), L.ptr(f), L.ptr(m), L.ptr(df), f.numel(), _code(f), L.stream())
        return df, None


class ProtoMaxCombine(torch.autograd.Function):
    """a = max(f, max_j bank[idx[:, j]]) with a TRAINABLE bank (GraphONE(freeze=False)): the gradient reaches f where
    f wins and the arg-max prototype otherwise.  `bank_cd` is the compute-dtype copy of `bank` the forward reads."""

    @staticmethod
    def forward(ctx, f, bank, bank_cd, idx):
        f = _c(f)
        m = proto_max_gather(bank_cd, idx)
        a = torch.empty_like(f)
        L.call("egp_max_combine_fwd", L.ptr(f), L.ptr(m), L.ptr(a), f.numel(), _code(f), L.stream())
        ctx.save_for_backward(f, m, bank_cd, idx)
        ctx.bank_shape = tuple(bank.shape)
        return a

    @staticmethod
    def backward(ctx, da):
        f, m, bank_cd, idx = ctx.saved_tensors
        da = _c(da)
        df = torch.empty_like(f)
        L.call("egp_max_combine_bwd", L.ptr(da), L.ptr(f), L.ptr(m), L.ptr(df), f.numel(), _code(f), L.stream())
        dbank = None
        if ctx.needs_input_grad[1]:
            dbank = torch.zeros(ctx.bank_shape, dtype=torch.float32, device=f.device)
            L.call("egp_proto_max_scatter_bwd", L.ptr(da), L.ptr(f), L.ptr(bank_cd), L.ptr(_i64(idx, "idx")), L.ptr(dbank),
                   f.shape[0], idx.shape[1], f.shape[1], _code(f), L.stream())
        return df, dbank, None, None


def class_sum_f64(x: Tensor, labels: Tensor, num_classes: int, out: Optional[Tensor] = None) -> Tensor:
    """out[label[i]] += x[i] in fp64 (rows with label < 0 are skipped)."""
    x, labels = _c(x), _i64(labels, "labels")
    if out is None:
        out = torch.zeros((num_classes, x.shape[1]), dtype=torch.float64, device=x.device)
    L.call("egp_class_sum_f64", L.ptr(x), L.ptr(labels), L.ptr(out), x.shape[0], x.shape[1], num_classes, _code(x), L.stream())
    return out


def label_rank(logits: Tensor, labels: Tensor, ignore_index: int = -1) -> Tensor:
    """int32 [N]: how many classes beat the label's logit (ties: lower class index wins); -1 for ignored rows.
    ``0 <= rank < k`` is a top-k hit (utils/meters/ego4d.py MulticlassAccuracy(top_k=k, ignore_index=-1))."""
    assert logits.dim() == 2 and labels.dim() == 1 and labels.shape[0] == logits.shape[0]
    logits = logits.detach()
    if logits.dtype != torch.float32:
        logits = logits.float()
    if logits.stride(1) != 1:
        logits = logits.contiguous()
    labels = labels.detach()
    assert labels.dtype == torch.int64
    rank = torch.empty(logits.shape[0], dtype=torch.int32, device=logits.device)
    if logits.shape[0]:
        L.call("egp_label_rank", L.ptr(logits), logits.stride(0), L.ptr(labels), labels.stride(0), logits.shape[0],
               logits.shape[1], int(ignore_index), L.ptr(rank), L.stream())
    return rank


def segment_argmax(values: Tensor, ptr: Tensor, apply_sigmoid: bool = False) -> Tensor:
    """int64 [G]: first arg-max of (sigmoid of) values inside every graph [ptr[g], ptr[g+1]), relative to ptr[g]."""
    values, ptr = _c(values.detach().float()), _i64(ptr, "ptr")
    out = torch.empty(ptr.shape[0] - 1, dtype=torch.int64, device=values.device)
    if out.shape[0]:
        L.call("egp_segment_argmax", L.ptr(values), L.ptr(ptr), out.shape[0], int(apply_sigmoid), L.ptr(out), L.stream())
    return out


def edit_distance_min(preds: Tensor, labels: Tensor) -> Tensor:
    """int32 [N]: min over the K samples of Levenshtein(preds[n, :, k], labels[n, :]); preds int64 [N, Z, K]."""
    assert preds.dim() == 3 and labels.shape == preds.shape[:2]
    preds, labels = _c(preds.detach().long()), _c(labels.detach().long())
    out = torch.empty(preds.shape[0], dtype=torch.int32, device=preds.device)
    if preds.shape[0]:
        L.call("egp_edit_distance_min", L.ptr(preds), L.ptr(labels), preds.shape[0], preds.shape[1], preds.shape[2],
               L.ptr(out), L.stream())
    return out


def row_normalize(x: Tensor, out_dtype: torch.dtype = torch.float32, with_round_err: bool = False):
    """x / ||x||_2 per row.  ``with_round_err`` also returns float [rows]: the L2 distance between the stored (rounded)
    row and the exact normalised row -- what the k-NN miss detector adds to its bound for a bf16 operand."""
    x = _c(x)
    out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
    err = torch.empty(x.shape[0], dtype=torch.float32, device=x.device) if with_round_err else None
    L.call("egp_row_normalize", L.ptr(x), L.ptr(out), x.shape[0], x.shape[1], _code(x), L.DTYPE_CODE[out_dtype], L.ptr(err),
           L.stream())
    return (out, err) if with_round_err else out


# k-NN guard statistics (cumulative): rows scored through the tensor-core path / rows the miss detector sent to the
# exact fp32 path.  bench.py and the tests report them.
KNN_STATS = {"rows": 0, "flagged": 0, "calls": 0}


class PendingTopk:
    """A guarded tensor-core k-NN whose flagged-row count has not been read back yet.  ``resolve_all`` reads the counts
    of several pending calls with ONE host round trip (GraphONE.interact queries one bank per task) and re-ranks the
    flagged rows exactly."""

    def __init__(self, idx, fn, pn, k, rows, count):
        self.idx, self.fn, self.pn, self.k, self.rows, self.count = idx, fn, pn, k, rows, count

    @staticmethod
    def resolve_all(pending) -> None:
        live = [p for p in pending if p is not None and p.count is not None]
        if not live:
            return
        counts = torch.cat([p.count for p in live]).tolist() if len(live) > 1 else [int(live[0].count.item())]
        for p, n_flag in zip(live, counts):
            b, c = p.fn.shape
            kp = p.pn.shape[0]
            KNN_STATS["rows"] += b
            KNN_STATS["flagged"] += int(n_flag)
            KNN_STATS["calls"] += 1
            if n_flag:
                sel = p.rows[:n_flag].long()
                sub = p.fn.index_select(0, sel)
                nb2 = L.size("egp_cos_topk_workspace", n_flag, kp, p.k)
                ws2 = L.workspace(nb2, p.fn.device, "topk")
                exact = torch.empty((n_flag, p.k), dtype=torch.int64, device=p.fn.device)
                L.call("egp_cos_topk", L.ptr(sub), L.ptr(p.pn), None, None, n_flag, kp, c, int(p.k), L.ptr(exact), None, 0.0,
                       None, None, L.ptr(ws2), nb2, L.stream())
                p.idx.index_copy_(0, sel, exact)
            p.count = None


def cos_topk(fn: Tensor, pn: Tensor, k: int, fn16: Optional[Tensor] = None, pn16: Optional[Tensor] = None,
             f_err: Optional[Tensor] = None, p_err: Optional[float] = None, guard: bool = True, defer: bool = False):
    """k nearest prototypes by cosine dissimilarity for every row of the NORMALISED fp32 features ``fn``
    (GraphONE.__compute_edges, graphONE.py:119-141): exact fp32 ranking, ties -> lower prototype index.

    With bf16 copies the similarity runs on the tensor cores and only the candidates kept in the GEMM epilogue are
    re-scored in fp32.  ``guard`` makes that safe: rows where bf16 rounding could have kept a true neighbour out of the
    candidate set (bound from the measured rounding errors ``f_err`` [B] and ``p_err`` = max over the bank) are
    re-run through the exact fp32 path.  Reading the flagged count is one host round trip; ``defer=True`` returns
    ``(idx, pending)`` instead so that several calls share it (``PendingTopk.resolve_all``); ``guard=False`` (or
    CUDA-graph capture) skips the detector."""
    fn, pn = _c(fn), _c(pn)
    assert fn.dtype == torch.float32 and pn.dtype == torch.float32
    b, c = fn.shape
    kp = pn.shape[0]
    idx = torch.empty((b, k), dtype=torch.int64, device=fn.device)
    nb = L.size("egp_cos_topk_workspace", b, kp, k)
    ws = L.workspace(nb, fn.device, "topk")
    tensor_path = fn16 is not None and pn16 is not None
    guard = guard and tensor_path and b > 0 and not torch.cuda.is_current_stream_capturing()
    rows = count = None
    if guard:
        rows = torch.empty(b, dtype=torch.int32, device=fn.device)
        count = torch.empty(1, dtype=torch.int32, device=fn.device)
    with _Traced("cos_topk", 2.0 * b * kp * c, "FLOP", f"{b}x{kp}x{c} k={k}"):
        L.call("egp_cos_topk", L.ptr(fn), L.ptr(pn), L.ptr(_c(fn16)), L.ptr(_c(pn16)), b, kp, c, int(k), L.ptr(idx),
               L.ptr(_c(f_err)), float(p_err if p_err is not None else 2.0 ** -8), L.ptr(rows), L.ptr(count),
               L.ptr(ws), nb, L.stream())
    pending = PendingTopk(idx, fn, pn, k, rows, count) if guard else None
    if defer:
        return idx, pending
    PendingTopk.resolve_all([pending])
    return idx


def proto_max_gather(bank: Tensor, idx: Tensor) -> Tensor:
    bank, idx = _c(bank), _i64(idx, "idx")
    b, k = idx.shape
    c = bank.shape[1]
    m = torch.empty((b, c), dtype=bank.dtype, device=bank.device)
    with _Traced("proto_max_gather", 1.0 * b * c * bank.element_size() + 8.0 * b * k, "B"):   # write m, read idx (bank: L2)
        L.call("egp_proto_max_gather", L.ptr(bank), L.ptr(idx), L.ptr(m), b, k, c, _code(bank), _code(bank), L.stream())
    return m
