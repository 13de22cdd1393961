"""Parameter containers that reproduce the reference's ``state_dict`` layout for the third-party layers it uses
(torch_geometric 2.3.0 ``SAGEConv`` / ``LayerNorm`` / ``PositionalEncoding`` / ``Sequential`` / ``Linear``), plus
their kernel-backed forwards.  Nothing here depends on torch_geometric.
"""
from __future__ import annotations


import torch
import torch.nn as nn
from torch import Tensor

from .. import ops
from ..ops import ACT_NONE, ACT_RELU, GraphStructure


def _batch_ptr(data, n: int, dev):
    batch = data.batch if getattr(data, "batch", None) is not None else torch.zeros(n, dtype=torch.long, device=dev)
    ptr = getattr(data, "ptr", None)
    if ptr is None:
        counts = torch.bincount(batch)
        ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    return batch, ptr


def structure_for(data, n: int) -> GraphStructure:
    """Aggregation structure of a batch, cached on the batch object.

    ``band_k`` (set by our RadiusGraph / LTATemporalConnectivity when positions are unit spaced) selects the
    sliding-window kernels -- with ``star`` (LTA) their band+star variants; any other ``edge_index`` (graphs built by
    real PyG, stars wider than radius 4) goes through a deterministic CSR.
    """
    cached = getattr(data, "_egp_structure", None)
    dev = data.pos.device if getattr(data, "pos", None) is not None else data.x.device
    if cached is not None and cached.n == n and cached.inv_deg.device == dev:
        return cached
    band_k = getattr(data, "band_k", None)
    if band_k is not None:
        batch, ptr = _batch_ptr(data, n, dev)
        gs = ops.band_structure(batch, ptr, int(band_k), getattr(data, "star", None))
    else:
        gs = ops.csr_structure(data.edge_index, n)
    try:
        data._egp_structure = gs
    except Exception:  # foreign batch types may refuse new attributes; rebuilding per call is still correct
        pass
    return gs


def structure_for_many(batches, sizes) -> GraphStructure:
    """One band(+star) structure over several batches laid out back to back (``Graph.forward_many``).  Requires every
    batch to carry the same ``band_k`` hint; returns None otherwise (the caller then runs the batches one by one)."""
    ks = {getattr(b, "band_k", None) for b in batches}
    if len(ks) != 1 or None in ks:
        return None
    k = int(ks.pop())
    if k > 4 and any(getattr(b, "star", None) is not None for b in batches):
        return None
    parts = []
    for b, n in zip(batches, sizes):
        dev = b.pos.device if getattr(b, "pos", None) is not None else b.x.device
        batch, ptr = _batch_ptr(b, n, dev)
        parts.append((batch, ptr, getattr(b, "star", None)))
    return ops.band_structure_many(parts, k)


class PositionalEncoding(nn.Module):
    """gnn.PositionalEncoding(out_channels): buffer ``frequency = logspace(0, 1, C/2, base=1e-4)``."""

    def __init__(self, out_channels: int, base_freq: float = 1e-4, granularity: float = 1.0):
        super().__init__()
        if out_channels % 2:
            raise ValueError(f"Cannot use sinusoidal positional encoding with odd 'out_channels' (got {out_channels}).")
        self.out_channels, self.base_freq, self.granularity = out_channels, base_freq, granularity
        self.register_buffer("frequency", torch.logspace(0, 1, out_channels // 2, base_freq))

    def add_to(self, x: Tensor, pos: Tensor) -> Tensor:
        if self.granularity != 1.0:
            raise NotImplementedError("the reference uses the default granularity of 1.0")
        return ops.PosEncAdd.apply(x, pos, self.frequency)


class SAGEConv(nn.Module):
    """Parameters of gnn.SAGEConv: ``lin`` (only if project, with bias), ``lin_l`` (bias = ctor bias), ``lin_r``
    (no bias).  Forward (mean aggregation): ``lin_l(mean_j relu(lin(x_j))) + lin_r(x)`` -- the projection ReLU
    is the GEMM epilogue, and ``lin_l``/``lin_r`` accumulate into one TMEM tile."""

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "mean", project: bool = False, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels, self.aggr, self.project = in_channels, out_channels, aggr, project
        if project:
            self.lin = nn.Linear(in_channels, in_channels, bias=True)
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_r = nn.Linear(in_channels, out_channels, bias=False)

    @property
    def fusable(self) -> bool:
        """One autograd node for the whole layer (ops.SageLayer): mean aggregation with projection, 16-byte rows."""
        return self.aggr == "mean" and self.project and self.in_channels % 8 == 0 and self.out_channels % 8 == 0

    def forward(self, x: Tensor, gs: GraphStructure) -> Tensor:
        if self.aggr != "mean":
            raise NotImplementedError("max aggregation is implemented by GraphONE's reduced stage")
        if self.fusable:
            return ops.SageLayer.apply(x, self.lin.weight, self.lin.bias, self.lin_l.weight, self.lin_l.bias,
                                       self.lin_r.weight, gs)
        xs = ops.linear(x, self.lin.weight, self.lin.bias, act=ACT_RELU) if self.project else x
        agg = ops.SageMean.apply(xs, gs)
        return ops.linear(agg, self.lin_l.weight, self.lin_l.bias, x2=x, w2=self.lin_r.weight)


class GraphLayerNorm(nn.Module):
    """Parameters of gnn.LayerNorm(C) (``weight``, ``bias``); graph-mode statistics over the whole call."""

    def __init__(self, in_channels: int, eps: float = 1e-5):
        super().__init__()
        self.in_channels, self.eps = in_channels, eps
        self.weight = nn.Parameter(torch.ones(in_channels))
        self.bias = nn.Parameter(torch.zeros(in_channels))

    def forward(self, x: Tensor, act: int = ACT_NONE, slope: float = 0.0, seg_rows=None) -> Tensor:
        return ops.GraphLayerNorm.apply(x, self.weight, self.bias, self.eps, act, slope, seg_rows)


def row_layernorm(ln: nn.LayerNorm, x: Tensor, act: int = ACT_NONE, dropout_p: float = 0.0) -> Tensor:
    """LayerNorm (+ReLU) (+the Dropout that follows, when training) in one kernel."""
    if dropout_p >= 1.0:
        return ops.RowLayerNorm.apply(x, ln.weight, ln.bias, ln.eps, act, 0.0) * 0
    return ops.RowLayerNorm.apply(x, ln.weight, ln.bias, ln.eps, act, float(dropout_p))
