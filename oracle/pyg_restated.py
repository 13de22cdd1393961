"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the third-party arithmetic EgoPack's hot path calls.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this module; the product package ``egopack_b200`` never does.

The reference (sapeirone/EgoPack) delegates its numerics to three libraries that are NOT vendored under
``/root/reference`` and are NOT installable in this image (no network, not in /opt/wheelhouse):

* torch_geometric 2.3.0      (reference ``environment.yml:174``)
* torch_cluster   1.6.1      (reference ``environment.yml:184``)
* torch_scatter   2.1.1      (reference ``environment.yml:188``)

so the published algorithm of every symbol the path touches is restated here in plain PyTorch (CPU, fp32 /
fp64), issuing the same aten op sequence those pinned versions use on CPU.  **Parity unpinned**: the reference
ships no tests, golden vectors or fixtures, and PyG cannot be imported here, so nothing upstream
pins these restatements; they are anchored on the reference's own call sites (cited per symbol) and on
hand-derived known-answer cases in ``tests/test_oracle_*.py``.  The first-party model code that sits on top
of these symbols IS exercised for real: ``tests/golden/make_golden.py`` imports ``/root/reference/models/*``
with this module standing in for ``torch_geometric`` and stores the outputs as golden fixtures.

Assumption tags A1..A9 follow SURVEY.md §8c.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor


# --------------------------------------------------------------------------------------------------
# torch_geometric.utils.scatter  (A2)           call sites: graphone.py:53, SAGEConv/MaxAggregation, pooling
# --------------------------------------------------------------------------------------------------
def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = "sum") -> Tensor:
    """PyG 2.3.0 ``utils.scatter`` (torch>=2 branch, CPU): sum/mean via ``scatter_add_``; min/max via
    ``new_zeros(size).scatter_reduce_(..., 'amax', include_self=False)`` (empty group -> 0)."""
    if index.dim() != 1:
        raise ValueError("index must be one-dimensional")
    dim = src.dim() + dim if dim < 0 else dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    size = list(src.size())
    size[dim] = dim_size
    shape = [1] * src.dim()
    shape[dim] = -1
    idx = index.view(shape).expand_as(src)
    if reduce in ("sum", "add"):
        return src.new_zeros(size).scatter_add_(dim, idx, src)
    if reduce == "mean":
        count = src.new_zeros(dim_size)
        count.scatter_add_(0, index, src.new_ones(src.size(dim)))
        count = count.clamp(min=1)
        out = src.new_zeros(size).scatter_add_(dim, idx, src)
        return out / count.view(shape)
    if reduce in ("min", "max", "amin", "amax"):
        red = "a" + reduce[-3:]
        return src.new_zeros(size).scatter_reduce_(dim, idx, src, reduce=red, include_self=False)
    raise ValueError(f"unsupported reduce {reduce}")


# --------------------------------------------------------------------------------------------------
# torch_geometric.nn.Linear (A5)                 call site: models/graphONE/graphONE.py:63
# --------------------------------------------------------------------------------------------------
class Linear(torch.nn.Module):
    """``gnn.Linear`` = ``F.linear``; kaiming-uniform(a=sqrt(5)) weight, U(+-1/sqrt(fan_in)) bias."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True,
                 weight_initializer: Optional[str] = None, bias_initializer: Optional[str] = None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    @property
    def in_features(self):
        return self.in_channels

    @property
    def out_features(self):
        return self.out_channels

    def reset_parameters(self):
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0
            torch.nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x: Tensor) -> Tensor:
        return F.linear(x, self.weight, self.bias)


# --------------------------------------------------------------------------------------------------
# torch_geometric.nn.SAGEConv (A1)               call sites: models/graph.py:42, graphONE.py:60
# --------------------------------------------------------------------------------------------------
class SAGEConv(torch.nn.Module):
    """``lin`` (with bias) exists only if ``project``; ``lin_l`` bias = ctor ``bias``; ``lin_r`` has no bias.
    message = x_j with j = edge_index[0], aggregated at i = edge_index[1] (flow source_to_target)."""

    def __init__(self, in_channels: int, out_channels: int, aggr: str = "mean", normalize: bool = False,
                 root_weight: bool = True, project: bool = False, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.aggr, self.normalize, self.root_weight, self.project = aggr, normalize, root_weight, project
        if project:
            self.lin = Linear(in_channels, in_channels, bias=True)
        self.lin_l = Linear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = Linear(in_channels, out_channels, bias=False)

    def forward(self, x: Tensor, edge_index: Tensor) -> Tensor:
        x_src, x_dst = x, x
        if self.project:
            x_src = self.lin(x_src).relu()
        src, dst = edge_index[0], edge_index[1]
        msg = x_src.index_select(0, src)                                  # [E, C] materialised, as PyG does
        out = scatter(msg, dst, dim=0, dim_size=x_dst.size(0), reduce=self.aggr)
        out = self.lin_l(out)
        if self.root_weight:
            out = out + self.lin_r(x_dst)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        return out


# --------------------------------------------------------------------------------------------------
# torch_geometric.nn.LayerNorm (A3)              call site: models/graph.py:43 ("x -> x": batch=None)
# --------------------------------------------------------------------------------------------------
class LayerNorm(torch.nn.Module):
    def __init__(self, in_channels: int, eps: float = 1e-5, affine: bool = True, mode: str = "graph"):
        super().__init__()
        self.in_channels, self.eps, self.mode = in_channels, eps, mode
        if affine:
            self.weight = torch.nn.Parameter(torch.ones(in_channels))
            self.bias = torch.nn.Parameter(torch.zeros(in_channels))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x: Tensor, batch: Optional[Tensor] = None, batch_size: Optional[int] = None) -> Tensor:
        if self.mode == "graph":
            if batch is None:
                x = x - x.mean()
                out = x / (x.std(unbiased=False) + self.eps)
            else:
                if batch_size is None:
                    batch_size = int(batch.max()) + 1
                norm = torch.bincount(batch, minlength=batch_size).clamp_(min=1).to(x.dtype)
                norm = norm.mul_(x.size(-1)).view(-1, 1)
                mean = scatter(x, batch, 0, batch_size, "sum").sum(dim=-1, keepdim=True) / norm
                x = x - mean.index_select(0, batch)
                var = scatter(x * x, batch, 0, batch_size, "sum").sum(dim=-1, keepdim=True) / norm
                out = x / (var + self.eps).sqrt().index_select(0, batch)
            if self.weight is not None and self.bias is not None:
                out = out * self.weight + self.bias
            return out
        if self.mode == "node":
            return F.layer_norm(x, (self.in_channels,), self.weight, self.bias, self.eps)
        raise ValueError(self.mode)


# --------------------------------------------------------------------------------------------------
# torch_geometric.nn.PositionalEncoding (A4)     call site: models/graph.py:37,63
# --------------------------------------------------------------------------------------------------
class PositionalEncoding(torch.nn.Module):
    def __init__(self, out_channels: int, base_freq: float = 1e-4, granularity: float = 1.0):
        super().__init__()
        if out_channels % 2 != 0:
            raise ValueError("out_channels must be even")
        self.out_channels, self.base_freq, self.granularity = out_channels, base_freq, granularity
        self.register_buffer("frequency", torch.logspace(0, 1, out_channels // 2, base_freq))

    def forward(self, x: Tensor) -> Tensor:
        x = x / self.granularity if self.granularity != 1.0 else x
        out = x.view(-1, 1) * self.frequency.view(1, -1)
        return torch.cat([torch.sin(out), torch.cos(out)], dim=-1)


class TemporalEncoding(torch.nn.Module):  # imported (unused) by models/temporal_pooling/pooling.py:2
    def __init__(self, out_channels: int):
        super().__init__()
        self.out_channels = out_channels
        sqrt = math.sqrt(out_channels)
        weight = 1.0 / 10 ** torch.linspace(0, 9, out_channels - 1).view(1, -1)
        self.register_buffer("sqrt", torch.tensor(sqrt))
        self.register_buffer("weight", weight)

    def forward(self, x: Tensor) -> Tensor:
        return (1.0 / self.sqrt) * torch.cat([torch.cos(x.view(-1, 1) @ self.weight), torch.zeros(x.numel(), 1)], -1)


# --------------------------------------------------------------------------------------------------
# torch_geometric.nn.Sequential (A6)             call sites: models/graph.py:48, graphONE.py:66-71
# --------------------------------------------------------------------------------------------------
class Sequential(torch.nn.Module):
    """Children are registered as ``module_{i}`` (state_dict prefix).  Signature strings ``"a, b -> c"``."""

    def __init__(self, input_args: str, modules: Sequence[Union[Tuple[torch.nn.Module, str], torch.nn.Module]]):
        super().__init__()
        self._in = [a.strip() for a in input_args.split(",")]
        self._plan: List[Tuple[str, List[str], List[str]]] = []
        last_out = [self._in[0]]
        for i, m in enumerate(modules):
            if isinstance(m, (tuple, list)):
                mod, desc = m
                lhs, rhs = [s.strip() for s in desc.split("->")]
                ins = [a.strip() for a in lhs.split(",")]
                outs = [a.strip() for a in rhs.split(",")]
            else:
                mod, ins, outs = m, list(last_out), list(last_out)
            name = f"module_{i}"
            setattr(self, name, mod)
            self._plan.append((name, ins, outs))
            last_out = outs

    def forward(self, *args):
        env: Dict[str, object] = dict(zip(self._in, args))
        out = None
        for name, ins, outs in self._plan:
            out = getattr(self, name)(*[env[a] for a in ins])
            if len(outs) == 1:
                env[outs[0]] = out
            else:
                for k, v in zip(outs, out):
                    env[k] = v
        return out


# --------------------------------------------------------------------------------------------------
# torch_geometric.nn.pool.global_max_pool (A9)   call sites: models/tasks/oscc.py:68,85
# --------------------------------------------------------------------------------------------------
def global_max_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.max(dim=-2, keepdim=x.dim() <= 2)[0]
    size = int(batch.max().item() + 1) if size is None else size
    return scatter(x, batch, dim=-2, dim_size=size, reduce="max")


class _Pool:
    global_max_pool = staticmethod(global_max_pool)


pool = _Pool()


# --------------------------------------------------------------------------------------------------
# torch_cluster.radius_graph via torch_geometric.nn.radius_graph (A7)
#   call sites: main_temporal.py:168 (RadiusGraph transform), lta_temp_connectivity.py:37
# --------------------------------------------------------------------------------------------------
def radius_graph(x: Tensor, r: float, batch: Optional[Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32, flow: str = "source_to_target", num_workers: int = 1) -> Tensor:
    """Edges (j -> i) for every pair in the same graph with strict ``|x_i-x_j|^2 < r^2``.

    torch_cluster semantics: for every centre i at most ``max_num_neighbors (+1 if not loop)`` matches are
    kept *before* the self match is removed.  The CPU KD-tree visiting order is unspecified; this oracle keeps
    the first matches in ascending node index (the CUDA variant's order), which only matters when a node has
    more than 33 matches (k > 16 for unit-spaced positions).  Output is centre-major, neighbour ascending --
    parity tests compare after a lexsort by (dst, src).
    """
    assert flow in ("source_to_target", "target_to_source")
    x = x.view(-1, 1) if x.dim() == 1 else x
    n = x.size(0)
    xf = x.to(torch.float32)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    rows: List[Tensor] = []
    cols: List[Tensor] = []
    # graphs are contiguous blocks in every caller; handle each block with a dense distance matrix
    if n > 0:
        counts = torch.bincount(batch)
        start = 0
        for c in counts.tolist():
            if c == 0:
                continue
            xs = xf[start:start + c]
            d2 = ((xs[:, None, :] - xs[None, :, :]) ** 2).sum(-1)
            hit = d2 < (r * r)                                          # [centre i, neighbour j]
            rank = hit.cumsum(dim=1)                                    # 1-based rank of each match
            hit = hit & (rank <= cap)
            ci, nj = hit.nonzero(as_tuple=True)
            rows.append(ci + start)
            cols.append(nj + start)
            start += c
    centre = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.long)
    neigh = torch.cat(cols) if cols else torch.zeros(0, dtype=torch.long)
    if flow == "source_to_target":
        row, col = neigh, centre
    else:
        row, col = centre, neigh
    if not loop:
        mask = row != col
        row, col = row[mask], col[mask]
    return torch.stack([row, col], dim=0)


# --------------------------------------------------------------------------------------------------
# torch_geometric.utils.add_remaining_self_loops (A8)      call site: graphONE.py:109
# --------------------------------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                             fill_value=None, num_nodes: Optional[int] = None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    mask = edge_index[0] != edge_index[1]
    loop_index = torch.arange(0, n, dtype=torch.long, device=edge_index.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    edge_index = torch.cat([edge_index[:, mask], loop_index], dim=1)
    return edge_index, None


def coalesce(edge_index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    """``RemoveDuplicatedEdges`` == ``coalesce``: sort by (row, col), drop duplicates
    (call site: lta_temp_connectivity.py:56)."""
    n = int(edge_index.max()) + 1 if num_nodes is None and edge_index.numel() else (num_nodes or 0)
    key = edge_index[0] * max(n, 1) + edge_index[1]
    key = torch.unique(key, sorted=True)
    return torch.stack([key // max(n, 1), key % max(n, 1)], dim=0)


# --------------------------------------------------------------------------------------------------
# torch_geometric.data.Data / Batch (A9)
# --------------------------------------------------------------------------------------------------
class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.__dict__["_store"] = {}
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos, **kwargs).items():
            if v is not None:
                self._store[k] = v

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self.__dict__["_store"].get(k, None)

    def __setattr__(self, k, v):
        if v is None:
            self._store.pop(k, None)
        else:
            self._store[k] = v

    def __contains__(self, k):
        return k in self._store

    def keys(self):
        return list(self._store.keys())

    @property
    def num_nodes(self):
        return self._store["x"].size(0) if "x" in self._store else int(self._store["pos"].size(0))

    def to(self, device, non_blocking: bool = False):
        for k, v in list(self._store.items()):
            if torch.is_tensor(v):
                self._store[k] = v.to(device, non_blocking=non_blocking)
        return self


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        out = cls()
        keys = data_list[0].keys()
        offs, batch, ptr = 0, [], [0]
        cat: Dict[str, List] = {k: [] for k in keys}
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = getattr(d, k)
                if k == "edge_index":
                    cat[k].append(v + offs)
                elif torch.is_tensor(v):
                    cat[k].append(v.unsqueeze(0) if v.dim() == 0 else v)
                else:
                    cat[k].append(v)
            batch.append(torch.full((n,), g, dtype=torch.long))
            offs += n
            ptr.append(offs)
        for k in keys:
            if k == "edge_index":
                out.edge_index = torch.cat(cat[k], dim=-1)
            elif torch.is_tensor(cat[k][0]):
                setattr(out, k, torch.cat(cat[k], dim=0))
            else:
                setattr(out, k, cat[k])
        out.batch = torch.cat(batch)
        out.ptr = torch.tensor(ptr, dtype=torch.long)
        return out


class RadiusGraph:
    """``torch_geometric.transforms.RadiusGraph`` (call site main_temporal.py:168)."""

    def __init__(self, r: float, loop: bool = False, max_num_neighbors: int = 32,
                 flow: str = "source_to_target", num_workers: int = 1):
        self.r, self.loop, self.max_num_neighbors, self.flow = r, loop, max_num_neighbors, flow

    def __call__(self, data):
        data.edge_attr = None
        batch = data.batch if "batch" in data else None
        data.edge_index = radius_graph(data.pos, self.r, batch, self.loop,
                                       max_num_neighbors=self.max_num_neighbors, flow=self.flow)
        return data


class RemoveDuplicatedEdges:
    def __call__(self, data):
        data.edge_index = coalesce(data.edge_index, data.num_nodes)
        return data
