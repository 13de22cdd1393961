"""Band ``edge_index`` construction -- the drop-in for ``torch_geometric.transforms.RadiusGraph`` as the reference
uses it: ``RadiusGraph(r=cfg.k + 0.5, loop=False)`` (main_temporal.py:168-169, main_egopack.py:196-197).

The reference runs torch_cluster's CPU KD-tree per sample inside DataLoader workers; here the edges of a whole
batch are produced by one count + scan + fill kernel sequence on the device.  When positions are unit-spaced
integers (every Ego4D dataset: data/ego4d_fho.py:224,307, data/ego4d_oscc.py:223,296) the transform also
records ``band_k`` on the data object so ``Graph`` can use the sliding-window aggregation kernel instead of
the generic CSR one.
"""
from __future__ import annotations

import math

import torch

from ... import ops


def _ptr_of(data, n: int, device):
    if getattr(data, "ptr", None) is not None and getattr(data, "batch", None) is not None:
        return data.batch, data.ptr
    if getattr(data, "batch", None) is not None:                   # real PyG batch without ptr
        counts = torch.bincount(data.batch)
        return data.batch, torch.cat([counts.new_zeros(1), counts.cumsum(0)])
    return torch.zeros(n, dtype=torch.long, device=device), torch.tensor([0, n], dtype=torch.long, device=device)


def unit_spaced(pos: torch.Tensor, batch: torch.Tensor) -> bool:
    """True when pos increases by exactly 1 inside every graph (then the radius graph is an index band).
    On a CUDA tensor this reads one flag back (a host sync); the feeder records the answer on the HOST copy before the
    upload (``data.pos_unit_spaced``) so the device-side transforms never have to."""
    if pos.numel() < 2:
        return True
    d = pos[1:] - pos[:-1]
    return bool(((d == 1) | (batch[1:] != batch[:-1])).all().item())


def unit_spaced_hint(data, pos: torch.Tensor, batch: torch.Tensor) -> bool:
    hint = getattr(data, "pos_unit_spaced", None)
    if hint is not None:
        return bool(hint)
    return unit_spaced(pos, batch)


def _supports_lazy(data) -> bool:
    return callable(getattr(data, "set_lazy", None))


class RadiusGraph:
    """``lazy=True`` (default): when the batch lives on the GPU and its positions are unit spaced, only the structural
    hint ``band_k`` is recorded and ``edge_index`` is registered as a lazy attribute -- it is built (count + scan + fill,
    one host read of the edge count) the first time somebody reads it.  The aggregation kernels never do."""

    def __init__(self, r: float, loop: bool = False, max_num_neighbors: int = 32, flow: str = "source_to_target",
                 num_workers: int = 1, device=None, lazy: bool = True):
        if loop:
            raise NotImplementedError("the reference only builds loop-free radius graphs")
        if flow != "source_to_target":
            raise NotImplementedError("only flow='source_to_target' is used by the reference")
        self.r, self.loop, self.max_num_neighbors, self.flow = r, loop, max_num_neighbors, flow
        self.device, self.lazy = device, lazy

    def _edges(self, data):
        pos = data.pos.view(-1)
        home = pos.device
        dev = torch.device(self.device) if self.device is not None else (home if home.type == "cuda" else torch.device("cuda"))
        batch, ptr = _ptr_of(data, pos.numel(), home)
        return ops.band_edge_index(pos.to(dev), batch.to(dev), ptr.to(dev), self.r, self.max_num_neighbors).to(home)

    def __call__(self, data):
        data.edge_attr = None
        pos = data.pos.view(-1)
        home = pos.device
        k = int(math.floor(self.r))
        band_shape = 2 * k + 1 <= self.max_num_neighbors + 1 and self.r > k
        if self.lazy and band_shape and home.type == "cuda" and _supports_lazy(data):
            batch, _ = _ptr_of(data, pos.numel(), home)
            if unit_spaced_hint(data, pos, batch):
                data.band_k = k                                      # hint consumed by Graph.forward
                data.set_lazy("edge_index", self._edges)
                return data
        data.edge_index = self._edges(data)
        if band_shape:
            batch, _ = _ptr_of(data, pos.numel(), home)
            if unit_spaced_hint(data, pos, batch):                   # checked where the data lives
                data.band_k = k
        return data

    def __repr__(self):
        return f"{type(self).__name__}(r={self.r})"
