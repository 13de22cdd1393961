"""Drop-in for ``models/transforms/lta_temp_connectivity.py:12-56`` (reference): temporal band plus directed edges
from the last ``floor(r)`` input clips to every forecast clip, de-duplicated and sorted by (src, dst).

Differences from the reference, both deliberate: the edges are built on the device by ``egp_lta_edge_*`` (the
reference uses torch_cluster's CPU KD-tree), and a batched ``Data`` is accepted (the reference refuses it at
:31-32 only because its arange-based construction assumes one graph).  ``strict=True`` restores the refusal.
The ``y[:, 0] > 0`` forecast count (:50) -- which skips verb label 0 -- is reproduced as is.
"""
from __future__ import annotations

import math

import torch

from ... import ops
from .radius_graph import _ptr_of, _supports_lazy, unit_spaced_hint


class LTATemporalConnectivity:
    """``lazy=True`` (default): for a GPU-resident batch with unit-spaced positions and radius <= 4 the transform only
    records ``band_k`` and the per-graph star descriptor ``star`` (int32 [G,3] = n_in, n_fc, first_src -- counted on the
    device, no host round trip); the band+star aggregation kernels work from those, and ``edge_index`` is a lazy
    attribute built on first read."""

    def __init__(self, r: float, loop: bool = False, max_num_neighbors: int = 32, flow: str = "source_to_target",
                 num_workers: int = 1, strict: bool = False, device=None, lazy: bool = True):
        if loop or flow != "source_to_target":
            raise NotImplementedError("the reference only uses loop=False, flow='source_to_target'")
        self.r, self.loop, self.max_num_neighbors, self.flow, self.num_workers = r, loop, max_num_neighbors, flow, num_workers
        self.strict, self.device, self.lazy = strict, device, lazy

    def _edges(self, data):
        pos = data.pos.view(-1)
        home = pos.device
        dev = torch.device(self.device) if self.device is not None else (home if home.type == "cuda" else torch.device("cuda"))
        batch, ptr = _ptr_of(data, pos.numel(), home)
        y = data.y if data.y.dim() > 1 else data.y.view(-1, 1)
        return ops.lta_edge_index(pos.to(dev), y.to(dev), batch.to(dev), ptr.to(dev), self.r, self.max_num_neighbors).to(home)

    def __call__(self, data):
        if self.strict and data.batch is not None:
            raise ValueError("This transform expects no batched graphs.")
        data.edge_attr = None
        pos = data.pos.view(-1)
        home = pos.device
        k = int(math.floor(self.r))
        data.band_k = None
        data.star = None
        band_shape = 2 * k + 1 <= self.max_num_neighbors + 1 and self.r > k and k <= 4
        if band_shape and home.type == "cuda":
            batch, ptr = _ptr_of(data, pos.numel(), home)
            if unit_spaced_hint(data, pos, batch):
                y = data.y if data.y.dim() > 1 else data.y.view(-1, 1)
                data.star = ops.lta_star_counts(y, ptr, self.r)
                data.band_k = k
                if self.lazy and _supports_lazy(data):
                    data.set_lazy("edge_index", self._edges)
                    return data
        data.edge_index = self._edges(data)
        return data
