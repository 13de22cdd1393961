"""Drop-in for ``models/transforms/lta_temp_connectivity.py:12-56`` (reference): temporal band plus directed edges
from the last ``floor(r)`` input clips to every forecast clip, de-duplicated and sorted by (src, dst).

Differences from the reference, both deliberate: the edges are built on the device by ``egp_lta_edge_*`` (the
reference uses torch_cluster's CPU KD-tree), and a batched ``Data`` is accepted (the reference refuses it at
:31-32 only because its arange-based construction assumes one graph).  ``strict=True`` restores the refusal.
The ``y[:, 0] > 0`` forecast count (:50) -- which skips verb label 0 -- is reproduced as is.
"""
from __future__ import annotations

import torch

from ... import ops
from .radius_graph import _ptr_of


class LTATemporalConnectivity:
    def __init__(self, r: float, loop: bool = False, max_num_neighbors: int = 32, flow: str = "source_to_target",
                 num_workers: int = 1, strict: bool = False, device=None):
        if loop or flow != "source_to_target":
            raise NotImplementedError("the reference only uses loop=False, flow='source_to_target'")
        self.r, self.loop, self.max_num_neighbors, self.flow, self.num_workers = r, loop, max_num_neighbors, flow, num_workers
        self.strict, self.device = strict, device

    def __call__(self, data):
        if self.strict and data.batch is not None:
            raise ValueError("This transform expects no batched graphs.")
        data.edge_attr = None
        pos = data.pos.view(-1)
        home = pos.device
        dev = torch.device(self.device) if self.device is not None else (home if home.type == "cuda" else torch.device("cuda"))
        batch, ptr = _ptr_of(data, pos.numel(), home)
        y = data.y if data.y.dim() > 1 else data.y.view(-1, 1)
        edge_index = ops.lta_edge_index(pos.to(dev), y.to(dev), batch.to(dev), ptr.to(dev), self.r, self.max_num_neighbors)
        data.edge_index = edge_index.to(home)
        data.band_k = None                                           # star edges: not a pure band
        return data
