"""Minimal stand-ins for ``torch_geometric.data.Data`` / ``Batch`` (PyG is not a dependency of this package).

The models only duck-type their input (``.x .pos .batch .ptr .edge_index .y``), so real PyG batches work too.
``Batch.from_data_list`` follows PyG's collate for the attributes the reference uses: tensors are concatenated
along dim 0 (0-dim tensors are stacked), ``edge_index`` along dim -1 with cumulative node offsets, and
``batch`` / ``ptr`` are added (data/ego4d_fho.py:242, data/ego4d_oscc.py:223 build the per-sample ``Data``).
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import torch


def replicated_base(x):
    """``[N, D]`` base of a feature tensor ``[N, R, D]`` whose R segments are the SAME memory (a stride-0 view, what
    ``base.unsqueeze(1).expand(-1, R, -1)`` gives), else None.

    The reference's PNR dataset hands over one 1536-vector per node repeated over the three segments
    (``x=...unsqueeze(1).repeat(1, 3, 1)``, data/ego4d_oscc.py:291).  A loader that writes ``expand`` instead of ``repeat``
    keeps the same values without the copies; collation, pinning, the feature-dtype conversion and the host -> device
    feed below preserve that form, so only the base crosses PCIe and the repeat happens on the device."""
    if not torch.is_tensor(x) or x.dim() != 3 or x.shape[1] < 2 or x.stride(1) != 0:
        return None
    return x[:, 0, :]


def expand_base(base, repeats: int):
    return base.unsqueeze(1).expand(-1, repeats, -1)


def _map_features(x, fn):
    """``fn`` applied to the feature tensor, or to its base when the segments are replicated (form preserved)."""
    base = replicated_base(x)
    return fn(x) if base is None else expand_base(fn(base.contiguous()), x.shape[1])


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        object.__setattr__(self, "_fields", {})
        object.__setattr__(self, "_lazy", {})
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos, **kwargs).items():
            if v is not None:
                self._fields[k] = v

    # attribute protocol: missing attributes read as None (PyG behaviour the reference relies on:
    # lta_temp_connectivity.py:31 tests ``data.batch is not None``)
    def __getattr__(self, key: str) -> Any:
        if key.startswith("__"):
            raise AttributeError(key)
        fields = object.__getattribute__(self, "_fields")
        if key in fields:
            return fields[key]
        lazy = object.__getattribute__(self, "_lazy")
        if key in lazy:                                      # materialise on first read, then it is a plain field
            value = lazy.pop(key)(self)
            if value is not None:
                fields[key] = value
            return value
        return None

    def __setattr__(self, key: str, value: Any) -> None:
        self._lazy.pop(key, None)
        if value is None:
            self._fields.pop(key, None)
        else:
            self._fields[key] = value

    def set_lazy(self, key: str, producer) -> None:
        """Register ``producer(data) -> value`` for an attribute that is expensive (or needs a host round trip) to build
        and that the kernels do not read: our transforms use it for ``edge_index`` when the graph is a band (+ star) --
        the aggregation kernels work from ``band_k`` / ``star``, so the int64 edge list is only built if someone asks."""
        self._fields.pop(key, None)
        self._lazy[key] = producer

    def is_materialized(self, key: str) -> bool:
        return key in self._fields

    def __contains__(self, key: str) -> bool:
        return key in self._fields or key in self._lazy

    def keys(self) -> List[str]:
        return list(self._fields) + list(self._lazy)

    @property
    def num_nodes(self) -> int:
        f = self._fields
        for k in ("x", "pos", "batch"):
            if k in f:
                return int(f[k].shape[0])
        return 0

    def to(self, device, non_blocking: bool = False) -> "Data":
        for k, v in list(self._fields.items()):
            if torch.is_tensor(v):
                if k == "x" and replicated_base(v) is not None:     # ship the base, repeat on the destination
                    v = _map_features(v, lambda t: t.to(device, non_blocking=non_blocking)).contiguous()
                else:
                    v = v.to(device, non_blocking=non_blocking)
                self._fields[k] = v
        self._fields.pop("_egp_structure", None)            # device-specific cache
        return self

    def to_feature_dtype(self, dtype: torch.dtype = torch.bfloat16) -> "Data":
        """Store the node features ``x`` in ``dtype`` (what a loader does once per sample, on the host).  With bf16 the
        batch is half the bytes over PCIe and the result is bit-identical: in the bf16 compute mode ``Graph.forward``
        rounds fp32 features to bf16 (round-to-nearest-even, like this conversion) before the first GEMM anyway."""
        x = self._fields.get("x")
        if x is not None and x.dtype != dtype:
            pinned = x.device.type == "cpu" and x.is_pinned()
            self._fields["x"] = _map_features(x, lambda t: t.to(dtype).pin_memory() if pinned else t.to(dtype))
        return self

    def pin_memory(self) -> "Data":
        for k, v in list(self._fields.items()):
            if torch.is_tensor(v):
                self._fields[k] = _map_features(v, lambda t: t.pin_memory()) if k == "x" else v.pin_memory()
        return self

    def __repr__(self) -> str:
        body = ", ".join(f"{k}={list(v.shape) if torch.is_tensor(v) else type(v).__name__}"
                         for k, v in self._fields.items() if not k.startswith("_"))
        return f"{type(self).__name__}({body})"


class Batch(Data):
    @classmethod
    def from_data_list(cls, graphs: Sequence[Data]) -> "Batch":
        out = cls()
        keys = [k for k in graphs[0].keys() if not k.startswith("_")]
        columns: Dict[str, list] = {k: [] for k in keys}
        sizes = []
        offset = 0
        band = None
        for d in graphs:
            n = d.num_nodes
            sizes.append(n)
            for k in keys:
                v = getattr(d, k)
                if k == "edge_index":
                    v = v + offset
                elif torch.is_tensor(v) and v.dim() == 0:
                    v = v.unsqueeze(0)
                columns[k].append(v)
            offset += n
        for k in keys:
            if k == "band_k":                                # structural hint from our transforms
                ks = set(int(v) for v in columns[k])
                band = ks.pop() if len(ks) == 1 else None
            elif k == "edge_index":
                out.edge_index = torch.cat(columns[k], dim=-1)
            elif k == "x" and all(replicated_base(v) is not None for v in columns[k]) \
                    and len({v.shape[1] for v in columns[k]}) == 1:
                # every sample's segments are replicated: concatenate the bases, keep the form
                setattr(out, k, expand_base(torch.cat([replicated_base(v) for v in columns[k]], dim=0), columns[k][0].shape[1]))
            elif torch.is_tensor(columns[k][0]):
                setattr(out, k, torch.cat(columns[k], dim=0))
            else:
                setattr(out, k, columns[k])
        if band is not None:
            out.band_k = band
        sz = torch.tensor(sizes, dtype=torch.long)
        out.batch = torch.repeat_interleave(torch.arange(len(sizes), dtype=torch.long), sz)
        out.ptr = torch.cat([torch.zeros(1, dtype=torch.long), sz.cumsum(0)])
        return out

    @property
    def num_graphs(self) -> int:
        return int(self.ptr.numel() - 1) if self.ptr is not None else int(self.batch.max()) + 1
