"""Drop-in for the reference's ``models/graph.py:15-65`` (``Graph``): TRN pooling MLP, positional encoding,
``depth`` x [SAGEConv(project, mean) -> graph-mode LayerNorm -> LeakyReLU(0.2)], final Linear, outer residual.

Same constructor and ``forward(data)`` signature, same attributes, same ``state_dict`` keys
(``temporal_pooling.proj.*``, ``positional_encoding.frequency``, ``net.module_{i}.*``) -- but the forward is a
fixed sequence of sm_100a kernels:

    x = TRNPooling(x)                                  3 GEMMs (tcgen05) + 2 row-LN/ReLU kernels
    z = x + PE(pos)                                    1 elementwise kernel
    per layer:  xs = relu(z Wp^T + bp)                 GEMM, ReLU in the epilogue
                agg = band/CSR mean(xs)                sliding-window aggregation kernel
                u = agg Wl^T + bl + z Wr^T             ONE dual-operand GEMM
                z = leaky_relu(graph_LN(u))            stats kernel + apply kernel
    out = x + z Wf^T + bf                              GEMM, residual in the epilogue
"""
from __future__ import annotations

import importlib
from typing import Any, List, Sequence

import torch
import torch.nn as nn

from .. import config, ops
from ..ops import ACT_LEAKY
from .layers import GraphLayerNorm, PositionalEncoding, SAGEConv, structure_for, structure_for_many


def _instantiate(spec: Any, *args):
    """hydra.utils.instantiate stand-in for ``temporal_pooling`` (models/graph.py:32-33 passes
    ``(input_size, hidden_size, num_segments)`` positionally): accepts a module, a callable, or a dict /
    DictConfig with ``_target_`` -- reference target paths are mapped onto this package."""
    if isinstance(spec, nn.Module):
        return spec
    if callable(spec):
        return spec(*args)
    cfg = dict(spec)
    target = cfg.pop("_target_", "models.temporal_pooling.trn_pooling.TRNPooling")
    cfg.pop("_recursive_", None)
    mod_name, cls_name = target.rsplit(".", 1)
    if mod_name.startswith("models."):
        mod_name = "egopack_b200." + mod_name
    return getattr(importlib.import_module(mod_name), cls_name)(*args, **cfg)


class TemporalNet(nn.Module):
    """Stand-in for the ``gnn.Sequential`` at models/graph.py:48: children are named ``module_{i}`` exactly
    as PyG registers them (SAGEConv, LayerNorm, LeakyReLU) x depth, then the final Linear."""

    def __init__(self, hidden_size: int, depth: int):
        super().__init__()
        self.depth = depth
        for d in range(depth):
            setattr(self, f"module_{3 * d}", SAGEConv(hidden_size, hidden_size, project=True))
            setattr(self, f"module_{3 * d + 1}", GraphLayerNorm(hidden_size))
            setattr(self, f"module_{3 * d + 2}", nn.LeakyReLU(negative_slope=0.2))
        setattr(self, f"module_{3 * depth}", nn.Linear(hidden_size, hidden_size))

    def forward(self, x, gs, pos, pe, seg_rows=None):
        """``x + net(x + PE(pos))`` (models/graph.py:63).  The first SAGE layer owns the positional-encoding add and the
        residual branch (ops.SageLayerPE), so the two gradient paths into x meet in its dgrad GEMM epilogue."""
        residual = x
        z = None
        for d in range(self.depth):
            conv = getattr(self, f"module_{3 * d}")
            norm = getattr(self, f"module_{3 * d + 1}")
            slope = getattr(self, f"module_{3 * d + 2}").negative_slope
            if d == 0:
                if conv.fusable:
                    if pe.granularity != 1.0:
                        raise NotImplementedError("the reference uses the default granularity of 1.0")
                    u, residual = ops.SageLayerPE.apply(x, conv.lin.weight, conv.lin.bias, conv.lin_l.weight,
                                                        conv.lin_l.bias, conv.lin_r.weight, gs, pos, pe.frequency)
                else:
                    u = conv(pe.add_to(x, pos), gs)
            else:
                u = conv(z, gs)
            z = norm(u, act=ACT_LEAKY, slope=slope, seg_rows=seg_rows)
        last = getattr(self, f"module_{3 * self.depth}")
        return ops.linear(z, last.weight, last.bias, residual=residual)


class Graph(nn.Module):
    def __init__(self, input_size: int, hidden_size: int = 1024, depth: int = 3, pre_dropout: float = 0,
                 temporal_pooling=None, num_segments: int = 8, *args, **kwargs):
        super().__init__()
        self.num_segments = num_segments
        self.pre_dropout = nn.Dropout(pre_dropout)
        self.temporal_pooling = _instantiate(temporal_pooling, input_size, hidden_size, num_segments) \
            if temporal_pooling else None
        self.positional_encoding = PositionalEncoding(hidden_size)
        if depth > 0:
            self.net = TemporalNet(hidden_size, depth)

    def configure_optimizers(self, _):
        return self.parameters()

    def forward(self, data, *args, **kwargs):
        x = data.x
        if not x.is_cuda:
            raise RuntimeError("egopack_b200.Graph runs on CUDA only (no CPU fallback); move the batch to the GPU")
        x = ops.Cast.apply(x, config.compute_dtype())
        x = ops.dropout(x, self.pre_dropout.p, self.training)
        if self.temporal_pooling is not None:
            x = self.temporal_pooling(x, data.batch, data.pos)
        elif x.dim() != 2:
            raise ValueError("without temporal pooling the node features must be [N, hidden]")
        if hasattr(self, "net"):
            gs = structure_for(data, x.shape[0])
            x = self.net(x, gs, data.pos, self.positional_encoding)
        return x

    def forward_many(self, batches: Sequence) -> List[torch.Tensor]:
        """``[self(b) for b in batches]`` as ONE pass over the stacked rows (main_temporal.py:87-90 runs the shared
        ``Graph`` once per task batch): every weight-shared Linear becomes a single GEMM over all task batches (and
        yields one weight gradient instead of one per batch to be summed), the aggregation runs over one band(+star)
        structure with global row indices, and graph-mode LayerNorm keeps its per-call statistics through row
        segments -- so the result equals the per-batch forwards up to summation order.  Falls back to per-batch
        forwards when the batches cannot share a structure (different radii, CSR graphs) or features need a gradient."""
        batches = list(batches)
        if len(batches) <= 1:
            return [self.forward(b) for b in batches]
        xs = [b.x for b in batches]
        if any(not x.is_cuda for x in xs):
            raise RuntimeError("egopack_b200.Graph runs on CUDA only (no CPU fallback); move the batch to the GPU")
        sizes = [int(x.shape[0]) for x in xs]
        gs = None
        if hasattr(self, "net"):
            gs = structure_for_many(batches, sizes)
        stackable = (not hasattr(self, "net") or gs is not None) and not any(x.requires_grad for x in xs) \
            and len({x.dim() for x in xs}) == 1 and len(batches) <= 8
        if not stackable:
            return [self.forward(b) for b in batches]
        cd = config.compute_dtype()
        xs = [ops.dropout(ops.Cast.apply(x, cd), self.pre_dropout.p, self.training) for x in xs]
        if self.temporal_pooling is not None and hasattr(self.temporal_pooling, "forward_many"):
            x = self.temporal_pooling.forward_many(xs)
        elif self.temporal_pooling is not None:
            x = torch.cat([self.temporal_pooling(xi, b.batch, b.pos) for xi, b in zip(xs, batches)], 0)
        else:
            if xs[0].dim() != 2:
                raise ValueError("without temporal pooling the node features must be [N, hidden]")
            x = torch.cat(xs, 0)
        if hasattr(self, "net"):
            seg_rows = [0]
            for n in sizes:
                seg_rows.append(seg_rows[-1] + n)
            pos = torch.cat([b.pos.view(-1) for b in batches], 0)
            x = self.net(x, gs, pos, self.positional_encoding, seg_rows=tuple(seg_rows))
        return list(ops.SplitRows.apply(x, *sizes))
