"""Drop-in for ``models/tasks/task.py:9-26`` (``ProjectionTask``) plus the head helpers shared by the four tasks.

``net`` keeps the reference layout ``Sequential(Dropout, Linear, LayerNorm, ReLU, Linear)`` (state_dict keys
``net.1.*``, ``net.2.*``, ``net.4.*``); ``forward_features`` runs GEMM -> fused row-LN+ReLU -> GEMM.
Late fusion (``stack([...]).sum(0)``, recognition.py:51-54, oscc.py:75-77, pnr.py:69-71) is folded into the
classifier GEMMs: each auxiliary classifier takes the running sum of logits as its epilogue residual, so no
[T, N, classes] stack is ever materialised.
"""
from __future__ import annotations

from typing import Literal, Optional

import torch
import torch.nn as nn
from torch import Tensor

from ... import config, ops
from ...ops import ACT_RELU
from ..layers import row_layernorm

TaskLiteral = Literal["ar", "oscc", "lta", "pnr", "ant"]


class ProjectionTask(nn.Module):
    def __init__(self, name: str, input_size: int, features_size: int = 1024, dropout: float = 0):
        super().__init__()
        self.name, self.input_size, self.features_size = name, input_size, features_size
        self.net = nn.Sequential(nn.Dropout(dropout), nn.Linear(input_size, features_size),
                                 nn.LayerNorm(features_size), nn.ReLU(), nn.Linear(features_size, features_size))

    def configure_optimizers(self, _):
        return self.parameters()

    def forward_features(self, x: Tensor, *args, **kwargs) -> Tensor:
        if not x.is_cuda:
            raise RuntimeError("egopack_b200 tasks run on CUDA only (no CPU fallback)")
        x = ops.Cast.apply(x, config.compute_dtype())
        net = self.net
        h = ops.dropout(x, net[0].p, self.training)
        h = ops.linear(h, net[1].weight, net[1].bias)
        h = row_layernorm(net[2], h, act=ACT_RELU)
        return ops.linear(h, net[4].weight, net[4].bias)

    # ---- head helpers -------------------------------------------------------------------------------
    def _build_head(self, head_dropout: float, n_out: int) -> nn.Sequential:
        return nn.Sequential(nn.Dropout(head_dropout), nn.Linear(self.features_size, n_out))

    def _head(self, head: nn.Sequential, features: Tensor, running: Optional[Tensor] = None) -> Tensor:
        """fp32 logits of one classifier, accumulated onto ``running`` (late-fusion sum) in the GEMM epilogue."""
        f = ops.Cast.apply(features, config.compute_dtype())
        f = ops.dropout(f, head[0].p, self.training)
        return ops.linear(f, head[1].weight, head[1].bias, residual=running, out_dtype=torch.float32)

    @staticmethod
    def _finish(total: Tensor, count: int, average: bool) -> Tensor:
        return ops.Scale.apply(total, 1.0 / count) if average and count > 1 else total
