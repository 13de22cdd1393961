"""Drop-in for ``models/temporal_pooling/trn_pooling.py:10-45`` (``TRNPooling``).

``(N, S, D) -> (N, S*D)`` is a free view of the contiguous features; then
Linear(S*D, HT) -> LayerNorm -> ReLU -> Dropout -> Linear(HT, HT) -> LayerNorm -> ReLU -> Dropout -> Linear(HT, H).
``self.proj`` keeps the reference's nn.Sequential layout (indices 0,1,4,5,8 hold parameters) so checkpoints
load unchanged; the forward issues 3 tcgen05 GEMMs and 2 fused row-LayerNorm+ReLU kernels.
"""
from __future__ import annotations

import logging

import torch.nn as nn

from ... import ops
from ...ops import ACT_RELU
from ..layers import row_layernorm
from .pooling import TemporalPooling

logger = logging.getLogger(__name__)


class TRNPooling(TemporalPooling):
    def __init__(self, input_size: int = 1024, output_size: int = 1024, num_segments: int = 8,
                 hidden_size: int = 1024, dropout: float = 0.0) -> None:
        super().__init__(input_size, output_size, num_segments)
        self.input_size, self.num_segments = input_size, num_segments
        logger.info("TRNPooling(input=%d, hidden=%d, output=%d, segments=%d, dropout=%s)", input_size, hidden_size,
                    output_size, num_segments, dropout)
        self.proj = nn.Sequential(
            nn.Linear(num_segments * input_size, hidden_size), nn.LayerNorm(hidden_size), nn.ReLU(inplace=True),
            nn.Dropout(dropout),
            nn.Linear(hidden_size, hidden_size), nn.LayerNorm(hidden_size), nn.ReLU(inplace=True),
            nn.Dropout(dropout),
            nn.Linear(hidden_size, output_size),
        )

    def _flat(self, x):
        if x.dim() == 3 and (x.shape[1] != self.num_segments or x.shape[2] != self.input_size):
            raise ValueError(f"expected [N, {self.num_segments}, {self.input_size}] features, got {tuple(x.shape)}")
        return x.reshape(x.shape[0], self.num_segments * self.input_size)

    def _tail(self, h):
        """Everything after the first Linear: LN+ReLU+Dropout, Linear, LN+ReLU+Dropout, Linear."""
        p = self.proj
        h = row_layernorm(p[1], h, act=ACT_RELU, dropout_p=p[3].p if self.training else 0.0)   # LN+ReLU+Dropout fused
        h = ops.linear(h, p[4].weight, p[4].bias)
        h = row_layernorm(p[5], h, act=ACT_RELU, dropout_p=p[7].p if self.training else 0.0)
        return ops.linear(h, p[8].weight, p[8].bias)

    def forward(self, x, *_):
        p = self.proj
        return self._tail(ops.linear(self._flat(x), p[0].weight, p[0].bias))

    def forward_many(self, xs):
        """Several feature tensors through the shared weights as ONE stacked pass ([sum N_t, output_size])."""
        p = self.proj
        return self._tail(ops.LinearCat.apply(p[0].weight, p[0].bias, *[self._flat(x) for x in xs]))
