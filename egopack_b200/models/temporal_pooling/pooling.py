"""Base class mirroring ``models/temporal_pooling/pooling.py:10-90``.  The reference's optional positional /
temporal / learnt encodings of the *segments* are never configured by any experiment (``encoding=None``
everywhere: configs/model/temporal_pooling/trn.yaml) so they are out of scope and asking
for one raises instead of silently ignoring it."""
from __future__ import annotations

from typing import Literal, Optional

import torch


class TemporalPooling(torch.nn.Module):
    def __init__(self, input_size: int, output_size: int, num_segments: int,
                 encoding: Optional[Literal["positional", "temporal", "learnt"]] = None,
                 encoding_level: Literal["frame", "action"] = "frame") -> None:
        super().__init__()
        if encoding is not None:
            raise NotImplementedError("segment-level encodings are unused by the reference's configs and not built")
        self.input_size, self.output_size, self.num_segments = input_size, output_size, num_segments
        self.encoding_level = encoding_level
        self.encoding = None
        self.encoding_mlp = None

    def apply_positional_embedding(self, x, batch, pos):
        return x

    def forward(self, x, batch, pos):
        raise NotImplementedError("TemporalPooling.forward is not implemented")
