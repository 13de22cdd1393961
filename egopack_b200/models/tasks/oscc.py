"""Drop-in for ``models/tasks/oscc.py:16-96`` (object state change classification).

``global_max_pool`` (:68,:85) is the segment-max kernel with its arg-max saved for the backward; the graph count
comes from ``ptr`` when the batch carries one, so the reference's ``batch.max().item()`` host sync disappears.
"""
from __future__ import annotations

import logging
from typing import Dict, Literal, Mapping, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import config, ops
from .task import ProjectionTask, TaskLiteral

logger = logging.getLogger(__name__)


def _ptr_from_batch(batch: torch.Tensor) -> torch.Tensor:
    counts = torch.bincount(batch)
    return torch.cat([counts.new_zeros(1), counts.cumsum(0)])


class OSCCTask(ProjectionTask):
    """OSCC task."""

    def __init__(self, input_size: int, features_size: int, dropout: float = 0, head_dropout: float = 0,
                 loss_func: Literal["ce", "bce", "focal"] = "ce", aux_tasks: Optional[Tuple[TaskLiteral, ...]] = None,
                 average_logits: bool = False):
        super().__init__("oscc", input_size, features_size, dropout)
        logger.info("OSCC task: loss=%s dropout=%s head_dropout=%s aux=%s", loss_func, dropout, head_dropout, aux_tasks)
        self.loss_func = loss_func
        self.classifier = self._build_classifier(head_dropout)
        if aux_tasks:
            self.aux_classifiers: Mapping[str, nn.Module] = nn.ModuleDict(
                {task: self._build_classifier(head_dropout) for task in aux_tasks})
            self.average_logits = average_logits

    def _build_classifier(self, head_dropout) -> nn.Module:
        return self._build_head(head_dropout, 2)

    def _pool(self, features: torch.Tensor, batch: torch.Tensor, ptr: Optional[torch.Tensor]) -> torch.Tensor:
        if ptr is None:
            ptr = _ptr_from_batch(batch)
        f = ops.Cast.apply(features, config.compute_dtype())
        return ops.SegmentMaxPool.apply(f, ptr, batch)

    def forward_logits(self, features: torch.Tensor, batch: torch.Tensor,
                       aux_features: Optional[Dict[TaskLiteral, torch.Tensor]] = None, *args, ptr=None, **kwargs):
        if ptr is None:
            ptr = _ptr_from_batch(batch)
        logits = self._head(self.classifier, self._pool(features, batch, ptr))
        if aux_features is not None:
            for task_name, task_features in aux_features.items():
                logits = self._head(self.aux_classifiers[task_name], self._pool(task_features, batch, ptr), running=logits)
            logits = self._finish(logits, 1 + len(aux_features), self.average_logits)
        return logits

    def forward_aux_logits(self, features: torch.Tensor, batch: torch.Tensor, t: TaskLiteral = "ar", *args, ptr=None, **kwargs):
        if not hasattr(self, "aux_classifiers"):
            raise ValueError("OSCC task has no auxiliary classifiers.")
        return self._head(self.aux_classifiers[t], self._pool(features, batch, ptr))

    def compute_loss(self, logits, targets):
        if self.loss_func == "ce":
            return ops.cross_entropy(logits, targets, ignore_index=-1, label_smoothing=0.1)
        if self.loss_func == "bce":
            return ops.bce_with_logits(logits, F.one_hot(targets, 2).float())
        raise NotImplementedError("the focal OSCC loss (torchvision) is not configured by any reference experiment")
