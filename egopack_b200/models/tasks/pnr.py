"""Drop-in for ``models/tasks/pnr.py:12-83`` (point of no return: one logit per node)."""
from __future__ import annotations

import logging
from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.nn as nn

from ... import ops
from .task import ProjectionTask, TaskLiteral

logger = logging.getLogger(__name__)


class PNRTask(ProjectionTask):
    """Point of No Return task."""

    def __init__(self, input_size: int, features_size: int, dropout: float = 0, head_dropout: float = 0,
                 aux_tasks: Optional[Tuple[TaskLiteral, ...]] = None, average_logits: bool = False):
        super().__init__("pnr", input_size, features_size, dropout)
        self.loss_fn = nn.BCEWithLogitsLoss(reduction="none")
        self.classifier = self._build_classifier(head_dropout)
        if aux_tasks:
            self.aux_classifiers: Mapping[str, nn.Module] = nn.ModuleDict(
                {task: self._build_classifier(head_dropout) for task in aux_tasks})
            self.average_logits = average_logits

    def _build_classifier(self, head_dropout) -> nn.Module:
        return self._build_head(head_dropout, 1)

    def forward(self, x: torch.Tensor, *args, **kwargs):
        features = self.forward_features(x)
        return self._head(self.classifier, features).squeeze(), features

    def forward_logits(self, features: torch.Tensor, aux_features: Optional[Dict[TaskLiteral, torch.Tensor]] = None,
                       *args, **kwargs):
        logits = self._head(self.classifier, features)                       # [N, 1]
        if aux_features is not None:
            for task_name, task_features in aux_features.items():
                logits = self._head(self.aux_classifiers[task_name], task_features, running=logits)
            logits = self._finish(logits, 1 + len(aux_features), self.average_logits)
        return logits.squeeze()

    def forward_aux_logits(self, features: torch.Tensor, t: TaskLiteral = "ar", *args, **kwargs):
        if not hasattr(self, "aux_classifiers"):
            raise ValueError("PNR task has no auxiliary classifiers.")
        return self._head(self.aux_classifiers[t], features)

    def compute_loss(self, logits: torch.Tensor, targets: torch.Tensor):
        return ops.bce_with_logits(logits, targets)
