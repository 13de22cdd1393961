"""Drop-in for ``models/tasks/recognition.py:10-72`` (multi-head action recognition: verb + noun)."""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import ModuleDict, ModuleList

from ... import ops
from .task import ProjectionTask, TaskLiteral


class _MultiHeadTask(ProjectionTask):
    """Shared by RecognitionTask and LTATask (lta.py:10-74 is head-for-head the same model)."""

    def __init__(self, name, input_size: int, features_size: int, heads: Tuple[int, ...], dropout: float = 0,
                 head_dropout: float = 0, aux_tasks: Optional[Tuple[TaskLiteral, ...]] = None,
                 average_logits: bool = False):
        super().__init__(name, input_size, features_size, dropout)
        self.loss_fn = nn.CrossEntropyLoss(ignore_index=-1, reduction="none")
        self.classifiers = self._build_classifier(head_dropout, heads)
        if aux_tasks:
            self.aux_classifiers: Mapping[str, nn.ModuleList] = ModuleDict(
                {task: self._build_classifier(head_dropout, heads) for task in aux_tasks})
            self.average_logits = average_logits

    def _build_classifier(self, head_dropout, heads) -> nn.ModuleList:
        return ModuleList([self._build_head(head_dropout, h) for h in heads])

    def forward_logits(self, features: torch.Tensor, batch: Optional[torch.Tensor] = None,
                       aux_features: Optional[Dict[TaskLiteral, torch.Tensor]] = None, *args, **kwargs):
        logits = [self._head(c, features) for c in self.classifiers]
        if aux_features is not None:
            for task_name, task_features in aux_features.items():
                heads = self.aux_classifiers[task_name]
                logits = [self._head(c, task_features, running=l) for c, l in zip(heads, logits)]
            logits = [self._finish(l, 1 + len(aux_features), self.average_logits) for l in logits]
        return tuple(logits)

    def forward_aux_logits(self, features: torch.Tensor, t="ar", *args, **kwargs):
        return tuple(self._head(c, features) for c in self.aux_classifiers[t])

    def loss_from_features(self, features: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """``compute_loss(forward_logits(features), targets)`` as one fused autograd node (``ops.LinearCrossEntropy``:
        head GEMMs + loss kernels, with the loss gradient emitted as the GEMM operand).  Used by the training-step glue
        (``steps.mtl_losses``); falls back to the two reference calls when head dropout is active."""
        from ... import config
        if any(c[0].p > 0 for c in self.classifiers) and self.training:
            return self.compute_loss(self.forward_logits(features), targets)
        f = ops.Cast.apply(features, config.compute_dtype())
        wb = []
        for c in self.classifiers:
            wb += [c[1].weight, c[1].bias]
        return ops.LinearCrossEntropy.apply(f, targets, self.loss_fn.ignore_index, 0.0, *wb)

    def compute_loss(self, logits: Tuple[torch.Tensor], targets: torch.Tensor, return_separate_losses: bool = False):
        if return_separate_losses:
            losses = [ops.cross_entropy(l, t, ignore_index=self.loss_fn.ignore_index)
                      for l, t in zip(logits, targets.unbind(1))]
            return torch.stack(losses).sum(0), tuple(losses)        # logging path (validate.py), not the training step
        # sum over the label heads inside the loss kernel (criterion/wrapper.py:80-82 semantics)
        return ops.cross_entropy(tuple(logits), targets, ignore_index=self.loss_fn.ignore_index)


class RecognitionTask(_MultiHeadTask):
    """Multihead recognition task."""

    def __init__(self, input_size: int, features_size: int, heads: Tuple[int, ...], dropout: float = 0,
                 head_dropout: float = 0, aux_tasks: Optional[Tuple[TaskLiteral, ...]] = None,
                 average_logits: bool = False):
        super().__init__("ar", input_size, features_size, heads, dropout, head_dropout, aux_tasks, average_logits)
