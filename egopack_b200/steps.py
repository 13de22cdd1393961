"""Host-side training-step glue mirroring the reference's hot loops (the loops themselves stay Python, as in the
reference): ``main_temporal.py:76-130`` (MTL pre-training) and ``main_egopack.py:45-61,102-155`` (novel task with
the GraphONE backpack).  The callers' semantics are kept exactly: all Graph forwards first, per-task
``w * loss.mean()`` over ALL nodes (ignore_index rows included in the denominator), one backward over the sum.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import ops

TASK_ORDER = ("ar", "lta", "oscc", "pnr")           # main_temporal.py:93-126 order of the loss terms


def _graph_features(model, batches: Dict[str, object]) -> Dict[str, torch.Tensor]:
    """``feat_t = model(data_t)`` for every task batch (main_temporal.py:87-90, main_egopack.py:113-117).  The native
    ``Graph`` runs them as one stacked pass over its shared weights (``forward_many``); any other module is called per
    batch exactly as the reference does."""
    many = getattr(model, "forward_many", None)
    if callable(many) and len(batches) > 1:
        return dict(zip(batches.keys(), many(list(batches.values()))))
    return {t: model(b) for t, b in batches.items()}


def multi_head_ce(logits, targets):
    """criterion/wrapper.py:80-82 around CrossEntropyLoss(ignore_index=-1, reduction='none') (main_temporal.py:285)."""
    return ops.cross_entropy(tuple(logits), targets, ignore_index=-1)


def mtl_losses(model, tasks: Dict[str, torch.nn.Module], batches: Dict[str, object],
               weights: Optional[Dict[str, float]] = None):
    """Forward + loss of one MTL step (main_temporal.py:87-126).  Returns (total, {task: per-sample loss})."""
    weights = weights or {}
    feats = _graph_features(model, batches)
    terms, per_task = [], {}
    for t in TASK_ORDER:
        if t not in batches:
            continue
        task, data = tasks[t], batches[t]
        f = task.forward_features(feats[t])
        if t == "oscc":
            loss = ops.cross_entropy(task.forward_logits(f, data.batch, ptr=getattr(data, "ptr", None)), data.y,
                                     ignore_index=-100)                    # plain nn.CrossEntropyLoss, main_temporal.py:291
        elif t == "pnr":
            loss = ops.bce_with_logits(task.forward_logits(f), data.y)
        elif callable(getattr(task, "loss_from_features", None)):
            loss = task.loss_from_features(f, data.y)                      # heads + criterion as one fused node
        else:
            loss = multi_head_ce(task.forward_logits(f), data.y)
        per_task[t] = loss
        terms.append(weights.get(t, 1.0))
    return ops.weighted_mean_sum(list(per_task.values()), terms), per_task


def egopack_task_loss(feat, batch, y, primary, others: Sequence[torch.nn.Module], graphone, late_fusion: bool = True,
                      ptr=None):
    """``train_step_task`` (main_egopack.py:45-61): secondary-task features are DETACHED before the interaction."""
    feat_primary = primary.forward_features(feat)
    secondary, _ = graphone.interact({t.name: t.forward_features(feat).detach() for t in others})
    kw = {"ptr": ptr} if primary.name == "oscc" else {}
    if late_fusion:
        logits = primary.forward_logits(features=feat_primary, batch=batch, aux_features=secondary, **kw)
    else:
        logits = primary.forward_logits(feat_primary, batch, **kw)
    return primary.compute_loss(logits, y)


def egopack_losses(model, tasks: Dict[str, torch.nn.Module], batches: Dict[str, object], graphone,
                   weights: Optional[Dict[str, float]] = None, late_fusion: bool = True,
                   backprop_temporal_graph: bool = True):
    """One EgoPack step (main_egopack.py:113-152) over whichever task batches are present."""
    weights = weights or {}
    with torch.set_grad_enabled(backprop_temporal_graph):
        feats = _graph_features(model, batches)
    terms, per_task = [], {}
    for t in ("ar", "oscc", "lta", "pnr"):                                   # main_egopack.py:121-149 order
        if t not in batches:
            continue
        data = batches[t]
        others = [tasks[o] for o in ("ar", "lta", "oscc", "pnr") if o != t and o in tasks]
        loss = egopack_task_loss(feats[t], data.batch, data.y, tasks[t], others, graphone, late_fusion,
                                 ptr=getattr(data, "ptr", None))
        per_task[t] = loss
        terms.append(weights.get(t, 1.0))
    return ops.weighted_mean_sum(list(per_task.values()), terms), per_task
