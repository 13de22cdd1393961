"""Prototype-bank builder: drop-in for the reference's ``graphone.py:16-63`` (``build_graphone``), SURVEY.md §8f-1.

For every batch of the AR training set: ``feat = model(data)``; keep the labelled nodes; for each task accumulate the
class-conditional sums of ``task.forward_features(feat)`` over the (verb, noun) pair id in **fp64**
(``egp_class_sum_f64`` replaces ``torch_geometric.utils.scatter``); finally divide by the label histogram and drop
the pairs never seen.  The reference appends the labels once per task inside the task loop (:47-52), so its
histogram -- and therefore the scale of every bank -- is ``len(tasks)`` times too large; that quirk is kept, because
the max-aggregation downstream is not scale invariant.
"""
from __future__ import annotations

from typing import Dict, Iterable, List

import torch

from . import ops


@torch.no_grad()
def build_graphone(model, ar_task, tasks: List[torch.nn.Module], dataloader: Iterable, device="cuda") -> Dict[str, torch.Tensor]:
    model.eval()
    for task in tasks:
        task.eval()
    feat_size = ar_task.net[-1].out_features
    n_classes = tuple(classifier[-1].out_features for classifier in ar_task.classifiers)
    size = n_classes[0] * n_classes[1]
    all_labels = []
    graphone = {task.name: torch.zeros((size, feat_size), dtype=torch.float64, device=device) for task in tasks}
    for data in dataloader:
        data = data.to(device)
        feat = model(data)
        keep = data.y[:, 0] != -1
        feat = feat[keep]                                    # row compaction: data movement only
        y = data.y[keep]
        labels = (y[:, 0] * n_classes[1] + y[:, 1]).contiguous()
        for task in tasks:
            task_feat = task.forward_features(feat)
            all_labels.append(labels)                        # once PER TASK, as upstream (graphone.py:47-52)
            ops.class_sum_f64(task_feat, labels, size, out=graphone[task.name])
    bincount = torch.cat(all_labels).bincount(minlength=size).float()
    seen = bincount > 0
    return {name: (bank[seen] / bincount[seen, None]).float() for name, bank in graphone.items()}
