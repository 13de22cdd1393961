"""Device-side headline metrics -- the mirror of ``utils/meters/ego4d.py`` (SURVEY.md section 8, rows a-M and (f)-4).

The reference accumulates these through torchmetrics 1.0.1 (``MulticlassAccuracy``, ``BinaryAccuracy/Recall/AUROC``,
``MeanMetric``) and ``editdistance`` 0.6.2 on the host, with ``.item()`` syncs per graph (PNR) and Python loops per
sequence (LTA).  Here every ``update`` is a handful of asynchronous kernel launches on the current stream
(``egp_label_rank``, ``egp_segment_argmax``, ``egp_edit_distance_min``) plus integer counter adds; nothing is read
back until ``get_logs`` / ``print_logs``.  Class names, ``update`` signatures and log keys follow the reference;
what is NOT mirrored is its logging surface: wandb tables, confusion matrices, calibration error, t-SNE plots.

Tie rule: a top-k hit means fewer than k classes beat the label's logit, where an equal logit beats it only from a
lower class index.  For k = 1 that is exactly ``argmax`` (first maximum), which is what torchmetrics uses; for k > 1
torchmetrics inherits ``torch.topk``'s unspecified tie order, so this rule is the deterministic refinement.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

from . import ops

__all__ = ["BaseMeter", "Ego4dRecognitionMeter", "Ego4dOSCCMeter", "Ego4dPNRMeter", "Ego4dLTAMeter"]


def _num_classes(dataset, num_verbs: Optional[int], num_nouns: Optional[int]):
    """(idx_verbs, idx_nouns, n_verbs, n_nouns) from a reference-style dataset (``label_names`` / ``class_labels``,
    data/ego4d_fho.py:88,124-130) or from explicit counts (label columns 0 = verbs, 1 = nouns)."""
    if dataset is not None and hasattr(dataset, "label_names"):
        iv, inn = dataset.label_names.index("verbs"), dataset.label_names.index("nouns")
        return iv, inn, len(dataset.class_labels[iv]), len(dataset.class_labels[inn])
    if num_verbs is None or num_nouns is None:
        raise ValueError("pass a dataset with label_names/class_labels or num_verbs and num_nouns")
    return 0, 1, int(num_verbs), int(num_nouns)


class BaseMeter:
    """utils/meters/base.py:11-34: running mean of the loss (MeanMetric) and a sample counter (SumMetric)."""

    def __init__(self, save_features: bool = False, device: torch.device = torch.device("cpu")) -> None:
        if save_features:
            raise NotImplementedError("feature dumps / t-SNE plots are part of the reference's logging UI, not of this path")
        self.save_features = save_features
        self.device = torch.device(device)
        self._loss_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        self._loss_cnt = torch.zeros((), dtype=torch.float64, device=self.device)
        self._count = 0

    def update(self, labels: Tensor, loss: Tensor, *args, **kwargs) -> None:
        loss = loss.detach()
        if torch.isnan(loss).any():                                 # MeanMetric(nan_strategy="error")
            raise RuntimeError("Encountered `nan` values in tensor")
        self._loss_sum += loss.sum().to(self._loss_sum)
        self._loss_cnt += loss.numel()
        self._count += int(labels.shape[0])

    @property
    def loss(self) -> Tensor:
        return (self._loss_sum / self._loss_cnt).float()

    def print_logs(self) -> List[str]:
        return [f"Loss: {float(self.loss):.4f}"]

    def get_logs(self, *args, **kwargs) -> Dict[str, Tensor]:
        return {"loss": self.loss}


class _TopK:
    """Counters behind MulticlassAccuracy(top_k in ks, average 'micro' and 'macro', ignore_index=-1)."""

    def __init__(self, num_classes: int, ks: Sequence[int], device) -> None:
        self.num_classes, self.ks = num_classes, tuple(ks)
        self.hits = torch.zeros(len(self.ks), dtype=torch.int64, device=device)
        self.valid = torch.zeros((), dtype=torch.int64, device=device)
        self.tp = torch.zeros(num_classes, dtype=torch.int64, device=device)       # top-1 hits per target class
        self.support = torch.zeros(num_classes, dtype=torch.int64, device=device)  # targets per class (tp + fn)
        self.predicted = torch.zeros(num_classes, dtype=torch.int64, device=device)  # arg-max predictions per class
        self._kvec = torch.tensor(self.ks, dtype=torch.int32, device=device)

    def update(self, logits: Tensor, target: Tensor) -> None:
        rank = ops.label_rank(logits, target, ignore_index=-1)
        valid = rank >= 0
        self.valid += valid.sum()
        self.hits += (valid[:, None] & (rank[:, None] < self._kvec[None, :])).sum(0)
        t = target[valid]
        self.support += torch.bincount(t, minlength=self.num_classes)
        self.tp += torch.bincount(t[rank[valid] == 0], minlength=self.num_classes)
        self.predicted += torch.bincount(logits.detach().argmax(-1)[valid], minlength=self.num_classes)

    def micro(self, k: int) -> Tensor:
        return self.hits[self.ks.index(k)].float() / self.valid.clamp(min=1).float()

    def macro(self) -> Tensor:
        """torchmetrics 1.0.1 `_accuracy_reduce(average='macro')`: mean of tp/(tp+fn) over the classes that occur
        as a target or as a prediction (classes with tp+fp+fn == 0 get weight 0)."""
        score = self.tp.float() / self.support.clamp(min=1).float()
        score = torch.where(self.support > 0, score, torch.zeros_like(score))
        seen = ((self.support + self.predicted) > 0).float()     # tp+fn = support, tp+fp = predicted
        return (score * seen).sum() / seen.sum().clamp(min=1)


class Ego4dRecognitionMeter(BaseMeter):
    """utils/meters/ego4d.py:34-133: verbs / nouns top-1/2/3/5 micro accuracy and mean-class accuracy."""

    def __init__(self, dataset=None, *args, num_verbs: Optional[int] = None, num_nouns: Optional[int] = None, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.dataset = dataset
        self.idx_verbs, self.idx_nouns, nv, nn_ = _num_classes(dataset, num_verbs, num_nouns)
        self.verbs = _TopK(nv, (1, 2, 3, 5), self.device)
        self.nouns = _TopK(nn_, (1, 2, 3, 5), self.device)

    @torch.no_grad()
    def update(self, logits, labels, *args, **kwargs) -> None:
        super().update(labels, *args, **kwargs)
        self.verbs.update(logits[self.idx_verbs], labels[:, self.idx_verbs])
        self.nouns.update(logits[self.idx_nouns], labels[:, self.idx_nouns])

    def get_logs(self, *args, **kwargs) -> Dict[str, Tensor]:
        out = {}
        for name, m in (("verbs", self.verbs), ("nouns", self.nouns)):
            for k in (1, 2, 3, 5):
                out[f"{name}_top{k}"] = m.micro(k)
            out[f"{name}_mc"] = m.macro()
        out.update(super().get_logs(*args, **kwargs))
        return out

    def print_logs(self) -> List[str]:
        lg = self.get_logs()
        return [
            "Verbs " + ", ".join(f"Top-{k}: {float(lg[f'verbs_top{k}']) * 100:.2f}" for k in (1, 2, 3, 5)),
            "Nouns " + ", ".join(f"Top-{k}: {float(lg[f'nouns_top{k}']) * 100:.2f}" for k in (1, 2, 3, 5)),
            f"Verbs Mean class: {float(lg['verbs_mc']) * 100:.2f}",
            f"Nouns Mean class: {float(lg['nouns_mc']) * 100:.2f}",
            *super().print_logs(),
        ]


class Ego4dOSCCMeter(BaseMeter):
    """utils/meters/ego4d.py:300-329: 2-class micro accuracy (ignore_index=-1)."""

    def __init__(self, dataset=None, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.dataset = dataset
        self._acc = _TopK(2, (1,), self.device)

    @torch.no_grad()
    def update(self, logits, labels, *args, **kwargs) -> None:
        super().update(labels, *args, **kwargs)
        self._acc.update(logits, labels)

    def get_logs(self, *args, **kwargs) -> Dict[str, Tensor]:
        return {"accuracy": self._acc.micro(1), **super().get_logs(*args, **kwargs)}

    def print_logs(self) -> List[str]:
        return [f"Accuracy: {float(self._acc.micro(1)) * 100:.2f}", *super().print_logs()]


class Ego4dPNRMeter(BaseMeter):
    """utils/meters/ego4d.py:332-389: binary accuracy / recall / AUROC of sigmoid(logit) against the one-hot key-frame
    labels, and the key-frame localisation error in seconds,
    ``| (end - start) / 16 * argmax_node(sigmoid(logit)) - (pnr - start) | / 30`` averaged over graphs (:356-366,376)."""

    def __init__(self, dataset=None, *args, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.dataset = dataset
        self._conf = torch.zeros(4, dtype=torch.int64, device=self.device)   # tp, fp, tn, fn at threshold 0.5
        self._err_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        self._err_cnt = 0
        self._probs: List[Tensor] = []
        self._targets: List[Tensor] = []

    @torch.no_grad()
    def update(self, logits, labels, batch, start_frame, end_frame, pnr_frame, *args, ptr: Optional[Tensor] = None,
               **kwargs) -> None:
        super().update(labels, *args, **kwargs)
        logits = logits.detach().float().reshape(-1)
        probs = torch.sigmoid(logits)
        pred, tgt = probs > 0.5, labels.reshape(-1) > 0.5
        self._conf += torch.stack([(pred & tgt).sum(), (pred & ~tgt).sum(), (~pred & ~tgt).sum(), (~pred & tgt).sum()])
        self._probs.append(probs)
        self._targets.append(tgt)
        if ptr is None:   # `batch` is sorted (PyG collate): graph boundaries from the per-graph counts
            counts = torch.bincount(batch, minlength=len(start_frame))
            ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)])
        loc = ops.segment_argmax(logits, ptr, apply_sigmoid=True)              # int64 [V], first arg-max
        # the reference's arithmetic, dtype for dtype: (ef - sf) / 16 is a default-dtype (fp32) tensor, times the python int
        # arg-max, .item()ed into a double; the ground-truth offset pf - sf stays integral (utils/meters/ego4d.py:357-365)
        sf, ef, pf = (torch.as_tensor(v, device=logits.device) for v in (start_frame, end_frame, pnr_frame))
        mapped = (torch.true_divide(ef - sf, 16) * loc).double()
        err = (mapped - (pf - sf).double()).abs() / 30
        self._err_sum += err.sum()
        self._err_cnt += int(err.numel())

    def _auroc(self) -> Tensor:
        """Exact ROC AUC (BinaryAUROC(thresholds=None)): Mann-Whitney U with average ranks for tied scores."""
        p, y = torch.cat(self._probs), torch.cat(self._targets)
        n1, n0 = y.sum().double(), (~y).sum().double()
        uniq, inv, cnt = torch.unique(p, return_inverse=True, return_counts=True)
        hi = cnt.cumsum(0).double()
        avg_rank = hi - (cnt.double() - 1) / 2
        u = avg_rank[inv][y].sum() - n1 * (n1 + 1) / 2
        return torch.where((n1 > 0) & (n0 > 0), u / (n1 * n0).clamp(min=1), torch.zeros_like(u)).float()

    def get_logs(self, *args, **kwargs) -> Dict[str, object]:
        tp, fp, tn, fn = self._conf.unbind(0)
        return {
            "accuracy": (tp + tn).float() / (tp + fp + tn + fn).clamp(min=1).float(),
            "recall": tp.float() / (tp + fn).clamp(min=1).float(),
            "auroc": self._auroc(),
            "localization_error": float(self._err_sum / max(self._err_cnt, 1)),
            **super().get_logs(*args, **kwargs),
        }

    def print_logs(self) -> List[str]:
        lg = self.get_logs()
        return [f"accuracy: {float(lg['accuracy']):.4f}", f"recall: {float(lg['recall']):.4f}",
                f"auroc: {float(lg['auroc']):.4f}", f"localization_error: {lg['localization_error']:.4f}",
                *super().print_logs()]


class Ego4dLTAMeter(BaseMeter):
    """utils/meters/ego4d.py:392-449: top-1 accuracy on labelled nodes and the LTA edit distance -- for every graph the
    minimum over the K sampled futures of Levenshtein(sample, ground truth) / Z, averaged over graphs.  The reference
    hard-codes 22 nodes per graph of which the first 2 are observed (:432-433); both are parameters here."""

    def __init__(self, dataset=None, *args, num_verbs: Optional[int] = None, num_nouns: Optional[int] = None,
                 nodes_per_graph: int = 22, num_input: int = 2, **kwargs) -> None:
        super().__init__(*args, **kwargs)
        self.dataset = dataset
        self.idx_verbs, self.idx_nouns, nv, nn_ = _num_classes(dataset, num_verbs, num_nouns)
        self.nodes_per_graph, self.num_input = nodes_per_graph, num_input
        self.verbs, self.nouns = _TopK(nv, (1,), self.device), _TopK(nn_, (1,), self.device)
        self._ed_sum = torch.zeros(2, dtype=torch.float64, device=self.device)
        self._ed_cnt = 0

    def _edit_distance(self, preds: Tensor, labels: Tensor) -> Tensor:
        """[N] lowest normalised edit distance among the K predictions (utils/meters/ego4d.py:410-422)."""
        return ops.edit_distance_min(preds, labels).double() / preds.shape[1]

    @torch.no_grad()
    def update(self, logits, labels, predictions, *args, **kwargs) -> None:
        super().update(labels, *args, **kwargs)
        n, ni = self.nodes_per_graph, self.num_input
        for j, (idx, acc) in enumerate(((self.idx_verbs, self.verbs), (self.idx_nouns, self.nouns))):
            keep = labels[:, idx] >= 0
            acc.update(logits[idx][keep], labels[keep, idx])
            k = predictions[idx].shape[-1]
            d = self._edit_distance(predictions[idx].reshape(-1, n, k)[:, ni:], labels[:, idx].reshape(-1, n)[:, ni:])
            self._ed_sum[j] += d.sum()
            if j == 0:
                self._ed_cnt += int(d.numel())

    def get_logs(self, *args, **kwargs) -> Dict[str, Tensor]:
        ed = (self._ed_sum / max(self._ed_cnt, 1)).float()
        return {"verbs_ed": ed[0], "nouns_ed": ed[1], "verbs_top1": self.verbs.micro(1), "nouns_top1": self.nouns.micro(1),
                **super().get_logs(*args, **kwargs)}

    def print_logs(self) -> List[str]:
        lg = self.get_logs()
        return [f"verbs_ed: {float(lg['verbs_ed']):.4f}", f"nouns_ed: {float(lg['nouns_ed']):.4f}",
                f"verbs_top1: {float(lg['verbs_top1']):.4f}", f"nouns_top1: {float(lg['nouns_top1']):.4f}",
                *super().print_logs()]
