"""TEST INFRASTRUCTURE ONLY -- CPU oracle for EgoPack's temporal-graph hot path.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the reference's first-party model code, layered on
``oracle/pyg_restated.py`` (the third-party ops).  ``state_dict`` keys equal the reference's so one set of
weights can be loaded into the reference modules (``tests/golden/make_golden.py``), this oracle and the
native ``egopack_b200`` modules.  Only tests, ``__graft_entry__.smoke()`` and bench.py's CPU-baseline legs may
import it.  **Parity unpinned upstream** (the reference has no tests); pinned here against outputs of the
reference's own ``models/*.py`` executed with the restated third-party layer -- see ``tests/golden``.

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

from math import floor
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import pyg_restated as pyg


# ---------------------------------------------------------------------------------------------- a3
class TRNPoolingOracle(nn.Module):
    """models/temporal_pooling/trn_pooling.py:12-45 -- flatten (N,S,D)->(N,S*D) then a 3-layer MLP."""

    def __init__(self, input_size=1024, output_size=1024, num_segments=8, hidden_size=1024, dropout=0.0):
        super().__init__()
        self.input_size, self.num_segments = input_size, num_segments
        dims = [(num_segments * input_size, hidden_size), (hidden_size, hidden_size)]
        layers: List[nn.Module] = []
        for fan_in, fan_out in dims:                                     # :28-37  Linear, LN, ReLU, Dropout
            layers += [nn.Linear(fan_in, fan_out), nn.LayerNorm(fan_out), nn.ReLU(), nn.Dropout(dropout)]
        layers.append(nn.Linear(hidden_size, output_size))               # :39
        self.proj = nn.Sequential(*layers)

    def forward(self, x: Tensor, *_):
        n = x.shape[0]
        return self.proj(x.reshape(n, self.num_segments * self.input_size))   # :44 (einops rearrange)


# ------------------------------------------------------------------------------------ a4..a8 (Graph)
class GraphOracle(nn.Module):
    """models/graph.py:15-65.  ``temporal_pooling`` is a dict of TRNPooling kwargs (hydra-free) or None."""

    def __init__(self, input_size: int, hidden_size: int = 1024, depth: int = 3, pre_dropout: float = 0,
                 temporal_pooling: Optional[dict] = None, num_segments: int = 8, **_):
        super().__init__()
        self.num_segments = num_segments
        self.pre_dropout = nn.Dropout(pre_dropout)                       # :29
        self.temporal_pooling = None
        if temporal_pooling:                                             # :32-33 positional (in, hidden, segs)
            kw = {k: v for k, v in dict(temporal_pooling).items() if k != "_target_"}
            self.temporal_pooling = TRNPoolingOracle(input_size, hidden_size, num_segments, **kw)
        self.positional_encoding = pyg.PositionalEncoding(hidden_size)  # :37
        if depth > 0:                                                    # :39-48
            chain = []
            for _ in range(depth):
                chain.append((pyg.SAGEConv(hidden_size, hidden_size, project=True), "x, edges -> x"))
                chain.append((pyg.LayerNorm(hidden_size), "x -> x"))     # batch never passed => global stats
                chain.append((nn.LeakyReLU(negative_slope=0.2), "x -> x"))
            chain.append((nn.Linear(hidden_size, hidden_size), "x -> x"))
            self.net = pyg.Sequential("x, edges, batch", chain)

    def forward(self, data, *_, **__):
        x = self.pre_dropout(data.x)                                     # :54-56
        if self.temporal_pooling is not None:
            x = self.temporal_pooling(x, data.batch, data.pos)           # :59-60
        if hasattr(self, "net"):                                         # :62-63
            x = x + self.net(x + self.positional_encoding(data.pos), data.edge_index, data.batch)
        return x


# ---------------------------------------------------------------------------------- a9..a13 (tasks)
class _TaskOracle(nn.Module):
    """models/tasks/task.py:9-26 (projection MLP) + the aux-classifier plumbing shared by the four tasks."""

    def __init__(self, name, input_size, features_size=1024, dropout=0.0):
        super().__init__()
        self.name, self.input_size, self.features_size = name, input_size, features_size
        self.net = nn.Sequential(nn.Dropout(dropout), nn.Linear(input_size, features_size),
                                 nn.LayerNorm(features_size), nn.ReLU(), nn.Linear(features_size, features_size))

    def forward_features(self, x, *_, **__):
        return self.net(x)

    def _head(self, p, n_out):
        return nn.Sequential(nn.Dropout(p), nn.Linear(self.features_size, n_out))

    @staticmethod
    def _fuse(stack: Tensor, average: bool) -> Tensor:                   # recognition.py:54, oscc.py:77
        return stack.mean(0) if average else stack.sum(0)


class _MultiHeadOracle(_TaskOracle):
    """RecognitionTask (models/tasks/recognition.py:10-72) and LTATask (models/tasks/lta.py:10-74)."""

    def __init__(self, name, input_size, features_size, heads, dropout=0, head_dropout=0, aux_tasks=None,
                 average_logits=False):
        super().__init__(name, input_size, features_size, dropout)
        self.classifiers = nn.ModuleList([self._head(head_dropout, h) for h in heads])
        if aux_tasks:
            self.aux_classifiers = nn.ModuleDict({
                t: nn.ModuleList([self._head(head_dropout, h) for h in heads]) for t in aux_tasks})
            self.average_logits = average_logits

    def forward_aux_logits(self, features, t="ar", *_, **__):
        return tuple(c(features) for c in self.aux_classifiers[t])

    def forward_logits(self, features, batch=None, aux_features=None, *_, **__):
        logits = tuple(c(features) for c in self.classifiers)            # recognition.py:42
        if aux_features is not None:                                     # :44-57
            aux = [self.forward_aux_logits(f, t) for t, f in aux_features.items()]
            logits = tuple(self._fuse(torch.stack([p, *a]), self.average_logits)
                           for p, a in zip(logits, zip(*aux)))
        return logits

    def compute_loss(self, logits, targets, return_separate_losses=False):
        per_head = torch.stack([F.cross_entropy(l, t, ignore_index=-1, reduction="none")
                                for l, t in zip(logits, targets.unbind(1))])      # recognition.py:62
        total = per_head.sum(0)
        return (total, per_head.unbind(0)) if return_separate_losses else total


class RecognitionTaskOracle(_MultiHeadOracle):
    def __init__(self, input_size, features_size, heads, **kw):
        super().__init__("ar", input_size, features_size, heads, **kw)


class LTATaskOracle(_MultiHeadOracle):
    def __init__(self, input_size, features_size, heads, **kw):
        super().__init__("lta", input_size, features_size, heads, **kw)

    def generate_from_logits(self, logits, K=5, *_, **__):              # lta.py:63-71
        preds = []
        for hl in logits:
            dist = torch.distributions.Categorical(logits=hl)
            preds.append(torch.stack([dist.sample() for _ in range(K)], dim=1))
        return preds, logits


class OSCCTaskOracle(_TaskOracle):
    """models/tasks/oscc.py:16-96."""

    def __init__(self, input_size, features_size, dropout=0, head_dropout=0, loss_func="ce", aux_tasks=None,
                 average_logits=False):
        super().__init__("oscc", input_size, features_size, dropout)
        self.loss_func = loss_func
        self.classifier = self._head(head_dropout, 2)
        if aux_tasks:
            self.aux_classifiers = nn.ModuleDict({t: self._head(head_dropout, 2) for t in aux_tasks})
            self.average_logits = average_logits

    def forward_aux_logits(self, features, batch, t="ar", *_, **__):
        return self.aux_classifiers[t](pyg.global_max_pool(features, batch))          # :85-86

    def forward_logits(self, features, batch, aux_features=None, *_, **__):
        logits = self.classifier(pyg.global_max_pool(features, batch))                 # :68-70
        if aux_features is not None:
            aux = [self.forward_aux_logits(f, batch, t) for t, f in aux_features.items()]
            logits = self._fuse(torch.stack([logits, *aux]), self.average_logits)
        return logits

    def compute_loss(self, logits, targets):
        if self.loss_func != "ce":
            raise NotImplementedError("only the default 'ce' OSCC loss is on the hot path")
        return F.cross_entropy(logits, targets, ignore_index=-1, reduction="none", label_smoothing=0.1)  # :90


class PNRTaskOracle(_TaskOracle):
    """models/tasks/pnr.py:12-83."""

    def __init__(self, input_size, features_size, dropout=0, head_dropout=0, aux_tasks=None, average_logits=False):
        super().__init__("pnr", input_size, features_size, dropout)
        self.classifier = self._head(head_dropout, 1)
        if aux_tasks:
            self.aux_classifiers = nn.ModuleDict({t: self._head(head_dropout, 1) for t in aux_tasks})
            self.average_logits = average_logits

    def forward(self, x, *_, **__):
        f = self.net(x)
        return self.classifier(f).squeeze(), f

    def forward_aux_logits(self, features, t="ar", *_, **__):
        return self.aux_classifiers[t](features)

    def forward_logits(self, features, aux_features=None, *_, **__):
        logits = self.classifier(features).squeeze()                     # :64
        if aux_features is not None:                                     # :66-72
            aux = [self.forward_aux_logits(f, t) for t, f in aux_features.items()]
            logits = self._fuse(torch.stack([logits.unsqueeze(1), *aux]), self.average_logits)
        return logits.squeeze()

    def compute_loss(self, logits, targets):
        return F.binary_cross_entropy_with_logits(logits, targets.float(), reduction="none")   # :83


# ------------------------------------------------------------------------------ a14..a16 (GraphONE)
def cos_dissimilarity(a: Tensor, b: Tensor) -> Tensor:
    """graphONE.py:148-151 -- no epsilon: a zero row yields NaN, as upstream."""
    a = a / a.norm(dim=1, keepdim=True)
    b = b / b.norm(dim=1, keepdim=True)
    return 1 - torch.mm(a, b.T)


class GraphONEOracle(nn.Module):
    """models/graphONE/graphONE.py:13-141, LITERAL form: per stage re-sort, cat([bank, feats]), self loops,
    max-SAGE over all K+B rows, slice the last B rows."""

    def __init__(self, graphone: Dict[str, Tensor], features_size=1024, hidden_size=1024, freeze=True, k=8,
                 depth=3, distance_func="cosine", residual=False, mix_strategy="max",
                 update_edges_interval=1, share_params=False, *_, **__):
        super().__init__()
        self.feature_size, self.k, self.depth = features_size, k, depth
        self.distance_func, self.residual, self.mix_strategy = distance_func, residual, mix_strategy
        self.update_edges_interval = update_edges_interval
        self.task_labels = sorted(graphone.keys())                       # :46
        self.embeddings = nn.ModuleDict({t: nn.Embedding.from_pretrained(graphone[t], freeze=freeze)
                                         for t in self.task_labels})     # :47-49
        stages = {}
        for t in self.task_labels:                                       # :53-73
            stages[t] = nn.ModuleList([
                pyg.Sequential("x, edge_index, weights", [
                    (pyg.SAGEConv(features_size, hidden_size, bias=False, project=False, aggr="max"),
                     "x, edge_index -> x"),
                    (nn.LayerNorm(hidden_size), "x -> x"),
                    (nn.ReLU(), "x -> x"),
                    (pyg.Linear(hidden_size, features_size), "x -> x"),
                ]) for _ in range(depth)])
        self.conv_stages = nn.ModuleDict(stages)

    @torch.no_grad()
    def knn(self, feats: Tensor, bank: Tensor) -> Tensor:
        """:119-141 -- returns the [B,k] nearest-prototype indices (``argsort`` of the dissimilarity)."""
        if self.distance_func == "cosine":
            d = cos_dissimilarity(feats, bank)
        elif self.distance_func == "l2":
            d = torch.cdist(feats, bank, p=2, compute_mode="donot_use_mm_for_euclid_dist") / 4096
        else:
            raise ValueError(f"Unknown distance function: {self.distance_func}")
        return d.argsort(dim=-1, descending=False)[:, : self.k]

    def _task(self, task: str, feats: Tensor, bank: Tensor) -> Tuple[Tensor, List[Tensor]]:
        f0, assignments, edges = feats, [], None
        K, B = bank.shape[0], feats.shape[0]
        for d, conv in enumerate(self.conv_stages[task]):               # :94
            nn_idx = self.knn(f0, bank)                                   # always the ORIGINAL features (:102)
            online = torch.stack([nn_idx.flatten(),
                                  torch.arange(K, K + B).repeat_interleave(nn_idx.shape[1])])   # :134
            assignments.append(nn_idx[:, 0])                             # :103
            if edges is None or (self.update_edges_interval and d % self.update_edges_interval == 0):
                edges = online                                           # :105-106
            graph = torch.cat([bank, feats], dim=0)                      # :108
            edges, _ = pyg.add_remaining_self_loops(edges, num_nodes=graph.shape[0])   # :109
            graph = conv(graph, edges, 0.5 * torch.ones(edges.shape[1]))               # :110
            feats = graph[-B:] + feats if self.residual else graph[-B:]  # :112-115
        return feats, assignments

    def interact(self, features: Dict[str, Tensor]):
        out, closest = {}, {}
        for t in features.keys():                                        # :80 dict order of the INPUT
            out[t], closest[t] = self._task(t, features[t], self.embeddings[t].weight)
        return out, closest

    def interact_reduced(self, features: Dict[str, Tensor]):
        """Algebraically reduced form -- what the CUDA path computes; the test-suite
        proves it equal to :meth:`interact`."""
        out, closest = {}, {}
        for t, f0 in features.items():
            bank = self.embeddings[t].weight
            nn_idx = self.knn(f0, bank)
            m = bank[nn_idx].max(dim=1).values                           # [B,C] constant across stages
            f = f0
            for conv in self.conv_stages[t]:
                sage, ln, _, proj = conv.module_0, conv.module_1, conv.module_2, conv.module_3
                h = sage.lin_l(torch.maximum(f, m)) + sage.lin_r(f)
                g = proj(torch.relu(ln(h)))
                f = g + f if self.residual else g
            out[t], closest[t] = f, [nn_idx[:, 0]] * self.depth
        return out, closest


# -------------------------------------------------------------------------------------- a2 (LTA edges)
def lta_temporal_connectivity(data, r: float, loop: bool = False, max_num_neighbors: int = 32):
    """models/transforms/lta_temp_connectivity.py:30-56 on ONE un-batched graph."""
    if data.batch is not None:
        raise ValueError("This transform expects no batched graphs.")    # :31-32
    band = pyg.radius_graph(data.pos, r, None, loop, max_num_neighbors=max_num_neighbors)
    n_in = int((data.y[:, 0] == -1).sum())                               # :48
    n_fc = int((data.y[:, 0] > 0).sum())                                 # :50 -- verb label 0 NOT counted
    lo = max(int(torch.tensor(n_in - r).ceil()), 0)
    src = torch.arange(lo, n_in, dtype=torch.long).repeat_interleave(n_fc)              # :52
    tgt = torch.arange(n_in, n_in + n_fc, dtype=torch.long).repeat(min(floor(r), n_in))  # :53
    data.edge_index = pyg.coalesce(torch.cat([torch.stack([src, tgt]), band], dim=-1), data.num_nodes)
    return data


# ------------------------------------------------------------------------------------ a11/a17 (steps)
def mtl_step(model: GraphOracle, tasks: Dict[str, nn.Module], batches: Dict[str, object],
             weights: Optional[Dict[str, float]] = None) -> Tuple[Tensor, Dict[str, Tensor]]:
    """One MTL training step without the optimiser -- main_temporal.py:76-128.  ``batches`` maps task name
    ('ar','lta','oscc','pnr') to a batched Data.  Returns (scalar loss, per-task per-sample losses)."""
    weights = weights or {}
    feats = {t: model(b) for t, b in batches.items()}                    # :87-90 (all forwards first)
    losses, per_task = [], {}
    for t in ("ar", "lta", "oscc", "pnr"):                               # order of :93-126
        if t not in batches:
            continue
        task, data = tasks[t], batches[t]
        f = task.forward_features(feats[t])
        if t == "oscc":
            logits = task.forward_logits(f, data.batch)
            loss = F.cross_entropy(logits, data.y, reduction="none")      # main_temporal.py:291 plain CE
        elif t == "pnr":
            logits = task.forward_logits(f)
            loss = F.binary_cross_entropy_with_logits(logits, data.y.float(), reduction="none")
        else:
            logits = task.forward_logits(f)
            loss = torch.stack([F.cross_entropy(l, y, ignore_index=-1, reduction="none")
                                for l, y in zip(logits, data.y.unbind(1))]).sum(0)   # criterion/wrapper.py:80-82
        per_task[t] = loss
        losses.append(weights.get(t, 1.0) * loss.mean())                 # :99 mean over ALL nodes
    return torch.stack(losses).sum(), per_task


def egopack_task_step(feat: Tensor, batch: Tensor, y: Tensor, primary, others: Sequence[nn.Module],
                      graphone: GraphONEOracle, late_fusion: bool = True) -> Tensor:
    """main_egopack.py:45-61."""
    fp = primary.forward_features(feat)
    secondary, _ = graphone.interact({t.name: t.forward_features(feat).detach() for t in others})   # :53
    if late_fusion:
        logits = primary.forward_logits(features=fp, batch=batch, aux_features=secondary)
    else:
        logits = primary.forward_logits(fp, batch)
    return primary.compute_loss(logits, y)


# ---------------------------------------------------------------------------------- §8f-1 bank builder
@torch.no_grad()
def build_graphone(model, ar_task, tasks: Sequence[nn.Module], batches: Iterable) -> Dict[str, Tensor]:
    """graphone.py:16-63, including the ``len(tasks)x`` bincount quirk (labels appended once per task)."""
    model.eval()
    for t in tasks:
        t.eval()
    feat_size = ar_task.net[-1].out_features
    n_cls = tuple(c[-1].out_features for c in ar_task.classifiers)
    size = n_cls[0] * n_cls[1]
    seen: List[Tensor] = []
    banks = {t.name: torch.zeros((size, feat_size), dtype=torch.float64) for t in tasks}
    for data in batches:
        feat = model(data)
        keep = data.y[:, 0] != -1
        feat, y = feat[keep], data.y[keep]
        for t in tasks:
            tf = t.forward_features(feat)
            labels = y[:, 0] * n_cls[1] + y[:, 1]
            seen.append(labels)                                           # inside the task loop (:47-52)
            banks[t.name] = banks[t.name] + pyg.scatter(tf, labels, dim=0, dim_size=size, reduce="sum")
    count = torch.cat(seen).bincount(minlength=size).float()
    return {n: (b[count > 0] / count[count > 0, None]).float() for n, b in banks.items()}


# ------------------------------------------------------------------------------------------- a-M metrics
def metric_ar(logits: Sequence[Tensor], y: Tensor) -> Tuple[float, float]:
    """Top-1 micro accuracy verbs / nouns, ignore_index=-1 (utils/meters/ego4d.py:46,60,93,107)."""
    out = []
    for l, t in zip(logits, y.unbind(1)):
        m = t != -1
        out.append(float((l.argmax(-1)[m] == t[m]).float().mean()) if m.any() else float("nan"))
    return out[0], out[1]


def metric_oscc(logits: Tensor, y: Tensor) -> float:
    return float((logits.argmax(-1) == y).float().mean())                # utils/meters/ego4d.py:306,313


def metric_pnr_localisation(logits: Tensor, batch: Tensor, pnr_node: Tensor, n_per_graph: int = 16,
                            clip_frames: float = 16.0, fps: float = 30.0) -> float:
    """|argmax_node(sigmoid(logit)) - true node| scaled to seconds (utils/meters/ego4d.py:356-366,376),
    for equal-length graphs spanning ``clip_frames`` frames each."""
    v = int(batch.max()) + 1
    pred = torch.sigmoid(logits).view(v, n_per_graph).argmax(-1)
    return float(((pred - pnr_node).abs().float() * clip_frames / n_per_graph / fps).mean())


def topk_accuracy(logits: Tensor, target: Tensor, k: int, ignore_index: int = -1) -> float:
    """MulticlassAccuracy(top_k=k, average='micro', ignore_index=-1) (utils/meters/ego4d.py:46-49 etc.; torchmetrics
    1.0.1, third party, absent).  A row is a hit when fewer than k classes beat the label's logit; an equal logit
    beats it only from a lower class index -- for k=1 this is torch.argmax (first maximum), which is what torchmetrics
    uses; for k>1 torchmetrics inherits torch.topk's unspecified tie order and this rule is the deterministic choice."""
    hits = tot = 0
    for row, t in zip(logits.tolist(), target.tolist()):
        if t == ignore_index:
            continue
        tot += 1
        beat = sum(1 for c, v in enumerate(row) if v > row[t] or (v == row[t] and c < t))
        hits += beat < k
    return hits / tot if tot else 0.0


def macro_accuracy(logits: Tensor, target: Tensor, ignore_index: int = -1) -> float:
    """MulticlassAccuracy(top_k=1, average='macro', ignore_index=-1): mean over classes of tp/(tp+fn), classes with
    tp+fp+fn == 0 excluded (torchmetrics 1.0.1 `_adjust_weights_safe_divide`)."""
    c = logits.shape[1]
    tp, sup, pred = [0] * c, [0] * c, [0] * c
    for row, t in zip(logits.tolist(), target.tolist()):
        if t == ignore_index:
            continue
        p = max(range(c), key=lambda j: (row[j], -j))
        sup[t] += 1
        pred[p] += 1
        tp[t] += p == t
    seen = [j for j in range(c) if sup[j] + pred[j] > 0]
    return sum((tp[j] / sup[j]) if sup[j] else 0.0 for j in seen) / len(seen) if seen else 0.0


def pnr_meter(logits: Tensor, labels: Tensor, batch: Tensor, start_frame: Tensor, end_frame: Tensor,
              pnr_frame: Tensor) -> Dict[str, float]:
    """Ego4dPNRMeter.update + get_logs (utils/meters/ego4d.py:347-389), literal: per-graph loop with .item()s."""
    probs = torch.sigmoid(logits)
    pred, tgt = probs > 0.5, labels > 0.5
    tp, fp = int((pred & tgt).sum()), int((pred & ~tgt).sum())
    tn, fn = int((~pred & ~tgt).sum()), int((~pred & tgt).sum())
    errs = []
    for g, (sf, ef, pf) in enumerate(zip(start_frame, end_frame, pnr_frame)):
        p = probs[batch == g]
        loc = torch.argmax(p).item()
        mapped = ((ef - sf) / 16 * loc).item()
        errs.append(abs(mapped - (pf.item() - sf.item())) / 30)
    # BinaryAUROC(thresholds=None): area under the exact ROC curve = P(score_pos > score_neg) + 0.5 P(equal)
    pos, neg = probs[tgt].tolist(), probs[~tgt].tolist()
    auc = (sum((a > b) + 0.5 * (a == b) for a in pos for b in neg) / (len(pos) * len(neg))) if pos and neg else 0.0
    return {"accuracy": (tp + tn) / max(tp + fp + tn + fn, 1), "recall": tp / max(tp + fn, 1), "auroc": auc,
            "localization_error": sum(errs) / max(len(errs), 1)}


def levenshtein(a: Sequence[int], b: Sequence[int]) -> int:
    """``editdistance.eval`` (third party, absent): classic two-row DP."""
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def metric_lta_edit_distance(samples: Tensor, target: Tensor) -> float:
    """min over K sampled sequences of Levenshtein / Z (utils/meters/ego4d.py:410-422): samples [V,Z,K], target [V,Z]."""
    v, z, k = samples.shape
    tot = 0.0
    for g in range(v):
        tgt = target[g].tolist()
        tot += min(levenshtein(samples[g, :, s].tolist(), tgt) for s in range(k)) / z
    return tot / v
