"""Generate golden fixtures by running the REAL reference model code (``/root/reference/models/*.py``,
``graphone.py``) in this container.

The reference's first-party code imports torch_geometric / hydra, which are not installable here, so those
two packages are replaced by thin stand-ins that route to ``oracle/pyg_restated.py`` (third-party restatement)
and to an importlib-based ``hydra.utils.instantiate``.  Everything else that runs is the reference's own,
unmodified code, read from where it lies (nothing is copied).  Run from the repo root:

    python tests/golden/make_golden.py            # rewrites tests/golden/*.pt

The fixtures are small (reduced widths) and are what ``tests/test_oracle_golden.py`` pins the oracle against
and what the ``-m gpu`` parity tests feed to the CUDA path.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("EGOPACK_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import pyg_restated as pyg  # noqa: E402


def install_stubs():
    tg = types.ModuleType("torch_geometric")
    tg_nn = types.ModuleType("torch_geometric.nn")
    for name in ("SAGEConv", "LayerNorm", "PositionalEncoding", "TemporalEncoding", "Sequential", "Linear",
                 "radius_graph", "pool"):
        setattr(tg_nn, name, getattr(pyg, name))
    tg_pool = types.ModuleType("torch_geometric.nn.pool")
    tg_pool.global_max_pool = pyg.global_max_pool
    tg_nn.pool = tg_pool
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data, tg_data.Batch = pyg.Data, pyg.Batch
    tg_utils = types.ModuleType("torch_geometric.utils")
    tg_utils.add_remaining_self_loops = pyg.add_remaining_self_loops
    tg_utils.scatter = pyg.scatter
    tg_tr = types.ModuleType("torch_geometric.transforms")

    class BaseTransform:
        pass

    tg_tr.BaseTransform, tg_tr.RadiusGraph = BaseTransform, pyg.RadiusGraph
    tg_rd = types.ModuleType("torch_geometric.transforms.remove_duplicated_edges")
    tg_rd.RemoveDuplicatedEdges = pyg.RemoveDuplicatedEdges
    tg.nn, tg.data, tg.utils, tg.transforms = tg_nn, tg_data, tg_utils, tg_tr
    sys.modules.update({
        "torch_geometric": tg, "torch_geometric.nn": tg_nn, "torch_geometric.nn.pool": tg_pool,
        "torch_geometric.data": tg_data, "torch_geometric.utils": tg_utils,
        "torch_geometric.transforms": tg_tr, "torch_geometric.transforms.remove_duplicated_edges": tg_rd,
    })

    hydra = types.ModuleType("hydra")
    hutils = types.ModuleType("hydra.utils")

    def instantiate(cfg, *args, **kwargs):
        cfg = dict(cfg)
        mod, cls = cfg.pop("_target_").rsplit(".", 1)
        return getattr(importlib.import_module(mod), cls)(*args, **cfg, **kwargs)

    hutils.instantiate = instantiate
    hydra.utils = hutils
    sys.modules.update({"hydra": hydra, "hydra.utils": hutils})


def cross_check_real_pyg():
    """If a REAL torch_geometric is importable (it is not in this image), hold the restated third-party layer to it before
    any fixture is written: the fixtures pin the reference's first-party code ON TOP of oracle/pyg_restated.py, so a wrong
    restatement (SURVEY 8c assumptions A1-A9) must fail loudly here rather than be baked into the golden files."""
    try:
        import torch_geometric                     # noqa: F401 -- only succeeds before install_stubs() replaces it
        from torch_geometric import nn as gnn
        from torch_geometric.nn import radius_graph
    except Exception:
        return "torch_geometric not importable: third-party layer unpinned (restated from the published algorithms)"
    g = torch.Generator().manual_seed(0)
    x = torch.randn(40, 16, generator=g)
    pos = torch.arange(40)
    batch = torch.arange(4).repeat_interleave(10)
    ei = radius_graph(pos.view(-1, 1).float(), 2.5, batch)
    mine = pyg.radius_graph(pos, 2.5, batch)
    canon = lambda e: sorted(zip(e[1].tolist(), e[0].tolist()))
    assert canon(ei) == canon(mine), "radius_graph restatement disagrees with torch_cluster (A7)"
    for kw in (dict(project=True), dict(aggr="max", bias=False)):
        real, ours = gnn.SAGEConv(16, 16, **kw), pyg.SAGEConv(16, 16, **kw)
        ours.load_state_dict(real.state_dict())
        assert torch.allclose(real(x, ei), ours(x, ei), atol=1e-6), f"SAGEConv{kw} restatement disagrees (A1/A2)"
    real, ours = gnn.LayerNorm(16), pyg.LayerNorm(16)
    assert torch.allclose(real(x), ours(x), atol=1e-6), "graph-mode LayerNorm restatement disagrees (A3)"
    real, ours = gnn.PositionalEncoding(16), pyg.PositionalEncoding(16)
    assert torch.allclose(real(pos), ours(pos), atol=1e-6), "PositionalEncoding restatement disagrees (A4)"
    return f"restated layer checked against torch_geometric {torch_geometric.__version__}"


def synth_graphs(gen, sizes, feat, segs, n_verb, n_noun, unlabeled=0.3, pos_shift=0):
    out = []
    for n in sizes:
        y = torch.stack([torch.randint(0, n_verb, (n,), generator=gen),
                         torch.randint(0, n_noun, (n,), generator=gen)], 1)
        drop = torch.rand(n, generator=gen) < unlabeled
        y[drop] = -1
        out.append(pyg.Data(x=torch.randn(n, segs, feat, generator=gen), y=y,
                            pos=torch.arange(n, dtype=torch.long) - pos_shift))
    return out


def grads_of(module):
    return {k: p.grad.clone() for k, p in module.named_parameters() if p.grad is not None}


def main():
    third_party = cross_check_real_pyg()
    import json
    json.dump({"torch": torch.__version__, "third_party_layer": third_party,
               "reference": REF, "note": "fixtures = outputs of the reference's own models/*.py on top of oracle/pyg_restated.py"},
              open(os.path.join(OUT, "VERSIONS.json"), "w"), indent=1)
    install_stubs()
    sys.path.insert(0, REF)
    from models.graph import Graph                                       # reference code, unmodified
    from models.tasks import RecognitionTask, OSCCTask, LTATask, PNRTask
    from models.graphONE.graphONE import GraphONE
    from models.transforms.lta_temp_connectivity import LTATemporalConnectivity
    import graphone as ref_graphone

    torch.manual_seed(1)
    gen = torch.Generator().manual_seed(1234)
    D, S, H, HT, NV, NN = 24, 3, 32, 40, 7, 11
    tp = {"_target_": "models.temporal_pooling.trn_pooling.TRNPooling", "dropout": 0.0, "hidden_size": HT}

    # ---- case 1: Graph fwd+bwd on band graphs (k=2), including a 1-node graph (isolated) -------------------
    model = Graph(D, hidden_size=H, depth=2, temporal_pooling=tp, num_segments=S)
    graphs = synth_graphs(gen, (5, 1, 9, 4), D, S, NV, NN, pos_shift=2)
    tr = pyg.RadiusGraph(r=2.5, loop=False)
    batch = pyg.Batch.from_data_list([tr(g) for g in graphs])
    batch.x.requires_grad_(True)
    out = model(batch)
    w = torch.randn(out.shape, generator=gen)
    (out * w).sum().backward()
    torch.save({
        "cfg": dict(input_size=D, hidden_size=H, depth=2, num_segments=S, trn_hidden=HT, k=2),
        "state": model.state_dict(), "x": batch.x.detach(), "pos": batch.pos, "batch": batch.batch,
        "ptr": batch.ptr, "edge_index": batch.edge_index, "y": batch.y, "out": out.detach(), "w": w,
        "grad_x": batch.x.grad.clone(), "grads": grads_of(model),
    }, os.path.join(OUT, "graph_band.pt"))

    # ---- case 2: LTA connectivity transform, incl. the verb-label-0 quirk ----------------------------------
    lta_cases = []
    for n_in, n_fc, r, zero_at in ((2, 6, 1.5, None), (2, 6, 1.5, 3), (3, 5, 2.5, 0), (1, 4, 3.5, None), (4, 3, 1.5, 6)):
        y = torch.full((n_in + n_fc, 2), -1, dtype=torch.long)
        y[n_in:, 0] = torch.randint(1, NV, (n_fc,), generator=gen)
        y[n_in:, 1] = torch.randint(0, NN, (n_fc,), generator=gen)
        if zero_at is not None:
            y[min(zero_at, n_in + n_fc - 1), 0] = 0 if zero_at >= n_in else -1
            if zero_at >= n_in:
                y[zero_at, 0] = 0
        d = pyg.Data(x=torch.zeros(n_in + n_fc, 1), y=y, pos=torch.arange(n_in + n_fc))
        d = LTATemporalConnectivity(r=r)(d)
        lta_cases.append({"y": y, "r": r, "edge_index": d.edge_index})
    torch.save(lta_cases, os.path.join(OUT, "lta_edges.pt"))

    # ---- case 3: task heads with late-fusion aux features ---------------------------------------------------
    C = 32
    feat = torch.randn(19, H, generator=gen)
    gbatch = torch.tensor([0] * 5 + [1] * 1 + [2] * 9 + [3] * 4)
    aux_all = {t: torch.randn(19, C, generator=gen) for t in ("pnr", "ar", "oscc", "lta")}
    ar = RecognitionTask(H, C, heads=(NV, NN), aux_tasks=("lta", "oscc", "pnr"))
    lta = LTATask(H, C, heads=(NV, NN), aux_tasks=("ar", "oscc", "pnr"), average_logits=False)
    oscc = OSCCTask(H, C, aux_tasks=("ar", "lta", "pnr"), average_logits=True)
    pnr = PNRTask(H, C, aux_tasks=("ar", "lta", "oscc"))
    y_ar = batch.y
    y_oscc = torch.tensor([1, 0, 1, 1])
    y_pnr = torch.zeros(19)
    y_pnr[[2, 5, 9, 16]] = 1
    heads = {}
    for name, task, kw, y in (("ar", ar, {}, y_ar), ("lta", lta, {}, y_ar),
                              ("oscc", oscc, {"batch": gbatch}, y_oscc), ("pnr", pnr, {}, y_pnr)):
        task.zero_grad()
        f = feat.clone().requires_grad_(True)
        a = {t: v.clone().requires_grad_(True) for t, v in aux_all.items() if t != name}   # dict order kept
        ff = task.forward_features(f)
        plain = task.forward_logits(ff, **kw)
        fused = task.forward_logits(features=ff, aux_features=a, **kw)
        loss = task.compute_loss(fused, y)
        loss.mean().backward()
        heads[name] = {"state": task.state_dict(), "features": ff.detach(),
                       "plain": [t.detach() for t in plain] if isinstance(plain, tuple) else plain.detach(),
                       "fused": [t.detach() for t in fused] if isinstance(fused, tuple) else fused.detach(),
                       "loss": loss.detach(), "grad_feat": f.grad.clone(),
                       "grad_aux": {t: v.grad.clone() for t, v in a.items()}, "grads": grads_of(task)}
    torch.save({"feat": feat, "batch": gbatch, "aux": aux_all, "y_ar": y_ar, "y_oscc": y_oscc, "y_pnr": y_pnr,
                "C": C, "H": H, "heads": (NV, NN), "tasks": heads}, os.path.join(OUT, "task_heads.pt"))

    # ---- case 4: GraphONE literal interaction --------------------------------------------------------------
    banks = {"ar": torch.randn(37, C, generator=gen) / 3, "lta": torch.randn(29, C, generator=gen) / 3}
    banks["lta"][5] = banks["lta"][4]                                    # duplicate prototype
    go_cases = []
    for residual, k, depth in ((False, 4, 2), (True, 3, 3)):
        go = GraphONE({t: b.clone() for t, b in banks.items()}, features_size=C, hidden_size=48, k=k,
                      depth=depth, residual=residual)
        feats = {"lta": torch.randn(19, C, generator=gen).requires_grad_(True),
                 "ar": torch.randn(19, C, generator=gen).requires_grad_(True)}
        out, closest = go.interact(feats)
        ws = {t: torch.randn(o.shape, generator=gen) for t, o in out.items()}
        sum((out[t] * ws[t]).sum() for t in out).backward()
        go_cases.append({"cfg": dict(features_size=C, hidden_size=48, k=k, depth=depth, residual=residual),
                         "state": go.state_dict(), "banks": banks,
                         "feats": {t: f.detach() for t, f in feats.items()},
                         "out": {t: o.detach() for t, o in out.items()},
                         "closest": {t: [c.clone() for c in cs] for t, cs in closest.items()},
                         "w": ws, "grad_feats": {t: f.grad.clone() for t, f in feats.items()},
                         "grads": grads_of(go)})
    torch.save(go_cases, os.path.join(OUT, "graphone.pt"))

    # ---- case 5: prototype bank builder (graphone.py) incl. the len(tasks)x bincount quirk -----------------
    model.eval()
    ar2 = RecognitionTask(H, C, heads=(3, 4))
    lta2 = LTATask(H, C, heads=(3, 4))
    pnr2 = PNRTask(H, C)
    loader = []
    for _ in range(3):
        gs = synth_graphs(gen, (6, 7, 3), D, S, 3, 4, unlabeled=0.4)
        loader.append(pyg.Batch.from_data_list([tr(g) for g in gs]))
    _tqdm = ref_graphone.tqdm
    ref_graphone.tqdm = lambda it, *a, **k: it
    built = ref_graphone.build_graphone(model, ar2, [ar2, lta2, pnr2], loader, device="cpu")
    ref_graphone.tqdm = _tqdm
    torch.save({"graph_state": model.state_dict(), "ar": ar2.state_dict(), "lta": lta2.state_dict(),
                "pnr": pnr2.state_dict(), "heads": (3, 4),
                "batches": [{"x": b.x, "pos": b.pos, "batch": b.batch, "ptr": b.ptr, "edge_index": b.edge_index,
                             "y": b.y} for b in loader],
                "banks": built}, os.path.join(OUT, "bank_builder.pt"))
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
