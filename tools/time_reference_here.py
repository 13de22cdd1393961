"""Times the reference's OWN model code (``/root/reference/models/*.py``, imported unmodified through the stub loader of
``tests/golden/make_golden.py``) against the oracle port on this machine's CPU, on the c2 workload of ``bench.py``.

``bench.py --impl reference`` has to run on the GPU box, where ``/root/reference`` does not exist, so it times the
oracle (``kind: "port"``).  This tool is the evidence that the two are interchangeable as a CPU baseline: same
state_dict, same batches, same aten operator sequence -- losses agree to fp32 rounding and the step times to within
run-to-run noise.  Not a pytest file; run where the reference checkout is present:

    python tools/time_reference_here.py [graphs_per_task] [steps]      # writes profiles/r2_reference_vs_oracle_cpu.json
"""
from __future__ import annotations

import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import torch  # noqa: E402

import make_golden as mg  # noqa: E402
from egopack_b200 import synthetic as syn  # noqa: E402
from oracle import egopack_oracle as eo  # noqa: E402
from oracle import pyg_restated as pyg  # noqa: E402

HIDDEN, DEPTH, TRN_HIDDEN, DROPOUT, K = 1024, 3, 1024, 0.5, 1


def batches(videos, nodes):
    gen = syn.generator(1, 2, 0)
    out = {}
    for t in ("ar", "lta", "pnr"):
        b = syn.make_batch(t, videos, nodes, gen)
        d = pyg.Data(x=b.x, pos=b.pos, y=b.y)
        d.batch, d.ptr = b.batch, b.ptr
        if t == "lta":
            d.edge_index = torch.cat([eo.lta_temporal_connectivity(
                pyg.Data(x=b.x[g * nodes:(g + 1) * nodes], pos=b.pos[g * nodes:(g + 1) * nodes], y=b.y[g * nodes:(g + 1) * nodes]),
                K + 0.5).edge_index + g * nodes for g in range(videos)], 1)
        else:
            d.edge_index = pyg.radius_graph(b.pos, K + 0.5, b.batch)
        out[t] = d
    return out


def reference_step(model, tasks, data):
    """The body of main_temporal.py:87-128, with the reference's own modules."""
    import torch.nn.functional as F
    feats = {t: model(b) for t, b in data.items()}
    losses = []
    for t in ("ar", "lta", "pnr"):
        task = tasks[t]
        f = task.forward_features(feats[t])
        logits = task.forward_logits(f)
        if t == "pnr":
            loss = F.binary_cross_entropy_with_logits(logits, data[t].y.float(), reduction="none")
        else:
            loss = task.compute_loss(logits, data[t].y)
        losses.append(loss.mean())
    return torch.stack(losses).sum()


def main():
    videos = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    nodes = 128
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    mg.install_stubs()
    sys.path.insert(0, mg.REF)
    from models.graph import Graph                                       # reference code, unmodified
    from models.tasks import LTATask, PNRTask, RecognitionTask

    torch.manual_seed(1)
    heads = (syn.N_VERBS, syn.N_NOUNS)
    tp = {"_target_": "models.temporal_pooling.trn_pooling.TRNPooling", "dropout": DROPOUT, "hidden_size": TRN_HIDDEN}
    ref_model = Graph(syn.FEATURE_DIM, hidden_size=HIDDEN, depth=DEPTH, temporal_pooling=tp, num_segments=syn.NUM_SEGMENTS)
    ref_tasks = {"ar": RecognitionTask(HIDDEN, HIDDEN, heads), "lta": LTATask(HIDDEN, HIDDEN, heads), "pnr": PNRTask(HIDDEN, HIDDEN)}
    ora_model = eo.GraphOracle(syn.FEATURE_DIM, HIDDEN, DEPTH, temporal_pooling={"hidden_size": TRN_HIDDEN, "dropout": DROPOUT},
                               num_segments=syn.NUM_SEGMENTS)
    ora_tasks = {"ar": eo.RecognitionTaskOracle(HIDDEN, HIDDEN, heads), "lta": eo.LTATaskOracle(HIDDEN, HIDDEN, heads),
                 "pnr": eo.PNRTaskOracle(HIDDEN, HIDDEN)}
    ora_model.load_state_dict(ref_model.state_dict())
    for t in ora_tasks:
        ora_tasks[t].load_state_dict(ref_tasks[t].state_dict())
    data = batches(videos, nodes)
    n_nodes = videos * nodes * 3

    def run(model, tasks, step_fn, train):
        params = list(model.parameters()) + [p for t in tasks.values() for p in t.parameters()]
        opt = torch.optim.Adam(params, lr=1e-5, weight_decay=1e-5)
        model.train(train)
        for t in tasks.values():
            t.train(train)
        times, loss = [], None
        for i in range(steps + 1):
            t0 = time.perf_counter()
            opt.zero_grad()
            loss = step_fn(model, tasks, data)
            loss.backward()
            opt.step()
            if i:
                times.append(time.perf_counter() - t0)
        return statistics.median(times), float(loss)

    # losses in eval mode (no dropout) on identical weights, before any update
    ref_model.eval(), ora_model.eval()
    for t in ref_tasks:
        ref_tasks[t].eval(), ora_tasks[t].eval()
    with torch.no_grad():
        l_ref = float(reference_step(ref_model, ref_tasks, data))
        l_ora = float(eo.mtl_step(ora_model, ora_tasks, data)[0])
    t_ref, _ = run(ref_model, ref_tasks, reference_step, True)
    t_ora, _ = run(ora_model, ora_tasks, lambda m, ts, d: eo.mtl_step(m, ts, d)[0], True)
    out = {"graphs_per_task": videos, "nodes_per_step": n_nodes, "threads": threads, "torch": torch.__version__,
           "eval_loss_reference_code": l_ref, "eval_loss_oracle_port": l_ora,
           "reference_code_s_per_step": round(t_ref, 4), "oracle_port_s_per_step": round(t_ora, 4),
           "reference_code_nodes_per_s": round(n_nodes / t_ref, 1), "oracle_port_nodes_per_s": round(n_nodes / t_ora, 1),
           "note": "reference = /root/reference/models/{graph,tasks/*,temporal_pooling/*}.py imported unmodified; its "
                   "torch_geometric layer is oracle/pyg_restated.py in BOTH arms (PyG is not installable here)"}
    print(json.dumps(out, indent=1))
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_reference_vs_oracle_cpu.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
