"""Pins the oracle against fixtures produced by the reference's own model code (tests/golden/make_golden.py)."""
import pytest
import torch

from oracle import egopack_oracle as eo
from oracle import pyg_restated as pyg


def _data(d):
    b = pyg.Data(x=d["x"].clone(), pos=d["pos"], edge_index=d["edge_index"], y=d["y"])
    b.batch, b.ptr = d["batch"], d["ptr"]
    return b


def close(a, b, tol=1e-6):
    assert a.shape == b.shape
    denom = b.abs().max().clamp(min=1e-12)
    assert float((a - b).abs().max() / denom) <= tol, float((a - b).abs().max() / denom)


def test_graph_forward_backward_matches_reference(golden):
    g = golden("graph_band.pt")
    c = g["cfg"]
    m = eo.GraphOracle(c["input_size"], c["hidden_size"], c["depth"], temporal_pooling={"hidden_size": c["trn_hidden"]},
                       num_segments=c["num_segments"])
    assert set(m.state_dict().keys()) == set(g["state"].keys())
    m.load_state_dict(g["state"])
    data = _data(g)
    data.x.requires_grad_(True)
    out = m(data)
    close(out, g["out"])
    (out * g["w"]).sum().backward()
    close(data.x.grad, g["grad_x"], 1e-5)
    for k, p in m.named_parameters():
        close(p.grad, g["grads"][k], 1e-5)


def test_band_edges_in_fixture_are_the_band(golden):
    g = golden("graph_band.pt")
    e = pyg.radius_graph(g["pos"], g["cfg"]["k"] + 0.5, g["batch"])
    key = lambda t: sorted(zip(t[1].tolist(), t[0].tolist()))
    assert key(e) == key(g["edge_index"])
    # brute force: same graph, 0 < |dpos| <= k
    want = [(i, j) for i in range(len(g["pos"])) for j in range(len(g["pos"]))
            if i != j and g["batch"][i] == g["batch"][j] and abs(int(g["pos"][i] - g["pos"][j])) <= g["cfg"]["k"]]
    assert key(e) == sorted(want)


def test_lta_connectivity_matches_reference(golden):
    for case in golden("lta_edges.pt"):
        n = case["y"].shape[0]
        d = pyg.Data(x=torch.zeros(n, 1), y=case["y"], pos=torch.arange(n))
        d = eo.lta_temporal_connectivity(d, case["r"])
        assert torch.equal(d.edge_index, case["edge_index"])


def _mk_task(name, g, aux):
    H, C, heads = g["H"], g["C"], g["heads"]
    if name == "ar":
        return eo.RecognitionTaskOracle(H, C, heads, aux_tasks=aux)
    if name == "lta":
        return eo.LTATaskOracle(H, C, heads, aux_tasks=aux)
    if name == "oscc":
        return eo.OSCCTaskOracle(H, C, aux_tasks=aux, average_logits=True)
    return eo.PNRTaskOracle(H, C, aux_tasks=aux)


def test_task_heads_match_reference(golden):
    g = golden("task_heads.pt")
    for name, ref in g["tasks"].items():
        aux_names = tuple(t for t in ("ar", "lta", "oscc", "pnr") if t != name)
        task = _mk_task(name, g, aux_names)
        assert set(task.state_dict().keys()) == set(ref["state"].keys())
        task.load_state_dict(ref["state"])
        f = g["feat"].clone().requires_grad_(True)
        aux = {t: v.clone().requires_grad_(True) for t, v in g["aux"].items() if t != name}
        kw = {"batch": g["batch"]} if name == "oscc" else {}
        y = {"ar": g["y_ar"], "lta": g["y_ar"], "oscc": g["y_oscc"], "pnr": g["y_pnr"]}[name]
        ff = task.forward_features(f)
        close(ff, ref["features"])
        plain = task.forward_logits(ff, **kw)
        fused = task.forward_logits(features=ff, aux_features=aux, **kw)
        if isinstance(plain, tuple):
            for a, b in zip(plain, ref["plain"]):
                close(a, b)
            for a, b in zip(fused, ref["fused"]):
                close(a, b)
        else:
            close(plain, ref["plain"])
            close(fused, ref["fused"])
        loss = task.compute_loss(fused, y)
        close(loss, ref["loss"])
        loss.mean().backward()
        close(f.grad, ref["grad_feat"], 1e-5)
        for t, v in aux.items():
            close(v.grad, ref["grad_aux"][t], 1e-5)
        for k, p in task.named_parameters():
            if k in ref["grads"]:
                close(p.grad, ref["grads"][k], 1e-5)


def test_graphone_literal_and_reduced_match_reference(golden):
    for case in golden("graphone.pt"):
        go = eo.GraphONEOracle({t: b.clone() for t, b in case["banks"].items()}, **case["cfg"])
        assert set(go.state_dict().keys()) == set(case["state"].keys())
        go.load_state_dict(case["state"])
        for fn, tol in ((go.interact, 1e-6), (go.interact_reduced, 2e-5)):
            go.zero_grad()
            feats = {t: f.clone().requires_grad_(True) for t, f in case["feats"].items()}
            out, closest = fn(feats)
            assert list(out.keys()) == list(case["out"].keys())
            for t in out:
                close(out[t], case["out"][t], tol)
                for a, b in zip(closest[t], case["closest"][t]):
                    assert torch.equal(a, b)
            sum((out[t] * case["w"][t]).sum() for t in out).backward()
            for t in feats:
                close(feats[t].grad, case["grad_feats"][t], 10 * tol)
            for k, p in go.named_parameters():
                if k in case["grads"]:
                    close(p.grad, case["grads"][k], 10 * tol)


def test_bank_builder_matches_reference(golden):
    g = golden("bank_builder.pt")
    st = g["graph_state"]
    D = st["temporal_pooling.proj.0.weight"].shape[1] // 3
    H = st["temporal_pooling.proj.8.weight"].shape[0]
    HT = st["temporal_pooling.proj.0.weight"].shape[0]
    depth = sum(1 for k in st if k.endswith("lin_r.weight"))
    m = eo.GraphOracle(D, H, depth, temporal_pooling={"hidden_size": HT}, num_segments=3)
    m.load_state_dict(st)
    C = g["ar"]["net.4.weight"].shape[0]
    ar = eo.RecognitionTaskOracle(H, C, g["heads"]); ar.load_state_dict(g["ar"])
    lta = eo.LTATaskOracle(H, C, g["heads"]); lta.load_state_dict(g["lta"])
    pnr = eo.PNRTaskOracle(H, C); pnr.load_state_dict(g["pnr"])
    banks = eo.build_graphone(m, ar, [ar, lta, pnr], [_data(b) for b in g["batches"]])
    for t in banks:
        close(banks[t], g["banks"][t])


# ---- headline metrics: hand-derived known answers (utils/meters/ego4d.py has no tests of its own) -------------------
def test_oracle_meters_known_answers():
    from oracle import egopack_oracle as eo
    assert eo.levenshtein([1, 2, 3], [1, 2, 3]) == 0
    assert eo.levenshtein([1, 2, 3], [1, 3]) == 1                       # one deletion
    assert eo.levenshtein([1, 2, 3, 4], [2, 3, 4, 5]) == 2              # delete front, insert back
    assert eo.levenshtein(list("kitten"), list("sitting")) == 3        # the textbook pair
    assert eo.levenshtein([], [7, 7]) == 2
    logits = torch.tensor([[0.1, 0.9, 0.0], [0.5, 0.5, 0.2], [0.3, 0.2, 0.9], [1.0, 0.0, 0.0]])
    y = torch.tensor([1, 1, 0, -1])
    # row 0: label is the arg-max; row 1: tie with class 0, lower index wins -> label ranks 2nd; row 2: label ranks 2nd
    assert eo.topk_accuracy(logits, y, 1) == pytest.approx(1 / 3)
    assert eo.topk_accuracy(logits, y, 2) == pytest.approx(1.0)
    # classes: 0 -> support 1, tp 0; 1 -> support 2, tp 1; 2 -> predicted once, never a target -> counted with score 0
    assert eo.macro_accuracy(logits, y) == pytest.approx((0.0 + 0.5 + 0.0) / 3)
    # PNR: graph 0 arg-max node 2 of 4; (ef - sf) / 16 * 2 = 20 frames vs pnr offset 25 -> 5 / 30 s
    lg = torch.tensor([-1.0, 0.0, 2.0, 1.0, 3.0, -3.0])
    lab = torch.tensor([0.0, 0.0, 1.0, 0.0, 0.0, 1.0])
    out = eo.pnr_meter(lg, lab, torch.tensor([0, 0, 0, 0, 1, 1]), torch.tensor([100, 0]), torch.tensor([260, 32]),
                       torch.tensor([125, 10]))
    assert out["localization_error"] == pytest.approx((5 / 30 + 10 / 30) / 2)
    assert out["accuracy"] == pytest.approx(3 / 6) and out["recall"] == pytest.approx(1 / 2)
    assert out["auroc"] == pytest.approx(3 / 8)   # positives {2, -3} vs negatives {-1, 0, 1, 3}: 2 beats three of them
