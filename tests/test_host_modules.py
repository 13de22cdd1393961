"""CPU: host-side logic of the drop-in modules -- state_dict compatibility with the reference checkpoints
(golden fixtures were produced by the reference's own classes), construction rules, Data/Batch collation."""
import pytest
import torch

from egopack_b200 import Batch, Data
from egopack_b200.models.graph import Graph
from egopack_b200.models.graphONE.graphONE import GraphONE
from egopack_b200.models.tasks import LTATask, OSCCTask, PNRTask, RecognitionTask
from oracle import pyg_restated as pyg

TP = {"_target_": "models.temporal_pooling.trn_pooling.TRNPooling", "dropout": 0.0}


def test_graph_state_dict_matches_reference(golden):
    g = golden("graph_band.pt")
    c = g["cfg"]
    m = Graph(c["input_size"], c["hidden_size"], c["depth"], temporal_pooling=dict(TP, hidden_size=c["trn_hidden"]),
              num_segments=c["num_segments"])
    assert set(m.state_dict()) == set(g["state"])
    m.load_state_dict(g["state"], strict=True)
    for k, v in m.state_dict().items():
        assert v.shape == g["state"][k].shape


def test_task_and_graphone_state_dicts_match_reference(golden):
    g = golden("task_heads.pt")
    H, C, heads = g["H"], g["C"], g["heads"]
    mk = {"ar": lambda a: RecognitionTask(H, C, heads, aux_tasks=a), "lta": lambda a: LTATask(H, C, heads, aux_tasks=a),
          "oscc": lambda a: OSCCTask(H, C, aux_tasks=a, average_logits=True), "pnr": lambda a: PNRTask(H, C, aux_tasks=a)}
    for name, ref in g["tasks"].items():
        aux = tuple(t for t in ("ar", "lta", "oscc", "pnr") if t != name)
        t = mk[name](aux)
        assert set(t.state_dict()) == set(ref["state"])
        t.load_state_dict(ref["state"], strict=True)
    for case in golden("graphone.pt"):
        go = GraphONE({t: b.clone() for t, b in case["banks"].items()}, **case["cfg"])
        assert set(go.state_dict()) == set(case["state"])
        go.load_state_dict(case["state"], strict=True)
        assert go.task_labels == sorted(case["banks"])
        assert all(not p.requires_grad for p in go.embeddings.parameters())


def test_graph_construction_variants():
    g = Graph(8, 16, depth=0, temporal_pooling=None, num_segments=3)
    assert not hasattr(g, "net") and g.temporal_pooling is None
    from egopack_b200.models.temporal_pooling.trn_pooling import TRNPooling
    g = Graph(8, 16, depth=1, temporal_pooling=lambda i, h, s: TRNPooling(i, h, s, hidden_size=24), num_segments=3)
    assert g.temporal_pooling.proj[0].in_features == 24 and g.temporal_pooling.proj[0].out_features == 24
    assert g.temporal_pooling.proj[8].out_features == 16
    assert list(g.configure_optimizers(None)) == list(g.parameters())


def test_modules_refuse_cpu_inputs():
    g = Graph(8, 16, depth=1, temporal_pooling=dict(TP, hidden_size=16), num_segments=3)
    d = Data(x=torch.zeros(4, 3, 8), pos=torch.arange(4))
    with pytest.raises(RuntimeError, match="CUDA"):
        g(d)


def test_batch_collation_matches_oracle_collate():
    gen = torch.Generator().manual_seed(0)
    graphs, ref = [], []
    for n in (3, 1, 5):
        x = torch.randn(n, 3, 4, generator=gen)
        ei = torch.randint(0, n, (2, 4), generator=gen)
        y = torch.randint(0, 5, (n, 2), generator=gen)
        graphs.append(Data(x=x, pos=torch.arange(n), y=y, edge_index=ei, band_k=1))
        ref.append(pyg.Data(x=x, pos=torch.arange(n), y=y, edge_index=ei))
    a, b = Batch.from_data_list(graphs), pyg.Batch.from_data_list(ref)
    for k in ("x", "pos", "y", "edge_index", "batch", "ptr"):
        assert torch.equal(getattr(a, k), getattr(b, k)), k
    assert a.band_k == 1 and a.num_graphs == 3
    scalar = Batch.from_data_list([Data(x=torch.zeros(2, 1), y=torch.tensor(1)), Data(x=torch.zeros(3, 1), y=torch.tensor(0))])
    assert scalar.y.tolist() == [1, 0]
