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


# ---- feed: multi-task loader interleaving (utils/dataloading.py:8-47) and the device feeder's host-side behaviour ----
def test_multiloader_restarts_short_loaders_until_all_completed():
    from egopack_b200.feed import multiloader
    a, b = [1, 2, 3], ["x", "y"]
    # b runs out first and is restarted; the epoch ends when a (the last one still pending) is exhausted
    assert list(multiloader([a, b, None], [1, 1, 1])) == [(1, "x", None), (2, "y", None), (3, "x", None)]
    # weight 0 disables a loader
    assert list(multiloader([a, b], [1, 0])) == [(1, None), (2, None), (3, None)]
    # equal lengths: one pass
    assert list(multiloader([[1, 2], [3, 4]], [1, 1])) == [(1, 3), (2, 4)]
    # (with no active loader at all the reference yields tuples of None forever; mirrored, not exercised)


def test_device_feeder_preserves_order_structure_and_errors():
    from egopack_b200 import Batch, Data
    from egopack_b200.feed import DeviceFeeder
    mk = lambda i: Batch.from_data_list([Data(x=torch.full((3, 2), float(i)), pos=torch.arange(3))])
    items = [{"ar": mk(i), "lta": mk(10 + i), "pnr": None} for i in range(4)]
    seen = []
    tf = {"ar": lambda b: (seen.append("ar"), b)[1], "lta": lambda b: (setattr(b, "tag", 7), b)[1]}
    out = list(DeviceFeeder(items, "cpu", transforms=tf))
    assert [float(o["ar"].x[0, 0]) for o in out] == [0.0, 1.0, 2.0, 3.0]
    assert [float(o["lta"].x[0, 0]) for o in out] == [10.0, 11.0, 12.0, 13.0]
    assert all(o["pnr"] is None and o["lta"].tag == 7 for o in out) and seen == ["ar"] * 4
    assert out[0]["ar"] is not items[0]["ar"] and out[0]["ar"].ptr.tolist() == [0, 3]     # a fresh Batch per step
    for i, _ in enumerate(DeviceFeeder(items, "cpu")):                                        # early exit stops the worker
        if i == 1:
            break

    def broken():
        yield items[0]
        raise ValueError("loader failed")
    with pytest.raises(ValueError, match="loader failed"):
        list(DeviceFeeder(broken(), "cpu"))


def test_index_dtype_is_validated_before_the_c_abi():
    """index tensors cross the C ABI as raw int64_t*; any other dtype must be refused on the host
    (a float32 pos or an int32 batch would otherwise be read out of bounds on the device)."""
    import pytest
    from egopack_b200 import ops
    pos = torch.arange(6, dtype=torch.float32)
    batch = torch.zeros(6, dtype=torch.int64)
    ptr = torch.tensor([0, 6])
    with pytest.raises(TypeError):
        ops.band_edge_index(pos, batch, ptr, 1.5)
    with pytest.raises(TypeError):
        ops.band_structure(batch.int(), ptr, 1)
    with pytest.raises(TypeError):
        ops.csr_structure(torch.zeros((2, 3), dtype=torch.int32), 6)
    with pytest.raises(TypeError):
        ops.lta_edge_index(pos.long(), torch.zeros((6, 2), dtype=torch.int32), batch, ptr, 1.5)


def test_weight_cache_keys_on_param_generation():
    from egopack_b200 import ops
    g0 = ops.param_generation()
    ops.bump_param_generation()
    assert ops.param_generation() == g0 + 1


def test_device_feeder_prefetch_thread_and_host_hints():
    """The helper thread only drives the HOST loader (one item ahead); order, structure, early exit and loader errors
    behave as in the single-thread mode, and the band test is answered on the host copy (pos_unit_spaced)."""
    from egopack_b200 import Batch, Data
    from egopack_b200.feed import DeviceFeeder, unit_spaced_host
    mk = lambda i, pos: Batch.from_data_list([Data(x=torch.full((4, 2), float(i)), pos=pos)])
    items = [(mk(i, torch.arange(4)), mk(10 + i, torch.tensor([0, 2, 4, 6])), None) for i in range(5)]
    out = list(DeviceFeeder(items, "cpu", prefetch_thread=True))
    assert [float(o[0].x[0, 0]) for o in out] == [0.0, 1.0, 2.0, 3.0, 4.0] and all(o[2] is None for o in out)
    assert all(o[0].pos_unit_spaced is True and o[1].pos_unit_spaced is False for o in out)
    assert unit_spaced_host(torch.tensor([0, 1, 2, 0, 1]), torch.tensor([0, 0, 0, 1, 1]))
    assert not unit_spaced_host(torch.tensor([0, 1, 3]), torch.tensor([0, 0, 0]))
    for i, _ in enumerate(DeviceFeeder(items, "cpu", prefetch_thread=True)):      # early exit stops the helper thread
        if i == 1:
            break

    def broken():
        yield items[0]
        raise ValueError("loader failed")
    with pytest.raises(ValueError, match="loader failed"):
        list(DeviceFeeder(broken(), "cpu", prefetch_thread=True))


def test_lazy_attributes_of_the_data_stand_in():
    from egopack_b200 import Data
    calls = []
    d = Data(x=torch.zeros(3, 1), pos=torch.arange(3))
    d.set_lazy("edge_index", lambda dd: (calls.append(1), torch.tensor([[0, 1], [1, 0]]))[1])
    assert "edge_index" in d and "edge_index" in d.keys() and not d.is_materialized("edge_index") and not calls
    assert d.edge_index.shape == (2, 2) and calls == [1] and d.is_materialized("edge_index")
    assert d.edge_index.shape == (2, 2) and calls == [1]                       # built once
    d.set_lazy("edge_index", lambda dd: torch.zeros((2, 0), dtype=torch.long))
    d.edge_index = None                                                         # assignment drops the producer
    assert d.edge_index is None and "edge_index" not in d


def test_replicated_segment_features_keep_their_form_on_the_host():
    """data.replicated_base / Batch.from_data_list / to_feature_dtype: the PNR loader's repeat (data/ego4d_oscc.py:291)
    carried as a stride-0 view, values identical to the materialised form."""
    from egopack_b200 import Batch, Data, synthetic as syn
    from egopack_b200.data import expand_base, replicated_base
    full = syn.make_batch("pnr", 2, 6, torch.Generator().manual_seed(3), feature_dim=8, num_segments=3)
    comp = syn.make_batch("pnr", 2, 6, torch.Generator().manual_seed(3), feature_dim=8, num_segments=3, compact=True,
                          feature_dtype=torch.bfloat16)
    assert replicated_base(full.x) is None and torch.equal(full.x[:, 0], full.x[:, 1])
    base = replicated_base(comp.x)
    assert base is not None and base.shape == (12, 8) and base.dtype == torch.bfloat16
    assert torch.equal(comp.x, full.x.to(torch.bfloat16)) and torch.equal(full.y, comp.y)
    assert replicated_base(torch.randn(4, 3, 8)) is None and replicated_base(torch.randn(4, 8)) is None
    assert replicated_base(expand_base(torch.randn(4, 8), 1)) is None          # one segment: nothing to save
    mixed = [Data(x=expand_base(torch.randn(3, 8), 3), pos=torch.arange(3)), Data(x=torch.randn(2, 3, 8), pos=torch.arange(2))]
    b = Batch.from_data_list(mixed)                                           # not all replicated: plain concatenation
    assert replicated_base(b.x) is None and b.x.shape == (5, 3, 8) and torch.equal(b.x[:3], mixed[0].x)
