"""GPU parity: integer work (edge_index construction, CSR) must be bit-exact against the oracle."""
import pytest
import torch

from egopack_b200 import Batch, Data, ops
from egopack_b200.models.transforms import LTATemporalConnectivity, RadiusGraph
from oracle import egopack_oracle as eo
from oracle import pyg_restated as pyg
from tests.gpu_util import DEV, canon_edges, graph_sizes_to_index

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("r", [0.5, 1.5, 2.5, 8.5, 16.5, 20.5])
def test_band_edges_match_radius_graph(r):
    sizes = [1, 2, 7, 40, 3, 100, 33, 34]
    batch, ptr = graph_sizes_to_index(sizes)
    pos = torch.cat([torch.arange(s) - 4 for s in sizes])           # AR-style shifted positions (ego4d_fho.py:224)
    want = pyg.radius_graph(pos, r, batch)
    got = ops.band_edge_index(pos.to(DEV), batch.to(DEV), ptr.to(DEV), r)
    assert got.dtype == torch.int64 and got.shape == want.shape
    assert canon_edges(got) == canon_edges(want)
    dst = got[1].cpu()
    assert bool((dst[1:] >= dst[:-1]).all()), "canonical order is dst-major"


def test_band_edges_unsorted_and_repeated_positions():
    g = torch.Generator().manual_seed(1)
    sizes = [9, 1, 30, 17]
    batch, ptr = graph_sizes_to_index(sizes)
    pos = torch.cat([torch.randint(0, 12, (s,), generator=g) for s in sizes])
    for r in (1.5, 3.5):
        want = pyg.radius_graph(pos, r, batch)
        got = ops.band_edge_index(pos.to(DEV), batch.to(DEV), ptr.to(DEV), r)
        assert canon_edges(got) == canon_edges(want)


def test_band_edges_empty_and_single():
    z = torch.zeros(0, dtype=torch.long, device=DEV)
    got = ops.band_edge_index(z, z, torch.zeros(1, dtype=torch.long, device=DEV), 1.5)
    assert got.shape == (2, 0)
    one = ops.band_edge_index(torch.zeros(1, dtype=torch.long, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV),
                              torch.tensor([0, 1], device=DEV), 1.5)
    assert one.shape == (2, 0)


def test_band_edges_long_video_count_and_order():
    """config 4 shape: 2048-node graphs, k=16 -- closed-form edge count, sortedness, symmetry."""
    v, n, k = 64, 2048, 16
    pos = torch.arange(n).repeat(v).to(DEV)
    batch = torch.arange(v).repeat_interleave(n).to(DEV)
    ptr = (torch.arange(v + 1) * n).to(DEV)
    e = ops.band_edge_index(pos, batch, ptr, k + 0.5)
    assert e.shape[1] == v * (2 * k * n - k * (k + 1))
    src, dst = e[0], e[1]
    assert bool((dst[1:] >= dst[:-1]).all())
    same = dst[1:] == dst[:-1]
    assert bool((src[1:][same] > src[:-1][same]).all())
    assert bool(((src - dst).abs() <= k).all()) and bool((src != dst).all())
    assert bool((batch[src] == batch[dst]).all())
    key = lambda a, b: a * (v * n) + b
    assert torch.equal(key(src, dst).sort().values, key(dst, src).sort().values)     # symmetric


def test_radius_graph_transform_sets_band_hint():
    d = Batch.from_data_list([Data(x=torch.zeros(5, 1), pos=torch.arange(5) - 2), Data(x=torch.zeros(3, 1), pos=torch.arange(3))])
    d = RadiusGraph(r=1.5)(d)
    assert d.band_k == 1 and d.edge_index.device.type == "cpu"
    want = pyg.radius_graph(d.pos, 1.5, d.batch)
    assert canon_edges(d.edge_index) == canon_edges(want)
    d2 = Data(x=torch.zeros(4, 1), pos=torch.tensor([0, 2, 4, 6]))
    assert RadiusGraph(r=2.5)(d2).band_k is None                     # not unit spaced -> generic CSR path


def test_lta_connectivity_golden_and_random(golden):
    for case in golden("lta_edges.pt"):
        n = case["y"].shape[0]
        d = LTATemporalConnectivity(r=case["r"])(Data(x=torch.zeros(n, 1), y=case["y"], pos=torch.arange(n)))
        assert torch.equal(d.edge_index, case["edge_index"])
    g = torch.Generator().manual_seed(2)
    graphs, want, off = [], [], 0
    for _ in range(12):
        n_in, n_fc = int(torch.randint(1, 5, (1,), generator=g)), int(torch.randint(1, 21, (1,), generator=g))
        y = torch.full((n_in + n_fc, 2), -1, dtype=torch.long)
        y[n_in:, 0] = torch.randint(0, 5, (n_fc,), generator=g)      # verb 0 appears: the `> 0` quirk is exercised
        y[n_in:, 1] = torch.randint(0, 5, (n_fc,), generator=g)
        n = n_in + n_fc
        graphs.append(Data(x=torch.zeros(n, 1), y=y, pos=torch.arange(n)))
        ref = eo.lta_temporal_connectivity(pyg.Data(x=torch.zeros(n, 1), y=y, pos=torch.arange(n)), 2.5)
        want.append(ref.edge_index + off)
        off += n
    b = LTATemporalConnectivity(r=2.5)(Batch.from_data_list(graphs))
    assert torch.equal(b.edge_index, torch.cat(want, 1))
    with pytest.raises(ValueError):
        LTATemporalConnectivity(r=2.5, strict=True)(b)


def test_csr_structure_is_deterministic_and_complete():
    g = torch.Generator().manual_seed(3)
    n, e = 200, 1500
    ei = torch.randint(0, n, (2, e), generator=g)
    gs = ops.csr_structure(ei.to(DEV), n)
    for rowptr, col, key, val in ((gs.rowptr_in, gs.col_in, ei[1], ei[0]), (gs.rowptr_out, gs.col_out, ei[0], ei[1])):
        rowptr, col = rowptr.cpu().long(), col.cpu().long()
        assert rowptr[0] == 0 and rowptr[-1] == e
        for i in range(n):
            want = sorted(val[key == i].tolist())
            assert col[rowptr[i]:rowptr[i + 1]].tolist() == want
    deg = torch.bincount(ei[1], minlength=n).clamp(min=1).float()
    assert torch.allclose(gs.inv_deg.cpu(), 1.0 / deg)


def test_device_feeder_uploads_and_builds_edges_on_device():
    """egopack_b200.feed.DeviceFeeder: values survive the pinned async copy, edges are built on the device behind it,
    and match the oracle's radius graph / LTA connectivity (bit-exact after canonical ordering)."""
    from egopack_b200 import synthetic as syn
    from egopack_b200.feed import DeviceFeeder
    from egopack_b200.models.transforms import LTATemporalConnectivity, RadiusGraph
    gen = torch.Generator().manual_seed(5)
    host = [{"ar": syn.make_batch("ar", 3, 9, gen, feature_dim=8, num_segments=2, band_k=1, n_verbs=5, n_nouns=7),
             "lta": syn.make_batch("lta", 2, 22, gen, feature_dim=8, num_segments=2, band_k=1, n_verbs=5, n_nouns=7)}
            for _ in range(3)]
    tf = {"ar": RadiusGraph(r=1.5), "lta": LTATemporalConnectivity(r=1.5)}
    feeder = DeviceFeeder(host, DEV, tf)
    n = 0
    for hb, db in zip(host, feeder):
        n += 1
        for t in ("ar", "lta"):
            assert db[t].x.is_cuda and torch.equal(db[t].x.cpu(), hb[t].x) and torch.equal(db[t].y.cpu(), hb[t].y)
            assert hb[t].x.device.type == "cpu"                      # the host batch is left alone
        want = pyg.radius_graph(hb["ar"].pos, 1.5, hb["ar"].batch)
        assert canon_edges(db["ar"].edge_index) == canon_edges(want) and db["ar"].band_k == 1
        want_lta = []
        off = 0
        for g in range(2):
            sl = slice(int(hb["lta"].ptr[g]), int(hb["lta"].ptr[g + 1]))
            d = pyg.Data(pos=hb["lta"].pos[sl], y=hb["lta"].y[sl])
            want_lta.append(eo.lta_temporal_connectivity(d, 1.5).edge_index + off)
            off += sl.stop - sl.start
        assert canon_edges(db["lta"].edge_index) == canon_edges(torch.cat(want_lta, 1))
    assert n == 3 and feeder.h2d_bytes == 3 * sum(v.numel() * v.element_size() for b in host[0].values()
                                                 for v in (b.x, b.pos, b.y, b.batch, b.ptr))


def test_transforms_are_lazy_and_sync_free_on_device_batches():
    """On a GPU-resident batch the transforms only record the structural hints (band_k, star); edge_index is built on
    first read and equals the eager construction.  A host-side hint (pos_unit_spaced) removes the last read-back."""
    g = torch.Generator().manual_seed(4)
    graphs = []
    for _ in range(6):
        n_in, n_fc = int(torch.randint(1, 4, (1,), generator=g)), int(torch.randint(1, 12, (1,), generator=g))
        y = torch.full((n_in + n_fc, 2), -1, dtype=torch.long)
        y[n_in:] = torch.randint(0, 5, (n_fc, 2), generator=g)
        graphs.append(Data(x=torch.zeros(n_in + n_fc, 1), y=y, pos=torch.arange(n_in + n_fc)))
    host = Batch.from_data_list(graphs)
    eager = LTATemporalConnectivity(r=1.5)(Batch.from_data_list(graphs))           # host batch: eager, returned on the host
    assert eager.is_materialized("edge_index") and eager.edge_index.device.type == "cpu"
    dev = Batch.from_data_list(graphs).to(DEV)
    dev.pos_unit_spaced = True
    dev = LTATemporalConnectivity(r=1.5)(dev)
    assert dev.band_k == 1 and dev.star is not None and not dev.is_materialized("edge_index")
    assert "edge_index" in dev
    assert torch.equal(dev.edge_index.cpu(), eager.edge_index) and dev.is_materialized("edge_index")
    dev2 = RadiusGraph(r=2.5)(Batch.from_data_list(graphs).to(DEV))
    assert dev2.band_k == 2 and not dev2.is_materialized("edge_index")
    assert canon_edges(dev2.edge_index) == canon_edges(pyg.radius_graph(host.pos, 2.5, host.batch))
    # not unit spaced: no hint, eager edges, generic CSR path
    odd = Batch.from_data_list([Data(x=torch.zeros(4, 1), pos=torch.tensor([0, 2, 4, 6]))]).to(DEV)
    odd = RadiusGraph(r=2.5)(odd)
    assert odd.band_k is None and odd.is_materialized("edge_index")


def test_device_feeder_ships_replicated_segments_once():
    """PNR features are one vector per node repeated over the segments (data/ego4d_oscc.py:291).  A loader that hands them
    over as a stride-0 view keeps that form through collation, the feature-dtype conversion and pinning; the feeder copies
    the base only and repeats it on the device: the consumer sees the same [N, R, D] tensor, in the fused allocation."""
    from egopack_b200 import Batch, Data, synthetic as syn
    from egopack_b200.data import replicated_base
    from egopack_b200.feed import DeviceFeeder
    from egopack_b200.models.transforms import RadiusGraph
    D, S = 16, 3
    full = syn.make_batch("pnr", 3, 8, torch.Generator().manual_seed(9), feature_dim=D, num_segments=S)
    comp = syn.make_batch("pnr", 3, 8, torch.Generator().manual_seed(9), feature_dim=D, num_segments=S, compact=True)
    assert full.x.is_contiguous() and replicated_base(full.x) is None and replicated_base(comp.x) is not None
    assert torch.equal(full.x, comp.x) and torch.equal(full.x[:, 0], full.x[:, 2])
    ar = syn.make_batch("ar", 2, 8, torch.Generator().manual_seed(10), feature_dim=D, num_segments=S, n_verbs=5, n_nouns=7)
    for dtype in (None, torch.bfloat16):
        feeder = DeviceFeeder([{"ar": ar, "pnr": comp}], DEV, RadiusGraph(r=1.5), feature_dtype=dtype)
        (out,) = list(feeder)
        want = full.x if dtype is None else full.x.to(dtype)
        assert out["pnr"].x.is_contiguous() and out["pnr"].x.shape == full.x.shape and torch.equal(out["pnr"].x.cpu(), want)
        item = 4 if dtype is None else 2
        per_batch = lambda b, nx: nx * item + sum(v.numel() * v.element_size() for v in (b.pos, b.y, b.batch, b.ptr))
        assert feeder.h2d_bytes == per_batch(ar, ar.x.numel()) + per_batch(comp, comp.x.numel() // S)
        # one allocation for the step's features: the PNR rows sit right behind the AR rows
        assert out["pnr"].x.data_ptr() == out["ar"].x.data_ptr() + ar.x.numel() * item
        assert replicated_base(comp.x) is not None                       # the host batch is left alone
    # per-sample Data objects collate without materialising the repeat; .to(device) repeats on the device
    samples = [Data(x=torch.randn(5, D).unsqueeze(1).expand(-1, S, -1), pos=torch.arange(5)) for _ in range(3)]
    b = Batch.from_data_list(samples)
    assert replicated_base(b.x) is not None and b.x.shape == (15, S, D)
    want = torch.cat([d.x for d in samples])
    b.to_feature_dtype(torch.bfloat16)
    assert replicated_base(b.x) is not None
    d = b.to(DEV)
    assert d.x.is_contiguous() and torch.equal(d.x.cpu(), want.to(torch.bfloat16))


def test_feeders_share_one_copy_stream_and_the_allocator_cache():
    """torch's caching allocator keeps free blocks per stream: a second feeder (next epoch) must find the staging buffers
    the first one returned, i.e. run without a single cudaMalloc."""
    from egopack_b200 import synthetic as syn
    from egopack_b200.feed import DeviceFeeder
    from egopack_b200.models.transforms import RadiusGraph
    gen = torch.Generator().manual_seed(11)
    items = [{"ar": syn.make_batch("ar", 4, 64, gen, feature_dim=256, num_segments=3, n_verbs=5, n_nouns=7, pin=True),
              "pnr": syn.make_batch("pnr", 4, 64, gen, feature_dim=256, num_segments=3, pin=True, compact=True)}
             for _ in range(4)]
    first = DeviceFeeder(items, DEV, RadiusGraph(r=1.5))
    for out in first:
        out["ar"].x.sum().item()
    torch.cuda.synchronize()
    before = torch.cuda.memory_stats(DEV)["num_device_alloc"]
    second = DeviceFeeder(items, DEV, RadiusGraph(r=1.5))
    assert second._stream is first._stream
    for out in second:
        out["ar"].x.sum().item()
    torch.cuda.synchronize()
    assert torch.cuda.memory_stats(DEV)["num_device_alloc"] == before
