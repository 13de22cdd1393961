"""GPU: BASELINE.json full-size shapes checked through size-independent properties (the CPU oracle cannot run
these sizes in seconds): linearity / conservation of the aggregation, statistics of the normalisations, GEMM rows
against fp64 on a sample, k-NN sets against exact fp64 on sampled rows."""
import pytest
import torch

from egopack_b200 import ops
from egopack_b200.ops import ACT_LEAKY, ACT_NONE
from tests.gpu_util import DEV, rel_max

pytestmark = pytest.mark.gpu


def _band(v, n, k):
    batch = torch.arange(v, device=DEV).repeat_interleave(n)
    ptr = torch.arange(v + 1, device=DEV) * n
    return ops.band_structure(batch, ptr, k), batch, ptr


@pytest.mark.parametrize("v,n,k,c,dtype", [(256, 128, 1, 1024, torch.bfloat16), (256, 128, 1, 1024, torch.float32),
                                           (64, 2048, 16, 1024, torch.bfloat16)])
def test_aggregation_properties_at_full_size(v, n, k, c, dtype):
    gs, batch, ptr = _band(v, n, k)
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(v * n, c, device=DEV, generator=g).to(dtype)
    y = ops.SageMean.apply(x, gs)
    # (1) a constant field is a fixed point of the mean (every node here has >= 1 neighbour)
    ones = torch.ones(v * n, c, device=DEV, dtype=dtype)
    assert torch.equal(ops.SageMean.apply(ones, gs), ones)
    # (2) conservation: sum_i deg_i * agg_i == sum_j deg_j * x_j for a symmetric band (checked per graph, fp64)
    deg = (1.0 / gs.inv_deg.double())
    lhs = (y.double() * deg[:, None]).view(v, n, c).sum(1)
    rhs = (x.double() * deg[:, None]).view(v, n, c).sum(1)
    assert rel_max(lhs, rhs) < (1e-5 if dtype == torch.float32 else 2e-2)
    # (3) adjointness: <A x, w> == <x, A^T w> ties the backward kernel to the forward one
    w = torch.randn(v * n, c, device=DEV, generator=g).to(dtype)
    xr = x.clone().requires_grad_(True)
    ops.SageMean.apply(xr, gs).backward(w)
    a = (y.double() * w.double()).sum()
    b = (x.double() * xr.grad.double()).sum()
    assert abs(float(a - b)) / abs(float(a)) < (1e-5 if dtype == torch.float32 else 2e-2)
    # (4) spot rows against a direct fp64 window mean
    for i in (0, 1, n - 1, n, v * n - 1, v * n // 2 + 7):
        lo, hi = int(gs.win_lo[i]), int(gs.win_hi[i])
        nb = [j for j in range(lo, hi + 1) if j != i]
        want = x[nb].double().mean(0)
        assert rel_max(y[i], want) < (1e-5 if dtype == torch.float32 else 2e-2)


def test_graph_layernorm_statistics_at_full_size():
    n, c = 32768, 1024
    g = torch.Generator(device=DEV).manual_seed(1)
    x = (torch.randn(n, c, device=DEV, generator=g) * 3 + 1).to(torch.bfloat16)
    w = torch.ones(c, device=DEV)
    b = torch.zeros(c, device=DEV)
    y = ops.GraphLayerNorm.apply(x, w, b, 1e-5, ACT_NONE, 0.0).double()
    assert abs(float(y.mean())) < 1e-3 and abs(float(y.std(unbiased=False)) - 1) < 2e-3
    yl = ops.GraphLayerNorm.apply(x, w, b, 1e-5, ACT_LEAKY, 0.2).double()
    assert rel_max(yl, torch.where(y > 0, y, 0.2 * y)) < 1e-2


def test_gemm_sampled_rows_at_full_size():
    m, k, n = 32768, 4608, 1024                                    # TRNPooling's first Linear at V=256
    g = torch.Generator(device=DEV).manual_seed(2)
    x = torch.randn(m, k, device=DEV, generator=g).to(torch.bfloat16)
    w = (torch.randn(n, k, device=DEV, generator=g) / 68).to(torch.bfloat16)
    bias = torch.randn(n, device=DEV, generator=g)
    y = ops.gemm(x, False, w, False, m, n, k, bias=bias, out_dtype=torch.float32)
    rows = torch.tensor([0, 1, 127, 128, 4095, 20000, m - 1], device=DEV)
    want = x[rows].double() @ w.double().t() + bias.double()
    assert rel_max(y[rows], want) < 2e-5
    dw = ops.gemm(y.to(torch.bfloat16), True, x, True, n, k, m, out_dtype=torch.float32)   # wgrad layout, split-K
    cols = torch.tensor([0, 5, 1023], device=DEV)
    want = y.to(torch.bfloat16)[:, cols].double().t() @ x.double()
    assert rel_max(dw[cols], want) < 1e-4


def test_knn_sampled_rows_at_full_size():
    b, kp, c, k = 32768, 4096, 1024, 4
    g = torch.Generator(device=DEV).manual_seed(3)
    f = torch.randn(b, c, device=DEV, generator=g)
    p = torch.randn(kp, c, device=DEV, generator=g) / 3
    fn, pn = ops.row_normalize(f), ops.row_normalize(p)
    idx = ops.cos_topk(fn, pn, k, ops.row_normalize(f, torch.bfloat16), ops.row_normalize(p, torch.bfloat16))
    rows = torch.arange(0, b, 97, device=DEV)
    d = 1 - (fn[rows].double() @ pn.double().t())
    srt, order = d.sort(1)
    clear = (srt[:, k] - srt[:, k - 1]) > 1e-6
    same = (idx[rows].sort(1).values == order[:, :k].sort(1).values).all(1)
    assert bool(same[clear].all()) and bool((idx[rows, 0] == order[:, 0])[(srt[:, 1] - srt[:, 0]) > 1e-6].all())


def test_band_star_equals_csr_at_the_stacked_c2_size():
    """Round 2: the stacked c2 step aggregates 98 304 nodes (AR | LTA | PNR, 768 graphs) through ONE band+star structure.
    Against the generic CSR kernel on the reference's edge list (egp_lta_edge_* / egp_band_edge_*, bit-exact vs the
    reference's own transform in test_gpu_edges.py): forward and backward, plus the adjointness property."""
    v, n, c = 256, 128, 1024
    g = torch.Generator(device=DEV).manual_seed(5)
    batch = torch.arange(v, device=DEV).repeat_interleave(n)
    ptr = torch.arange(v + 1, device=DEV) * n
    pos = torch.arange(n, device=DEV).repeat(v)
    y = torch.randint(0, 100, (v * n, 2), device=DEV, generator=g)          # verb 0 appears: the `> 0` quirk
    y.view(v, n, 2)[:, :2] = -1
    star = ops.lta_star_counts(y, ptr, 1.5)
    parts = [(batch, ptr, None), (batch, ptr, star), (batch, ptr, None)]
    gs = ops.band_structure_many(parts, 1)
    e_band = ops.band_edge_index(pos, batch, ptr, 1.5)
    e_lta = ops.lta_edge_index(pos, y, batch, ptr, 1.5)
    edges = torch.cat([e_band, e_lta + v * n, e_band + 2 * v * n], 1)
    gc = ops.csr_structure(edges, 3 * v * n)
    assert torch.equal(gs.inv_deg, gc.inv_deg)
    x = torch.randn(3 * v * n, c, device=DEV, generator=g).bfloat16()
    w = torch.randn(3 * v * n, c, device=DEV, generator=g).bfloat16()
    outs = []
    for s in (gs, gc):
        xr = x.clone().requires_grad_(True)
        o = ops.SageMean.apply(xr, s)
        o.backward(w)
        outs.append((o.detach(), xr.grad))
    assert rel_max(outs[0][0], outs[1][0]) < 1e-2 and rel_max(outs[0][1], outs[1][1]) < 1e-2
    a = (outs[0][0].double() * w.double()).sum()
    b = (x.double() * outs[0][1].double()).sum()
    assert abs(float(a - b)) / abs(float(a)) < 2e-2


def test_fused_heads_loss_and_adam_at_full_size():
    """32 768 nodes x (115 verbs, 478 nouns): the fused heads + cross entropy against torch on the same bf16 operands, and
    one FlatAdam step over the ~27 M parameters of the MTL model against torch.optim.Adam."""
    from egopack_b200.models.tasks import RecognitionTask
    from egopack_b200.optim import FlatAdam
    import egopack_b200
    egopack_b200.set_precision("bf16")
    n, H = 32768, 1024
    g = torch.Generator(device=DEV).manual_seed(6)
    torch.manual_seed(6)
    task = RecognitionTask(H, H, (115, 478)).to(DEV)
    f = torch.randn(n, H, device=DEV, generator=g).bfloat16().requires_grad_(True)
    y = torch.stack([torch.randint(0, 115, (n,), device=DEV, generator=g), torch.randint(0, 478, (n,), device=DEV, generator=g)], 1)
    y[torch.rand(n, device=DEV, generator=g) < 0.5] = -1
    loss = task.loss_from_features(f, y)
    loss.mean().backward()
    fr = f.detach().float().requires_grad_(True)
    want = 0
    for h, c in enumerate(task.classifiers):
        lg = torch.nn.functional.linear(fr, c[1].weight.bfloat16().float(), c[1].bias)
        want = want + torch.nn.functional.cross_entropy(lg, y[:, h], ignore_index=-1, reduction="none")
    want.mean().backward()
    assert rel_max(loss, want) < 1e-3 and rel_max(f.grad, fr.grad) < 2e-2
    # optimiser: 27 M parameters in one kernel
    shapes = [(1024, 4608)] + [(1024, 1024)] * 21 + [(1024,)] * 30 + [(478, 1024), (115, 1024), (478,), (115,), (1,)]
    ours = [torch.nn.Parameter(torch.randn(*s, device=DEV, generator=g) * 0.02) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    opt, ropt = FlatAdam(ours, lr=1e-3, weight_decay=1e-5), torch.optim.Adam(ref, lr=1e-3, weight_decay=1e-5)
    for _ in range(2):
        for p, q in zip(ours, ref):
            gr = torch.randn(p.shape, device=DEV, generator=g)
            p.grad, q.grad = gr, gr.clone()
        opt.step()
        ropt.step()
    assert sum(p.numel() for p in ours) > 26_000_000
    for p, q in zip(ours, ref):
        assert rel_max(p, q) < 2e-6
        assert torch.equal(ops.weight_cache.get(p, torch.bfloat16), p.detach().bfloat16())
