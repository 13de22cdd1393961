"""GPU parity of the cosine k-NN (GraphONE.__compute_edges): top-k prototype indices must equal the oracle's.

The reference's ``argsort`` of fp32 ``1 - cos`` is itself ambiguous where two dissimilarities are closer than fp32
resolution, so rows whose fp64 gap at the k/(k+1) boundary (or between the top two, for the nearest-prototype
output) is below 1e-6 are excluded and counted; every other row must match exactly -- as a set for the k
neighbours (max-aggregation is order-free) and exactly for the nearest prototype."""
import pytest
import torch

from egopack_b200 import ops
from oracle import egopack_oracle as eo
from tests.gpu_util import DEV

pytestmark = pytest.mark.gpu


def reference(f, p, k):
    d64 = eo.cos_dissimilarity(f.double(), p.double())
    srt, order = d64.sort(dim=1)
    return order[:, :k], (srt[:, k] - srt[:, k - 1]) > 1e-6, (srt[:, 1] - srt[:, 0]) > 1e-6, eo.cos_dissimilarity(f, p).argsort(-1)[:, :k]


@pytest.mark.parametrize("b,kp,c,k", [(19, 37, 32, 4), (300, 500, 128, 4), (257, 1000, 64, 8), (2048, 4096, 1024, 4),
                                      (64, 40, 16, 20)])
@pytest.mark.parametrize("tensor", [False, True])
def test_cos_topk_matches_oracle(b, kp, c, k, tensor):
    g = torch.Generator().manual_seed(b)
    f = torch.randn(b, c, generator=g)
    p = torch.randn(kp, c, generator=g) / 3
    if kp > 5:
        p[5] = p[4]                                        # duplicated prototype -> exact tie -> lower index first
    want, clear_k, clear_1, want32 = reference(f, p, k)
    fn = ops.row_normalize(f.to(DEV))
    pn = ops.row_normalize(p.to(DEV))
    if tensor:
        idx = ops.cos_topk(fn, pn, k, ops.row_normalize(f.to(DEV), torch.bfloat16), ops.row_normalize(p.to(DEV), torch.bfloat16))
    else:
        idx = ops.cos_topk(fn, pn, k)
    idx = idx.cpu()
    assert idx.dtype == torch.int64 and idx.shape == (b, k)
    assert all(len(set(r)) == k for r in idx.tolist()), "indices within a row are distinct"
    same_set = (idx.sort(1).values == want.sort(1).values).all(1)
    assert bool(same_set[clear_k].all()), f"{int((~same_set[clear_k]).sum())} clear rows differ"
    assert bool((idx[:, 0] == want[:, 0])[clear_1].all())
    # and against the fp32 oracle exactly as the reference computes it, wherever that is unambiguous
    assert bool((idx.sort(1).values == want32.sort(1).values).all(1)[clear_k].all())


def test_row_normalize_matches_reference_formula():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(100, 96, generator=g) * 3
    want = x / x.norm(dim=1, keepdim=True)
    got = ops.row_normalize(x.to(DEV))
    assert float((got.cpu() - want).abs().max()) < 2e-7


def _fp64_reference(f, p, k, chunk=4096):
    """top-k by fp64 cosine on the device (chunked), plus the fp64 gaps that decide which rows are unambiguous."""
    fn = (f.double() / f.double().norm(dim=1, keepdim=True))
    pn = (p.double() / p.double().norm(dim=1, keepdim=True))
    order, gap_k, gap_1 = [], [], []
    for r0 in range(0, f.shape[0], chunk):
        d = 1 - fn[r0:r0 + chunk] @ pn.t()
        srt, o = d.topk(k + 1, dim=1, largest=False)
        order.append(o[:, :k])
        gap_k.append(srt[:, k] - srt[:, k - 1])
        gap_1.append(srt[:, 1] - srt[:, 0])
    return torch.cat(order), torch.cat(gap_k), torch.cat(gap_1)


def _guarded_topk(f, p, k):
    fn, pn = ops.row_normalize(f), ops.row_normalize(p)
    f16, f_err = ops.row_normalize(f, torch.bfloat16, with_round_err=True)
    p16, p_err = ops.row_normalize(p, torch.bfloat16, with_round_err=True)
    before = dict(ops.KNN_STATS)
    idx = ops.cos_topk(fn, pn, k, f16, p16, f_err=f_err, p_err=float(p_err.max()))
    flagged = ops.KNN_STATS["flagged"] - before["flagged"]
    assert ops.KNN_STATS["rows"] - before["rows"] == f.shape[0]
    return idx, flagged, f_err, p_err


@pytest.mark.parametrize("bank", ["class_means", "near_duplicates", "gaussian"])
def test_knn_miss_detector_on_adversarial_banks(bank, capsys):
    """Real prototype banks are class means of correlated features (graphone.py:16-63): cosine gaps between
    neighbouring prototypes fall far below bf16 resolution, exactly where a bf16 candidate pass can lose a true
    neighbour.  The miss detector must send every such row to the exact fp32 path: set-exact top-k and exact top-1 at
    B = 32 768 on (a) class means of a low-rank model, (b) near-duplicate prototypes spaced 1e-3 in cosine, and (c) the
    Gaussian bank of the benchmark (where only a few rows may be flagged, or the guard would cost the speed-up)."""
    b, kp, c, k = 32768, 4096, 1024, 4
    g = torch.Generator(device=DEV).manual_seed(7)
    rn = lambda *s: torch.randn(*s, device=DEV, generator=g)
    if bank == "class_means":
        basis, mean = rn(16, c), 2.0 * rn(1, c)
        p = (rn(kp, 16) @ basis + 0.3 * rn(kp, c) + mean) / 3       # class means: shared mean + low-rank + small noise
        f = rn(b, 16) @ basis + 0.3 * rn(b, c) + mean
    elif bank == "near_duplicates":
        # every prototype is a small perturbation of ONE direction: pairwise cosine ~0.96, and a node's similarities to
        # the 4096 prototypes spread over ~1e-3 -- rank gaps of 1e-4, below the bf16 error of a similarity (~4e-4)
        u = rn(1, c)
        u = u / u.norm()
        unit = lambda t: t / t.norm(dim=1, keepdim=True)
        p = u + 0.2 * unit(rn(kp, c))
        f = u + 0.2 * unit(rn(b, c))
    else:
        p, f = rn(kp, c) / 3, rn(b, c)
    want, gap_k, gap_1 = _fp64_reference(f, p, k)
    idx, flagged, f_err, p_err = _guarded_topk(f, p, k)
    # a row is unambiguous when its fp64 gap exceeds what ANY fp32 evaluation of a 1024-term dot product can resolve:
    # ~1e-6 for the small similarities of the other banks, ~1e-5 when every similarity is ~0.96 (near-duplicates)
    amb = 1e-5 if bank == "near_duplicates" else 1e-6
    clear_k, clear_1 = gap_k > amb, gap_1 > amb
    same = (idx.sort(1).values == want.sort(1).values).all(1)
    with capsys.disabled():
        print(f"\\n[knn {bank}] rows={b} flagged->exact={flagged} ({100.0 * flagged / b:.1f}%)  ambiguous rows excluded: "
              f"top-k {int((~clear_k).sum())}, top-1 {int((~clear_1).sum())}  bound: |df| max {float(f_err.max()):.2e} "
              f"|dp| max {float(p_err.max()):.2e}")
    assert bool(same[clear_k].all()), f"{int((~same[clear_k]).sum())} unambiguous rows differ"
    assert bool((idx[:, 0] == want[:, 0])[clear_1].all())
    assert int((~clear_k).sum()) < b // 8, "the fp64 reference itself is ambiguous on too many rows to mean anything"
    if bank == "gaussian":
        assert flagged < b // 10, "the guard must stay cheap on the benchmark's banks"
    # the unguarded pass really is unsafe on the adversarial banks (otherwise this test proves nothing)
    if bank == "near_duplicates":
        fn, pn = ops.row_normalize(f), ops.row_normalize(p)
        raw = ops.cos_topk(fn, pn, k, ops.row_normalize(f, torch.bfloat16), ops.row_normalize(p, torch.bfloat16), guard=False)
        wrong = int((~(raw.sort(1).values == want.sort(1).values).all(1))[clear_k].sum())
        with capsys.disabled():
            print(f"[knn {bank}] without the guard: {wrong} unambiguous rows wrong")
