"""GPU parity of the device-side headline metrics (egopack_b200/meters.py; utils/meters/ego4d.py) against the oracle's
literal restatement.  Ranks, arg-max positions and edit distances are integers: bit-exact."""
import pytest
import torch

from egopack_b200 import meters, ops
from oracle import egopack_oracle as eo
from tests.gpu_util import DEV

pytestmark = pytest.mark.gpu


def _logits(n, c, gen, ties=True):
    x = torch.randn(n, c, generator=gen)
    if ties:  # coarse grid => many exactly equal logits, including with the label's
        x = (x * 2).round() / 2
    return x


@pytest.mark.parametrize("n,c", [(1, 2), (37, 5), (300, 115), (257, 478), (64, 1000)])
def test_label_rank_matches_bruteforce(n, c):
    g = torch.Generator().manual_seed(n * 1000 + c)
    x = _logits(n, c, g)
    y = torch.randint(0, c, (n,), generator=g)
    y[torch.rand(n, generator=g) < 0.2] = -1
    want = []
    for row, t in zip(x.tolist(), y.tolist()):
        want.append(-1 if t < 0 else sum(1 for j, v in enumerate(row) if v > row[t] or (v == row[t] and j < t)))
    got = ops.label_rank(x.to(DEV), y.to(DEV))
    assert got.dtype == torch.int32 and got.cpu().tolist() == want
    # strided label column (labels[:, idx]) and a padded logits pitch, as the task heads hand them over
    y2 = torch.stack([y, y.flip(0)], 1).to(DEV)
    xp = torch.zeros(n, c + 3, device=DEV)
    xp[:, :c] = x.to(DEV)
    assert ops.label_rank(xp[:, :c], y2[:, 0]).cpu().tolist() == want


def test_label_rank_empty_and_out_of_range():
    assert ops.label_rank(torch.zeros(0, 7, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV)).numel() == 0
    x = torch.randn(4, 6, device=DEV)
    y = torch.tensor([0, 6, -1, 5], device=DEV)          # 6 is outside [0, C): treated like an ignored row
    r = ops.label_rank(x, y).cpu().tolist()
    assert r[1] == -1 and r[2] == -1 and r[0] >= 0 and r[3] >= 0


def test_recognition_meter_matches_oracle_over_several_updates():
    g = torch.Generator().manual_seed(3)
    nv, nn_ = 115, 478
    m = meters.Ego4dRecognitionMeter(num_verbs=nv, num_nouns=nn_, device=DEV)
    lv, ln, ys, losses = [], [], [], []
    for n in (64, 1, 200):
        a, b = _logits(n, nv, g), _logits(n, nn_, g)
        y = torch.stack([torch.randint(0, nv, (n,), generator=g), torch.randint(0, nn_, (n,), generator=g)], 1)
        y[torch.rand(n, generator=g) < 0.5] = -1            # AR labels 1 node in 9 (data/ego4d_fho.py:222-223)
        loss = torch.rand(n, generator=g)
        m.update((a.to(DEV), b.to(DEV)), y.to(DEV), loss.to(DEV))
        lv.append(a), ln.append(b), ys.append(y), losses.append(loss)
    lv, ln, y, loss = torch.cat(lv), torch.cat(ln), torch.cat(ys), torch.cat(losses)
    logs = m.get_logs()
    for k in (1, 2, 3, 5):
        assert float(logs[f"verbs_top{k}"]) == pytest.approx(eo.topk_accuracy(lv, y[:, 0], k), abs=1e-7)
        assert float(logs[f"nouns_top{k}"]) == pytest.approx(eo.topk_accuracy(ln, y[:, 1], k), abs=1e-7)
    assert float(logs["verbs_mc"]) == pytest.approx(eo.macro_accuracy(lv, y[:, 0]), abs=1e-6)
    assert float(logs["nouns_mc"]) == pytest.approx(eo.macro_accuracy(ln, y[:, 1]), abs=1e-6)
    assert float(logs["loss"]) == pytest.approx(float(loss.double().mean()), rel=1e-6)
    # the oracle's own headline helper (argmax-based) agrees with the k=1 rule
    v1, n1 = eo.metric_ar([lv, ln], y)
    assert float(logs["verbs_top1"]) == pytest.approx(v1, abs=1e-7) and float(logs["nouns_top1"]) == pytest.approx(n1, abs=1e-7)
    assert len(m.print_logs()) == 5


def test_oscc_meter_matches_oracle():
    g = torch.Generator().manual_seed(4)
    m = meters.Ego4dOSCCMeter(device=DEV)
    x, y = _logits(333, 2, g), torch.randint(0, 2, (333,), generator=g)
    m.update(x.to(DEV), y.to(DEV), torch.rand(333, generator=g).to(DEV))
    assert float(m.get_logs()["accuracy"]) == pytest.approx(eo.metric_oscc(x, y), abs=1e-7)


@pytest.mark.parametrize("v,n", [(1, 16), (40, 16), (25, 7)])
def test_pnr_meter_matches_oracle(v, n):
    g = torch.Generator().manual_seed(v * 100 + n)
    logits = torch.randn(v * n, generator=g) * 3
    logits[::5] = 25.0                                      # saturates sigmoid to exactly 1.0: tie -> first node wins
    node = torch.randint(0, n, (v,), generator=g)
    labels = torch.zeros(v * n)
    labels[torch.arange(v) * n + node] = 1.0                # one-hot key frame per graph (data/ego4d_oscc.py:284-286)
    batch = torch.arange(v).repeat_interleave(n)
    sf = torch.randint(0, 1000, (v,), generator=g)
    ef = sf + torch.randint(100, 300, (v,), generator=g)
    pf = sf + torch.randint(0, 100, (v,), generator=g)
    m = meters.Ego4dPNRMeter(device=DEV)
    half = (v // 2) * n                                      # two updates: counters and stored scores accumulate
    if half:
        m.update(logits[:half].to(DEV), labels[:half].to(DEV), batch[:half].to(DEV), sf[:v // 2], ef[:v // 2], pf[:v // 2],
                 torch.rand(half, generator=g).to(DEV))
    m.update(logits[half:].to(DEV), labels[half:].to(DEV), (batch[half:] - v // 2).to(DEV), sf[v // 2:].to(DEV),
             ef[v // 2:].to(DEV), pf[v // 2:].to(DEV), torch.rand(v * n - half, generator=g).to(DEV))
    want = eo.pnr_meter(logits, labels, batch, sf, ef, pf)
    got = m.get_logs()
    assert got["localization_error"] == pytest.approx(want["localization_error"], rel=1e-12, abs=1e-12)
    for k in ("accuracy", "recall", "auroc"):
        assert float(got[k]) == pytest.approx(want[k], abs=1e-6), k
    # arg-max positions themselves (sigmoid applied before the arg-max, as the meter does)
    ptr = torch.arange(v + 1) * n
    loc = ops.segment_argmax(logits.to(DEV), ptr.to(DEV), apply_sigmoid=True).cpu()
    assert loc.tolist() == torch.sigmoid(logits).view(v, n).argmax(-1).tolist()


def test_segment_argmax_ragged_and_empty_graphs():
    vals = torch.tensor([1.0, 3.0, 3.0, -2.0, 5.0, 0.5], device=DEV)
    ptr = torch.tensor([0, 3, 3, 4, 6], device=DEV)         # graph 1 is empty
    assert ops.segment_argmax(vals, ptr).cpu().tolist() == [1, -1, 0, 0]


@pytest.mark.parametrize("n,z,k,classes", [(1, 20, 5, 115), (33, 20, 5, 4), (50, 20, 5, 478), (7, 1, 1, 3), (9, 64, 3, 2),
                                           (300, 20, 5, 10)])
def test_edit_distance_min_matches_levenshtein(n, z, k, classes):
    g = torch.Generator().manual_seed(n + z + k)
    preds = torch.randint(0, classes, (n, z, k), generator=g)
    labels = torch.randint(0, classes, (n, z), generator=g)
    preds[0, :, 0] = labels[0]                               # an exact match: distance 0
    got = ops.edit_distance_min(preds.to(DEV), labels.to(DEV)).cpu().tolist()
    want = [min(eo.levenshtein(preds[i, :, s].tolist(), labels[i].tolist()) for s in range(k)) for i in range(n)]
    assert got == want and got[0] == 0


def test_lta_meter_matches_oracle():
    g = torch.Generator().manual_seed(8)
    v, n, nv, nn_, k = 12, 22, 20, 30, 5
    lv, ln = _logits(v * n, nv, g), _logits(v * n, nn_, g)
    y = torch.stack([torch.randint(0, nv, (v * n,), generator=g), torch.randint(0, nn_, (v * n,), generator=g)], 1)
    y.view(v, n, 2)[:, :2] = -1                              # the two observed nodes carry no label
    preds = (torch.randint(0, nv, (v * n, k), generator=g), torch.randint(0, nn_, (v * n, k), generator=g))
    m = meters.Ego4dLTAMeter(num_verbs=nv, num_nouns=nn_, device=DEV)
    m.update((lv.to(DEV), ln.to(DEV)), y.to(DEV), tuple(p.to(DEV) for p in preds), torch.rand(v * n, generator=g).to(DEV))
    logs = m.get_logs()
    for j, name in enumerate(("verbs", "nouns")):
        want = eo.metric_lta_edit_distance(preds[j].view(v, n, k)[:, 2:], y[:, j].view(v, n)[:, 2:])
        assert float(logs[f"{name}_ed"]) == pytest.approx(want, rel=1e-6)
        assert float(logs[f"{name}_top1"]) == pytest.approx(eo.topk_accuracy((lv, ln)[j], y[:, j], 1), abs=1e-7)
