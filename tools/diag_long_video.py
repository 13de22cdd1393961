"""Diagnostic (not a pytest file): fp32 gradient error of the long-video graph against the fp64 oracle over several
weight draws, next to the fp32 oracle's own error against the same fp64 truth.  Shows that the O(1e-3) element-wise
gradient differences at this size are activation-kink flips (both fp32 implementations sit at the SAME distance from
fp64 for most draws) and not a defect of either.  Loops over the fp32 GEMM evaluations (FFMA kernel, bf16x6 and bf16x3 on
the tensor cores): each flips different kinks.    python tools/diag_long_video.py"""
import sys, copy, torch
sys.path.insert(0, '.')
import egopack_b200
from egopack_b200 import Batch, synthetic as syn
from egopack_b200.models.graph import Graph
from egopack_b200.models.transforms import RadiusGraph
from oracle import egopack_oracle as eo, pyg_restated as pyg  # noqa: E401
from tests.gpu_util import DEV, rel_max, rel_l2
TP = dict(name="trn")
import tests.test_gpu_models as tm
TP = tm.TP
egopack_b200.set_precision("fp32")
from egopack_b200 import config as _cfg
import itertools
for kind, seed in itertools.product(("ffma", "bf16x6", "bf16x3"), range(12)):
    _cfg.set_fp32_gemm(kind)
    torch.manual_seed(1000 + seed)
    gen = torch.Generator().manual_seed(31)
    D, S, H, HT, k, depth = 32, 3, 128, 96, 16, 4
    b = syn.make_batch("ar", 2, 2048, gen, feature_dim=D, num_segments=S, band_k=k, n_verbs=5, n_nouns=7)
    ref = eo.GraphOracle(D, H, depth, temporal_pooling={"hidden_size": HT}, num_segments=S)
    ref64 = copy.deepcopy(ref).double()
    w = torch.randn(4096, H, generator=gen)
    edges = pyg.radius_graph(b.pos, k + 0.5, b.batch)
    def run_oracle(model, dt):
        rb = pyg.Data(x=b.x.clone().to(dt).requires_grad_(True), pos=b.pos)
        rb.batch, rb.ptr, rb.edge_index = b.batch, b.ptr, edges
        ry = model(rb)
        (ry * w.to(dt)).sum().backward()
        return ry.detach(), rb.x.grad
    ry, rgx = run_oracle(ref, torch.float32)
    ry64, rgx64 = run_oracle(ref64, torch.float64)
    m = Graph(D, H, depth, temporal_pooling=dict(TP, hidden_size=HT), num_segments=S).to(DEV)
    m.load_state_dict(ref.state_dict())
    nb = Batch()
    for key in ("x", "pos", "y", "batch", "ptr"):
        setattr(nb, key, getattr(b, key).to(DEV))
    nb = RadiusGraph(k + 0.5)(nb)
    nb.x.requires_grad_(True)
    y = m(nb)
    (y * w.to(DEV)).sum().backward()
    out = [f"{kind} seed {seed} y {rel_max(y, ry):.1e} | x max {rel_max(nb.x.grad, rgx64):.1e}/{rel_max(rgx, rgx64):.1e} l2 {rel_l2(nb.x.grad, rgx64):.1e}/{rel_l2(rgx, rgx64):.1e}"]
    worst = (0, None)
    for (name, p), (_, rp), (_, rp64) in zip(m.named_parameters(), ref.named_parameters(), ref64.named_parameters()):
        for norm in (rel_max, rel_l2):
            yard = max(norm(rp.grad, rp64.grad), 1e-4 / 3)
            r = norm(p.grad, rp64.grad) / yard
            if r > worst[0]: worst = (r, f"{name} {norm.__name__} {norm(p.grad, rp64.grad):.1e}/{norm(rp.grad, rp64.grad):.1e}")
    print(out[0], "| worst param ratio %.2f %s" % worst, flush=True)
