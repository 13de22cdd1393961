"""GPU bring-up diagnostics (not a pytest file).  Every check runs in its own subprocess with a timeout so that a
trap or a hang in one kernel cannot take the others down; results go to ``gpurun_out/diag.jsonl``.

    python tests/gpu_diag.py              # run everything
    python tests/gpu_diag.py gemm_tc_kk   # one check, in-process
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")

import torch  # noqa: E402


def _err(got: torch.Tensor, want: torch.Tensor):
    got, want = got.double().cpu(), want.double().cpu()
    d = (got - want).abs()
    return {"max_abs": float(d.max()), "rel": float(d.max() / want.abs().max().clamp(min=1e-30)),
            "nan": int(torch.isnan(got).sum())}


def _block_map(got, want, tol, bs=32):
    """Which bs x bs blocks are wrong -- helps to spot swizzle / major / descriptor mistakes."""
    got, want = got.double().cpu(), want.double().cpu()
    bad = ((got - want).abs() > tol * want.abs().max()).float()
    m, n = bad.shape
    rows = []
    for i in range(0, min(m, 256), bs):
        rows.append("".join("X" if bad[i:i + bs, j:j + bs].any() else "." for j in range(0, min(n, 512), bs)))
    return rows


def _gemm_case(dtype, m, n, k, a_trans=False, b_trans=False, k2=0, bias=False, act=0, residual=False, out_dtype=None,
               seed=0):
    from egopack_b200 import ops
    g = torch.Generator().manual_seed(seed)
    dev = "cuda"
    A = torch.randn((k, m) if a_trans else (m, k), generator=g)
    B = torch.randn((k, n) if b_trans else (n, k), generator=g)
    A2 = torch.randn((k2, m) if a_trans else (m, k2), generator=g) if k2 else None
    B2 = torch.randn((k2, n) if b_trans else (n, k2), generator=g) if k2 else None
    bi = torch.randn(n, generator=g) if bias else None
    out_dtype = out_dtype or dtype
    R = torch.randn(m, n, generator=g).to(out_dtype) if residual else None
    q = lambda t: None if t is None else t.to(dtype)
    Aq, Bq, A2q, B2q = q(A), q(B), q(A2), q(B2)
    ref = (Aq.double().t() if a_trans else Aq.double()) @ (Bq.double() if b_trans else Bq.double().t())
    if k2:
        ref = ref + (A2q.double().t() if a_trans else A2q.double()) @ (B2q.double() if b_trans else B2q.double().t())
    if bias:
        ref = ref + bi.double()
    if act == 1:
        ref = ref.relu()
    elif act == 2:
        ref = torch.where(ref > 0, ref, 0.2 * ref)
    if residual:
        ref = ref + R.double()
    d = lambda t: None if t is None else t.to(dev)
    out = ops.gemm(d(Aq), a_trans, d(Bq), b_trans, m, n, k, a2=d(A2q), b2=d(B2q), k2=k2, bias=d(bi), residual=d(R),
                   act=act, slope=0.2, out_dtype=out_dtype)
    torch.cuda.synchronize()
    e = _err(out, ref)
    tol = 2e-2 if out_dtype == torch.bfloat16 else (1e-5 if dtype == torch.float32 else 1e-5)
    e["ok"] = e["rel"] < tol and e["nan"] == 0
    if not e["ok"]:
        e["map"] = _block_map(out, ref, tol)
    return e


def check_device():
    from egopack_b200 import _lib
    import ctypes
    a, b, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.call("egp_device_info", ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return {"sms": a.value, "cc": f"{b.value}.{c.value}", "name": torch.cuda.get_device_name(0), "ok": True}


def check_gemm_simt():
    out = {}
    f = torch.float32
    out["nn_small"] = _gemm_case(f, 100, 70, 50)
    out["tails_bt"] = _gemm_case(f, 300, 130, 77, b_trans=True, bias=True, act=1)
    out["at_bt"] = _gemm_case(f, 64, 200, 300, a_trans=True, b_trans=True)
    out["dual_res"] = _gemm_case(f, 256, 128, 96, k2=64, bias=True, residual=True)
    out["bf16_in"] = _gemm_case(torch.bfloat16, 100, 7, 36, out_dtype=torch.float32)   # unaligned ld -> FFMA path
    out["ok"] = all(v["ok"] for v in out.values())
    return out


def _tc(name_cases):
    out = {}
    for name, kw in name_cases.items():
        out[name] = _gemm_case(torch.bfloat16, **kw)
    out["ok"] = all(v["ok"] for v in out.values() if isinstance(v, dict))
    return out


def check_gemm_tc_kk():
    """K-major A and B (the forward layout)."""
    return _tc({
        "one_tile_one_kb": dict(m=128, n=256, k=64, out_dtype=torch.float32),
        "one_tile_4kb": dict(m=128, n=256, k=256, out_dtype=torch.float32),
        "bn128": dict(m=128, n=128, k=128, out_dtype=torch.float32),
        "bn64": dict(m=128, n=64, k=128, out_dtype=torch.float32),
        "bn32": dict(m=128, n=32, k=128, out_dtype=torch.float32),
        "bn16": dict(m=128, n=16, k=128, out_dtype=torch.float32),
        "multi_tile": dict(m=1024, n=1024, k=512, out_dtype=torch.float32),
        "tails": dict(m=300, n=520, k=200, out_dtype=torch.float32),
        "persistent": dict(m=4096, n=1024, k=1024),
        "bf16_out_epilogue": dict(m=384, n=512, k=192, bias=True, act=1),
        "residual": dict(m=384, n=512, k=192, bias=True, residual=True),
        "leaky_f32out": dict(m=200, n=115, k=1024, bias=True, act=2, out_dtype=torch.float32),
        "dual": dict(m=512, n=256, k=256, k2=320, bias=True),
        "small_k": dict(m=19, n=32, k=72, bias=True),
    })


def check_gemm_tc_bt():
    """MN-major B (dgrad layout)."""
    return _tc({
        "one_tile": dict(m=128, n=256, k=64, b_trans=True, out_dtype=torch.float32),
        "kb4": dict(m=128, n=128, k=256, b_trans=True, out_dtype=torch.float32),
        "multi": dict(m=1024, n=1024, k=512, b_trans=True),
        "tails": dict(m=300, n=200, k=120, b_trans=True, out_dtype=torch.float32),
        "dual": dict(m=256, n=128, k=128, k2=192, b_trans=True),
    })


def check_gemm_tc_at():
    """MN-major A and B (wgrad layout, split-K)."""
    return _tc({
        "one_tile": dict(m=128, n=256, k=64, a_trans=True, b_trans=True, out_dtype=torch.float32),
        "at_only": dict(m=128, n=128, k=256, a_trans=True, out_dtype=torch.float32),
        "splitk": dict(m=256, n=512, k=8192, a_trans=True, b_trans=True, out_dtype=torch.float32),
        "tails": dict(m=120, n=1000, k=2048, a_trans=True, b_trans=True, out_dtype=torch.float32),
        "dual": dict(m=128, n=128, k=512, k2=512, a_trans=True, b_trans=True, out_dtype=torch.float32),
    })


def check_kernels():
    """Memory-bound kernels against plain torch on the same device data."""
    from egopack_b200 import ops
    from egopack_b200.ops import ACT_LEAKY, ACT_RELU
    out = {}
    dev = "cuda"
    g = torch.Generator().manual_seed(3)
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        tag = "f32" if dt == torch.float32 else "bf16"
        n, c = 777, 256
        sizes = [5, 1, 300, 64, 407]
        assert sum(sizes) == n
        batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes)).to(dev)
        ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.long, device=dev)
        x = torch.randn(n, c, generator=g).to(dt).to(dev)
        for k in (1, 2, 6, 16):
            gs = ops.band_structure(batch, ptr, k)
            xr = x.detach().clone().requires_grad_(True)
            y = ops.SageMean.apply(xr, gs)
            w = torch.randn(n, c, generator=g).to(dt).to(dev)
            y.backward(w)
            # reference: dense band mean in fp64
            idx = torch.arange(n, device=dev)
            A = ((idx[:, None] - idx[None, :]).abs() <= k) & (batch[:, None] == batch[None, :]) & (idx[:, None] != idx[None, :])
            A = A.double()
            deg = A.sum(1).clamp(min=1)
            want = (A @ x.double()) / deg[:, None]
            wantg = A.t() @ (w.double() / deg[:, None])
            e1, e2 = _err(y, want), _err(xr.grad, wantg)
            out[f"sage_band_{tag}_k{k}"] = {"fwd": e1["rel"], "bwd": e2["rel"], "ok": e1["rel"] < tol and e2["rel"] < tol}
            # CSR path on the same graph must agree
            ei = A.nonzero().t().contiguous()            # [i (dst), j (src)] -> edge_index = [src; dst]
            ei = torch.stack([ei[1], ei[0]])
            gc = ops.csr_structure(ei, n)
            xr2 = x.detach().clone().requires_grad_(True)
            y2 = ops.SageMean.apply(xr2, gc)
            y2.backward(w)
            e3, e4 = _err(y2, want), _err(xr2.grad, wantg)
            out[f"sage_csr_{tag}_k{k}"] = {"fwd": e3["rel"], "bwd": e4["rel"], "ok": e3["rel"] < tol and e4["rel"] < tol}
        # graph LN + leaky relu
        wgt = (torch.randn(c, generator=g) * 0.5 + 1).to(dev).requires_grad_(True)
        bia = (torch.randn(c, generator=g) * 0.1).to(dev).requires_grad_(True)
        xr = (x.detach().clone() * 1.7 + 0.3).requires_grad_(True)
        y = ops.GraphLayerNorm.apply(xr, wgt, bia, 1e-5, ACT_LEAKY, 0.2)
        wv = torch.randn(n, c, generator=g).to(dt).to(dev)
        y.backward(wv)
        xd = xr.detach().double().requires_grad_(True)
        wd, bd = wgt.detach().double().requires_grad_(True), bia.detach().double().requires_grad_(True)
        xc = xd - xd.mean()
        yd = torch.nn.functional.leaky_relu(xc / (xc.std(unbiased=False) + 1e-5) * wd + bd, 0.2)
        yd.backward(wv.double())
        es = [_err(y, yd.detach())["rel"], _err(xr.grad, xd.grad)["rel"], _err(wgt.grad, wd.grad)["rel"], _err(bia.grad, bd.grad)["rel"]]
        out[f"graph_ln_{tag}"] = {"errs": es, "ok": all(e < (tol if i < 2 else 5 * tol) for i, e in enumerate(es))}
        # row LN + relu
        for cc in (256, 40, 1024):
            xx = torch.randn(n, cc, generator=g).to(dt).to(dev).requires_grad_(True)
            ww = (torch.randn(cc, generator=g) * 0.5 + 1).to(dev).requires_grad_(True)
            bb = (torch.randn(cc, generator=g) * 0.1).to(dev).requires_grad_(True)
            y = ops.RowLayerNorm.apply(xx, ww, bb, 1e-5, ACT_RELU, 0.0)
            wv = torch.randn(n, cc, generator=g).to(dt).to(dev)
            y.backward(wv)
            xd = xx.detach().double().requires_grad_(True)
            wd, bd = ww.detach().double().requires_grad_(True), bb.detach().double().requires_grad_(True)
            yd = torch.nn.functional.layer_norm(xd, (cc,), wd, bd, 1e-5).relu()
            yd.backward(wv.double())
            es = [_err(y, yd.detach())["rel"], _err(xx.grad, xd.grad)["rel"], _err(ww.grad, wd.grad)["rel"], _err(bb.grad, bd.grad)["rel"]]
            out[f"row_ln_{tag}_{cc}"] = {"errs": es, "ok": all(e < (tol if i < 2 else 5 * tol) for i, e in enumerate(es))}
        # segment max pool
        xx = torch.randn(n, c, generator=g).to(dt).to(dev).requires_grad_(True)
        y = ops.SegmentMaxPool.apply(xx, ptr, batch)
        wv = torch.randn(len(sizes), c, generator=g).to(dt).to(dev)
        y.backward(wv)
        xd = xx.detach().double().requires_grad_(True)
        yd = torch.stack([xd[ptr[i]:ptr[i + 1]].max(0).values for i in range(len(sizes))])
        yd.backward(wv.double())
        out[f"segmax_{tag}"] = {"fwd": _err(y, yd.detach())["rel"], "bwd": _err(xx.grad, xd.grad)["rel"]}
        out[f"segmax_{tag}"]["ok"] = out[f"segmax_{tag}"]["fwd"] < 1e-6 and out[f"segmax_{tag}"]["bwd"] < 1e-6
        # posenc
        pos = torch.randint(-4, 200, (n,), generator=g).to(dev)
        freq = torch.logspace(0, 1, c // 2, 1e-4).to(dev)
        y = ops.PosEncAdd.apply(x, pos, freq)
        o = pos.float().view(-1, 1) * freq.view(1, -1)
        want = x.double() + torch.cat([o.sin(), o.cos()], -1).double()
        out[f"posenc_{tag}"] = _err(y, want)
        out[f"posenc_{tag}"]["ok"] = out[f"posenc_{tag}"]["rel"] < tol
        # colsum, cast, max combine, proto gather
        cs = ops.colsum(x)
        out[f"colsum_{tag}"] = _err(cs, x.double().sum(0))
        out[f"colsum_{tag}"]["ok"] = out[f"colsum_{tag}"]["rel"] < 1e-5
    x7 = torch.randn(501, 7, generator=g).to(dev)
    out["colsum_ragged"] = _err(ops.colsum(x7), x7.double().sum(0))
    out["colsum_ragged"]["ok"] = out["colsum_ragged"]["rel"] < 1e-5
    out["ok"] = all(v.get("ok", False) for v in out.values() if isinstance(v, dict))
    return out


def check_edges():
    from egopack_b200 import ops
    from oracle import pyg_restated as pyg
    out = {}
    g = torch.Generator().manual_seed(5)
    sizes = [1, 2, 7, 40, 3, 100]
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.long)
    pos = torch.cat([torch.arange(s) - 4 for s in sizes])
    canon = lambda e: sorted(zip(e[1].tolist(), e[0].tolist()))
    for r in (1.5, 2.5, 16.5, 20.5):
        want = pyg.radius_graph(pos, r, batch)
        got = ops.band_edge_index(pos.cuda(), batch.cuda(), ptr.cuda(), r).cpu()
        out[f"band_r{r}"] = {"E": int(got.shape[1]), "ok": canon(got) == canon(want) and bool((got[1][1:] >= got[1][:-1]).all())}
    # non-monotone positions -> whole-graph scan
    posr = torch.cat([torch.randperm(s, generator=g) for s in sizes])
    want = pyg.radius_graph(posr, 2.5, batch)
    got = ops.band_edge_index(posr.cuda(), batch.cuda(), ptr.cuda(), 2.5).cpu()
    out["band_shuffled"] = {"ok": canon(got) == canon(want)}
    # LTA golden cases, batched together
    gold = torch.load(os.path.join(ROOT, "tests", "golden", "lta_edges.pt"), weights_only=False)
    by_r = {}
    for c in gold:
        by_r.setdefault(c["r"], []).append(c)
    ok = True
    for r, cases in by_r.items():
        ys = torch.cat([c["y"] for c in cases])
        ns = [c["y"].shape[0] for c in cases]
        b = torch.repeat_interleave(torch.arange(len(ns)), torch.tensor(ns))
        p = torch.tensor([0] + list(torch.tensor(ns).cumsum(0)), dtype=torch.long)
        ps = torch.cat([torch.arange(n) for n in ns])
        got = ops.lta_edge_index(ps.cuda(), ys.cuda(), b.cuda(), p.cuda(), r).cpu()
        want = torch.cat([c["edge_index"] + int(p[i]) for i, c in enumerate(cases)], dim=1)
        ok = ok and torch.equal(got, want)
    out["lta_golden"] = {"ok": ok}
    out["ok"] = all(v["ok"] for v in out.values())
    return out


def check_topk():
    from egopack_b200 import ops
    out = {}
    g = torch.Generator().manual_seed(11)
    for b, kp, c, k in ((300, 500, 128, 4), (2048, 4096, 1024, 4), (257, 1000, 64, 8)):
        f = torch.randn(b, c, generator=g)
        p = torch.randn(kp, c, generator=g) / 3
        fn64 = f.double() / f.double().norm(dim=1, keepdim=True)
        pn64 = p.double() / p.double().norm(dim=1, keepdim=True)
        d = 1 - fn64 @ pn64.t()
        srt, order = d.sort(dim=1)
        gap = srt[:, k] - srt[:, k - 1]
        clear = gap > 1e-5
        fn = ops.row_normalize(f.cuda(), torch.float32)
        pn = ops.row_normalize(p.cuda(), torch.float32)
        res = {}
        for mode in ("exact", "tensor"):
            if mode == "tensor":
                idx = ops.cos_topk(fn, pn, k, ops.row_normalize(f.cuda(), torch.bfloat16), ops.row_normalize(p.cuda(), torch.bfloat16))
            else:
                idx = ops.cos_topk(fn, pn, k)
            idx = idx.cpu()
            same_set = (idx.sort(1).values == order[:, :k].sort(1).values).all(1)
            top1 = idx[:, 0] == order[:, 0]
            res[mode] = {"set_match_clear_rows": float(same_set[clear].float().mean()), "top1": float(top1.float().mean()),
                         "ambiguous_rows": int((~clear).sum())}
            res[mode]["ok"] = res[mode]["set_match_clear_rows"] == 1.0 and res[mode]["top1"] > 0.999
        out[f"b{b}_kp{kp}_c{c}_k{k}"] = res
    out["ok"] = all(m["ok"] for v in out.values() if isinstance(v, dict) for m in v.values())
    return out


def check_models():
    """Golden fixtures through the native modules, fp32 and bf16."""
    import egopack_b200
    from egopack_b200 import Data
    from egopack_b200.models.graph import Graph
    from egopack_b200.models.graphONE.graphONE import GraphONE
    gold = lambda n: torch.load(os.path.join(ROOT, "tests", "golden", n), weights_only=False)
    out = {}
    g = gold("graph_band.pt")
    c = g["cfg"]
    for mode, tol in (("fp32", 1e-4), ("bf16", 3e-2)):
        egopack_b200.set_precision(mode)
        m = Graph(c["input_size"], c["hidden_size"], c["depth"],
                  temporal_pooling={"hidden_size": c["trn_hidden"], "dropout": 0.0}, num_segments=c["num_segments"]).cuda()
        m.load_state_dict(g["state"])
        for struct in ("band", "csr"):
            d = Data(x=g["x"].clone().cuda().requires_grad_(True), pos=g["pos"].cuda(), edge_index=g["edge_index"].cuda())
            d.batch, d.ptr = g["batch"].cuda(), g["ptr"].cuda()
            if struct == "band":
                d.band_k = c["k"]
            m.zero_grad()
            y = m(d)
            (y.float() * g["w"].cuda()).sum().backward()
            e = {"out": _err(y, g["out"])["rel"], "grad_x": _err(d.x.grad, g["grad_x"])["rel"]}
            e["grad_w_max"] = max(_err(p.grad, g["grads"][k])["rel"] for k, p in m.named_parameters())
            e["ok"] = all(v < tol for v in e.values())
            out[f"graph_{mode}_{struct}"] = e
        for i, case in enumerate(gold("graphone.pt")):
            go = GraphONE({t: b.clone() for t, b in case["banks"].items()}, **case["cfg"]).cuda()
            go.load_state_dict(case["state"])
            feats = {t: f.clone().cuda().requires_grad_(True) for t, f in case["feats"].items()}
            o, closest = go.interact(feats)
            sum((o[t].float() * case["w"][t].cuda()).sum() for t in o).backward()
            e = {"out": max(_err(o[t], case["out"][t])["rel"] for t in o),
                 "grad_f": max(_err(feats[t].grad, case["grad_feats"][t])["rel"] for t in o),
                 "grad_w": max(_err(p.grad, case["grads"][k])["rel"] for k, p in go.named_parameters() if k in case["grads"])}
            if mode == "fp32":
                e["closest_equal"] = all(torch.equal(a.cpu(), b) for t in o for a, b in zip(closest[t], case["closest"][t]))
            e["ok"] = all(v < tol for k, v in e.items() if k != "closest_equal") and e.get("closest_equal", True)
            out[f"graphone_{mode}_{i}"] = e
    egopack_b200.set_precision("bf16")
    out["ok"] = all(v["ok"] for v in out.values())
    return out


def check_gemm_tc_tiny():
    """Shapes of the golden-fixture models: every dim smaller than one tile / one swizzle atom."""
    f32 = torch.float32
    return _tc({
        "fwd_k40": dict(m=19, n=32, k=40, bias=True),
        "fwd_k32_f32": dict(m=19, n=40, k=32, out_dtype=f32),
        "dgrad_n72_k40": dict(m=19, n=72, k=40, b_trans=True),
        "dgrad_n40_k32": dict(m=19, n=40, k=32, b_trans=True, out_dtype=f32),
        "dgrad_n32_k40": dict(m=19, n=32, k=40, b_trans=True, out_dtype=f32),
        "wgrad_m40_n72_k19": dict(m=40, n=72, k=19, a_trans=True, b_trans=True, out_dtype=f32),
        "wgrad_m32_n40_k19": dict(m=32, n=40, k=19, a_trans=True, b_trans=True, out_dtype=f32),
        "wgrad_m32_n32_k19": dict(m=32, n=32, k=19, a_trans=True, b_trans=True, out_dtype=f32),
        "wgrad_m64_n64_k19": dict(m=64, n=64, k=19, a_trans=True, b_trans=True, out_dtype=f32),
        "wgrad_m64_n64_k64": dict(m=64, n=64, k=64, a_trans=True, b_trans=True, out_dtype=f32),
        "wgrad_m40_n72_k64": dict(m=40, n=72, k=64, a_trans=True, b_trans=True, out_dtype=f32),
        "dual_dgrad": dict(m=19, n=32, k=32, k2=32, b_trans=True),
        "dual_fwd": dict(m=19, n=32, k=32, k2=32, bias=True),
    })


def check_graph_bf16_locate():
    """Per-parameter gradient error of the golden Graph in bf16, with the tensor-core and the FFMA GEMM."""
    import egopack_b200
    from egopack_b200 import Data
    from egopack_b200.models.graph import Graph
    g = torch.load(os.path.join(ROOT, "tests", "golden", "graph_band.pt"), weights_only=False)
    c = g["cfg"]
    egopack_b200.set_precision("bf16")
    m = Graph(c["input_size"], c["hidden_size"], c["depth"],
              temporal_pooling={"hidden_size": c["trn_hidden"], "dropout": 0.0}, num_segments=c["num_segments"]).cuda()
    m.load_state_dict(g["state"])
    d = Data(x=g["x"].clone().cuda().requires_grad_(True), pos=g["pos"].cuda(), edge_index=g["edge_index"].cuda())
    d.batch, d.ptr, d.band_k = g["batch"].cuda(), g["ptr"].cuda(), c["k"]
    y = m(d)
    (y.float() * g["w"].cuda()).sum().backward()
    out = {"force_simt": os.environ.get("EGP_FORCE_SIMT", "0"), "out": _err(y, g["out"])["rel"],
           "grad_x": _err(d.x.grad, g["grad_x"])["rel"]}
    out["params"] = {k: round(_err(p.grad, g["grads"][k])["rel"], 5) for k, p in m.named_parameters()}
    out["ok"] = True
    return out


def check_graph_bf16_locate_simt():
    os.environ["EGP_FORCE_SIMT"] = "1"
    return check_graph_bf16_locate()


CHECKS = {
    "device": check_device, "gemm_simt": check_gemm_simt, "gemm_tc_kk": check_gemm_tc_kk,
    "gemm_tc_bt": check_gemm_tc_bt, "gemm_tc_at": check_gemm_tc_at, "kernels": check_kernels,
    "edges": check_edges, "topk": check_topk, "models": check_models, "gemm_tc_tiny": check_gemm_tc_tiny,
    "graph_bf16_locate": check_graph_bf16_locate, "graph_bf16_locate_simt": check_graph_bf16_locate_simt,
}


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] in CHECKS:
        name = sys.argv[1]
        try:
            res = CHECKS[name]()
        except Exception as ex:  # noqa: BLE001
            res = {"ok": False, "error": f"{type(ex).__name__}: {ex}", "trace": traceback.format_exc()[-1500:]}
        print("DIAG_RESULT " + json.dumps({"check": name, **res}))
        return
    names = sys.argv[2:] if len(sys.argv) > 2 and sys.argv[1] == "only" else list(CHECKS)
    with open(os.path.join(OUT, "diag.jsonl"), "a") as log:
        for name in names:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                                   timeout=300, env=dict(os.environ, EGP_NO_REBUILD="1"))
                line = next((l for l in r.stdout.splitlines() if l.startswith("DIAG_RESULT ")), None)
                res = json.loads(line[len("DIAG_RESULT "):]) if line else {
                    "check": name, "ok": False, "rc": r.returncode, "stdout": r.stdout[-1500:], "stderr": r.stderr[-2500:]}
            except subprocess.TimeoutExpired:
                res = {"check": name, "ok": False, "error": "timeout"}
            res["seconds"] = round(time.time() - t0, 1)
            log.write(json.dumps(res) + "\n")
            log.flush()
            print(f"[{'OK ' if res.get('ok') else 'BAD'}] {name} ({res['seconds']}s)")
            if not res.get("ok") or os.environ.get("DIAG_VERBOSE"):
                print(json.dumps(res, indent=1)[:6000])


if __name__ == "__main__":
    main()
