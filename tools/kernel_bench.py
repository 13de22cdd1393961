"""Per-kernel microbenchmark at the bench workload's shapes (not a pytest file).

    python tools/kernel_bench.py [filter] [--reps 10]

Times each C-ABI op in isolation with CUDA events on the launching stream, flushing L2 between repetitions, and
prints achieved TFLOP/s (GEMMs) or GB/s of ALGORITHMIC bytes (memory-bound kernels) against MEASURED_PEAKS.json.
Also usable under ``ncu -k regex:<kernel>`` to capture one kernel.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from egopack_b200 import ops  # noqa: E402
from egopack_b200.ops import ACT_LEAKY, ACT_RELU  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16
N, H, K0 = int(os.environ.get("EGP_KB_N", 32768)), 1024, 4608   # EGP_KB_N: rows (default = one task batch of c2)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j["hbm_gbs"], j["bf16_tflops"]
    return 6650.0, 1590.0


def timeit(fn, reps):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    filt = args[0] if args else ""
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 10
    hbm, tf = peaks()
    g = torch.Generator(device=DEV).manual_seed(0)
    rnd = lambda *s, dt=BF: torch.randn(*s, device=DEV, generator=g).to(dt)
    cases = []

    def gemm_case(name, m, n, k, a_trans=False, b_trans=False, k2=0, out_dtype=BF, bias=False, act=0, residual=False):
        def build():
            A = rnd(k, m) if a_trans else rnd(m, k)
            B = rnd(k, n) if b_trans else rnd(n, k)
            A2 = (rnd(k2, m) if a_trans else rnd(m, k2)) if k2 else None
            B2 = (rnd(k2, n) if b_trans else rnd(n, k2)) if k2 else None
            bi = rnd(n, dt=torch.float32) if bias else None
            R = rnd(m, n, dt=out_dtype) if residual else None
            out = torch.empty(m, n, dtype=out_dtype, device=DEV)
            return lambda: ops.gemm(A, a_trans, B, b_trans, m, n, k, a2=A2, b2=B2, k2=k2, bias=bi, residual=R, act=act,
                                    slope=0.2, out_dtype=out_dtype, out=out)
        cases.append((name, build, 2.0 * m * n * (k + k2), "TFLOP/s"))

    gemm_case("gemm_fwd_k1024", N, H, H, bias=True)
    gemm_case("gemm_fwd_k1024_relu", N, H, H, bias=True, act=ACT_RELU)
    gemm_case("gemm_fwd_k4608", N, H, K0, bias=True)
    gemm_case("gemm_fwd_dual", N, H, H, k2=H, bias=True)
    gemm_case("gemm_fwd_residual", N, H, H, bias=True, residual=True)
    gemm_case("gemm_dgrad_k1024", N, H, H, b_trans=True)
    gemm_case("gemm_dgrad_dual", N, H, H, b_trans=True, k2=H)
    gemm_case("gemm_wgrad_1024x1024", H, H, N, a_trans=True, b_trans=True, out_dtype=torch.float32)
    gemm_case("gemm_wgrad_1024x4608", H, K0, N, a_trans=True, b_trans=True, out_dtype=torch.float32)
    gemm_case("gemm_head_478_f32out", N, 478, H, bias=True, out_dtype=torch.float32)
    gemm_case("gemm_head_115_f32out", N, 115, H, bias=True, out_dtype=torch.float32)
    gemm_case("gemm_epilogue_only_k64", N, H, 64, bias=True)          # one k-block: launch + epilogue cost of a fwd tile sweep
    gemm_case("gemm_one_tile_256x256x64 (launch + prologue + drain)", 256, 256, 64, bias=True)
    gemm_case("gemm_small_batch_fwd", 2048, H, H, bias=True)
    gemm_case("gemm_small_batch_wgrad", H, H, 2048, a_trans=True, b_trans=True, out_dtype=torch.float32)

    # the library baseline at the same shapes (cuBLAS through torch.mm, bias / activation NOT included)
    def cublas_case(name, m, n, k, a_trans=False, b_trans=False, out_dtype=BF):
        def build():
            A = rnd(k, m).t() if a_trans else rnd(m, k)
            B = rnd(k, n) if b_trans else rnd(n, k).t()
            out = torch.empty(m, n, dtype=BF, device=DEV)
            return lambda: torch.mm(A, B, out=out)
        cases.append((name, build, 2.0 * m * n * k, "TFLOP/s"))

    cublas_case("cublas_fwd_k1024", N, H, H)
    cublas_case("cublas_fwd_k4608", N, H, K0)
    cublas_case("cublas_dgrad_k1024", N, H, H, b_trans=True)
    cublas_case("cublas_wgrad_1024x1024 (bf16 out)", H, H, N, a_trans=True, b_trans=True)
    cublas_case("cublas_wgrad_1024x4608 (bf16 out)", H, K0, N, a_trans=True, b_trans=True)

    def mem_case(name, build, nbytes):
        cases.append((name, build, float(nbytes), "GB/s"))

    def band(k):
        v = N // 128
        batch = torch.arange(v, device=DEV).repeat_interleave(128)
        ptr = torch.arange(v + 1, device=DEV) * 128
        return ops.band_structure(batch, ptr, k)

    for k in (1, 16):
        def build(k=k):
            gs, x = band(k), rnd(N, H)
            return lambda: ops._aggregate(x, gs, False)
        mem_case(f"sage_mean_band_k{k}_bf16", build, 2 * N * H * 2)

    def build_csr():
        gs = band(1)
        idx = torch.arange(N, device=DEV)
        src = torch.cat([idx[:-1], idx[1:]])
        dst = torch.cat([idx[1:], idx[:-1]])
        keep = (src // 128) == (dst // 128)
        gc = ops.csr_structure(torch.stack([src[keep], dst[keep]]), N)
        x = rnd(N, H)
        return lambda: ops._aggregate(x, gc, False)
    mem_case("sage_mean_csr_k1_bf16", build_csr, 2 * N * H * 2)

    def build_rln_fwd():
        x, w, b = rnd(N, H), rnd(H, dt=torch.float32), rnd(H, dt=torch.float32)
        return lambda: ops.RowLayerNorm.apply(x, w, b, 1e-5, ACT_RELU, 0.0)
    mem_case("row_ln_relu_fwd_bf16", build_rln_fwd, 2 * N * H * 2)

    def build_rln_bwd():
        x = rnd(N, H).requires_grad_(True)
        w, b = rnd(H, dt=torch.float32).requires_grad_(True), rnd(H, dt=torch.float32).requires_grad_(True)
        y = ops.RowLayerNorm.apply(x, w, b, 1e-5, ACT_RELU, 0.0)
        dy = rnd(N, H)
        return lambda: torch.autograd.grad(y, (x, w, b), dy, retain_graph=True)
    mem_case("row_ln_relu_bwd_bf16", build_rln_bwd, 4 * N * H * 2)

    def build_gln_fwd():
        x, w, b = rnd(N, H), rnd(H, dt=torch.float32), rnd(H, dt=torch.float32)
        return lambda: ops.GraphLayerNorm.apply(x, w, b, 1e-5, ACT_LEAKY, 0.2)
    mem_case("graph_ln_leaky_fwd_bf16", build_gln_fwd, 3 * N * H * 2)

    def build_gln_bwd():
        x = rnd(N, H).requires_grad_(True)
        w, b = rnd(H, dt=torch.float32).requires_grad_(True), rnd(H, dt=torch.float32).requires_grad_(True)
        y = ops.GraphLayerNorm.apply(x, w, b, 1e-5, ACT_LEAKY, 0.2)
        dy = rnd(N, H)
        return lambda: torch.autograd.grad(y, (x, w, b), dy, retain_graph=True)
    mem_case("graph_ln_leaky_bwd_bf16", build_gln_bwd, 5 * N * H * 2)

    def build_colsum():
        x = rnd(N, H)
        return lambda: ops.colsum(x)
    mem_case("colsum_bf16", build_colsum, N * H * 2)

    def build_cast():
        x = rnd(N, K0, dt=torch.float32)
        return lambda: ops.cast(x, BF)
    mem_case("cast_f32_bf16_4608", build_cast, N * K0 * 6)

    def build_tiny():
        x = rnd(8, 8, dt=torch.float32)
        return lambda: ops.cast(x, BF)
    mem_case("cast_64_elements (launch floor of this harness)", build_tiny, 64 * 6)

    def build_actbwd():
        dy, y = rnd(N, H), rnd(N, H)
        return lambda: ops.act_bwd(dy, y, ACT_RELU, 0.0)
    mem_case("relu_bwd_bf16", build_actbwd, 3 * N * H * 2)

    def build_drop():
        x = rnd(N, H)
        mask = torch.ones(N, H, dtype=torch.uint8, device=DEV)
        out = torch.empty_like(x)
        from egopack_b200 import _lib as L
        return lambda: L.call("egp_mask_scale", x.data_ptr(), mask.data_ptr(), out.data_ptr(), x.numel(), 2.0, 1, L.stream())
    mem_case("dropout_apply_bf16", build_drop, N * H * 5)

    def build_posenc():
        x = rnd(N, H)
        pos = torch.arange(128, device=DEV).repeat(N // 128)
        freq = torch.logspace(0, 1, H // 2, 1e-4).to(DEV)
        return lambda: ops.PosEncAdd.apply(x, pos, freq)
    mem_case("posenc_add_bf16", build_posenc, 2 * N * H * 2)

    def build_topk(guard=True):
        f, p = rnd(N, H, dt=torch.float32), rnd(4096, H, dt=torch.float32)
        fn, pn = ops.row_normalize(f), ops.row_normalize(p)
        f16, fe = ops.row_normalize(f, BF, with_round_err=True)
        p16, pe = ops.row_normalize(p, BF, with_round_err=True)
        pmax = float(pe.max())
        return lambda: ops.cos_topk(fn, pn, 4, f16, p16, f_err=fe, p_err=pmax, guard=guard)
    for guard in (True, False):
        cases.append((f"cos_topk_k4_kp4096 guard={guard} (GEMM flops)", (lambda g_=guard: build_topk(g_)), 2.0 * N * 4096 * H, "TFLOP/s"))

    # ---- round 2 additions: LTA band+star, pooling / GraphONE pieces, losses, fused dropout, fp32 GEMM on tensor cores
    def lta_structure(n_nodes):
        v = n_nodes // 128
        batch = torch.arange(v, device=DEV).repeat_interleave(128)
        ptr = torch.arange(v + 1, device=DEV) * 128
        y = torch.randint(1, 100, (n_nodes, 2), device=DEV)
        y.view(v, 128, 2)[:, :2] = -1                                    # 2 input clips, 126 forecast clips per graph
        star = ops.lta_star_counts(y, ptr, 1.5)
        return ops.band_structure(batch, ptr, 1, star)

    for nn in (N, 3 * N):
        def build_star(backward, nn=nn):
            gs, x = lta_structure(nn), rnd(nn, H)
            return lambda: ops._aggregate(x, gs, backward)
        mem_case(f"sage_mean_band_star_k1_fwd_bf16_n{nn}", lambda nn=nn: build_star(False, nn), 2 * nn * H * 2)
        mem_case(f"sage_mean_band_star_k1_bwd_bf16_n{nn}", lambda nn=nn: build_star(True, nn), 2 * nn * H * 2)

    def build_band_bwd():
        gs, x = band(1), rnd(N, H)
        return lambda: ops._aggregate(x, gs, True)
    mem_case("sage_mean_band_k1_bwd_bf16", build_band_bwd, 2 * N * H * 2)

    def build_rln_drop():
        x, w, b = rnd(N, H), rnd(H, dt=torch.float32), rnd(H, dt=torch.float32)
        return lambda: ops.RowLayerNorm.apply(x, w, b, 1e-5, ACT_RELU, 0.5)
    mem_case("row_ln_relu_dropout_fwd_bf16", build_rln_drop, 2 * N * H * 2)

    def build_pool(backward):
        v = N // 128
        x = rnd(N, H).requires_grad_(True)
        batch = torch.arange(v, device=DEV).repeat_interleave(128)
        ptr = torch.arange(v + 1, device=DEV) * 128
        if not backward:
            return lambda: ops.SegmentMaxPool.apply(x.detach(), ptr, batch)
        y = ops.SegmentMaxPool.apply(x, ptr, batch)
        dy = rnd(v, H)
        return lambda: torch.autograd.grad(y, x, dy, retain_graph=True)
    mem_case("segment_max_pool_fwd_bf16 (read x)", lambda: build_pool(False), N * H * 2)
    mem_case("segment_max_pool_bwd_bf16 (write dx)", lambda: build_pool(True), N * H * 2)

    def build_proto_gather():
        bank = rnd(4096, H)
        idx = torch.randint(0, 4096, (N, 4), device=DEV)
        return lambda: ops.proto_max_gather(bank, idx)
    mem_case("proto_max_gather_k4_bf16 (write m + idx; bank in L2)", build_proto_gather, N * H * 2 + N * 4 * 8)

    def build_maxc(backward):
        f, m = rnd(N, H).requires_grad_(True), rnd(N, H)
        if not backward:
            return lambda: ops.MaxCombine.apply(f.detach(), m)
        a = ops.MaxCombine.apply(f, m)
        da = rnd(N, H)
        return lambda: torch.autograd.grad(a, f, da, retain_graph=True)
    mem_case("max_combine_fwd_bf16", lambda: build_maxc(False), 3 * N * H * 2)
    mem_case("max_combine_bwd_bf16", lambda: build_maxc(True), 4 * N * H * 2)

    def build_class_sum():
        x = rnd(N, H)
        labels = torch.randint(0, 115 * 478, (N,), device=DEV)
        out = torch.zeros(115 * 478, H, dtype=torch.float64, device=DEV)
        return lambda: ops.class_sum_f64(x, labels, 115 * 478, out)
    mem_case("class_sum_f64_bf16 (read x + fp64 atomics on 8 B/elt)", build_class_sum, N * H * 2 + N * H * 8)

    def build_ce(backward):
        lv, ln = rnd(N, 115, dt=torch.float32).requires_grad_(True), rnd(N, 478, dt=torch.float32).requires_grad_(True)
        y = torch.stack([torch.randint(0, 115, (N,), device=DEV), torch.randint(0, 478, (N,), device=DEV)], 1)
        if not backward:
            return lambda: ops.cross_entropy((lv.detach(), ln.detach()), y, ignore_index=-1)
        loss = ops.cross_entropy((lv, ln), y, ignore_index=-1)
        g_ = torch.ones((), device=DEV).expand(N)
        return lambda: torch.autograd.grad(loss, (lv, ln), g_, retain_graph=True)
    mem_case("ce_2heads_fwd (read logits)", lambda: build_ce(False), N * (115 + 478) * 4)
    mem_case("ce_2heads_bwd (read logits, write dlogits)", lambda: build_ce(True), 2 * N * (115 + 478) * 4)

    from egopack_b200 import config as _cfg

    def fp32_case(kind, name, m, n, k, **kw):
        def build():
            A = rnd(k, m, dt=torch.float32) if kw.get("a_trans") else rnd(m, k, dt=torch.float32)
            B = rnd(k, n, dt=torch.float32) if kw.get("b_trans") else rnd(n, k, dt=torch.float32)
            bi = rnd(n, dt=torch.float32)
            out = torch.empty(m, n, dtype=torch.float32, device=DEV)

            def run():
                old = _cfg.get_fp32_gemm()
                _cfg.set_fp32_gemm(kind)
                try:
                    ops.gemm(A, bool(kw.get("a_trans")), B, bool(kw.get("b_trans")), m, n, k, bias=bi, out=out)
                finally:
                    _cfg.set_fp32_gemm(old)
            return run
        cases.append((name, build, 2.0 * m * n * k, "TFLOP/s"))

    for kind in ("bf16x6", "bf16x3", "ffma"):
        fp32_case(kind, f"gemm_fp32_{kind}_fwd_k1024 (fp32-equivalent flops, split kernels included)", N, H, H)
    fp32_case("bf16x6", "gemm_fp32_bf16x6_fwd_k4608 (fp32-equivalent flops, split kernels included)", N, H, K0)
    fp32_case("bf16x6", "gemm_fp32_bf16x6_wgrad_1024x1024 (fp32-equivalent flops, split kernels included)", H, H, N,
              a_trans=True, b_trans=True)

    print(f"{'kernel':40s} {'ms':>9s} {'achieved':>12s} {'of peak':>8s}")
    out = {}
    for name, build, work, unit in cases:
        if filt and filt not in name:
            continue
        fn = build()
        ms = timeit(fn, reps)
        ach = work / (ms / 1e3) / (1e12 if unit == "TFLOP/s" else 1e9)
        frac = ach / (tf if unit == "TFLOP/s" else hbm)
        out[name] = {"ms": round(ms, 4), "achieved": round(ach, 1), "unit": unit, "frac_of_burst_peak": round(frac, 3)}
        print(f"{name:40s} {ms:9.4f} {ach:9.1f} {unit[:5]:>5s} {100 * frac:7.1f}%")
        del fn
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
