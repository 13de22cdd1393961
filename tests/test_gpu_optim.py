"""GPU: egopack_b200.optim.FlatAdam (egp_adam_step: one kernel over flat parameter / moment buffers that also writes the
bf16 weight copies) against torch.optim.Adam -- the optimiser of main_temporal.py:265-271 / main_egopack.py:317-324."""
import copy

import pytest
import torch

import egopack_b200
from egopack_b200 import ops
from egopack_b200.optim import FlatAdam
from tests.gpu_util import DEV, rel_max

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(64, 48), (64,), (115, 64), (115,), (1, 64), (1,), (478, 64), (7, 9, 5), (4097,)]
    return [torch.nn.Parameter(torch.randn(*s, generator=g).to(DEV)) for s in shapes]


@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_flat_adam_matches_torch_adam(wd):
    ours, ref = _params(0), _params(0)
    opt = FlatAdam(ours, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    ropt = torch.optim.Adam(ref, lr=3e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    sched = torch.optim.lr_scheduler.StepLR(opt, 3, 0.5)
    rsched = torch.optim.lr_scheduler.StepLR(ropt, 3, 0.5)
    g = torch.Generator().manual_seed(1)
    for it in range(8):
        for i, (p, q) in enumerate(zip(ours, ref)):
            if it % 3 == 1 and i == 2:                      # a parameter without gradient this step is skipped
                p.grad, q.grad = None, None
                continue
            gr = torch.randn(p.shape, generator=g).to(DEV)
            if i == 3:                                       # a misaligned view of a bucket (data-parallel gradients)
                buf = torch.zeros(gr.numel() + 3, device=DEV)
                buf[3:] = gr.reshape(-1)
                p.grad = buf[3:].view(p.shape)
            else:
                p.grad = gr.clone()
            q.grad = gr.clone()
        opt.step()
        ropt.step()
        sched.step()
        rsched.step()
    for p, q in zip(ours, ref):
        assert rel_max(p, q) < 2e-6, p.shape
    st, rst = opt.state[ours[0]], ropt.state[ref[0]]
    assert rel_max(st["exp_avg"], rst["exp_avg"]) < 3e-6 and rel_max(st["exp_avg_sq"], rst["exp_avg_sq"]) < 3e-6
    assert int(st["step"]) == 8


def test_flat_adam_keeps_parameter_identity_and_bf16_shadows():
    egopack_b200.set_precision("bf16")
    lin = torch.nn.Linear(64, 32).to(DEV)
    w_obj, before = lin.weight, lin.weight.detach().clone()
    opt = FlatAdam(lin.parameters(), lr=1e-2)
    assert lin.weight is w_obj and torch.equal(lin.weight, before)          # same Parameter, same values, new storage
    shadow = ops.weight_cache.get(lin.weight, torch.bfloat16)
    assert torch.equal(shadow, before.bfloat16())
    x = torch.randn(40, 64, device=DEV).bfloat16()
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        ops.linear(x, lin.weight, lin.bias).float().square().mean().backward()
        opt.step()
        got = ops.weight_cache.get(lin.weight, torch.bfloat16)
        assert got.data_ptr() == shadow.data_ptr(), "the GEMM operand is the optimizer's bf16 copy (no cast kernel)"
        assert torch.equal(got, lin.weight.detach().bfloat16())
    with torch.no_grad():
        lin.weight.mul_(2.0)                                                 # touched by something else: shadow is stale
    got = ops.weight_cache.get(lin.weight, torch.bfloat16)
    assert got.data_ptr() != shadow.data_ptr() and torch.equal(got, lin.weight.detach().bfloat16())
    opt.zero_grad(set_to_none=True)
    ops.linear(x, lin.weight, lin.bias).float().square().mean().backward()
    opt.step()                                                                # ... and valid again after the next step
    assert ops.weight_cache.get(lin.weight, torch.bfloat16).data_ptr() == shadow.data_ptr()


def test_flat_adam_state_dict_round_trip_with_torch_adam():
    ours, ref = _params(2), _params(2)
    ropt = torch.optim.Adam(ref, lr=1e-3, weight_decay=1e-3)
    g = torch.Generator().manual_seed(5)
    grads = [[torch.randn(p.shape, generator=g).to(DEV) for p in ref] for _ in range(4)]
    for it in range(2):
        for q, gr in zip(ref, grads[it]):
            q.grad = gr.clone()
        ropt.step()
    opt = FlatAdam(ours, lr=1e-3, weight_decay=1e-3)
    with torch.no_grad():
        for p, q in zip(ours, ref):
            p.copy_(q)
    opt.load_state_dict(copy.deepcopy(ropt.state_dict()))                     # resume from a torch.optim.Adam checkpoint
    for it in range(2, 4):
        for p, q, gr in zip(ours, ref, grads[it]):
            p.grad, q.grad = gr.clone(), gr.clone()
        opt.step()
        ropt.step()
    for p, q in zip(ours, ref):
        assert rel_max(p, q) < 2e-6
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
