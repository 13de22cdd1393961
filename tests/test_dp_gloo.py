"""CPU, world_size 2 over gloo: the data-parallel host logic (graph sharding + bucketed, hook-driven gradient
all-reduce).  The GPU path swaps gloo for NCCL; the collective code is the same."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egopack_b200.dp import GradientAllReduce, shard_graphs


def test_shard_graphs_partitions_every_graph_once():
    for n in (1, 7, 16, 255, 256):
        for w in (1, 2, 4, 8):
            got = [g for r in range(w) for g in shard_graphs(n, r, w)]
            assert got == list(range(n))
            sizes = [len(shard_graphs(n, r, w)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, overlap, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    for p in model[3].parameters():                      # an unused sub-module: its grads never fire a hook
        p.requires_grad_(True)
    sync = GradientAllReduce(model.parameters(), bucket_bytes=256, overlap=overlap)
    assert len(sync.buckets) > 1
    for step in range(2):                                # two steps: buckets must re-arm
        model.zero_grad(set_to_none=True)
        x = torch.full((3, 8), float(rank + 1 + step))
        model[2](model[1](model[0](x))).sum().backward()
        sync.finish()
    grads = [p.grad.clone() for p in model.parameters()]
    # reference: average of both ranks' local gradients
    ref_model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 4))
    ref_model.load_state_dict(model.state_dict())
    acc = None
    for r in range(world):
        ref_model.zero_grad(set_to_none=True)
        x = torch.full((3, 8), float(r + 1 + 1))
        ref_model[2](ref_model[1](ref_model[0](x))).sum().backward()
        g = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in ref_model.parameters()]
        acc = g if acc is None else [a + b for a, b in zip(acc, g)]
    ok = all(torch.allclose(a, b / world, atol=1e-6) for a, b in zip(grads, acc))
    out[rank] = ok
    dist.destroy_process_group()


def _run(overlap):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), overlap, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_gradient_allreduce_overlapped():
    _run(True)


def test_gradient_allreduce_after_backward():
    _run(False)


def _worker_accum(rank, world, port, out):
    """Gradient accumulation (two micro-batches, the first under no_sync) with a parameter that only rank 0 uses:
    the collectives must still be issued in bucket order on both ranks and reduce the ACCUMULATED gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(16, 4))
    sync = GradientAllReduce(model.parameters(), bucket_bytes=128, overlap=True)

    def fwd(x):
        h = model[0](x)
        y = model[1](h).sum()
        if rank == 0:                                    # model[2] is unused on rank 1: its hooks never fire there
            y = y + model[2](h).sum()
        return y

    model.zero_grad(set_to_none=True)
    with sync.no_sync():
        fwd(torch.full((2, 8), 1.0 + rank)).backward()
    fwd(torch.full((2, 8), 3.0 + rank)).backward()
    sync.finish()
    got = [p.grad.clone() for p in model.parameters()]
    # a second backward without no_sync after the reduce of this step must be refused, not silently lost
    raised = False
    model.zero_grad(set_to_none=True)
    fwd(torch.ones(2, 8)).backward()
    try:
        fwd(torch.ones(2, 8)).backward()
    except RuntimeError:
        raised = True
    sync.finish()

    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Linear(16, 4), torch.nn.Linear(16, 4))
    ref.load_state_dict(model.state_dict())
    acc = None
    for r in range(world):
        ref.zero_grad(set_to_none=True)
        for base in (1.0, 3.0):
            h = ref[0](torch.full((2, 8), base + r))
            y = ref[1](h).sum()
            if r == 0:
                y = y + ref[2](h).sum()
            y.backward()
        g = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in ref.parameters()]
        acc = g if acc is None else [a + b for a, b in zip(acc, g)]
    out[rank] = all(torch.allclose(a, b / world, atol=1e-5) for a, b in zip(got, acc)) and raised
    dist.destroy_process_group()


def test_gradient_accumulation_and_bucket_order():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_accum, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
