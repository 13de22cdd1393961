"""GPU, world size 2 over NCCL: one data-parallel step (graphs sharded across two ranks, bucketed gradient all-reduce
launched from autograd hooks) reproduces the gradients a single GPU gets when it runs the two shards one after the other
and averages -- which is what data parallelism over video graphs means here: graph-mode LayerNorm statistics are per
forward call, i.e. per shard.  Skipped when fewer than two GPUs are visible."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(dev):
    import egopack_b200
    from egopack_b200.models.graph import Graph
    from egopack_b200.models.tasks import LTATask, PNRTask, RecognitionTask
    egopack_b200.set_precision("bf16")
    torch.manual_seed(0)                                      # identical replicas on both ranks
    D, S, H = 64, 3, 128
    model = Graph(D, H, 2, temporal_pooling={"hidden_size": H, "dropout": 0.0}, num_segments=S).to(dev)
    tasks = {"ar": RecognitionTask(H, H, (9, 13)).to(dev), "lta": LTATask(H, H, (9, 13)).to(dev), "pnr": PNRTask(H, H).to(dev)}
    params = list(model.parameters()) + [p for t in tasks.values() for p in t.parameters()]
    return model, tasks, params


def _shard_batches(rank, dev):
    from egopack_b200 import synthetic as syn
    from egopack_b200.models.transforms import LTATemporalConnectivity, RadiusGraph
    gen = torch.Generator().manual_seed(100 + rank)           # every rank draws its own shard of graphs
    out = {}
    for t in ("ar", "lta", "pnr"):
        b = syn.make_batch(t, 6, 24, gen, feature_dim=64, num_segments=3, band_k=1, n_verbs=9, n_nouns=13).to(dev)
        out[t] = LTATemporalConnectivity(1.5)(b) if t == "lta" else RadiusGraph(1.5)(b)
    return out


def _local_grads(model, tasks, params, batches):
    from egopack_b200 import steps
    for p in params:
        p.grad = None
    loss, _ = steps.mtl_losses(model, tasks, batches)
    loss.backward()
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in params])


def _worker(rank, world, port, out):
    from egopack_b200 import steps
    from egopack_b200.dp import GradientAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    model, tasks, params = _build(dev)
    mine = _shard_batches(rank, dev)
    local = _local_grads(model, tasks, params, mine)          # this rank's shard, no collective
    # rank 0 also computes the OTHER rank's shard on its own GPU: the single-GPU reference of the same two shards
    # (before the all-reduce hooks exist: a backward on one rank only must not launch collectives)
    other = _local_grads(model, tasks, params, _shard_batches(1, dev)) if rank == 0 else None
    sync = GradientAllReduce(params, bucket_bytes=256 << 10)  # several buckets, launched from the backward hooks
    assert len(sync.buckets) > 1
    for p in params:
        p.grad = None
    loss, _ = steps.mtl_losses(model, tasks, mine)
    loss.backward()
    sync.finish()
    averaged = torch.cat([p.grad.reshape(-1).float() for p in params])
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    want = torch.stack(gathered).mean(0)
    scale = float(want.abs().max())
    err_avg = float((averaged - want).abs().max()) / scale
    sync.remove()
    err_cross = 0.0
    if rank == 0:
        err_cross = float((other - gathered[1]).abs().max()) / float(gathered[1].abs().max())
    out[rank] = (err_avg, err_cross, float(loss))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_step_matches_single_gpu_shard_average():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = dict(out)
    assert set(res) == {0, 1}
    for rank, (err_avg, err_cross, loss) in res.items():
        # NCCL's AVG of two fp32 buffers is exact up to one rounding; split-K wgrad atomics reorder fp32 sums
        assert err_avg < 1e-5, (rank, err_avg)
        assert loss == loss
    assert res[0][1] < 1e-4, res[0][1]
