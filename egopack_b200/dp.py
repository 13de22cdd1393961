"""Data parallelism over video graphs: one process per GPU, every rank holds a full replica and
its own shard of graphs; the only exchange is one gradient all-reduce (mean) per step over NCCL / NVLink.

Graphs never span ranks (no edge crosses graphs), prototype banks are replicated, and graph-mode LayerNorm
statistics stay per-rank (they are per forward call in the reference too), so no other collective exists.
Gradients are packed into a few large buckets in reverse parameter order and each bucket's all-reduce is
launched from an autograd hook as soon as its last gradient is produced, so communication overlaps the rest of
the backward pass.  Works with ``gloo`` on CPU tensors for the host-logic tests.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def shard_graphs(num_graphs: int, rank: int, world_size: int) -> range:
    """Contiguous, near-even split of graph ids; every graph lands on exactly one rank."""
    base, extra = divmod(num_graphs, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


class _Bucket:
    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = params
        self.pending = len(params)
        self.flat: Optional[torch.Tensor] = None
        self.work = None


class GradientAllReduce:
    """Bucketed, backward-overlapped gradient averaging.

        sync = GradientAllReduce(params)        # once
        loss.backward(); sync.finish()          # every step (finish() waits and writes the means back)
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None,
                 overlap: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[_Bucket] = []
        cur, size = [], 0
        for p in reversed(self.params):                       # gradients arrive roughly in reverse order
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(_Bucket(cur))
                cur, size = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        self._owner = {}
        self._handles = []
        self._avg = False
        if overlap and self.world > 1:
            for b in self.buckets:
                for p in b.params:
                    self._owner[p] = b
                    self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _launch(self, b: _Bucket):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b.params]
        b.flat = torch.cat([g.reshape(-1) for g in grads])
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG, so it sums and finish() divides
        self._avg = b.flat.is_cuda
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def _on_grad(self, p):
        b = self._owner[p]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    def finish(self):
        """Wait for every bucket; afterwards every ``p.grad`` IS its slice of the averaged bucket (a view: no copy-back
        kernels; ``zero_grad(set_to_none=True)`` drops the views before the next backward)."""
        if self.world == 1:
            return
        for b in self.buckets:
            if b.work is None:                                # params without a hook firing (unused / no overlap)
                self._launch(b)
        for b in self.buckets:
            b.work.wait()
            if not self._avg:
                b.flat.div_(self.world)
            off = 0
            for p in b.params:
                n = p.numel()
                p.grad = b.flat[off:off + n].view_as(p)
                off += n
            b.pending, b.flat, b.work = len(b.params), None, None

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
