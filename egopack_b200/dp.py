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
        self.ready = False


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
        self._next = 0                                        # buckets launch in index order on EVERY rank
        self._enabled = True
        self._seen = set()                                    # parameters whose gradient arrived in this step
        if overlap and self.world > 1:
            for b in self.buckets:
                for p in b.params:
                    self._owner[p] = b
                    self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _launch(self, b: _Bucket):
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in b.params]
        pieces = []
        for g in grads:                                       # every gradient starts on a 16-byte boundary of the bucket,
            pieces.append(g.reshape(-1))                      # so the optimizer kernel can read the views vectorised
            if g.numel() % 4:
                pieces.append(g.new_zeros(4 - g.numel() % 4))
        b.flat = torch.cat(pieces)
        # NCCL averages inside the collective; gloo (CPU tests) has no AVG, so it sums and finish() divides
        self._avg = b.flat.is_cuda
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.group, async_op=True)

    def _on_grad(self, p):
        if not self._enabled:                                 # inside no_sync(): gradients only accumulate locally
            return
        b = self._owner[p]
        if b.work is not None or p in self._seen:
            raise RuntimeError("GradientAllReduce: a gradient arrived after its bucket was reduced -- more than one "
                               "backward per finish(); wrap the earlier micro-batches in `with sync.no_sync():`")
        self._seen.add(p)
        b.pending -= 1
        if b.pending == 0:
            b.ready = True
            # a bucket is launched only after every earlier bucket: the collective order is then the bucket order on
            # all ranks, whatever order the hooks fire in (and whichever parameters a rank leaves unused)
            while self._next < len(self.buckets) and self.buckets[self._next].ready:
                self._launch(self.buckets[self._next])
                self._next += 1

    def no_sync(self):
        """Context manager for gradient accumulation: backward passes inside it do not start the all-reduce; the first
        backward outside it reduces the accumulated gradients."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old, self._enabled = self._enabled, False
            try:
                yield
            finally:
                self._enabled = old
        return ctx()

    def finish(self):
        """Wait for every bucket; afterwards every ``p.grad`` IS its slice of the averaged bucket (a view: no copy-back
        kernels; ``zero_grad(set_to_none=True)`` drops the views before the next backward)."""
        if self.world == 1:
            return
        for b in self.buckets[self._next:]:                   # not launched from a hook (unused params / no overlap):
            self._launch(b)                                   # still in bucket order
        for b in self.buckets:
            b.work.wait()
            if not self._avg:
                b.flat.div_(self.world)
            off = 0
            for p in b.params:
                n = p.numel()
                p.grad = b.flat[off:off + n].view_as(p)
                off += (n + 3) // 4 * 4
            b.pending, b.flat, b.work, b.ready = len(b.params), None, None, False
        self._next = 0
        self._seen.clear()

    def remove(self):
        for h in self._handles:
            h.remove()
        self._handles = []
