"""Batch feed: multi-task loader interleaving and the pinned, double-buffered host -> device hand-over
(utils/dataloading.py:8-70 and the per-sample transforms of main_temporal.py:168-169).

The reference builds ``edge_index`` on the CPU inside DataLoader workers (torch_cluster KD-tree per sample), collates
int64 edges, and ships everything with ``batch.to(device, non_blocking=True)`` on the compute stream.  Here the host
side only moves ``x, pos, y, batch, ptr``; a ``DeviceFeeder`` uploads batch i+1 on a copy stream while step i computes,
and runs the graph transforms (``RadiusGraph`` / ``LTATemporalConnectivity``: count + scan + fill kernels) on the device
right behind the copy.  At 18.4 KB of fp32 features per node the feed is PCIe-bound, so hiding it behind the step is
what matters; the edges never cross the bus.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Mapping, Optional, Sequence, Tuple, Union

import queue
import threading

import torch

from .data import Batch, Data

__all__ = ["multiloader", "DeviceFeeder"]


class multiloader:
    """One tuple of batches per step from several task loaders (utils/dataloading.py:8-47).  A loader that is ``None``
    or has weight 0 contributes ``None``; a loader that runs out is restarted and keeps contributing until EVERY
    active loader has been exhausted once -- the epoch ends on the exhaustion that completes the set."""

    def __init__(self, loaders: Sequence[Optional[Iterable]], weights: Sequence[float]):
        self.loaders, self.weights = loaders, weights
        self.iterators = [iter(ld) if ld is not None and w > 0 else None for ld, w in zip(loaders, weights)]
        self.completed = [it is None for it in self.iterators]

    def __iter__(self) -> "multiloader":
        return self

    def __next__(self) -> Tuple:
        out = []
        for i, it in enumerate(self.iterators):
            if it is None:
                out.append(None)
                continue
            item = next(it, _DONE)
            if item is _DONE:
                self.completed[i] = True
                if all(self.completed):
                    raise StopIteration
                self.iterators[i] = iter(self.loaders[i])      # restart: the others still have batches to give
                item = next(self.iterators[i])
            out.append(item)
        return tuple(out)


_DONE = object()

BatchLike = Union[Data, Sequence[Optional[Data]], Mapping[str, Optional[Data]]]


class DeviceFeeder:
    """Iterates a host loader and yields the same structure with every batch resident on ``device`` and its graph
    structure built there.

    * ``transforms``: one callable for every batch, or a dict / sequence matching the loader's items
      (e.g. ``{"ar": RadiusGraph(1.5), "lta": LTATemporalConnectivity(1.5)}``); applied ON THE DEVICE after the copy.
    * a worker thread copies item i+1 (and builds its edges) on a private stream while the consumer runs step i; the
      consumer's stream waits on the copy's event, and the tensors are ``record_stream``-ed so the caching allocator
      does not recycle them while the step still reads them.  Up to three batches are alive at a time (one being
      consumed, one landed, one in flight).
    * host tensors that are not pinned yet are pinned once per batch (``pin=True``); pass loaders built with
      ``pin_memory=True`` (utils/dataloading.py:64) to skip that.
    """

    def __init__(self, loader: Iterable, device, transforms=None, pin: bool = True):
        self.loader, self.device, self.transforms, self.pin = loader, torch.device(device), transforms, pin
        self._cuda = self.device.type == "cuda"
        if self._cuda and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._stream = torch.cuda.Stream(self.device) if self._cuda else None
        self.h2d_bytes = 0                                      # bytes copied so far (what bench.py reports)

    # -- structure helpers ---------------------------------------------------------------------------------
    def _transform_for(self, key):
        t = self.transforms
        if t is None or callable(t):
            return t
        if isinstance(t, Mapping):
            return t.get(key)
        return t[key]

    def _copy_one(self, b: Optional[Data]) -> Optional[Data]:
        if b is None:
            return None
        if not isinstance(b, Data):                             # a real PyG batch: its own .to keeps every attribute
            return b.to(self.device, non_blocking=True)
        out = Batch()
        for k in b.keys():
            if k.startswith("_"):
                continue
            v = getattr(b, k)
            if torch.is_tensor(v):
                if self._cuda and self.pin and not v.is_pinned() and v.device.type == "cpu":
                    v = v.pin_memory()
                self.h2d_bytes += v.numel() * v.element_size() if v.device != self.device else 0
                v = v.to(self.device, non_blocking=True)
            setattr(out, k, v)
        return out

    def _move(self, item: BatchLike):
        """All copies of the item are enqueued first; only then the transforms run (they read edge counts back, and a
        host wait between two copies would leave the bus idle)."""
        tf = self._transform_for
        if item is None or isinstance(item, Data) or hasattr(item, "edge_index") or hasattr(item, "x"):
            out = self._copy_one(item)
            return tf(0)(out) if out is not None and tf(0) is not None else out
        if isinstance(item, Mapping):
            moved = {k: self._copy_one(v) for k, v in item.items()}
            return {k: (tf(k)(v) if v is not None and tf(k) is not None else v) for k, v in moved.items()}
        moved = [self._copy_one(v) for v in item]
        return tuple(tf(i)(v) if v is not None and tf(i) is not None else v for i, v in enumerate(moved))

    def _upload(self, item):
        if not self._cuda:
            return self._move(item), None
        with torch.cuda.stream(self._stream):
            moved = self._move(item)
            done = torch.cuda.Event()
            done.record(self._stream)
        return moved, done

    @staticmethod
    def _tensors(moved):
        items = moved.values() if isinstance(moved, Mapping) else (moved if isinstance(moved, tuple) else (moved,))
        for b in items:
            if b is None:
                continue
            for k in (b.keys() if hasattr(b, "keys") else ()):
                v = getattr(b, k)
                if torch.is_tensor(v):
                    yield v

    def __iter__(self) -> Iterator:
        # A worker thread drives the loader and the copy stream: the transforms size their outputs on the host (edge
        # counts are read back), and those waits must not hold up the consumer, which is enqueueing the current step.
        # The queue holds one landed item, so at most: one being consumed, one landed, one in flight.
        q: "queue.Queue" = queue.Queue(maxsize=1)
        stop = threading.Event()

        def work():
            try:
                if self._cuda:
                    torch.cuda.set_device(self.device)
                for item in self.loader:
                    if stop.is_set():
                        return
                    q.put(self._upload(item))
                q.put(_DONE)
            except BaseException as ex:  # noqa: BLE001 -- re-raised in the consumer
                q.put(ex)

        worker = threading.Thread(target=work, name="egopack-b200-feed", daemon=True)
        worker.start()
        try:
            while True:
                got = q.get()
                if got is _DONE:
                    return
                if isinstance(got, BaseException):
                    raise got
                moved, done = got
                if done is not None:
                    cur = torch.cuda.current_stream(self.device)
                    cur.wait_event(done)
                    for v in self._tensors(moved):
                        v.record_stream(cur)
                yield moved
        finally:
            stop.set()
            while worker.is_alive():     # unblock a producer waiting on the full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                worker.join(timeout=0.05)
