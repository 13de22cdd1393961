"""Batch feed: multi-task loader interleaving and the pinned, pipelined host -> device hand-over
(utils/dataloading.py:8-70 and the per-sample transforms of main_temporal.py:168-169).

The reference builds ``edge_index`` on the CPU inside DataLoader workers (torch_cluster KD-tree per sample), collates
int64 edges, and ships everything with ``batch.to(device, non_blocking=True)`` on the compute stream.  Here the host
side only moves ``x, pos, y, batch, ptr``; a ``DeviceFeeder`` enqueues the copies of batch i+1 on a copy stream BEFORE
it hands batch i to the training loop, and runs the graph transforms (``RadiusGraph`` / ``LTATemporalConnectivity``) on
the device right behind them -- as structural hints (``band_k``, ``star``) with a lazy ``edge_index``, so nothing in the
hand-over waits on the GPU.  The feed is PCIe-bound: 18.4 KB of fp32 features per node, 9.2 KB when the loader stores
them as bf16 (``Batch.to_feature_dtype`` -- bit-identical to the cast the first GEMM's operand goes through anyway);
``bind_host_memory_to_gpu`` keeps the pinned buffers on the GPU's own NUMA node, which is what an 8-GPU box needs to
run eight such feeds at once.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Mapping, Optional, Sequence, Tuple, Union

import os
import queue
import threading

import torch

from .data import Batch, Data, _map_features, expand_base, replicated_base

__all__ = ["multiloader", "DeviceFeeder", "bind_host_memory_to_gpu", "unit_spaced_host"]


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


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_memory_to_gpu(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers allocated
    afterwards (first touch) live in that node's memory and the H2D copies do not cross the socket interconnect.

    On a 2-socket 8-GPU box every rank otherwise allocates wherever the launcher happened to run (round 1: all eight
    feeds came out of NUMA node 0 and shared 184 GB/s).  Returns what was done -- ``{"numa_node": n, "cpus": k}`` or
    ``{"skipped": reason}``; never raises (containers often hide /sys or forbid sched_setaffinity)."""
    try:
        import torch.cuda as tc
        props = tc.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    except Exception as ex:  # noqa: BLE001
        return {"skipped": f"no PCI id: {ex}"}
    try:
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return {"skipped": f"{bus}: numa_node=-1 (single node or hidden)", "pci": bus}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cpus) & allowed) or None
        if not use:
            return {"skipped": f"node {node} has no CPU this process may use", "pci": bus, "numa_node": node}
        os.sched_setaffinity(0, use)
        return {"pci": bus, "numa_node": node, "cpus": len(use)}
    except Exception as ex:  # noqa: BLE001
        return {"skipped": f"{type(ex).__name__}: {ex}", "pci": bus}


def unit_spaced_host(pos: torch.Tensor, batch: Optional[torch.Tensor]) -> bool:
    """Host-side version of the transforms' band test (pos increases by exactly 1 inside every graph), evaluated on the
    CPU copy BEFORE the upload so that the device-side transforms need no read-back."""
    pos = pos.view(-1)
    if pos.numel() < 2:
        return True
    d = pos[1:] - pos[:-1]
    ok = d == 1
    if batch is not None:
        ok = ok | (batch[1:] != batch[:-1])
    return bool(ok.all())


_COPY_STREAMS: dict = {}


def _copy_stream(device: torch.device) -> "torch.cuda.Stream":
    """ONE copy stream per device for every feeder of the process.  torch's caching allocator keeps its free blocks per
    stream: a feeder with a stream of its own (say, one feeder per epoch) finds none of the staging buffers its predecessor
    returned and sends every allocation of its first steps to cudaMalloc (8 calls, and a first step of 50-330 ms instead of
    26, in bench.py's end-to-end loop) while the old stream's blocks stay cached for nobody."""
    key = (device.type, device.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device)
    return _COPY_STREAMS[key]


class DeviceFeeder:
    """Iterates a host loader and yields the same structure with every batch resident on ``device`` and its graph
    structure described there.

    * ``transforms``: one callable for every batch, or a dict / sequence matching the loader's items
      (e.g. ``{"ar": RadiusGraph(1.5), "lta": LTATemporalConnectivity(1.5)}``); applied ON THE DEVICE after the copy.
      They are sync-free: the band test is answered on the host copy (``pos_unit_spaced``), the LTA star is counted by a
      kernel, and ``edge_index`` stays lazy.
    * pipelining: the copies (and transform kernels) of item i+1 are enqueued on a private copy stream before item i is
      yielded, so they overlap step i whatever the consumer does with it (including ``loss.item()``); the consumer's
      stream waits on the copy's event and the tensors are ``record_stream``-ed so the caching allocator does not
      recycle them while the step still reads them.  All CUDA calls are made by the consuming thread; with
      ``prefetch_thread=True`` a helper thread only drives the HOST loader (collation, pinning) one item ahead.
    * host tensors that are not pinned yet are pinned once per batch (``pin=True``); pass loaders built with
      ``pin_memory=True`` (utils/dataloading.py:64) to skip that.
    * ``feature_dtype``: convert ``x`` on the host before the copy (a loader that stores bf16 features should do this
      once per sample instead: ``Batch.to_feature_dtype``).
    * replicated segments: an ``x`` whose segments are one memory (``data.replicated_base``: what a PNR loader that writes
      ``expand`` for the reference's ``repeat`` hands over, data/ego4d_oscc.py:291) crosses PCIe as its ``[N, D]`` base and is
      repeated on the device; collation, pinning and ``feature_dtype`` keep that form.
    * ``fuse_features``: the ``x`` tensors of the task batches of one step land in ONE device allocation, back to back
      (each batch still sees its own rows), so ``Graph.forward_many`` can feed them to the first Linear as a single
      GEMM operand.
    """

    def __init__(self, loader: Iterable, device, transforms=None, pin: bool = True, prefetch_thread: bool = False,
                 feature_dtype: Optional[torch.dtype] = None, fuse_features: bool = True):
        self.loader, self.device, self.transforms, self.pin = loader, torch.device(device), transforms, pin
        self.prefetch_thread, self.feature_dtype, self.fuse_features = prefetch_thread, feature_dtype, fuse_features
        self._cuda = self.device.type == "cuda"
        if self._cuda and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._stream = _copy_stream(self.device) if self._cuda else None
        self.h2d_bytes = 0                                      # bytes copied so far (what bench.py reports)

    # -- structure helpers ---------------------------------------------------------------------------------
    def _transform_for(self, key):
        t = self.transforms
        if t is None or callable(t):
            return t
        if isinstance(t, Mapping):
            return t.get(key)
        return t[key]

    def _prepare_host(self, b):
        """Host-side work of one batch (may run in the helper thread): feature dtype, pinning, the band hint."""
        if b is None or not isinstance(b, Data):
            return b
        if self.feature_dtype is not None and b.x is not None and b.x.dtype != self.feature_dtype and b.x.device.type == "cpu":
            b.x = _map_features(b.x, lambda t: t.to(self.feature_dtype))
        if b.pos is not None and b.pos.device.type == "cpu" and getattr(b, "pos_unit_spaced", None) is None \
                and not b.pos.is_floating_point():
            b.pos_unit_spaced = unit_spaced_host(b.pos, b.batch)
        if self._cuda and self.pin:
            for k in list(b._fields):
                v = b._fields[k]
                if torch.is_tensor(v) and v.device.type == "cpu" and not v.is_pinned():
                    b._fields[k] = _map_features(v, lambda t: t.pin_memory()) if k == "x" else v.pin_memory()
        return b

    def _prepare_item(self, item):
        if item is None or isinstance(item, Data) or hasattr(item, "edge_index") or hasattr(item, "x"):
            return self._prepare_host(item)
        if isinstance(item, Mapping):
            return {k: self._prepare_host(v) for k, v in item.items()}
        return tuple(self._prepare_host(v) for v in item)

    def _copy_one(self, b: Optional[Data], x_slot: Optional[torch.Tensor] = None) -> Optional[Data]:
        if b is None:
            return None
        if not isinstance(b, Data):                             # a real PyG batch: its own .to keeps every attribute
            return b.to(self.device, non_blocking=True)
        out = Batch()
        for k, v in b._fields.items():                          # lazy attributes of the host batch are not forced
            if k.startswith("_"):
                continue
            if torch.is_tensor(v):
                base = replicated_base(v) if k == "x" and v.device.type == "cpu" else None
                if base is not None:
                    # replicated segments (the PNR loader, data/ego4d_oscc.py:291): only the base crosses the bus, the
                    # repeat is a device-side copy on this stream -- the consumer sees the same [N, R, D] tensor
                    self.h2d_bytes += base.numel() * base.element_size()
                    dev_base = base.to(self.device, non_blocking=True)
                    if x_slot is None:
                        x_slot = torch.empty(v.shape, dtype=v.dtype, device=self.device)
                    x_slot.copy_(expand_base(dev_base, v.shape[1]))
                    v = x_slot
                else:
                    self.h2d_bytes += v.numel() * v.element_size() if v.device != self.device else 0
                    if k == "x" and x_slot is not None:
                        x_slot.copy_(v, non_blocking=True)
                        v = x_slot
                    else:
                        v = v.to(self.device, non_blocking=True)
            setattr(out, k, v)
        return out

    def _feature_slots(self, batches):
        """One device allocation for the ``x`` of all batches of an item (rows back to back), or None per batch when
        they cannot share one (different trailing shape / dtype, foreign batch types, already on the device)."""
        xs = [getattr(b, "x", None) if isinstance(b, Data) else None for b in batches]
        live = [x for x in xs if x is not None]
        if not self.fuse_features or not self._cuda or len(live) < 2 or any(x.device.type != "cpu" for x in live) \
                or len({(tuple(x.shape[1:]), x.dtype) for x in live}) != 1:
            return [None] * len(batches)
        buf = torch.empty((sum(x.shape[0] for x in live), *live[0].shape[1:]), dtype=live[0].dtype, device=self.device)
        slots, off = [], 0
        for x in xs:
            if x is None:
                slots.append(None)
            else:
                slots.append(buf[off:off + x.shape[0]])
                off += x.shape[0]
        return slots

    def _move(self, item: BatchLike):
        """All copies of the item are enqueued first, then the transforms' kernels."""
        tf = self._transform_for
        if item is None or isinstance(item, Data) or hasattr(item, "edge_index") or hasattr(item, "x"):
            out = self._copy_one(item)
            return tf(0)(out) if out is not None and tf(0) is not None else out
        if isinstance(item, Mapping):
            slots = self._feature_slots(list(item.values()))
            moved = {k: self._copy_one(v, sl) for (k, v), sl in zip(item.items(), slots)}
            return {k: (tf(k)(v) if v is not None and tf(k) is not None else v) for k, v in moved.items()}
        slots = self._feature_slots(list(item))
        moved = [self._copy_one(v, sl) for v, sl in zip(item, slots)]
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
            fields = getattr(b, "_fields", None)
            for v in (fields.values() if fields is not None else ()):
                if torch.is_tensor(v):
                    yield v

    def _host_items(self) -> Iterator:
        """The host loader, prepared; optionally one item ahead in a helper thread (host work only -- no CUDA calls)."""
        if not self.prefetch_thread:
            for item in self.loader:
                yield self._prepare_item(item)
            return
        q: "queue.Queue" = queue.Queue(maxsize=2)
        stop = threading.Event()

        def work():
            try:
                for item in self.loader:
                    if stop.is_set():
                        return
                    q.put(self._prepare_item(item))
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
                yield got
        finally:
            stop.set()
            while worker.is_alive():     # unblock a producer waiting on the full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                worker.join(timeout=0.05)

    def __iter__(self) -> Iterator:
        host = self._host_items()
        try:
            pending = None
            first = next(host, _DONE)
            if first is not _DONE:
                pending = self._upload(first)
            while pending is not None:
                moved, done = pending
                nxt = next(host, _DONE)
                pending = self._upload(nxt) if nxt is not _DONE else None      # item i+1 is on the bus before i is used
                if done is not None:
                    cur = torch.cuda.current_stream(self.device)
                    cur.wait_event(done)
                    for v in self._tensors(moved):
                        v.record_stream(cur)
                yield moved
        finally:
            close = getattr(host, "close", None)
            if close is not None:
                close()
