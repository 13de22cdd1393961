"""CUDA-graph capture of a whole training step for the launch-bound regime (the reference's real batches are 16
graphs of 4-22 nodes: ~300 kernel launches of microsecond-scale work per step).

    runner = GraphedStep(step_fn, static_inputs)     # warm-up on a side stream, then capture
    loss = runner(new_inputs)                        # copies the new tensors into the static buffers and replays

``step_fn(inputs) -> loss`` must be sync-free (no ``.item()``), allocate nothing outside torch's caching allocator and
use an optimizer built with ``capturable=True``, and the batch SHAPES and graph STRUCTURE (``edge_index``, graph sizes)
must not change between replays -- features, labels and positions may.  (The LTA connectivity depends on the labels
through ``y[:,0] > 0``; a batch whose edge set differs needs a new capture.)  Everything in this package satisfies
the kernel-side requirements: kernels are launched on
the current stream with caller-owned buffers, TMA descriptors are cached by address, workspaces are grown during the
warm-up, and the fused-dropout Philox stream reads its per-step counter from device memory (``ops.RNG_STATE``).
"""
from __future__ import annotations

from typing import Callable, Dict

import torch

from . import ops

_TENSOR_FIELDS = ("x", "pos", "y", "batch", "ptr", "edge_index")


def _materialized(b, key: str) -> bool:
    probe = getattr(b, "is_materialized", None)
    return probe(key) if callable(probe) else getattr(b, key, None) is not None


class GraphedStep:
    def __init__(self, step_fn: Callable[[Dict[str, object]], torch.Tensor], static_inputs: Dict[str, object],
                 warmup: int = 3):
        self.static = static_inputs
        dev = next(getattr(b, "x") for b in static_inputs.values()).device
        self.rng = torch.tensor([int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF, 0], dtype=torch.int64, device=dev)
        ops.RNG_STATE = self.rng
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn(self.static)
                self.rng[1] += 1
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = step_fn(self.static)
            self.rng[1] += 1                                  # captured: every replay advances the dropout stream
        ops.RNG_STATE = None

    def __call__(self, inputs: Dict[str, object] = None) -> torch.Tensor:
        if inputs is not None and inputs is not self.static:
            for t, b in inputs.items():
                dst = self.static[t]
                if getattr(b, "band_k", None) != getattr(dst, "band_k", None):
                    raise ValueError(f"GraphedStep: 'band_k' of task '{t}' changed; capture a new graph")
                for k in _TENSOR_FIELDS + ("star",):
                    if k == "edge_index" and not (_materialized(b, k) and _materialized(dst, k)):
                        continue                              # band(+star) batches: band_k / star describe the structure
                    src, cur = getattr(b, k, None), getattr(dst, k, None)
                    if src is None or cur is None:
                        continue
                    if src.shape != cur.shape or (k in ("edge_index", "ptr", "batch", "star") and not torch.equal(src.to(cur.device), cur)):
                        raise ValueError(f"GraphedStep: '{k}' of task '{t}' changed shape/structure; capture a new graph")
                    if k in ("x", "y", "pos"):
                        cur.copy_(src, non_blocking=True)
        self.graph.replay()
        # the replayed optimizer kernels changed the parameters on the device, but no Python version counter moved:
        # invalidate every cache derived from them (bf16 weight copies, normalised prototype banks) for eager code
        # that runs between replays (periodic validation)
        ops.bump_param_generation()
        return self.loss
