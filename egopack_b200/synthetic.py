"""Seeded synthetic Ego4D-shaped batches: the real features are licensed and absent, so the
benchmarks and parity tests use ``x ~ N(0,1)`` fp32 ``[N, 3, 1536]`` Omnivore-shaped features, unit-spaced ``pos``,
uniform AR/LTA labels over the Ego4D-v1 taxonomy sizes (115 verbs, 478 nouns), one-hot PNR labels per graph,
binary OSCC labels per graph, and N(0,1)/3 prototype banks.  Everything is created on the HOST (the H2D copy
is part of the end-to-end measurement).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .data import Batch, expand_base

N_VERBS, N_NOUNS = 115, 478
FEATURE_DIM, NUM_SEGMENTS = 1536, 3


def generator(seed: int, config_id: int = 0, rank: int = 0) -> torch.Generator:
    return torch.Generator().manual_seed(seed + 1000 * config_id + rank)


def make_batch(task: str, num_graphs: int, nodes_per_graph: int, gen: torch.Generator, *, feature_dim: int = FEATURE_DIM,
               num_segments: int = NUM_SEGMENTS, band_k: Optional[int] = 1, unlabeled: float = 0.0,
               n_verbs: int = N_VERBS, n_nouns: int = N_NOUNS, lta_inputs: int = 2, pin: bool = False,
               feature_dtype: Optional[torch.dtype] = None, compact: bool = False) -> Batch:
    """One collated task batch WITHOUT ``edge_index`` (the transform adds it; ``band_k`` is the structural hint
    our RadiusGraph would record).  task in {'ar','lta','oscc','pnr'}.

    PNR features are ONE vector per node repeated over the segments, as the reference's dataset builds them
    (``x=features.unsqueeze(1).repeat(1, 3, 1)``, data/ego4d_oscc.py:291): materialised like there by default, or with
    ``compact=True`` as the stride-0 view a loader that defers the repeat hands over (``data.replicated_base``)."""
    n = num_graphs * nodes_per_graph
    b = Batch()
    if task == "pnr" and num_segments > 1:
        base = torch.randn(n, feature_dim, generator=gen)
        b.x = expand_base(base, num_segments) if compact else base.unsqueeze(1).repeat(1, num_segments, 1)
    else:
        b.x = torch.randn(n, num_segments, feature_dim, generator=gen)
    b.pos = torch.arange(nodes_per_graph, dtype=torch.long).repeat(num_graphs)
    b.batch = torch.arange(num_graphs, dtype=torch.long).repeat_interleave(nodes_per_graph)
    b.ptr = torch.arange(num_graphs + 1, dtype=torch.long) * nodes_per_graph
    if task in ("ar", "lta"):
        y = torch.stack([torch.randint(0, n_verbs, (n,), generator=gen),
                         torch.randint(0, n_nouns, (n,), generator=gen)], 1)
        if task == "lta":                                    # first `lta_inputs` clips of each graph are inputs
            y = y.view(num_graphs, nodes_per_graph, 2)
            y[:, :lta_inputs] = -1
            y = y.view(n, 2)
        elif unlabeled > 0:
            y[torch.rand(n, generator=gen) < unlabeled] = -1
        b.y = y
    elif task == "oscc":
        b.y = torch.randint(0, 2, (num_graphs,), generator=gen)
    elif task == "pnr":
        hot = torch.randint(0, nodes_per_graph, (num_graphs,), generator=gen)
        y = torch.zeros(num_graphs, nodes_per_graph)
        y[torch.arange(num_graphs), hot] = 1
        b.y = y.view(n)
    else:
        raise ValueError(task)
    if band_k is not None and task != "lta":
        b.band_k = band_k
    if feature_dtype is not None:                            # a loader that stores bf16 features (Batch.to_feature_dtype)
        b.to_feature_dtype(feature_dtype)
    if pin:
        b.pin_memory()
    return b


def make_banks(tasks, num_protos: int, channels: int, gen: torch.Generator) -> Dict[str, torch.Tensor]:
    return {t: torch.randn(num_protos, channels, generator=gen) / 3 for t in tasks}
