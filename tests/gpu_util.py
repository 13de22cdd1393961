"""Helpers shared by the ``-m gpu`` parity tests (all of which call the CUDA path through the C ABI)."""
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda"

# tolerances (BASELINE.json north_star): fp32 path 1e-4 relative, bf16 tensor-core path 2e-2 relative.
# "relative" = max |a-b| / max |b| per tensor for fp32; for bf16 the same on activations, and relative L2
# (||a-b|| / ||b||) on gradients, whose max-norm is dominated by isolated ReLU/arg-max flips under bf16 rounding.
TOL_F32 = 1e-4
TOL_BF16 = 2e-2


def rel_max(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp(min=1e-30))


def rel_l2(got, want):
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).norm() / want.norm().clamp(min=1e-30))


def canon_edges(e):
    e = e.cpu()
    return sorted(zip(e[1].tolist(), e[0].tolist()))


def graph_sizes_to_index(sizes):
    sizes_t = torch.tensor(sizes)
    batch = torch.repeat_interleave(torch.arange(len(sizes)), sizes_t)
    ptr = torch.cat([torch.zeros(1, dtype=torch.long), sizes_t.cumsum(0)])
    return batch, ptr
