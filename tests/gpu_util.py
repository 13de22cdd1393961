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


def grads_close(got, want, tol=TOL_F32, outlier_frac=1e-4):
    """Gradient check that tolerates activation-kink flips.  ReLU / LeakyReLU (and arg-max) make the gradient a
    discontinuous function of the pre-activations: with millions of activations a few sit within fp32 rounding of a
    kink, and two correct fp32 implementations (different summation orders) then disagree on isolated elements by O(1).
    So: either every element is within `tol` of max|want|, or all but `outlier_frac` of them are and the relative L2
    error stays below 50*tol."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    scale = want.abs().max().clamp(min=1e-30)
    err = (got - want).abs() / scale
    if float(err.max()) < tol:
        return True
    frac = float((err > tol).double().mean())
    l2 = float((got - want).norm() / want.norm().clamp(min=1e-30))
    return frac < outlier_frac and l2 < 50 * tol
