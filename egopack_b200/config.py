"""Global numeric mode.

``bf16`` (default, the product path): activations are stored in bf16, every Linear runs on the tcgen05 tensor
cores with fp32 accumulation, normalisation statistics / losses / gradients of parameters are fp32.
``fp32`` (parity mode): fp32 activations and FFMA GEMMs, within 1e-4 relative of the fp32 oracle.
"""
from __future__ import annotations

import contextlib

import torch

_PRECISION = "bf16"


def set_precision(mode: str) -> None:
    global _PRECISION
    if mode not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = mode


def get_precision() -> str:
    return _PRECISION


def compute_dtype() -> torch.dtype:
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


@contextlib.contextmanager
def precision(mode: str):
    old = get_precision()
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(old)
