"""Global numeric mode.

``bf16`` (default, the product path): activations are stored in bf16, every Linear runs on the tcgen05 tensor
cores with fp32 accumulation, normalisation statistics / losses / gradients of parameters are fp32.
``fp32`` (parity mode): fp32 activations, within 1e-4 relative of the fp32 oracle.  Its Linear layers run on the
tensor cores too: every fp32 operand is split into three bf16 terms (24 significand bits) and the six largest
term products are accumulated in fp32 by ONE launch of the same tcgen05 kernel (``fp32_gemm = "bf16x6"``, error
~2^-22; ``"bf16x3"`` keeps three products, ~2^-15, for 2x the speed).  ``"ffma"`` selects the SIMT FFMA kernel.
"""
from __future__ import annotations

import contextlib

import torch

_PRECISION = "bf16"
_FP32_GEMM = "bf16x6"


def set_fp32_gemm(kind: str) -> None:
    """How fp32 GEMMs are evaluated: 'bf16x6' (default), 'bf16x3' (both on the tensor cores) or 'ffma' (SIMT)."""
    global _FP32_GEMM
    if kind not in ("bf16x6", "bf16x3", "ffma"):
        raise ValueError("fp32_gemm must be 'bf16x6', 'bf16x3' or 'ffma'")
    _FP32_GEMM = kind


def get_fp32_gemm() -> str:
    return _FP32_GEMM


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


_DETERMINISTIC = False


def set_deterministic(on: bool) -> None:
    """Bit-reproducible weight gradients: split-K GEMMs write per-split slabs and sum them in a fixed order instead of
    reduce-adding into the output (``egp_set_deterministic``).  Everything else on the training path is deterministic
    already.  Costs one extra pass over ``splits x [M, N]`` fp32 per split-K weight gradient."""
    global _DETERMINISTIC
    from . import _lib
    _lib.call("egp_set_deterministic", int(bool(on)))
    _DETERMINISTIC = bool(on)


def is_deterministic() -> bool:
    return _DETERMINISTIC
