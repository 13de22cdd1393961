"""ctypes binding of ``libegopack_b200.so`` (the C ABI declared in ``include/egopack_b200.h``).

The product path has NO fallback: if the shared library cannot be loaded (and cannot be built because nvcc is
absent) importing the ops raises, and every op raises on a non-CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libegopack_b200.so")

P, I64, I, F, SZ = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list EVERY symbol of include/egopack_b200.h (tests/test_abi.py checks)
SIGNATURES: Dict[str, tuple] = {
    "egp_version": (I, []),
    "egp_last_error": (I, [C.c_char_p, SZ]),
    "egp_device_info": (I, [P, P, P]),
    "egp_set_deterministic": (I, [I]),
    "egp_get_deterministic": (I, []),
    "egp_band_edge_count": (I, [P, P, P, I64, F, I, I, P, P]),
    "egp_band_edge_fill": (I, [P, P, P, I64, F, I, I, P, I64, P, P]),
    "egp_exclusive_scan_i32": (I, [P, I64, P, P]),
    "egp_lta_edge_count": (I, [P, P, I64, P, P, I64, F, I, P, P]),
    "egp_lta_edge_fill": (I, [P, P, I64, P, P, I64, F, I, P, I64, P, P]),
    "egp_band_windows": (I, [P, P, I64, I, P, P, P, P]),
    "egp_lta_star_counts": (I, [P, I64, P, I64, F, P, P]),
    "egp_band_star_windows": (I, [P, P, I64, I64, I, P, I64, I64, P, P, P, P, P, P, P, P]),
    "egp_csr_build": (I, [P, I64, I64, I, P, P, P, P]),
    "egp_csr_inv_degree": (I, [P, I64, P, P]),
    "egp_sage_mean_band": (I, [P, P, I64, I64, I64, I64, I, P, P, P, P, I, P]),
    "egp_sage_mean_band_star_workspace": (SZ, [I64, I64, I64]),
    "egp_sage_mean_band_star": (I, [P, P, I64, I64, I64, I64, I, P, P, P, P, P, P, P, P, I64, I, P, SZ, P]),
    "egp_sage_mean_csr": (I, [P, P, I64, I64, I64, I64, P, P, P, P, I, P]),
    "egp_graph_layernorm_workspace": (SZ, [I64, I64]),
    "egp_graph_layernorm_fwd": (I, [P, P, P, P, P, I64, I64, F, I, F, I, P, SZ, P]),
    "egp_graph_layernorm_bwd": (I, [P, P, P, P, P, P, P, P, P, I64, I64, F, I, F, I, P, SZ, P]),
    "egp_graph_layernorm_seg_workspace": (SZ, [I64, I64, I]),
    "egp_graph_layernorm_seg_fwd": (I, [P, P, P, P, P, I64, I64, I, P, F, I, F, I, P, SZ, P]),
    "egp_graph_layernorm_seg_bwd": (I, [P, P, P, P, P, P, P, P, P, I64, I64, I, P, F, I, F, I, P, SZ, P]),
    "egp_graph_layernorm_seg_fwd_rowstats": (I, [P, P, P, P, P, I64, I64, I, P, P, F, I, F, I, P]),
    "egp_gemm_rowstats_bytes": (SZ, [I64, I64]),
    "egp_gemm_rowstats": (I, [P, I64, I, P, I64, I, P, I64, P, I64, I64, P, P, I64, P, I64, I64, I64, I64, I, F, I, P, P]),
    "egp_row_layernorm_workspace": (SZ, [I64, I64]),
    "egp_row_layernorm_fwd": (I, [P, P, P, P, P, P, I64, I64, F, I, F, C.c_uint64, C.c_uint64, P, I, P]),
    "egp_row_layernorm_bwd": (I, [P, P, P, P, P, P, P, P, P, P, I64, I64, I, F, I, P, SZ, P]),
    "egp_posenc_add": (I, [P, P, P, P, I64, I64, I, P]),
    "egp_cast": (I, [P, P, I64, I, I, P]),
    "egp_cast_pad": (I, [P, I64, P, I64, I64, I64, I, I, P]),
    "egp_add": (I, [P, P, P, I64, I, P]),
    "egp_axpby": (I, [P, F, P, F, P, I64, I, P]),
    "egp_act_bwd": (I, [P, P, P, I64, I, F, I, P]),
    "egp_act_bwd_colsum_workspace": (SZ, [I64, I64]),
    "egp_act_bwd_colsum": (I, [P, P, P, P, I64, I64, I, F, I, P, SZ, P]),
    "egp_colsum_workspace": (SZ, [I64, I64]),
    "egp_colsum": (I, [P, P, I64, I64, I64, I, P, SZ, P]),
    "egp_mask_scale": (I, [P, P, P, I64, F, I, P]),
    "egp_gemm_workspace": (SZ, [I64, I64, I64]),
    "egp_gemm": (I, [P, I64, I, P, I64, I, P, I64, P, I64, I64, P, P, I64, P, I64, I64, I64, I64, I, F, I, I, I, P,
                     SZ, P]),
    "egp_row_normalize": (I, [P, P, I64, I64, I, I, P, P]),
    "egp_row_inv_norm": (I, [P, P, I64, I64, I64, I, P]),
    "egp_cos_topk_workspace": (SZ, [I64, I64, I64]),
    "egp_cos_topk": (I, [P, P, P, P, I64, I64, I64, I, P, P, F, P, P, P, SZ, P]),
    "egp_proto_max_gather": (I, [P, P, P, I64, I64, I64, I, I, P]),
    "egp_proto_max_scatter_bwd": (I, [P, P, P, P, P, I64, I64, I64, I, P]),
    "egp_class_sum_f64": (I, [P, P, P, I64, I64, I64, I, P]),
    "egp_max_combine_fwd": (I, [P, P, P, I64, I, P]),
    "egp_max_combine_bwd": (I, [P, P, P, P, I64, I, P]),
    "egp_segment_max_pool_fwd": (I, [P, P, P, P, I64, I64, I, P]),
    "egp_segment_max_pool_bwd": (I, [P, P, P, P, I64, I64, I, P]),
    "egp_adam_step": (I, [P, P, P, P, P, P, P, P, I64, P, I, P, P, F, F, F, F, P]),
    "egp_split_bf16": (I, [P, I64, I64, I64, I64, I64, P, I64, I64, I, P, P]),
    "egp_ce_loss_fwd": (I, [P, I64, P, I64, I64, I64, I64, F, P, I, P, P]),
    "egp_ce_loss_bwd": (I, [P, I64, P, P, I64, P, I64, I64, I64, I64, F, P, I64, I, P]),
    "egp_bce_logits_fwd": (I, [P, P, P, I64, P]),
    "egp_bce_logits_bwd": (I, [P, P, P, I64, P, I64, P]),
    "egp_weighted_mean": (I, [P, I64, F, P, I, P]),
    "egp_label_rank": (I, [P, I64, P, I64, I64, I64, I64, P, P]),
    "egp_segment_argmax": (I, [P, P, I64, I, P, P]),
    "egp_edit_distance_min": (I, [P, P, I64, I64, I64, P, P]),
}

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load (building first when the .so is absent and nvcc exists) and bind every entry point."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        if os.path.exists(_build.NVCC) and os.environ.get("EGP_NO_REBUILD") != "1":
            _build.build()                   # no-op when the source digest matches the built library
        elif not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing and nvcc is not available to build it; "
                              "egopack_b200 has no non-CUDA fallback")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the library does not export the symbol
            fn.restype, fn.argtypes = res, args
        _lib = lib
        return lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().egp_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses host tensors: there is no CPU path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("egopack_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(rc: int, name: str):
    if rc != 0:
        raise RuntimeError(f"{name} failed (status {rc}): {last_error()}")


CALL_COUNTS: Dict[str, int] = {}

# kernels launched per entry-point call (memsets excluded) -- used by bench.py's `gpu_launches` claim
KERNELS_PER_CALL = {
    "egp_band_edge_count": 1, "egp_band_edge_fill": 1, "egp_exclusive_scan_i32": 1, "egp_lta_edge_count": 1,
    "egp_lta_edge_fill": 1, "egp_band_windows": 1, "egp_csr_build": 4, "egp_csr_inv_degree": 1,
    "egp_sage_mean_band": 1, "egp_sage_mean_csr": 1, "egp_lta_star_counts": 1, "egp_band_star_windows": 1,
    "egp_sage_mean_band_star": 1,   # +1 fix-up launch in the backward direction (counted in ops._aggregate_launch)
    "egp_graph_layernorm_fwd": 2, "egp_graph_layernorm_bwd": 4,
    "egp_graph_layernorm_seg_fwd": 2, "egp_graph_layernorm_seg_bwd": 4, "egp_graph_layernorm_seg_fwd_rowstats": 2,
    "egp_gemm_rowstats": 1,
    "egp_row_layernorm_fwd": 1, "egp_row_layernorm_bwd": 2, "egp_posenc_add": 1, "egp_cast": 1, "egp_add": 1,
    "egp_axpby": 1, "egp_act_bwd": 1, "egp_act_bwd_colsum": 2, "egp_colsum": 2, "egp_mask_scale": 1, "egp_gemm": 1, "egp_row_normalize": 1,
    "egp_row_inv_norm": 1, "egp_cos_topk": 3, "egp_proto_max_gather": 1, "egp_max_combine_fwd": 1,
    "egp_max_combine_bwd": 1, "egp_segment_max_pool_fwd": 1, "egp_segment_max_pool_bwd": 1,
    "egp_label_rank": 1, "egp_segment_argmax": 1, "egp_edit_distance_min": 1,
    "egp_adam_step": 2, "egp_split_bf16": 1, "egp_ce_loss_fwd": 1, "egp_ce_loss_bwd": 1, "egp_bce_logits_fwd": 1, "egp_bce_logits_bwd": 1, "egp_weighted_mean": 1,
}


# EGP_NVTX=1: every C-ABI call is wrapped in an NVTX range named after the entry point, so a timeline (nsys, or ncu's
# --nvtx filters: `ncu --nvtx --nvtx-include "egp_sage_mean_band_star/"`) attributes kernels to the operator that
# launched them.  Off by default: a push/pop pair costs ~1 us of host time per call.
NVTX = os.environ.get("EGP_NVTX", "0") not in ("", "0")


def call(name: str, *args):
    if NVTX:
        torch.cuda.nvtx.range_push(name)
        try:
            rc = getattr(load(), name)(*args)
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        rc = getattr(load(), name)(*args)
    CALL_COUNTS[name] = CALL_COUNTS.get(name, 0) + 1
    check(rc, name)


def kernel_launches() -> int:
    return sum(n * KERNELS_PER_CALL.get(k, 1) for k, n in CALL_COUNTS.items())


def size(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))


DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1}

_workspaces: Dict[tuple, torch.Tensor] = {}


def workspace(nbytes: int, device, tag: str = "default") -> torch.Tensor:
    """A grow-only scratch buffer per (device, tag).  Kernels using it are ordered on the current stream."""
    key = (str(device), tag)
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w
