"""CPU: the C-ABI library builds/loads and exports every symbol ``include/egopack_b200.h`` declares, and the
ctypes prototypes agree with the header's argument counts.  No compute call is made (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "egopack_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|size_t)\s+(egp_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(2)] = n
    return out


def test_header_declares_the_path():
    fns = header_functions()
    for must in ("egp_band_edge_fill", "egp_lta_edge_fill", "egp_sage_mean_band", "egp_sage_mean_csr",
                 "egp_graph_layernorm_fwd", "egp_graph_layernorm_bwd", "egp_row_layernorm_fwd", "egp_gemm",
                 "egp_cos_topk", "egp_proto_max_gather", "egp_segment_max_pool_fwd", "egp_posenc_add"):
        assert must in fns


def test_library_exports_every_declared_symbol_and_prototypes_match():
    from egopack_b200 import _lib
    lib = _lib.load()
    fns = header_functions()
    assert set(fns) == set(_lib.SIGNATURES), set(fns) ^ set(_lib.SIGNATURES)
    for name, nargs in fns.items():
        assert hasattr(lib, name), f"{name} not exported"
        assert len(_lib.SIGNATURES[name][1]) == nargs, name
    assert lib.egp_version() == 2


def test_ops_refuse_host_tensors():
    import torch
    from egopack_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.cast(torch.zeros(8), torch.bfloat16)
