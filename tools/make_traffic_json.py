"""Builds profiles/r2_ncu_traffic.json from the `ncu --set full` extracts of tools/gpu_round.sh (gpurun_out/r2_ncu_*.csv):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of every kernel that bench.py reports a roofline
for, keyed by the kernel names of the bench line, next to the algorithmic bytes of that launch.  Not a pytest file.

    python tools/make_traffic_json.py [gpurun_out] > profiles/r2_ncu_traffic.json
"""
import csv
import json
import os
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
N, C, b = 98304, 1024, 2            # stacked rows of the c2 step, channels, bytes per element (bf16)
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def load(name):
    path = os.path.join(SRC, f"r2_ncu_{name}.csv")
    if not os.path.exists(path):
        return None
    rows = {r[0]: (r[1], r[2]) for r in csv.reader(open(path)) if len(r) > 2}
    if "dram__bytes_read.sum" not in rows:
        return None
    val = lambda k: float(rows[k][1].replace(",", "")) * UNIT.get(rows[k][0], 1)
    return {"bytes": int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")), "kernel": rows["Kernel Name"][1][:90],
            "us": float(rows["gpu__time_duration.sum"][1]), "source": f"r2_ncu_{name}.csv"}


def entry(parts, launch, algorithmic, note=""):
    got = [load(p) for p in parts]
    if any(g is None for g in got):
        return None
    return {"bytes": sum(g["bytes"] for g in got), "launch": launch, "algorithmic_bytes": int(algorithmic),
            "source": " + ".join(g["source"] for g in got), "cold_cache_us": round(sum(g["us"] for g in got), 1),
            "kernels": [g["kernel"] for g in got], "note": note}


out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (ncu --set full --clock-control none, cold cache, "
                   "serialised) of the kernels that ship in this tree; bench.py copies `bytes` into roofline.traffic / "
                   "roofline_hbm[i].traffic.  Written bytes can sit below the algorithmic figure: part of the output is still dirty "
                   "in the 126 MB L2 when the kernel ends."}
spec = {
    "tc_gemm_kernel": (["gemm_fwd"], "98304x1024x4608 NN -> bf16 +bias (first Linear of the stacked TRN pooling)",
                       N * 4608 * b + C * 4608 * b + N * C * b),
    "sage_mean_band_star k=1": (["sage_mean_band_star_fwd"], "98304 nodes x 1024 ch bf16, radius 1, forward (band + star extension rows)", 2 * N * C * b),
    "sage_mean_band_star k=1 (backward)": (["sage_mean_band_star_bwd", "sage_hub_fixup"], "same, backward: fused hub sums + fix-up kernel", 2 * N * C * b),
    "sage_mean_band k=1": (["sage_mean_band_k1"], "32768 nodes x 1024 ch bf16, radius 1 (tools/kernel_bench.py)", 2 * 32768 * C * b),
    "sage_mean_band k=16": (["sage_mean_band_run_c4"], "524288 nodes x 1024 ch bf16, radius 16 (tools/probe_band_wide.py, the c4 shape)", 2 * 524288 * C * b),
    "graph_layernorm_fwd": (["gln_apply"], "98304 x 1024 bf16, 3 segments: normalise pass (statistics come from the GEMM epilogue)", 2 * N * C * b),
    "graph_layernorm_bwd": (["gln_bwd_reduce", "gln_bwd_apply"], "98304 x 1024 bf16, 3 segments: reduce + apply", 5 * N * C * b),
    "row_layernorm_fwd": (["rln_fwd"], "98304 x 1024 bf16 + ReLU + dropout 0.5 (TRN pooling)", 2 * N * C * b),
    "row_layernorm_bwd": (["rln_bwd_block"], "first row-LN backward of the step (task net, 32768 x 1024 bf16)", 4 * 32768 * C * b),
    "act_bwd_colsum": (["act_bwd_colsum"], "98304 x 1024 bf16: ReLU backward + column sums", 3 * N * C * b),
    "segment_max_pool_fwd": (["segment_max_pool"], "32768 x 1024 bf16, 256 graphs (tools/kernel_bench.py)", 32768 * C * b),
    "proto_max_gather": (["proto_max_gather"], "32768 nodes, k=4, 4096 x 1024 bf16 bank (tools/kernel_bench.py)", 32768 * C * b + 32768 * 4 * 8),
}
for key, (parts, launch, alg) in spec.items():
    e = entry(parts, launch, alg)
    if e is not None:
        out[key] = e
print(json.dumps(out, indent=1))
