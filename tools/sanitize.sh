#!/bin/bash
# compute-sanitizer passes over the parity tests of the hand-written kernels (race / memory checking is
# absent in the reference).  memcheck over the memory-bound kernels, losses, optimiser, band+star and small GEMMs;
# racecheck over the kernels that communicate through shared memory without async-proxy traffic (the TMA / mbarrier
# GEMM is excluded: racecheck does not model the async proxy).  Output: gpurun_out/r2_sanitizer_*.log
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MEM_K="sage_mean or wide_band or band_star or (layernorm and not epilogue) or cross_entropy or bce or flat_adam or posenc or segment_max or proto or max_combine or lta or band_edges or csr or label_rank or edit_distance or row_normalize"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_edges.py tests/test_gpu_optim.py tests/test_gpu_meters.py -m gpu -q -x -k "$MEM_K" \
    > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_gemm.py -m gpu -q -x -k "forward_layout or dual_and_epilogues or ffma_fp32" \
    > gpurun_out/r2_sanitizer_memcheck_gemm.log 2>&1
echo "memcheck gemm rc=$?" | tee -a gpurun_out/r2_sanitizer_memcheck_gemm.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "graph_layernorm_leaky or row_layernorm or cross_entropy or band_star or colsum or wide_band" \
    > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/r2_sanitizer_racecheck.log
for f in memcheck memcheck_gemm racecheck; do tail -n 4 gpurun_out/r2_sanitizer_$f.log; done
