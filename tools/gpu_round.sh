#!/bin/bash
# One GPU-box round: parity tests, the bench line (+ per-launch trace), every kernel alone, the ncu launch list of the
# same bench command and `--set full` captures of the kernels that ship.  Everything lands in gpurun_out/.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tests|bench|kbench|launches|ncu]...'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
want() { [ $# -eq 0 ] && return 0; for a in "${ARGS[@]}"; do [ "$a" = "$1" ] && return 0; done; return 1; }
ARGS=("$@"); [ ${#ARGS[@]} -eq 0 ] && ARGS=(tests bench kbench launches ncu)
if want tests; then python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2_pytest.log; tail -4 gpurun_out/r2_pytest.log; fi
if want bench; then
  python bench.py --steps 10 --warmup 3 --trace-out gpurun_out/r2_trace_per_op.json > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
  head -c 1500 gpurun_out/r2_bench_n1.json; echo; tail -2 gpurun_out/r2_bench_n1.err
  python bench.py --impl reference --steps 3 --warmup 1 --cpu-sweep > gpurun_out/r2_bench_reference.json 2>/dev/null; head -c 600 gpurun_out/r2_bench_reference.json; echo
fi
if want kbench; then
  python tools/kernel_bench.py > gpurun_out/r2_kernel_bench.txt 2>&1; tail -70 gpurun_out/r2_kernel_bench.txt
  # the memory-bound kernels again at the row count of the stacked c2 pass (the default table is one task batch)
  EGP_KB_N=98304 python tools/kernel_bench.py 2>&1 | grep -v "gemm\|cublas\|cos_topk" > gpurun_out/r2_kernel_bench_n98304.txt; cat gpurun_out/r2_kernel_bench_n98304.txt
fi
if want launches; then
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launch_list_ncu.csv \
      python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline > gpurun_out/r2_launch_bench.log 2>&1
  python tools/summarize_launches.py gpurun_out/r2_launch_list_ncu.csv 2 > gpurun_out/r2_launch_list_summary.csv; head -40 gpurun_out/r2_launch_list_summary.csv
fi
if want ncu; then
  cap() {  # name regex extra-args... : one full capture of the first launch matching the regex
    local name=$1 re=$2; shift 2
    ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$re" -c 1 -f -o gpurun_out/r2_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1
    python tools/extract_ncu.py gpurun_out/r2_$name.ncu-rep > gpurun_out/r2_ncu_$name.csv 2>/dev/null
    rm -f gpurun_out/r2_$name.ncu-rep
  }
  B="python bench.py --steps 1 --warmup 1 --quick --no-cpu-baseline"
  cap sage_mean_band_star_fwd 'sage_mean_band_reg_kernel<[^>]*8, \(int\)1>' $B
  cap sage_mean_band_star_bwd 'sage_mean_band_reg_kernel<[^>]*8, \(int\)2>' $B
  cap sage_hub_fixup 'sage_hub_fixup' $B
  cap rln_bwd_block 'rln_bwd_block_kernel' $B
  cap rln_fwd 'rln_fwd_full_kernel' $B
  cap gln_bwd_reduce 'gln_bwd_reduce_kernel' $B
  cap gln_bwd_apply 'gln_bwd_apply_kernel' $B
  cap gln_stats 'gln_stats_kernel' $B
  cap gln_apply 'gln_apply_kernel' $B
  cap colsum 'colsum_partial_kernel' $B
  cap act_bwd_colsum 'act_bwd_colsum_kernel' $B
  cap gemm_fwd 'tc_gemm_kernel' $B
  cap sage_mean_band_k1 'sage_mean_band_reg_kernel<[^>]*8, \(int\)0>' python tools/kernel_bench.py sage_mean_band_k1_bf16 --reps 1
  cap sage_mean_band_run_c4 'sage_mean_band_run_kernel' python tools/probe_band_wide.py
  cap segment_max_pool 'segment_max_fwd_kernel' python tools/kernel_bench.py segment_max_pool_fwd --reps 1
  cap proto_max_gather 'proto_max_gather' python tools/kernel_bench.py proto_max_gather --reps 1
  cap ce_loss_fwd 'ce_loss_fwd' python tools/kernel_bench.py ce_2heads_fwd --reps 1
  ls gpurun_out/r2_ncu_*.csv | head -30
fi
