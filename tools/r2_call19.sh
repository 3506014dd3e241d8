#!/bin/bash
# Round 2, GPU call 19: reference thread count under the cgroup quota; default bench after the allocation-free pass; ncu metrics of the tail / screen / lane kernels.
set -u
OUT=gpurun_out/r2c19
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 4 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-700)" | tee -a "$OUT/summary.txt"; }
step ref_threads 200 python tools/ref_threads.py 100000 "16,128,32,16"
step bench 1200 python bench.py
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,launch__grid_size,launch__registers_per_thread
step ncu 400 ncu --metrics $M --clock-control none -k regex:'k_tail_final|k_tail_pass|k_prescreen|k_nwlane|k_kord|k_gapless' -s 300 -c 240 --csv --log-file "$OUT/tail_metrics.csv" python tools/run_once.py 1000000
