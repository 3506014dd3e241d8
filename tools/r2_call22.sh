#!/bin/bash
# Round 2, GPU call 22 (final single-GPU capture): per-launch list with DRAM bytes and pipe utilisation of every kernel of one 1e6 run; default bench (legs included).
set -u
OUT=gpurun_out/r2c22
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | cut -c1-500)" | tee -a "$OUT/summary.txt"; }
step bench 1500 python bench.py
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active
step ncu_list 600 ncu --metrics $M --clock-control none -c 1700 --csv --log-file "$OUT/launches_1e6.csv" python tools/run_once.py 1000000
ls -la "$OUT"
step pytest_gpu 600 python -m pytest tests -m gpu -x -q
