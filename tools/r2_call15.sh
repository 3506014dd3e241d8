#!/bin/bash
# Round 2, GPU call 15: DRAM traffic and pipe utilisation of the NW family in large rounds (explicit metric list: the --set full replay of a large k_nwrow<BOUND> launch failed in call 14).
set -u
OUT=gpurun_out/r2c15
mkdir -p "$OUT"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,launch__grid_size,launch__registers_per_thread,lts__t_bytes.sum
timeout 900 ncu --metrics $M --clock-control none -k regex:"k_nwrow|k_nwlane|k_gapless|k_prescreen|k_kord" -c 400 --csv --log-file "$OUT/nw_metrics.csv" python tools/run_once.py 1000000 > "$OUT/run.log" 2>&1
echo rc=$?; tail -2 "$OUT/run.log" | cut -c1-200; wc -l "$OUT/nw_metrics.csv"
