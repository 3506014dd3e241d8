#!/bin/bash
# Round 2, GPU call 23: full launch list (device time only) of one 1e6 pass with the final kernels.
set -u
OUT=gpurun_out/r2c23
mkdir -p "$OUT"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file "$OUT/launches_1e6.csv" python tools/run_once.py 1000000 > "$OUT/run.log" 2>&1
echo "rc=$?"; tail -n 2 "$OUT/run.log" | cut -c1-300; wc -l "$OUT/launches_1e6.csv"
