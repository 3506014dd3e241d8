#!/bin/bash
# Round 2, GPU call 1: ncu evidence on the round-1 NW variants (why fewer SASS instructions bought nothing) + launch list at 1e6.
set -u
OUT=gpurun_out/r2c1
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 2 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-400)" | tee -a "$OUT/summary.txt"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > "$OUT/smi.txt"
DADA2B_TWOPHASE=1 DADA2B_BOUND16=1 step ncu_bound16 600 ncu --set full --clock-control none --import-source on -k regex:k_nw -s 60 -c 9 -o "$OUT/k_bound16_full" python tools/run_once.py 100000
DADA2B_NWFWD_V2=1 step ncu_nwfwd2 600 ncu --set full --clock-control none --import-source on -k regex:k_nwfwd2 -s 30 -c 4 -o "$OUT/k_nwfwd2_full" python tools/run_once.py 100000
step ncu_nwfwd 600 ncu --set full --clock-control none --import-source on -k regex:k_nwfwd -s 30 -c 4 -o "$OUT/k_nwfwd_full" python tools/run_once.py 100000
step launches_1e6 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches_1e6.csv" python tools/run_once.py 1000000
cat "$OUT/summary.txt"
