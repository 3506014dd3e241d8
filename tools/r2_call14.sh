#!/bin/bash
# Round 2, GPU call 14: final single-GPU validation and measurements: whole GPU suite, default bench line (with legs), ncu evidence.
set -u
OUT=gpurun_out/r2c14
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-400)" | tee -a "$OUT/summary.txt"; }
step pytest_gpu 1500 python -m pytest tests -m gpu -q -x
step smoke 300 python __graft_entry__.py --smoke
step bench_default 1700 python bench.py --steps 20 --warmup 5
step ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k_nwrow|k_nwlane|k_prescreen|k_kord|k_tail_final" -s 24 -c 30 -o "$OUT/k_r2_final_full" python tools/run_once.py 1000000
step launches_1e6 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches_1e6.csv" python tools/run_once.py 1000000
tail -c 1500 "$OUT/bench_default.log"
