#!/bin/bash
# Round 2, GPU call 20: full GPU suite + bench after (mask AND removal, p-value shortcut, deferred list loads in k_prescreen, diagonal shortcut in k_nwlane); source-level ncu of k_nwlane / k_prescreen.
set -u
OUT=gpurun_out/r2c20
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 4 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-700)" | tee -a "$OUT/summary.txt"; }
step bench 900 python bench.py --no-legs --steps 8 --warmup 3
step pytest_gpu 1500 python -m pytest tests -m gpu -x -q
step ncu_full 500 ncu --set full --import-source on --clock-control none -k regex:'k_nwlane|k_prescreen|k_tail_final' -s 60 -c 9 -o "$OUT/k_r2_lane_screen_tail" python tools/run_once.py 1000000
ls -la "$OUT"
