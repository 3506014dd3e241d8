#!/bin/bash
# Round 2, GPU call 4: first hardware run of k_nwlane (small rounds), k_prescreen (TMA tier 0) ; fused tail A/B at 1e5 and 1e6.
set -u
OUT=gpurun_out/r2c4
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-600)" | tee -a "$OUT/summary.txt"; }
step pytest_parity 900 python -m pytest tests/test_gpu_parity.py -x -q
DADA2B_VERBOSE=1 step run_1e5 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 DADA2B_FUSED_TAIL=1 step run_1e5_fused 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 step run_1e6 900 python tools/run_big.py 1000000
DADA2B_VERBOSE=1 DADA2B_FUSED_TAIL=1 step run_1e6_fused 900 python tools/run_big.py 1000000
step launches_1e5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches_1e5.csv" python tools/run_once.py 100000
step ncu_new 900 ncu --set full --clock-control none --import-source on -k regex:"k_prescreen|k_nwlane" -s 20 -c 6 -o "$OUT/k_new_full" python tools/run_once.py 100000
grep -h "loop NW\|one-shot\|PARITY" "$OUT"/run_*.log | cut -c1-900
