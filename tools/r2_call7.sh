#!/bin/bash
# Round 2, GPU call 7: rewritten lane kernel (G = 8, 5 slots per lane); host-side phase timing of the round loop; launch list at 1e6.
set -u
OUT=gpurun_out/r2c7
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-600)" | tee -a "$OUT/summary.txt"; }
step pytest_parity 900 python -m pytest tests/test_gpu_parity.py -x -q
DADA2B_VERBOSE=1 step run_1e5 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 step run_1e6 900 python tools/run_big.py 1000000
step launches_1e6 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches_1e6.csv" python tools/run_once.py 1000000
grep -h "loop NW\|one-shot\|PARITY\|loop done" "$OUT"/run_*.log | cut -c1-1100
