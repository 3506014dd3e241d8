#!/bin/bash
# Round 2, GPU call 3: ALU micro-benchmarks + first hardware run of the exact / final row kernels + ncu on large rounds.
set -u
OUT=gpurun_out/r2c3
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-600)" | tee -a "$OUT/summary.txt"; }
step ubench 120 tools/ubench/alu
step pytest_parity 900 python -m pytest tests/test_gpu_parity.py -x -q
DADA2B_VERBOSE=1 step run_1e5 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 step run_1e6 900 python tools/run_big.py 1000000
step ncu_nwrow 900 ncu --set full --clock-control none --import-source on -k regex:k_nwrow -s 0 -c 5 -o "$OUT/k_nwrow_1e6_full" python tools/run_once.py 1000000
cat "$OUT/ubench.log"
grep -h "loop NW\|one-shot\|PARITY" "$OUT"/run_1e5.log "$OUT"/run_1e6.log | cut -c1-700
