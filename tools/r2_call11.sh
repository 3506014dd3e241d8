#!/bin/bash
# Round 2, GPU call 11: exact min-sum in the streaming screen + k_kord; whole GPU suite; timings.
set -u
OUT=gpurun_out/r2c11
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-400)" | tee -a "$OUT/summary.txt"; }
step pytest_gpu 1500 python -m pytest tests -m gpu -q -x
DADA2B_VERBOSE=1 step run_1e5 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 step run_1e6 900 python tools/run_big.py 1000000 cpu
grep -h "loop NW\|one-shot\|PARITY\|loop done\|upload:" "$OUT"/run_*.log | cut -c1-1100
