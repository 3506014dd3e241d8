#!/bin/bash
# Round 2, GPU call 8: async host replay, survivors through the lane kernel; kernel-level loop aligner test; full default bench.py at N=1.
set -u
OUT=gpurun_out/r2c8
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-600)" | tee -a "$OUT/summary.txt"; }
step pytest_parity 900 python -m pytest tests/test_gpu_parity.py -x -q
DADA2B_VERBOSE=1 step run_1e5 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 step run_1e6 900 python tools/run_big.py 1000000
step bench_default 1700 python bench.py --steps 10 --warmup 3
grep -h "loop NW\|one-shot\|PARITY\|loop done" "$OUT"/run_*.log | cut -c1-1000
tail -c 6000 "$OUT/bench_default.log"
