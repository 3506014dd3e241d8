#!/bin/bash
# Round 2, GPU call 5: streaming traceback in the exact kernels; whole GPU suite; bench.py smoke at 1e5.
set -u
OUT=gpurun_out/r2c5
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-600)" | tee -a "$OUT/summary.txt"; }
step pytest_parity 900 python -m pytest tests/test_gpu_parity.py -x -q
DADA2B_VERBOSE=1 step run_1e5 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 step run_1e6 900 python tools/run_big.py 1000000
step launches_1e5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches_1e5.csv" python tools/run_once.py 100000
step pytest_gpu 1500 python -m pytest tests -m gpu -q -x
step bench_1e5 900 python bench.py --nuniques 100000 --steps 5 --warmup 3 --no-legs
grep -h "loop NW\|one-shot\|PARITY" "$OUT"/run_*.log | cut -c1-900
tail -c 3000 "$OUT/bench_1e5.log"
