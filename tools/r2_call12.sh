#!/bin/bash
# Round 2, GPU call 12: 48-entry repeated-5-mer lists, thread-per-pair gapless loop kernel, sharded quality upload.
set -u
OUT=gpurun_out/r2c12
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-400)" | tee -a "$OUT/summary.txt"; }
step pytest_parity 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zzz_edge.py -q -x
DADA2B_VERBOSE=1 step run_1e5 600 python tools/run_big.py 100000 cpu
DADA2B_VERBOSE=1 step run_1e6 900 python tools/run_big.py 1000000 cpu
step launches_1e6 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file "$OUT/launches_1e6.csv" python tools/run_once.py 1000000
grep -h "loop NW\|one-shot\|PARITY\|loop done\|upload:" "$OUT"/run_*.log | cut -c1-1100
