#!/bin/bash
# Round 2, GPU call 17: stall probe (who stalls: CPU thread, GPU context, driver calls?) + bench at 1e6 with the diagonal-path split of the bound pass.
set -u
OUT=gpurun_out/r2c17
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-600)" | tee -a "$OUT/summary.txt"; }
step stall 60 tools/ubench/stall 12
(ps -eo pid,ppid,pcpu,etime,args --sort=-pcpu | head -25) > "$OUT/ps.log" 2>&1
step bench 900 python bench.py --steps 12 --warmup 3 --no-legs
step run_verbose 200 env DADA2B_VERBOSE=1 python tools/run_once.py 1000000
grep -c "round" "$OUT/run_verbose.log"
