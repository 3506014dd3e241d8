#!/bin/bash
# Round 2, GPU call 13 (8 GPUs): owner-mode strong scaling at 1e6 uniques; parity on the goldens on every rank; bench line at N = 8.
set -u
OUT=gpurun_out/r2c13
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 4 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-600)" | tee -a "$OUT/summary.txt"; }
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551"
DADA2B_VERBOSE=1 step multi8 900 $TR8 tools/run_sharded.py 1000000
step bench8 1500 $TR8 bench.py --gpus 8 --steps 10 --warmup 3
grep -h "sharded world" "$OUT"/multi8.log | cut -c1-500
grep -h "loop done" "$OUT"/multi8.log | tail -4 | cut -c1-400
grep -c "PARITY OK" "$OUT"/multi8.log; grep -c MISMATCH "$OUT"/multi8.log
tail -c 3500 "$OUT/bench8.log"
