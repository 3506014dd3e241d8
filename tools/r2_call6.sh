#!/bin/bash
# Round 2, GPU call 6 (2 GPUs): owner mode (default of sharded runs) on hardware: goldens, 1e6 strong scaling, bench at N=2.
set -u
OUT=gpurun_out/r2c6
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 4 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-900)" | tee -a "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
step multi_owner 900 $TR tools/run_sharded.py 1000000
DADA2B_REPLICATED=1 step multi_replicated 900 $TR tools/run_sharded.py 1000000
step bench_n2_1e5 900 $TR bench.py --gpus 2 --nuniques 100000 --steps 5 --warmup 3
step single_1e6 600 python tools/run_big.py 1000000
grep -h "sharded world\|PARITY\|MISMATCH\|one-shot" "$OUT"/*.log | cut -c1-600
tail -c 2500 "$OUT/bench_n2_1e5.log"
