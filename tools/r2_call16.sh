#!/bin/bash
# Round 2, GPU call 16 (2 GPUs): sharded re-upload (own reads packed, all-gather over NVLink) on real NCCL; goldens; bench at N = 2 on 1e6 with the CPU parity diff.
set -u
OUT=gpurun_out/r2c16
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-500)" | tee -a "$OUT/summary.txt"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
step pytest_multi 600 python -m pytest tests/test_gpu_multi.py -q -x
step bench_n2 1500 $TR bench.py --gpus 2 --steps 10 --warmup 3
tail -c 2500 "$OUT/bench_n2.log"
