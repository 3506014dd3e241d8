#!/bin/bash
# Round 2, GPU call 24 (2 GPUs, last of the budget): the final library sharded over two ranks on real NCCL: bench at 1e5 (parity diff inside), then the multi-GPU goldens.
set -u
OUT=gpurun_out/r2c24
mkdir -p "$OUT"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 75 $TR bench.py --gpus 2 --nuniques 100000 --steps 3 --warmup 3 > "$OUT/bench_n2_1e5.log" 2>&1; echo "bench rc=$?"
tail -c 1500 "$OUT/bench_n2_1e5.log"
timeout 40 python -m pytest tests/test_gpu_multi.py -q -x > "$OUT/pytest_multi.log" 2>&1; echo "pytest rc=$?"; tail -n 2 "$OUT/pytest_multi.log"
