#!/bin/bash
# Round 2, GPU call 21: A/B of three builds on one box (A = before the prescreen / tail / lane changes, B = call 20's state + 4 raws per
# thread in k_tail_pass, C = working tree: + k_tail_final 4 raws per thread, group-wide lambda in k_nwlane); GPU suite on C.
set -u
OUT=gpurun_out/r2c21
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 7 "$OUT/$name.log" | cut -c1-330)" | tee -a "$OUT/summary.txt"; }
step ab 500 python tools/ab_probe.py 1000000 8 tools/ab/libdada2b_A.so tools/ab/libdada2b_B.so dada2_b200/libdada2b.so
step pytest_gpu 900 python -m pytest tests -m gpu -x -q
