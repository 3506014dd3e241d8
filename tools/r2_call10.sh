#!/bin/bash
set -u
OUT=gpurun_out/r2c10
mkdir -p "$OUT"
timeout 900 python tools/stall_probe.py 1000000 24 > "$OUT/probe.out" 2> "$OUT/probe.err"
grep -n "PASS" "$OUT/probe.err" | head -30
