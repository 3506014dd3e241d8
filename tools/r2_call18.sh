#!/bin/bash
# Round 2, GPU call 18: which call of the resident pass absorbs the 0.3-1 s stalls (DADA2B_STALLWATCH), and is the whole process stalled?
set -u
OUT=gpurun_out/r2c18
mkdir -p "$OUT"
timeout 300 python tools/stall_probe.py 1000000 60 > "$OUT/probe.log" 2>&1
echo "rc=$?"; grep -c PASS "$OUT/probe.log"; grep -v PASS "$OUT/probe.log" | head -60
