#!/bin/bash
# First GPU call of a round, one gpurun:   gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'
# Everything that round 1 could only validate on the host SIMT emulator, in priority order, each step under its own timeout;
# results land in gpurun_out/r2/ (copy the summaries worth keeping into profiles/).
set -u
OUT=gpurun_out/r2
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 2 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-400)" | tee -a "$OUT/summary.txt"; }
# 1. the GPU suite: validated default path + xfail-marked first runs (experimental variants, long reads, bimera kernels); -rxX lists them
step pytest_gpu 1500 python -m pytest tests -m gpu -q -rxX
# 2. the bench line: value / e2e / roofline / cpu_baseline + A/B of the experimental variants + bimera leg
step bench_n1 900 python bench.py --gpus 1 --steps 5 --warmup 3
# 3. launch list of one 1e5 pass (shares of the step) and one full capture of the dominant kernel
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file "$OUT/launches.csv" python tools/run_once.py 100000
step ncu_nwfwd_full 900 ncu --set full --clock-control none --import-source on -k regex:k_nwfwd -s 40 -c 3 -o "$OUT/k_nwfwd_full" python tools/run_once.py 100000
# 4. bimera kernels: launch list + full capture of both alignment kernels
step ncu_bimera 900 ncu --set full --clock-control none --import-source on -k regex:k_bim -c 12 -o "$OUT/k_bim_full" python tools/bimera_leg.py 1500 8
# 5. BASELINE configs[4] flavour (1.5 kb, band 32, homopolymer gaps): default vs DADA2B_NWFWD_V2
step config5 1800 python tools/run_config5.py 20000 1500
# 6. the accelerated steps chained on the device: derep (resident) -> dada -> bimera, per-step wall / device times
step pipeline 600 python tools/pipeline_demo.py 400000
cat "$OUT/summary.txt"
