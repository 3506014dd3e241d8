#!/bin/bash
# Round 2, GPU call 9: thread-per-raw prescreen; are the 250-450 ms step outliers caused by the NVML sampler thread?
set -u
OUT=gpurun_out/r2c9
mkdir -p "$OUT"
step() { local name=$1 t=$2; shift 2; echo "=== $name" | tee -a "$OUT/summary.txt"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "rc=$? $(tail -n 3 "$OUT/$name.log" | tr '\n' ' ' | cut -c1-300)" | tee -a "$OUT/summary.txt"; }
step pytest_parity 900 python -m pytest tests/test_gpu_parity.py -x -q
DADA2B_BENCH_NOCLOCKS=1 step bench_noclocks 900 python bench.py --steps 20 --warmup 3 --no-legs --no-cpu-baseline
step bench_clocks 900 python bench.py --steps 20 --warmup 3 --no-legs --no-cpu-baseline
step ncu_prescreen 600 ncu --set full --clock-control none -k regex:"k_prescreen|k_classify" -s 10 -c 4 -o "$OUT/k_screen_full" python tools/run_once.py 1000000
python - <<'PY'
import json
for f in ("bench_noclocks", "bench_clocks"):
    l = [x for x in open("gpurun_out/r2c9/%s.log" % f) if x.startswith('{"metric')]
    d = json.loads(l[-1])
    print(f, "value", round(d["value"]), "ms/step", round(d["ms_per_step"], 1), "median", d["ms_per_step_median"], "step_ms", d["step_ms"], "e2e", d["e2e_step_ms"], "prescreen", d["roofline_screen"]["achieved"], d["kernel_ms"])
PY
