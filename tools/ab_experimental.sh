#!/bin/bash
# A/B of the experimental paths on a GPU box (one gpurun call):  gpurun --timeout 1500 -- 'bash tools/ab_experimental.sh'
# For every flag set: the GPU parity suite with the flag exported, then one bench line.  Everything lands in
# gpurun_out/ab/; each step has its own timeout so that a hang in one variant cannot eat the whole call.
set -u
OUT=gpurun_out/ab
mkdir -p "$OUT"
STEPS=${STEPS:-5}
run_variant() {
  local tag=$1; shift
  echo "=== $tag: $*" | tee -a "$OUT/summary.txt"
  ( export "$@" DADA2B_AB_TAG="$tag"
    timeout 420 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_ties.py -x -q > "$OUT/$tag.pytest.log" 2>&1
    echo "pytest rc=$?" | tee -a "$OUT/summary.txt"
    timeout 300 python bench.py --gpus 1 --steps "$STEPS" --warmup 3 > "$OUT/$tag.bench.json" 2> "$OUT/$tag.bench.err"
    echo "bench rc=$?" | tee -a "$OUT/summary.txt"
    python - "$OUT/$tag.bench.json" <<'PY' | tee -a "$OUT/summary.txt"
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  value %.0f %s  ms/step %.2f  e2e %.0f  launches %s  kernels %s  parity %s" % (
        d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"),
        {k: round(v, 2) for k, v in (d.get("kernel_ms") or {}).items()}, d.get("parity")))
except Exception as e:
    print("  no bench line:", e)
PY
  )
}
run_variant base DADA2B_NONE=1
run_variant nwfwd2 DADA2B_NWFWD_V2=1
run_variant small16x4 DADA2B_NWFWD_SMALL=1
run_variant nwfwd2_small16x4 DADA2B_NWFWD_V2=1 DADA2B_NWFWD_SMALL=1
run_variant fused DADA2B_FUSED_TAIL=1
run_variant pivot DADA2B_PIVOT=1
run_variant twophase DADA2B_TWOPHASE=1
run_variant nwfwd2_twophase DADA2B_NWFWD_V2=1 DADA2B_TWOPHASE=1
run_variant nwfwd2_twophase_bound16 DADA2B_NWFWD_V2=1 DADA2B_TWOPHASE=1 DADA2B_BOUND16=1
run_variant all DADA2B_NWFWD_V2=1 DADA2B_FUSED_TAIL=1 DADA2B_PIVOT=1
run_variant all_np2 DADA2B_NWFWD_V2=1 DADA2B_FUSED_TAIL=1 DADA2B_PIVOT=1 DADA2B_NP=2
# multi-GPU variants (only when the box has more than one GPU): replicated control vs fused tail vs owner mode
NG=$(python -c "import torch; print(torch.cuda.device_count())" 2>/dev/null || echo 1)
if [ "${NG:-1}" -gt 1 ]; then
  for v in "mg_base DADA2B_NONE=1" "mg_fused DADA2B_FUSED_TAIL=1" "mg_owner DADA2B_FUSED_TAIL=1 DADA2B_OWNER=1" "mg_owner_all DADA2B_FUSED_TAIL=1 DADA2B_OWNER=1 DADA2B_NWFWD_V2=1 DADA2B_PIVOT=1"; do
    set -- $v; tag=$1; shift
    echo "=== $tag ($NG GPUs): $*" | tee -a "$OUT/summary.txt"
    ( export "$@"
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 --master-port 29541 tools/run_sharded.py > "$OUT/$tag.parity.log" 2>&1
      echo "sharded parity rc=$? ($(grep -c 'PARITY OK' "$OUT/$tag.parity.log") PARITY OK, $(grep -c MISMATCH "$OUT/$tag.parity.log") MISMATCH)" | tee -a "$OUT/summary.txt"
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NG" --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus "$NG" --steps "$STEPS" --warmup 3 > "$OUT/$tag.bench.json" 2> "$OUT/$tag.bench.err"
      echo "bench rc=$?" | tee -a "$OUT/summary.txt"
      tail -n 1 "$OUT/$tag.bench.json" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  value %.0f %s  ms/step %.2f  e2e %.0f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value']))" 2>/dev/null | tee -a "$OUT/summary.txt"
    )
  done
fi
cat "$OUT/summary.txt"
