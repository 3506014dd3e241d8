#!/bin/bash
# Round 2, GPU call 25 (the last 40 s of the budget): the upload pipeline / exact rounding / representative rows read back from the
# device, on hardware: the new -m gpu test, config 1, the edge-case table (incl. half-integer qualities).
set -u
OUT=gpurun_out/r2c25
mkdir -p "$OUT"
timeout 31 python -m pytest -q -p no:cacheprovider tests/test_gpu_parity.py::test_upload_pipeline_many_blocks tests/test_gpu_parity.py::test_config1_bit_identical tests/test_gpu_zzz_edge.py > "$OUT/pytest.log" 2>&1; echo "pytest rc=$?"; tail -n 5 "$OUT/pytest.log"
