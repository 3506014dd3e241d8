"""One-off: run a large synthetic workload on the GPU (and optionally the CPU reference) and compare."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
from tools import synth
from tests import cases
import dada2_b200
n = int(sys.argv[1]); do_cpu = len(sys.argv) > 2 and sys.argv[2] == "cpu"
t = time.time(); seqs, ab, q, truth = synth.illumina(n, seed=12345); print("gen s", round(time.time() - t, 1), "reads", int(ab.sum()), flush=True)
err = cases.tperr1()
call = dada2_b200.PackedCall(seqs, ab, None, err, q)
for i in range(3):
    r, ms = call.run(unpack=(i == 2))
    st = r["stats"]
    print("one-shot wall ms", round(ms, 1), {k: (round(v, 1) if isinstance(v, float) else v) for k, v in st.items()}, flush=True)
print("nclust", len(r["clustering"]["sequence"]))
if do_cpu:
    import os
    from oracle import ref
    ref.set_threads(os.cpu_count())
    t = time.time(); c = ref.dada_uniques(seqs, ab, None, err, q, multithread=True); dt = time.time() - t
    print("cpu ref s", round(dt, 1), "threads", os.cpu_count(), "uniques/s", round(n / dt))
    try:
        cases.assert_same(r, c, rtol=1e-10, label="big"); print("PARITY OK")
    except AssertionError as e:
        print("PARITY MISMATCH", e)
