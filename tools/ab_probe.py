"""A/B of library builds on ONE box (host speed differs between GPU boxes): resident passes of the same workload with each given
libdada2b*.so in turn (one process, one data set); prints the median host phases and kernel-event sums per build.
  python tools/ab_probe.py <n_uniques> <passes> lib1.so lib2.so ..."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import dada2_b200.api as api
import dada2_b200
from tools import synth
from tests import cases

n, reps, libs = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3:]
seqs, ab, q, _ = synth.illumina(n, seed=12345)
err = cases.tperr1()
keys = ["ms_setup", "ms_loop", "ms_final", "ms_total", "ms_device", "ms_k_classify", "ms_k_prescreen", "ms_k_align_nw", "ms_k_nw_bound", "ms_k_nw_exact", "ms_k_align_gl",
        "ms_k_align_final", "ms_k_tail"]
for rnd in range(2):
    for lib in libs:
        api._LIBPATH = os.path.abspath(lib); api._LIB = None
        res = dada2_b200.Resident(seqs, ab, None, q)
        for _ in range(3):
            res.run(err)
        rows = [res.run(err)["stats"] for _ in range(reps)]
        del res
        d = {k: float(np.median([r[k] for r in rows])) for k in keys if k in rows[0]}
        timed = d.get("ms_k_classify", 0) + d.get("ms_k_align_nw", 0) + d.get("ms_k_align_gl", 0) + d.get("ms_k_tail", 0)
        print("%-20s round %d: total %.1f  setup %.1f loop %.1f final %.1f | screen %.1f (pre %.1f) nw %.1f (bound %.1f exact %.1f) gl %.1f tail %.1f | loop - timed kernels %.1f" % (
            os.path.basename(lib), rnd, d["ms_total"], d["ms_setup"], d["ms_loop"], d["ms_final"], d["ms_k_classify"], d.get("ms_k_prescreen", 0), d["ms_k_align_nw"],
            d.get("ms_k_nw_bound", 0), d.get("ms_k_nw_exact", 0), d["ms_k_align_gl"], d["ms_k_tail"], d["ms_loop"] - timed), flush=True)
