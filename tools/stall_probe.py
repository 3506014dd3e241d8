"""Diagnostic: N resident passes with the library's verbose timeline, to find which host call absorbs the periodic ~350 ms stall."""
import sys, time, os
sys.path.insert(0, '.')
os.environ["DADA2B_VERBOSE"] = "1"
import numpy as np
from tools import synth
from tests import cases
import dada2_b200
n = int(sys.argv[1]); reps = int(sys.argv[2])
seqs, ab, q, _ = synth.illumina(n, seed=12345)
err = cases.tperr1()
res = dada2_b200.Resident(seqs, ab, None, q)
for i in range(reps):
    t0 = time.perf_counter()
    r = res.run(err)
    print("PASS %d wall %.1f ms setup %.1f loop %.1f final %.1f" % (i, (time.perf_counter() - t0) * 1e3, r["stats"]["ms_setup"], r["stats"]["ms_loop"], r["stats"]["ms_final"]), file=sys.stderr, flush=True)
