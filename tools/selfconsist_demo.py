"""BASELINE configs[2] flavour: the selfConsist error-learning loop of dada() (R/dada.R:256-391) driven from Python
over ONE resident upload (dada2b_upload once, dada2b_run_resident per iteration with a new error matrix).

The refit function of the reference is R's loess (R/errorModels.R:28-67: out of scope); the stand-in
below is a smoothed maximum-likelihood estimate per (transition, quality) from the `$subqual` tallies, enough to
exercise the loop structure: iteration 0 with an all-ones matrix and MAX_CLUST=1 (R/dada.R:297-299,342), then
refits until the matrix repeats or MAX_CONSIST=10.

    python tools/selfconsist_demo.py 100000
"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import dada2_b200
from tools import synth


def refit(trans, pseudo=1.0):
    """16 x Q transition counts -> 16 x Q error rates (rows sum to 1 per source nucleotide and quality)."""
    t = np.asarray(trans, dtype=np.float64) + pseudo
    err = np.zeros_like(t)
    for a in range(4):
        tot = t[4 * a:4 * a + 4].sum(axis=0)
        err[4 * a:4 * a + 4] = t[4 * a:4 * a + 4] / tot
    # R/dada.R:385-388: self-transitions of the initial estimate are forced to 1.0 only in iteration 0
    return err


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    seqs, ab, q, _ = synth.illumina(n, seed=12345)
    res = dada2_b200.Resident(seqs, ab, None, q)
    t0 = time.perf_counter()
    out = res.run(np.ones((16, 41)), max_clust=1)                 # initializeErr pass
    err = refit(out["subqual"])
    for a in range(4):
        err[5 * a] = 1.0
    seen = [err]
    times = [time.perf_counter() - t0]
    for it in range(10):                                           # MAX_CONSIST
        t1 = time.perf_counter()
        out = res.run(err)
        times.append(time.perf_counter() - t1)
        new = refit(out["subqual"])
        print("selfConsist step %d: %d partitions, %.1f ms" % (it + 1, len(out["clustering"]["sequence"]), 1e3 * times[-1]))
        if any(np.array_equal(new, e) for e in seen):              # identical(err, previous) -> converged (R/dada.R:391)
            break
        seen.append(new)
        err = new
    print("passes %d, total %.1f ms, %.0f uniques/s per pass" % (len(times), 1e3 * sum(times), n * len(times) / sum(times)))
    res.close()


if __name__ == "__main__":
    main()
