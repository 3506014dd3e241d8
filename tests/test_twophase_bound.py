"""CPU evidence for the (experimental) two-phase loop NW of DESIGN.md 9.3: for every real alignment the
reference performs, lambda <= S_r * rho_r^nsubs, so `bound * total_reads <= E_minmax` never drops a comparison
the reference would store (cluster.cpp:192).  Uses the oracle's alignment trace (PORT_TRACE)."""
import os
import subprocess
import sys
import tempfile
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys
    sys.path.insert(0, %r)
    from oracle import port
    from tests import cases
    seqs, ab, pri, err, q, opts = cases.build_case("syn2000_default")
    port.dada_uniques(seqs, ab, pri, err, q, **opts)
''') % ROOT


def test_bound_never_drops_a_stored_comparison():
    from tests import cases
    with tempfile.TemporaryDirectory() as d:
        trace = os.path.join(d, "trace.bin")
        env = dict(os.environ, PORT_TRACE=trace)
        subprocess.check_call([sys.executable, "-c", SCRIPT], env=env, cwd=ROOT)
        rec = np.fromfile(trace, dtype=np.dtype([("i", "u4"), ("index", "u4"), ("nsubs", "u4"), ("stored", "u4"),
                                                 ("lam", "f8"), ("emm", "f8"), ("breads", "f8")]))
    seqs, ab, pri, err, q, opts = cases.build_case("syn2000_default")
    ncol = err.shape[1]
    ratio = np.zeros((4, ncol))
    for b in range(4):
        for qq in range(ncol):
            ratio[b, qq] = max(err[4 * a + b, qq] for a in range(4) if a != b) / err[5 * b, qq]
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    qi = np.floor(np.nan_to_num(q) + 0.5).astype(int)      # C round(): halves away from zero (containers.cpp:34)
    S = np.ones(len(seqs)); RHO = np.zeros(len(seqs))
    for r in np.unique(rec["index"]):
        s = np.array([code[c] for c in seqs[r]])
        S[r] = np.prod(err[5 * s, qi[r, :len(s)]])
        RHO[r] = ratio[s, qi[r, :len(s)]].max()
    idx = rec["index"]
    bound = S[idx] * RHO[idx] ** rec["nsubs"]
    assert np.all(rec["lam"] <= bound * (1 + 1e-12))                      # the inequality itself
    skip = (bound * rec["breads"] * (1 + 1e-9) <= rec["emm"]) & ~(bound < 1e-280)
    assert not np.any(skip & (rec["stored"] == 1))                          # no comparison the reference stores is dropped
    later = rec["i"] > 0
    assert later.sum() > 1000 and skip[later].mean() > 0.5                  # and the filter is worth having
