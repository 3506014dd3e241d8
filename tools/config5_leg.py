"""BASELINE configs[4] flavour on ONE GPU, bounded for bench.py: PacBio-like ~1.5 kb uniques of variable length, BAND_SIZE = 32,
HOMOPOLYMER_GAP_PENALTY = -1 (nwalign_endsfree_homo, /root/reference/src/nwalign_endsfree.cpp:220-396), 94 quality columns.
Times resident passes through the C-ABI and diffs a subsample against the CPU restatement (the reference's scalar aligner
mallocs two (L+1)^2 matrices per pair: ~5 ms per pair, so only a few hundred uniques are affordable as a check).
Prints one line `C5LEG {json}`.      python tools/config5_leg.py [n_uniques=20000] [oracle_subsample=300]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    from tools import synth
    from tests import cases
    import dada2_b200
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    t0 = time.perf_counter()
    seqs, ab, q = synth.pacbio(n, L=1500, nvar=30, seed=5)
    gen_s = time.perf_counter() - t0
    err = synth.extend_err(cases.tperr1(), 94)
    opts = dict(band_size=32, vectorized_alignment=False, homo_gap=-1)
    res = dada2_b200.Resident(seqs, ab, None, q)
    ts, out = [], None
    for _ in range(3):
        t0 = time.perf_counter()
        out = res.run(err, **opts)
        ts.append((time.perf_counter() - t0) * 1e3)
    res.close()
    st = out["stats"]
    lens = np.array([len(s) for s in seqs])
    line = {"workload": "%d PacBio-like uniques, %d-%d nt (tools/synth.py pacbio seed 5), BAND_SIZE=32, HOMOPOLYMER_GAP_PENALTY=-1, 94 quality columns"
                        % (len(seqs), lens.min(), lens.max()),
            "ms_per_pass": [round(x, 1) for x in ts], "uniques_per_s": len(seqs) / (min(ts) / 1e3), "nclust": len(out["clustering"]["sequence"]),
            "nw_pairs": int(st["n_nw"]), "nw_cells": int(st["nw_cells"]),
            "nw_gcups": st["nw_cells"] / 1e9 / (max(st["ms_k_align_nw"] + st["ms_k_align_final"], 1e-9) / 1e3),
            "kernel_ms": {k: round(st[k], 2) for k in st if k.startswith("ms_k_")}, "gpu_launches": int(st["gpu_launches"]), "generator_s": round(gen_s, 1)}
    if nsub:
        from oracle import port
        s2, a2 = seqs[:nsub], ab[:nsub]
        q2 = q[:nsub, :max(len(x) for x in s2)]          # the quality matrix must be exactly as wide as the longest read of the call
        got = dada2_b200.dada_uniques(s2, a2, None, err, q2, **opts)
        t0 = time.perf_counter()
        want = port.dada_uniques(s2, a2, None, err, q2, **opts)
        dt = time.perf_counter() - t0
        line["cpu_port"] = {"subsample": nsub, "s": round(dt, 1), "uniques_per_s": nsub / dt, "cores": 1}
        try:
            cases.assert_same(got, want, rtol=1e-10, label="config5")
            line["parity_subsample"] = True
        except AssertionError as e:
            line["parity_subsample"] = "MISMATCH: " + str(e)[:200]
    print("C5LEG " + json.dumps(line))


if __name__ == "__main__":
    main()
