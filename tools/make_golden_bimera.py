"""Generates tests/golden/bimera.npz from the REFERENCE's own chimera.cpp (oracle/_ref, compiled unmodified from
/root/reference/src; build container only): a (query, parent) pair corpus with get_lr / get_ham_endsfree values for several
max_shift settings, and sequence tables with C_table_bimera2's nflag / nsam and C_is_bimera's verdicts for several option
sets.  Run:  python tools/make_golden_bimera.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref          # noqa: E402
from tools import synth         # noqa: E402

PAIR_SHIFTS = [16, 0, 3, 40, 300, -1]
TABLE_OPTS = [dict(), dict(allow_one_off=True), dict(allow_one_off=True, min_one_off_par_dist=2, min_fold=1.0, min_abund=1),
              dict(min_fold=0.5), dict(max_shift=4), dict(match=4, mismatch=-5, gap_p=-6, allow_one_off=True)]
TABLES = {"t120": dict(nseq=120, nsample=6, seed=0), "t160_ragged": dict(nseq=160, nsample=10, seed=1, lenvar=12),
          "t150_short": dict(nseq=150, nsample=3, seed=2, L=120), "t40_one_sample": dict(nseq=40, nsample=1, seed=3, L=60)}


def pair_corpus(seed=3, n=300):
    rng = np.random.default_rng(seed)

    def rs(L):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, L))

    def mut(s, k):
        s = list(s)
        for p in rng.choice(len(s), k, replace=False):
            s[p] = "ACGT"[("ACGT".index(s[p]) + 1 + rng.integers(0, 3)) % 4]
        return "".join(s)

    def indel(s):
        i = int(rng.integers(0, len(s)))
        return s[:i] + rs(int(rng.integers(1, 4))) + s[i:] if rng.random() < 0.5 else s[:i] + s[i + int(rng.integers(1, 4)):]

    seqs, q, p = [], [], []
    for it in range(n):
        L = int(rng.integers(20, 200)); a = rs(L); b = rs(int(rng.integers(20, 200)))
        mode = it % 6
        if mode == 0: x, y = a[:L // 2] + b[len(b) // 2:], a                      # true bimera vs its left parent
        elif mode == 1: x, y = mut(a, int(rng.integers(0, 5))), a                 # substitution variant
        elif mode == 2: x, y = indel(a), a                                        # internal indel
        elif mode == 3: x, y = a[int(rng.integers(0, 20)):], a[:len(a) - int(rng.integers(0, 15))]   # shifted ends
        elif mode == 4: x, y = a, b                                               # unrelated
        else: x, y = indel(mut(a[:L // 2] + b[len(b) // 2:], 1)), indel(a)
        if len(x) < 6 or len(y) < 6:
            continue
        seqs += [x, y]; q.append(len(seqs) - 2); p.append(len(seqs) - 1)
    return seqs, np.array(q, np.int32), np.array(p, np.int32)


def main():
    out = {}
    seqs, q, p = pair_corpus()
    out["pairs_seqs"] = np.array(seqs); out["pairs_q"] = q; out["pairs_p"] = p
    for ms in PAIR_SHIFTS:
        v = np.zeros((len(q), 5), np.int32)
        for n, (a, b) in enumerate(zip(q, p)):
            r = ref.bimera_pair(seqs[a], seqs[b], allow_one_off=True, max_shift=ms)
            v[n] = [r["left"], r["right"], r["left_oo"], r["right_oo"], r["ham"]]
        out["pairs_ms%d" % ms] = v
    for name, gen in TABLES.items():
        sq, mat = synth.bimera_table(**gen)
        out[name + "_seqs"] = np.array(sq); out[name + "_mat"] = mat
        tot = mat.sum(axis=0)
        for oi, o in enumerate(TABLE_OPTS):
            nflag, nsam = ref.table_bimera(mat, sq, **o)
            out["%s_o%d_nflag" % (name, oi)] = nflag; out["%s_o%d_nsam" % (name, oi)] = nsam
        for oo in (0, 1):                                   # isBimeraDenovo's parent rule (R/chimeras.R:127) + C_is_bimera
            v = np.zeros(len(sq), bool)
            for j in range(len(sq)):
                pars = [sq[k] for k in range(len(sq)) if tot[k] > 2 * tot[j] and tot[k] > 8]
                v[j] = len(pars) >= 2 and ref.is_bimera(sq[j], pars, allow_one_off=bool(oo))
            out["%s_isbim_oo%d" % (name, oo)] = v
    out["meta"] = np.array(json.dumps({"pair_shifts": PAIR_SHIFTS, "table_opts": TABLE_OPTS, "tables": TABLES}))
    path = os.path.join(ROOT, "tests", "golden", "bimera.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
