"""Generates tests/golden/merge.npz from the REFERENCE's own evaluate.cpp + nwalign_endsfree.cpp (oracle/_ref, compiled unmodified;
build container only): forward / rc(reverse) ASV pairs as mergePairs sees them (true overlaps of 12..200 nt, substitutions and
indels in the overlap, overhangs past the partner's start, unrelated reads) with C_eval_pair's counts and C_pair_consensus'
sequence for several option sets.  Run:  python tools/make_golden_merge.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref          # noqa: E402

OPTS = [dict(), dict(mismatch=-8, gap_p=-8), dict(trim_overhang=True), dict(mismatch=-8, gap_p=-8, homo_gap_p=-1),
        dict(band=16, trim_overhang=True), dict(match=5, mismatch=-4, gap_p=-8, band=-1)]


def corpus(seed=8, n=160):
    rng = np.random.default_rng(seed)

    def rs(L):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, L))

    def mut(s, k):
        s = list(s)
        for p in rng.choice(len(s), k, replace=False):
            s[p] = "ACGT"[("ACGT".index(s[p]) + 1 + rng.integers(0, 3)) % 4]
        return "".join(s)

    seqs, a, b = [], [], []
    for it in range(n):
        amp = rs(int(rng.integers(60, 460)))
        lf, lr = int(rng.integers(30, 251)), int(rng.integers(30, 251))
        f, r = amp[:lf], amp[max(0, len(amp) - lr):]
        mode = it % 6
        if mode == 1: r = mut(r, int(rng.integers(1, 4)))
        elif mode == 2:
            i = int(rng.integers(1, len(r) - 1)); r = r[:i] + rs(int(rng.integers(1, 3))) + r[i:]
        elif mode == 3: f = rs(int(rng.integers(1, 9))) + f          # forward read overhanging the reverse read's end after rc
        elif mode == 4: r = rs(len(r))
        elif mode == 5: f = f[:-1] + f[-1] * int(rng.integers(3, 7)); r = r[0] * 3 + r       # homopolymers at the junction
        seqs += [f, r]; a.append(len(seqs) - 2); b.append(len(seqs) - 1)
    return seqs, np.array(a, np.int32), np.array(b, np.int32)


def main():
    seqs, a, b = corpus()
    prefer = (1 + (np.arange(len(a)) % 2)).astype(np.int32)
    out = {"seqs": np.array(seqs), "s1": a, "s2": b, "prefer": prefer}
    for oi, o in enumerate(OPTS):
        cnt = np.zeros((len(a), 3), np.int32)
        cons = []
        for x, (i, j) in enumerate(zip(a, b)):
            r = ref.merge_pair(seqs[i], seqs[j], prefer=int(prefer[x]), **o)
            cnt[x] = [r["nmatch"], r["nmismatch"], r["nindel"]]; cons.append(r["sequence"])
        out["o%d_counts" % oi] = cnt; out["o%d_cons" % oi] = np.array(cons)
    out["meta"] = np.array(json.dumps({"opts": OPTS}))
    path = os.path.join(ROOT, "tests", "golden", "merge.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
