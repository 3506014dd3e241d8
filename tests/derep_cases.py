"""Shared dereplication parity checks: the product against oracle/derep.py (restatement of derepFastq /
qtables2; PARITY UNPINNED against R itself, pinned on the reference's sam1F fixture) and against the committed config-1 input."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sam1F_reads():
    z = np.load(os.path.join(GOLDEN, "sam1F_reads.npz"))
    seqs = z["seqs"].tolist()
    return seqs, [q for q in z["quals"]]


def synthetic(nreads, seed, L=60, nvar=25, ragged=True, zero_len=0):
    """Reads with heavy duplication, several lengths, prefix relations and exact abundance ties."""
    rng = np.random.default_rng(seed)
    base = ["".join("ACGT"[i] for i in rng.integers(0, 4, L)) for _ in range(nvar)]
    if ragged:
        base += [b[:L - int(k)] for b, k in zip(base[:nvar // 2], rng.integers(1, 9, nvar // 2))]      # proper prefixes of other uniques
    w = 1.0 / np.arange(1, len(base) + 1)
    pick = rng.choice(len(base), nreads, p=w / w.sum())
    seqs, quals = [], []
    for v in pick:
        s = list(base[v])
        if rng.random() < 0.3:
            p = int(rng.integers(0, len(s))); s[p] = "ACGT"[int(rng.integers(0, 4))]
        seqs.append("".join(s))
        quals.append(rng.integers(2, 42, len(s)).astype(np.uint8))
    for k in range(zero_len):
        p = int(rng.integers(0, len(seqs))); seqs.insert(p, ""); quals.insert(p, np.zeros(0, np.uint8))
    return seqs, quals


def assert_same(got, want, label=""):
    assert got["uniques"] == want["uniques"], label + " uniques / order"
    assert np.array_equal(got["abundances"], want["abundances"]), label + " abundances"
    assert np.array_equal(got["map"], want["map"]), label + " map"
    a, b = np.asarray(got["quals"]), np.asarray(want["quals"])
    assert a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)), label + " NA pattern"
    assert np.array_equal(a[~np.isnan(b)], b[~np.isnan(b)]), label + " mean qualities (bit-exact: integer sums / count)"


def check_all(derep_fn, sizes=(3000,)):
    """derep_fn(seqs, quals, n) -> dict like oracle.derep.derep_reads"""
    from oracle import derep as O
    seqs, quals = sam1F_reads()
    got = derep_fn(seqs, quals, 1000000)
    assert_same(got, O.derep_reads(seqs, quals), "sam1F")
    z = np.load(os.path.join(GOLDEN, "config1_sam1F_input.npz"))                 # the committed config-1 input of dada()
    assert got["uniques"] == z["seqs"].tolist() and np.array_equal(got["abundances"], z["abund"])
    assert np.array_equal(got["quals"], z["quals"], equal_nan=True)
    assert_same(derep_fn(seqs, quals, 400), O.derep_reads(seqs, quals, n=400), "sam1F chunks of 400")
    for k, nreads in enumerate(sizes):
        s, q = synthetic(nreads, seed=40 + k, zero_len=3 * (k % 2))
        for n in (1000000, 257):
            assert_same(derep_fn(s, q, n), O.derep_reads(s, q, n=n), "synthetic %d n=%d" % (nreads, n))


def product_fn(seqs, quals, n):
    from dada2_b200 import derep
    return derep.derep_reads(seqs, quals, n=n)
