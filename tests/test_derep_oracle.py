"""Dereplication oracle (oracle/derep.py, PARITY UNPINNED against R: see its header): what can be pinned is pinned here --
on the reference's own sam1F fixture it reproduces the committed config-1 input of dada() (896 uniques of 1 500 reads,
SURVEY.md 8c/8d) and its chunked form only reorders exact abundance ties."""
import numpy as np

from oracle import derep as O
from tests import derep_cases as D


def test_oracle_reproduces_the_config1_input_from_the_reference_fixture():
    seqs, quals = D.sam1F_reads()
    r = O.derep_reads(seqs, quals)
    z = np.load(D.os.path.join(D.GOLDEN, "config1_sam1F_input.npz"))
    assert len(r["uniques"]) == 896 and int(r["abundances"].sum()) == 1500
    assert r["uniques"] == z["seqs"].tolist() and np.array_equal(r["abundances"], z["abund"]) and np.array_equal(r["quals"], z["quals"], equal_nan=True)
    assert np.all(np.diff(r["abundances"]) <= 0) and r["map"].min() == 1 and r["map"].max() == 896
    assert [r["uniques"][m - 1] for m in r["map"][:50]] == seqs[:50]


def test_chunking_only_reorders_ties_and_zero_length_reads_are_ignored():
    seqs, quals = D.synthetic(2000, seed=3, zero_len=4)
    a, b = O.derep_reads(seqs, quals), O.derep_reads(seqs, quals, n=300)
    assert sorted(zip(a["uniques"], a["abundances"])) == sorted(zip(b["uniques"], b["abundances"]))
    assert np.array_equal(a["abundances"], b["abundances"]) and a["uniques"] != b["uniques"]
    na = np.iinfo(np.int32).min
    assert (a["map"] == na).sum() == 4 and all(seqs[i] == "" for i in np.nonzero(a["map"] == na)[0])
    for u, ab in zip(a["uniques"][:20], a["abundances"][:20]):
        assert seqs.count(u) == ab


def test_combine_dereps_pools_like_the_reference():
    """combineDereps2 (R/multiSample.R:165-203): pooling the per-sample dereps of a split read set gives the uniques, pooled
    abundances and weighted quality means of dereplicating the whole set at once (the order differs only inside abundance
    ties: first-seen order there); the translated maps point every read at its own sequence."""
    import numpy as np
    from oracle import derep as od
    from dada2_b200.derep import combineDereps2
    rng = np.random.default_rng(4)
    base = ["".join(rng.choice(list("ACGT"), 40)) for _ in range(12)]
    reads = [base[int(i)] for i in rng.integers(0, 12, 300)]
    quals = [rng.integers(2, 41, 40).astype(np.uint8) for _ in reads]
    parts = [od.derep_reads(reads[a:b], quals[a:b]) for a, b in ((0, 90), (90, 210), (210, 300))]
    whole = od.derep_reads(reads, quals)
    pool = combineDereps2(parts)
    assert sorted(pool["uniques"]) == sorted(whole["uniques"])
    w = dict(zip(whole["uniques"], whole["abundances"]))
    assert all(w[s] == a for s, a in zip(pool["uniques"], pool["abundances"]))
    assert list(pool["abundances"]) == sorted(pool["abundances"], reverse=True)
    wq = dict(zip(whole["uniques"], whole["quals"]))
    for s, q in zip(pool["uniques"], pool["quals"]):
        assert np.allclose(q, wq[s], rtol=1e-12, atol=0)
    assert [pool["uniques"][m - 1] for m in pool["map"]] == reads
