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
