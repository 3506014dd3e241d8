"""Size-independent invariants of a dada_uniques() result (used at BASELINE sizes on the GPU and, to pin the
checker itself, on the CPU oracle at small sizes).  Each follows from the reference's output builders
(/root/reference/src/error.cpp:9-300, Rmain.cpp:239-279)."""
import numpy as np

NA_INT = -2147483648


def check_invariants(seqs, abund, res, omegaA=1e-40, has_priors=False):
    abund = np.asarray(abund, dtype=np.int64)
    nraw = len(seqs)
    cl = res["clustering"]
    nclust = len(cl["sequence"])
    m = np.asarray(res["map"])
    pv = np.asarray(res["pval"])
    assert m.shape == (nraw,) and pv.shape == (nraw,)
    ok = m != NA_INT
    assert np.all((m[ok] >= 1) & (m[ok] <= nclust))
    # abundance / nunq are sums over the corrected raws of each partition (error.cpp:53-66; map at Rmain.cpp:269-279)
    ab_from_map = np.bincount(m[ok] - 1, weights=abund[ok], minlength=nclust).astype(np.int64)
    nunq_from_map = np.bincount(m[ok] - 1, minlength=nclust)
    assert np.array_equal(ab_from_map, np.asarray(cl["abundance"], dtype=np.int64))
    assert np.array_equal(nunq_from_map, np.asarray(cl["nunq"], dtype=np.int64))
    # n0 + n1 <= abundance ; every representative sequence is one of the uniques and maps to its own partition
    n0, n1 = np.asarray(cl["n0"], dtype=np.int64), np.asarray(cl["n1"], dtype=np.int64)
    assert np.all(n0 >= 0) and np.all(n1 >= 0) and np.all(n0 + n1 <= ab_from_map)
    index_of = {}
    for i, s in enumerate(seqs):
        index_of.setdefault(s, []).append(i)
    unique_input = len(index_of) == nraw
    for i, s in enumerate(cl["sequence"]):
        assert s in index_of
        cands = [r for r in index_of[s] if m[r] == i + 1 and pv[r] == 1.0]   # the centre: correct, p = 1 (Rmain.cpp:244-245)
        assert cands
        assert n0[i] >= min(abund[r] for r in cands)          # the centre itself has zero substitutions
    assert cl["sequence"][0] == seqs[int(np.argmax(abund))]   # initial centre = first most abundant raw
    if unique_input:
        assert len(set(cl["sequence"])) == nclust
    # p-values are probabilities; NA rows only where pval < omegaC
    assert np.all((pv >= 0) & (pv <= 1.0 + 1e-12))
    # births: row 0 is NA, later rows were born from an earlier partition with a significant p-value
    bf = np.asarray(cl["birth_from"])
    assert bf[0] == NA_INT and np.isnan(cl["birth_pval"][0])
    if nclust > 1:
        assert np.all((bf[1:] >= 1) & (bf[1:] <= np.arange(1, nclust)))
        if not has_priors:
            assert np.all(np.asarray(cl["birth_pval"])[1:] < omegaA)
        assert np.all(np.asarray(cl["birth_ham"])[1:] >= 1)
    # birth_subs rows reference valid partitions / positions
    bs = res["birth_subs"]
    if len(bs["pos"]):
        assert np.all((np.asarray(bs["clust"]) >= 2) & (np.asarray(bs["clust"]) <= nclust))
        assert np.all(np.asarray(bs["pos"]) >= 1)
        assert all(a != b for a, b in zip(bs["ref"], bs["sub"]))
        counts = np.bincount(np.asarray(bs["clust"]) - 1, minlength=nclust)
        assert np.array_equal(counts[1:] > 0, np.asarray(cl["birth_ham"])[1:] >= 0)
    # transition tallies: non-negative, self transitions dominate, total = corrected reads x aligned positions
    sq = np.asarray(res["subqual"], dtype=np.int64)
    assert sq.shape[0] == 16 and np.all(sq >= 0)
    lens = np.array([len(s) for s in seqs])
    tot = int(sq.sum())
    assert tot <= int((abund[ok] * lens[ok]).sum())
    assert sq[[0, 5, 10, 15]].sum() > 0.9 * tot
    # cluster quality profile: NA exactly beyond each centre's length, otherwise within the quality range
    cq = np.asarray(res["clusterquals"])
    assert cq.shape[1] == nclust
    for i, s in enumerate(cl["sequence"]):
        col = cq[:, i]
        assert np.all(np.isnan(col[len(s):])) and not np.any(np.isnan(col[:len(s)]))
        assert np.all((col[:len(s)] >= 0) & (col[:len(s)] <= 93))
