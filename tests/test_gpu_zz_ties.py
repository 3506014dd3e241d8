"""Large exact-tie sets in the b_bud scan (more candidates than the device reports inline): the host has to fetch the
whole set and apply the reference's (cluster, slot) scan order.  Checked against the CPU oracle."""
import pytest

from tests import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("opts", [dict(), dict(greedy=False), dict(max_clust=30)], ids=["default", "nogreedy", "maxclust30"])
def test_large_tie_sets_follow_scan_order(opts):
    import dada2_b200
    from oracle import port
    seqs, ab, q = cases.tie_case()
    err = cases.tperr1()
    got = dada2_b200.dada_uniques(seqs, ab, None, err, q, **opts)
    want = port.dada_uniques(seqs, ab, None, err, q, **dict(opts, homo_gap=-8))
    cases.assert_same(got, want, rtol=1e-10, label="ties " + str(opts))
    assert len(got["clustering"]["abundance"]) == (30 if "max_clust" in opts else 151)
