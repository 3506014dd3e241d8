"""Pins the invariant checker (tests/properties.py) on the CPU oracle so that the GPU run at BASELINE
sizes can rely on it."""
import pytest

from oracle import port
from tests import cases
from tests.properties import check_invariants


@pytest.mark.parametrize("name", ["syn800_default", "syn800_nogreedy", "syn700_ragged", "syn800_priors", "syn800_omegaC"])
def test_invariants_hold_on_oracle(name):
    seqs, ab, pri, err, q, opts = cases.build_case(name)
    opts.setdefault("homo_gap", opts.get("gap", -8))
    res = port.dada_uniques(seqs, ab, pri, err, q, **opts)
    check_invariants(seqs, ab, res, omegaA=opts.get("omegaA", 1e-40), has_priors=pri is not None)


def test_invariants_hold_on_config1():
    seqs, ab, q = cases.load_config1()
    res = port.dada_uniques(seqs, ab, None, cases.tperr1(), q)
    check_invariants(seqs, ab, res)
