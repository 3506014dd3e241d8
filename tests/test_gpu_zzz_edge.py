"""Corners of the interface on the GPU (one raw, minimum length, ragged lengths 6..300, unbanded / over-wide bands,
singletons, omegaA=1, unsorted and zero abundances, extreme scores, all priors, N, homopolymer indels): results and
errors must equal the CPU oracle's.
(Named zzz: these in-process cases were written after round 1's GPU budget was spent, so they run after every other GPU test --
with `-x` a first-run failure here cannot hide the results of the files before it.)"""
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
CASES = cases.edge_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_edge_case_matches_oracle(case):
    import dada2_b200
    cases.check_edge_case(case, dada2_b200.dada_uniques)
