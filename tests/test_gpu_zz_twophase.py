"""Experimental two-phase loop NW (DADA2B_TWOPHASE=1, off by default).  Written after round 1's GPU
budget was spent, so its first execution on hardware is this test: xfail(strict=False), isolated in a subprocess.
The default path is unaffected (its kernels are byte-identical in SASS to the validated build)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys
    sys.path.insert(0, %r)
    import numpy as np
    from tests import cases
    from tests.test_oracle import load_golden
    import dada2_b200
    for name in ("syn2000_default", "syn800_nogreedy", "syn700_ragged", "syn800_scores", "syn800_priors", "syn500_usequals0"):
        seqs, ab, pri, err, q, opts = cases.build_case(name)
        got = dada2_b200.dada_uniques(seqs, ab, pri, err, q, **opts)
        want = load_golden(name)
        pb = None
        if pri is not None:
            pb = np.zeros(len(want["clustering"]["sequence"]), dtype=bool)
            pb[1:] = want["clustering"]["birth_pval"][1:] >= opts.get("omegaA", 1e-40)
        cases.assert_same(got, want, rtol=1e-10, prior_born=pb, label=name)
    print("TWOPHASE OK")
''') % ROOT


@pytest.mark.xfail(strict=False, reason="experimental path, first run on hardware happens at round end")
def test_two_phase_loop_nw_matches_goldens():
    env = dict(os.environ, DADA2B_TWOPHASE="1")
    out = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "TWOPHASE OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
