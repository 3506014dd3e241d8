"""Dereplication kernels (dd_derep.cu) on hardware through the C-ABI of include/dada2b_derep.h: the
reference's sam1F fixture (must reproduce the committed config-1 input of dada()), chunked and synthetic inputs against
the oracle, then 2e5 reads timed.  Written after round 1's GPU budget was spent (emulator-validated only,
tests/test_emu_derep.py): a plain hardware gate since round 2 (XPASSed on a B200 in round 1); subprocess with a timeout."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys, time
    sys.path.insert(0, %r)
    import numpy as np
    from tests import derep_cases as D
    from dada2_b200 import derep
    D.check_all(D.product_fn, sizes=(3000, 20000))
    print("oracle cases ok", flush=True)
    from oracle import derep as O
    s, q = D.synthetic(200000, seed=77, L=250, nvar=400)
    t0 = time.time(); want = O.derep_reads(s, q); t1 = time.time() - t0
    got = derep.derep_reads(s, q, return_stats=True)
    D.assert_same(got, want, "2e5 reads")
    st = got["stats"]
    print("2e5 x 250 nt reads -> %%d uniques: gpu %%.1f ms (device %%.1f, sort %%.1f, %%d launches), oracle %%.1f s" %% (
        len(got["uniques"]), st["ms_total"], st["ms_device"], st["ms_sort"], st["gpu_launches"], t1), flush=True)
    import dada2_b200
    from tests import cases
    seqs, quals = D.sam1F_reads()                       # derep -> dada without the host round trip == the two-call form
    d, res = derep.derep_reads(seqs, quals, resident=True)
    a = res.run(cases.tperr1()); b = dada2_b200.dada_uniques(d["uniques"], d["abundances"], None, cases.tperr1(), d["quals"])
    cases.assert_same(a, b, rtol=0, label="derep_resident")
    from tests.test_oracle import load_golden
    cases.assert_same(a, load_golden("config1"), rtol=1e-10, label="derep_resident vs config-1 golden")
    res.close()
    print("DEREP OK")
''') % ROOT


def test_derep_kernels_match_oracle_and_config1_input():
    out = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "DEREP OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
