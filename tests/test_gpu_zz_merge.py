"""mergePairs' fused align / evaluate / consensus kernel (dd_merge.cu) on hardware, through the C-ABI of
include/dada2b_merge.h, against the goldens produced by the reference's own evaluate.cpp + nwalign_endsfree.cpp and, at a
larger size, against the CPU oracle.  Written after round 1's GPU budget was spent (emulator-validated only,
tests/test_emu_merge.py): a plain hardware gate since round 2 (XPASSed on a B200 in round 1); subprocess with a timeout."""
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
    from tests import merge_cases as M
    from dada2_b200 import merge
    M.check(M.product_fn)
    print("goldens ok", flush=True)
    from oracle import port
    rng = np.random.default_rng(31)
    seqs, a, b = [], [], []
    for it in range(3000):                      # 2 x 250 nt reads over 253..480 nt amplicons, like a V4 / V3-V4 run
        amp = "".join("ACGT"[i] for i in rng.integers(0, 4, int(rng.integers(253, 481))))
        f, r = amp[:250], amp[-250:]
        if it %% 3 == 1:
            p = int(rng.integers(0, 250)); r = r[:p] + "ACGT"[("ACGT".index(r[p]) + 1) %% 4] + r[p + 1:]
        seqs += [f, r]; a.append(len(seqs) - 2); b.append(len(seqs) - 1)
    pref = (1 + (np.arange(len(a)) %% 2)).astype(np.int32)
    for o in (dict(), dict(mismatch=-8, gap_p=-8, trim_overhang=True)):
        got = merge.merge_align(seqs, a, b, pref, return_stats=True, **o)
        t0 = time.time()
        for x in range(0, len(a), 7):
            w = port.merge_pair(seqs[a[x]], seqs[b[x]], prefer=int(pref[x]), **o)
            g = (int(got["nmatch"][x]), int(got["nmismatch"][x]), int(got["nindel"][x]), got["sequence"][x])
            assert g == (w["nmatch"], w["nmismatch"], w["nindel"], w["sequence"]), (o, x, g, w)
        st = got["stats"]
        print("3000 pairs", o, "gpu %%.1f ms (kernel %%.1f ms, %%.1f GCUPS), oracle %%.2f s for 1/7 of them" %% (
            st["ms_total"], st["ms_k_merge"], st["n_cells"] / 1e6 / max(st["ms_k_merge"], 1e-9), time.time() - t0), flush=True)
    print("MERGE OK")
''') % ROOT


def test_merge_kernel_matches_reference_goldens_and_oracle():
    out = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "MERGE OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
