"""Bimera detection kernels (dd_bimera.cu) on hardware, through the C-ABI of include/dada2b_bimera.h,
against the goldens produced by the reference's own chimera.cpp and, at a larger size, against the CPU oracle.
Written after round 1's GPU budget was spent: validated on the host SIMT emulator only (tests/test_emu_bimera.py), so the
validated on a B200 in round 1 (XPASS) and a plain hardware gate since round 2; run in a subprocess with a timeout."""
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
    from tests import bimera_cases as B
    from dada2_b200 import bimera
    B.check_pairs(B.product_pair_fn)
    B.check_pairs_vs_oracle(B.product_pair_fn, npairs=400)
    for name in B.table_names():
        B.check_table(name, B.product_table_fn)
        B.check_is_bimera(name, B.product_denovo_fn)
    print("goldens ok", flush=True)
    # larger table against the CPU oracle (restatement of chimera.cpp, pinned on the same goldens)
    from oracle import port
    from tools import synth
    import os
    seqs, mat = synth.bimera_table(1500, 12, seed=9, lenvar=6) if os.environ.get("DADA2B_BIMFWD", "0") == "0" else synth.bimera_table(700, 8, seed=9, lenvar=6)
    for o in (dict(), dict(allow_one_off=True)):
        t0 = time.time(); want = port.table_bimera(mat, seqs, **o); t1 = time.time()
        got = bimera.C_table_bimera2(mat, seqs, return_stats=True, **o)
        assert np.array_equal(got["nflag"], want[0]) and np.array_equal(got["nsam"], want[1]), o
        st = got["stats"]
        print("table %%dx%%d" %% (len(seqs), mat.shape[0]), o, "pairs", st["n_pairs"], "gpu %%.1f ms (align %%.1f ms), oracle %%.1f s" %% (st["ms_total"], st["ms_k_align"], t1 - t0), flush=True)
    print("BIMERA OK")
''') % ROOT


@pytest.mark.parametrize("variant", ["traceback", "register", "simd16"])
def test_bimera_kernels_match_reference_goldens_and_oracle(variant):
    env = dict(os.environ)
    env["DADA2B_BIMFWD"] = "0"                      # k_bim_align (warp-per-pair traceback): the kernel of last resort
    if variant == "register":                       # dd_bimfwd.cu (register-resident wavefront + per-pair traceback)
        env["DADA2B_BIMFWD"] = "1"
    if variant == "simd16":                         # dd_bimfwd16.cu (two jobs per lane group on the 16-bit SIMD datapath)
        env["DADA2B_BIMFWD"] = "2"
    out = subprocess.run([sys.executable, "-c", SCRIPT], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "BIMERA OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
