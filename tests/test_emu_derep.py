"""The dereplication kernels (dd_derep.cu: key packing, stable LSD radix sort, segment heads, abundance ordering, quality
sums, map) on the host SIMT emulator against the oracle and the committed config-1 input -- CPU suite."""
import os
import platform
import sys

import numpy as np
import pytest

from tests import derep_cases as D

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is x86-64 only")


@pytest.fixture()
def emu(monkeypatch):
    import build_emu
    import dada2_b200.api as api
    import dada2_b200.derep as dr
    lib = build_emu.build()
    monkeypatch.setattr(api, "_LIBPATH", lib)
    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setattr(dr, "_BOUND", False)
    yield lib


def test_emu_derep_matches_oracle_and_config1_input(emu):
    D.check_all(D.product_fn, sizes=(3000, 5000) if os.environ.get("DADA2B_EMU_FULL") else (2500,))


def test_emu_derep_feeds_dada_and_corners(emu):
    import dada2_b200
    from dada2_b200 import derep, Dada2bError
    from tests import cases
    from tests.test_oracle import load_golden
    seqs, quals = D.sam1F_reads()
    d = derep.derep_reads(seqs, quals)
    got = dada2_b200.dada_uniques(d["uniques"], d["abundances"], None, cases.tperr1(), d["quals"])       # derep -> dada, both on the "device"
    cases.assert_same(got, load_golden("config1"), rtol=1e-10, label="derep+dada config1")
    # the same without the host round trip: uniques stay packed on the device, dada() runs on that context
    d2, res = derep.derep_reads(seqs, quals, resident=True, want_quals=False)
    assert d2["quals"] is None and d2["uniques"] == d["uniques"] and np.array_equal(d2["abundances"], d["abundances"])
    cases.assert_same(res.run(cases.tperr1()), got, rtol=0, label="derep_resident + run_resident")
    res.close()
    s2, q2 = D.synthetic(900, seed=5, L=70, nvar=20, zero_len=2)            # ragged lengths, fractional means, NA map entries
    dd, rr = derep.derep_reads(s2, q2, n=300, resident=True)
    err = np.full((16, 45), 0.01); err[[0, 5, 10, 15]] = 0.97
    cases.assert_same(rr.run(err), dada2_b200.dada_uniques(dd["uniques"], dd["abundances"], None, err, dd["quals"]), rtol=0, label="resident ragged")
    rr.close()
    with pytest.raises(Dada2bError, match="A/C/G/T"):
        derep.derep_reads(["ACGTN", "ACGTA"], [np.full(5, 30, np.uint8)] * 2)
    with pytest.raises(Dada2bError, match="Only zero-length"):
        derep.derep_reads(["", ""], [np.zeros(0, np.uint8)] * 2)
    one = derep.derep_reads(["ACGTACGT"], [np.arange(8, dtype=np.uint8)])
    assert one["uniques"] == ["ACGTACGT"] and one["abundances"].tolist() == [1] and one["map"].tolist() == [1]
    assert np.array_equal(one["quals"], np.arange(8, dtype=float)[None, :])
