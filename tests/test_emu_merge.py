"""mergePairs' fused align / evaluate / consensus kernel (dd_merge.cu) on the host SIMT emulator against the reference
goldens -- CPU suite; same assertions as tests/test_gpu_zz_merge.py on the shorter pairs."""
import os
import platform
import sys

import numpy as np
import pytest

from tests import merge_cases as M

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is x86-64 only")


@pytest.fixture()
def emu(monkeypatch):
    import build_emu
    import dada2_b200.api as api
    import dada2_b200.merge as mg
    lib = build_emu.build()
    monkeypatch.setattr(api, "_LIBPATH", lib)
    monkeypatch.setattr(api, "_LIB", None)
    monkeypatch.setattr(mg, "_BOUND", False)
    yield lib


def test_emu_merge_matches_reference_goldens(emu):
    full = bool(os.environ.get("DADA2B_EMU_FULL"))
    M.check(M.product_fn, limit=None if full else 40, maxlen=None if full else 160)


def test_emu_merge_core_and_corners(emu):
    from dada2_b200 import merge, Dada2bError
    amp = "GTATCGGCTACCGCAAAAATAGTACCCTATTTACGCGGGATGTCCTAACGATCAGTTTCATGTAAGCCTAGCCTACAACGGCATTAACATGGCGTATACTTATTCC"
    sub = "A" if amp[50] != "A" else "C"
    F, R = [amp[:60], amp[:55]], [amp[40:], amp[40:50] + sub + amp[51:]]            # R[1]: one substitution inside the overlap
    r = merge.mergePairs_core(F, R, [0, 1, 0], [0, 0, 1], n0F=[5, 1], n0R=[2, 9])
    assert r["accept"].tolist() == [True, True, False] and r["sequence"][0] == amp and r["sequence"][1] == amp and r["sequence"][2] == ""
    assert r["nmatch"].tolist()[:2] == [20, 15] and r["prefer"].tolist() == [1, 2, 2]
    r8 = merge.mergePairs_core(F, R, [0], [1], n0F=[5, 1], n0R=[2, 9], maxMismatch=1)           # one mismatch allowed, reverse preferred
    assert r8["accept"].tolist() == [True] and r8["nmismatch"].tolist() == [1] and r8["nmatch"].tolist() == [19]
    assert r8["sequence"][0] == amp[:50] + sub + amp[51:]
    assert merge.merge_align(F + R, [], [], [])["sequence"] == []
    with pytest.raises(Dada2bError, match="A/C/G/T"):
        merge.merge_align(["ACGTNACGTA", "ACGTAACGTA"], [0], [1], [1])
    with pytest.raises(Dada2bError, match="bad pair index"):
        merge.merge_align(["ACGTAACGTA", "ACGTAACGTA"], [0], [2], [1])
