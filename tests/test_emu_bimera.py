"""The bimera kernels (dd_bimera.cu) on the host SIMT emulator against the reference goldens -- CPU suite.  Same assertions
as tests/test_gpu_zz_bimera.py, at sizes the emulator finishes in seconds (about 5 ms per emulated alignment)."""
import os
import platform
import sys

import numpy as np
import pytest

from tests import bimera_cases as B

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))
pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is x86-64 only")


@pytest.fixture()
def emu(monkeypatch):
    monkeypatch.setenv("DADA2B_BIMFWD", "0")        # tests of the traceback kernel pin it; the variant tests below override
    import build_emu
    import dada2_b200.api as api
    lib = build_emu.build()
    monkeypatch.setattr(api, "_LIBPATH", lib)
    monkeypatch.setattr(api, "_LIB", None)
    import dada2_b200.bimera as bm
    monkeypatch.setattr(bm, "_BOUND", False)
    yield lib


def test_emu_bimera_pairs(emu):
    B.check_pairs(B.product_pair_fn, limit=None if os.environ.get("DADA2B_EMU_FULL") else 120)


@pytest.mark.parametrize("name,opt_ids", [("t40_one_sample", None)] if not os.environ.get("DADA2B_EMU_FULL")
                         else [(n, None) for n in B.table_names()])
def test_emu_bimera_table(emu, name, opt_ids):
    B.check_table(name, B.product_table_fn, opt_ids)


def test_emu_bimera_denovo_and_sharding(emu):
    from dada2_b200 import bimera
    B.check_is_bimera("t40_one_sample", B.product_denovo_fn)
    g = B.golden()
    seqs, mat = g["t40_one_sample_seqs"].tolist(), g["t40_one_sample_mat"]
    tot = [np.zeros(len(seqs), np.int32), np.zeros(len(seqs), np.int32)]
    for r in range(3):                                   # three shards, summed like multi.table_bimera_sharded does
        x = bimera.C_table_bimera2(mat, seqs, shard_rank=r, shard_world=3)
        owned = np.arange(len(seqs)) % 3 == r
        assert not x["nflag"][~owned].any() and not x["nsam"][~owned].any()
        tot[0] += x["nflag"]; tot[1] += x["nsam"]
    assert np.array_equal(tot[0], g["t40_one_sample_o0_nflag"]) and np.array_equal(tot[1], g["t40_one_sample_o0_nsam"])


def test_emu_bimera_errors_and_corners(emu):
    from dada2_b200 import bimera, Dada2bError
    with pytest.raises(Dada2bError, match="A/C/G/T"):
        bimera.C_table_bimera2(np.ones((1, 2), np.int32), ["ACGTNACGTA", "ACGTAACGTA"])
    with pytest.raises(Dada2bError, match="valid sequence table"):
        bimera.C_table_bimera2(np.ones((1, 3), np.int32), ["ACGTAACGTA", "ACGTTACGTA"])
    r = bimera.C_table_bimera2(np.array([[5], [0], [7]], np.int32), ["ACGTAACGTAGG"])        # one sequence: no parents
    assert r["nflag"].tolist() == [0] and r["nsam"].tolist() == [2]
    assert bimera.C_is_bimera("ACGTAACGTAGG", []) is False
    # exact two-parent bimera, and the same with the parents given in the other order
    a = "GTATCGGCTACCGCAAAAATAGTACCCTATTTACGCGGGATGTCCTAACGATCAGTTTCATGTAAGCCTAGCCTACAACG"
    b = "GCATTAACATGGCGTATACTTATTCCTTGCATCGGACGAGTGGCTACGGAAGTGTCCTAAACATTCAGAGATGTGGCTAA"
    chim = a[:36] + b[36:]
    assert bimera.C_is_bimera(chim, [a, b]) is True and bimera.C_is_bimera(chim, [b, a]) is True
    assert bimera.C_is_bimera(chim, [a]) is False


def test_emu_bimera_register_kernel(emu, monkeypatch):
    """DADA2B_BIMFWD=1: the register-resident kernel (dd_bimfwd.cu) in every instantiation (<4,10> .. <32,8>), with and
    without its interior fast path, against the CPU oracle and the goldens; pairs it cannot hold fall back to k_bim_align."""
    import ctypes
    h = ctypes.CDLL(emu)
    h.cuemu_launches.restype = ctypes.c_long
    h.cuemu_launches.argtypes = [ctypes.c_char_p]
    monkeypatch.setenv("DADA2B_BIMFWD", "1")
    n0 = h.cuemu_launches(b"k_bimfwd")
    B.check_pairs_vs_oracle(B.product_pair_fn, npairs=24)
    assert h.cuemu_launches(b"k_bimfwd") - n0 == 5
    monkeypatch.setenv("DADA2B_NO_FAST", "1")
    B.check_pairs_vs_oracle(B.product_pair_fn, shifts=(16,), npairs=24)
    monkeypatch.delenv("DADA2B_NO_FAST")
    B.check_pairs(B.product_pair_fn, shifts=[16], limit=60)            # ragged corpus: everything falls back
    B.check_table("t40_one_sample", B.product_table_fn)
    B.check_table("t150_short", B.product_table_fn, [1])
    B.check_is_bimera("t40_one_sample", B.product_denovo_fn)


def test_emu_bimera_simd_kernel(emu, monkeypatch):
    """DADA2B_BIMFWD=2: two jobs per lane group on the 16-bit SIMD datapath (dd_bimfwd16.cu); neighbours with another query or
    parent length are handed to k_bimfwd, bands that do not fit to k_bim_align -- every route against goldens and oracle."""
    import ctypes
    from oracle import port
    from dada2_b200 import bimera
    h = ctypes.CDLL(emu)
    h.cuemu_launches.restype = ctypes.c_long
    h.cuemu_launches.argtypes = [ctypes.c_char_p]
    monkeypatch.setenv("DADA2B_BIMFWD", "2")
    n0 = h.cuemu_launches(b"k_bimfwd16")
    B.check_pairs_vs_oracle(B.product_pair_fn, npairs=24)                       # random neighbours: mostly handed over
    g = B.golden()
    seqs = g["t120_seqs"].tolist()                                              # all 250 nt: runs of one query with equally long parents
    q, p = np.repeat(np.arange(0, 8), 6), np.tile(np.arange(20, 26), 8)
    for ms, fast in ((16, True), (16, False), (3, True)):
        if not fast:
            monkeypatch.setenv("DADA2B_NO_FAST", "1")
        got = bimera.test_bimera_pairs(seqs, q, p, allow_one_off=True, max_shift=ms)
        monkeypatch.delenv("DADA2B_NO_FAST", raising=False)
        for n, (a, b) in enumerate(zip(q, p)):
            r = port.bimera_pair(seqs[a], seqs[b], allow_one_off=True, max_shift=ms)
            assert list(got[n]) == [r["left"], r["right"], r["left_oo"], r["right_oo"], r["ham"]], (ms, fast, n)
    assert h.cuemu_launches(b"k_bimfwd16") - n0 == 8
    B.check_pairs(B.product_pair_fn, shifts=[16], limit=40)                     # ragged corpus: falls through to k_bim_align
    B.check_table("t40_one_sample", B.product_table_fn)
    B.check_table("t150_short", B.product_table_fn, [1])
