"""The CUDA sources run on the host SIMT emulator (tests/emu) against the reference goldens -- CPU suite.

tests/emu compiles dada2_b200/csrc/*.cu (kernels, round driver, C-ABI) for the host: every CUDA thread is a fiber, warp
shuffles / ballots / barriers are real rendezvous points, shared memory and atomics behave as on the device.  These
tests therefore check the *kernel logic itself* without a GPU: the same assertions as tests/test_gpu_parity.py, on the
emulated library.  They do not replace the `-m gpu` parity tests (hardware scheduling, memory model and the device
libm are not emulated) and nothing under dada2_b200/ ever loads the emulated library.

DADA2B_EMU_FULL=1 runs all e2e cases (about five minutes) instead of the quick subset.
"""
import ctypes
import os
import platform
import sys

import pytest

from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is x86-64 only")

QUICK = ["syn800_default", "syn800_nogreedy", "syn800_nokmers", "syn600_band0", "syn800_maxclust5", "syn800_kdist",
         "syn800_ones_err", "syn700_ragged", "syn500_usequals0"]
E2E = list(cases.E2E_CASES) if os.environ.get("DADA2B_EMU_FULL") else QUICK


@pytest.fixture(scope="module")
def emu_lib():
    import build_emu
    return build_emu.build()


@pytest.fixture()
def emu(emu_lib, monkeypatch):
    """Points dada2_b200.api at the emulated library for one test; yields the launch-counter handle."""
    import dada2_b200.api as api
    monkeypatch.setattr(api, "_LIBPATH", emu_lib)
    monkeypatch.setattr(api, "_LIB", None)
    h = ctypes.CDLL(emu_lib)
    h.cuemu_launches.restype = ctypes.c_long
    h.cuemu_launches.argtypes = [ctypes.c_char_p]
    h.cuemu_reset_launches()
    yield h


def _gpu_tests():
    import tests.test_gpu_parity as T
    return T


def test_emu_calc_pA(emu):
    _gpu_tests().test_device_calc_pA_matches_oracle()


def test_emu_pair_corpus(emu):
    _gpu_tests().test_pair_corpus_kernels_match_reference()
    assert emu.cuemu_launches(b"k_classify") > 0 and emu.cuemu_launches(b"k_align") > 0


def test_emu_config1(emu):
    _gpu_tests().test_config1_bit_identical()
    assert emu.cuemu_launches(b"k_nwfwd<") > 0 and emu.cuemu_launches(b"k_nwfwd2") == 0


def test_emu_error_paths(emu):
    _gpu_tests().test_error_paths()


@pytest.mark.parametrize("name", E2E)
def test_emu_e2e(emu, name):
    _gpu_tests().test_e2e_matches_reference_golden(name)


@pytest.mark.parametrize("name", ["syn800_default", "syn800_band8", "syn700_ragged"] if not os.environ.get("DADA2B_EMU_FULL") else list(cases.E2E_CASES))
def test_emu_e2e_nwfwd_v2(emu, monkeypatch, name):
    """The restructured loop-NW kernel (dd_nwfwd2.cu, DADA2B_NWFWD_V2=1) gives the reference's results."""
    monkeypatch.setenv("DADA2B_NWFWD_V2", "1")
    _gpu_tests().test_e2e_matches_reference_golden(name)
    opts = cases.E2E_CASES[name][1]
    if opts.get("band_size", 16) >= 0 and "homo_gap" not in opts:
        assert emu.cuemu_launches(b"k_nwfwd2") > 0 and emu.cuemu_launches(b"k_nwfwd<") == 0


@pytest.mark.parametrize("name", ["syn800_default", "syn700_ragged"] if not os.environ.get("DADA2B_EMU_FULL") else list(cases.E2E_CASES))
def test_emu_e2e_twophase(emu, monkeypatch, name):
    """The two-phase loop NW (exact lambda bound first, DADA2B_TWOPHASE=1) gives the reference's results."""
    monkeypatch.setenv("DADA2B_TWOPHASE", "1")
    _gpu_tests().test_e2e_matches_reference_golden(name)
