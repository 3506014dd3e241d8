"""Memory-safety check of the CUDA sources on the host SIMT emulator built with AddressSanitizer (tests/emu,
build_emu.build(asan=True)): "device" buffers are exact-size heap blocks and the unused tail of the dynamic shared memory
window is poisoned, so any out-of-bounds global or shared access of a kernel aborts with the source line -- the CPU-side
stand-in for compute-sanitizer's memcheck.  Runs in a subprocess (ASan must be preloaded into the interpreter)."""
import os
import platform
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))

pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is x86-64 only")

SCRIPT = """
import sys
sys.path.insert(0, %r)
import dada2_b200.api as api
api._LIBPATH = %r; api._LIB = None
import tests.test_gpu_parity as T
T.test_pair_corpus_kernels_match_reference()
for name in %r:
    T.test_e2e_matches_reference_golden(name)
T.test_error_paths()
print('ASAN RUN OK')
"""


def _preload():
    """libasan plus libstdc++ (the interpreter does not link libstdc++; without it ASan cannot intercept __cxa_throw)."""
    for cxx in (os.environ.get("CXX", "g++"), "/usr/bin/g++"):
        try:
            libs = [subprocess.run([cxx, "-print-file-name=" + n], capture_output=True, text=True).stdout.strip()
                    for n in ("libasan.so", "libstdc++.so.6")]
        except OSError:
            continue
        if all(os.path.isabs(f) and os.path.exists(f) for f in libs):
            return " ".join(os.path.realpath(f) for f in libs)
    return None


@pytest.mark.parametrize("flags", [{}, {"DADA2B_NWFWD_V2": "1", "DADA2B_FUSED_TAIL": "1", "DADA2B_PIVOT": "1", "DADA2B_TWOPHASE": "1"}],
                         ids=["default", "experimental"])
def test_kernels_are_memory_clean_under_asan(flags):
    asan = _preload()
    if asan is None:
        pytest.skip("libasan not found")
    import build_emu
    lib = build_emu.build(asan=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", **flags)
    names = ["syn500_usequals0", "syn600_band0"]
    out = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, lib, names)], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert "AddressSanitizer" not in out.stderr, out.stderr[-4000:]
    assert out.returncode == 0 and "ASAN RUN OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
