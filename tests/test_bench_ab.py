"""bench.py's post-measurement A/B of the experimental kernel variants (experimental_ab + tools/ab_leg.py), exercised on
the host SIMT emulator: every variant must report parity with the default path's outputs, a hung or crashing leg must
be reported and never raise."""
import os
import platform
import sys

import pytest

from tests import cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "emu"))
pytestmark = pytest.mark.skipif(platform.machine() != "x86_64", reason="the emulator's fiber switch is x86-64 only")

WRAP = ("import sys, runpy; sys.path.insert(0, %r); import dada2_b200.api as api; api._LIBPATH = %r; "
        "sys.argv = ['ab_leg.py'] + sys.argv[1:]; runpy.run_path(%r, run_name='__main__')")


def test_ab_leg_reports_parity_for_every_variant(monkeypatch):
    import build_emu
    import bench
    import dada2_b200.api as api
    lib = build_emu.build()
    monkeypatch.setattr(api, "_LIBPATH", lib)
    monkeypatch.setattr(api, "_LIB", None)
    seqs, ab, pri, err, q, opts = cases.build_case("syn800_default")
    seqs, ab, q = seqs[:300], ab[:300], q[:300]
    import dada2_b200
    last = dada2_b200.dada_uniques(seqs, ab, None, err, q)
    leg = [sys.executable, "-c", WRAP % (ROOT, lib, os.path.join(ROOT, "tools", "ab_leg.py"))]
    variants = [v for v in bench.AB_VARIANTS if v[0] in ("nwfwd2", "all")]       # every variant: tests/test_emu_parity.py
    res = bench.experimental_ab(seqs, ab, q, err, last, 600, 0, leg_cmd=leg, steps=1, warmup=0, variants=variants)
    assert set(res) == {"nwfwd2", "all"}
    for tag, r in res.items():
        assert r.get("parity_vs_default") is True, (tag, r)
        assert r["gpu_launches"] > 0 and r["switches"]


def test_ab_leg_failures_are_contained():
    import bench
    seqs, ab, pri, err, q, opts = cases.build_case("syn800_default")
    from oracle import port
    last = port.dada_uniques(seqs, ab, None, err, q, homo_gap=-8)
    res = bench.experimental_ab(seqs, ab, q, err, last, 600, 0, leg_cmd=[sys.executable, "-c", "import sys; sys.exit('boom')"])
    assert all("failed" in r for r in res.values())
    res = bench.experimental_ab(seqs, ab, q, err, last, 10, 0, leg_cmd=[sys.executable, "-c", "pass"])
    assert all("skipped" in r for r in res.values())


def test_bimera_leg_runs_on_the_emulator_and_failures_are_contained():
    import build_emu
    import bench
    lib = build_emu.build()
    wrap = ("import sys, runpy; sys.path.insert(0, %r); import dada2_b200.api as api; api._LIBPATH = %r; "
            "sys.argv = ['bimera_leg.py', '40', '2']; runpy.run_path(%r, run_name='__main__')")
    r = bench.bimera_leg(300, 0, cmd=[sys.executable, "-c", wrap % (ROOT, lib, os.path.join(ROOT, "tools", "bimera_leg.py"))])
    assert all(r[t]["parity_vs_cpu"] is True for t in ("traceback", "register", "simd16")), r
    assert r["traceback"]["pairs"] == r["register"]["pairs"] > 0 and r["cpu_baseline"]["kind"] in ("reference", "port")
    assert "failed" in bench.bimera_leg(30, 0, cmd=[sys.executable, "-c", "import sys; sys.exit('boom')"])
