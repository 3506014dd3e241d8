"""Test tooling: one large synthetic sample through the CUDA sources on the host SIMT emulator (tests/emu) against the reference's own
C++ (oracle/_ref), full output compared -- sizes at which the multi-block upload pipeline, the thread-per-pair row kernels and the
large-round list handling run as they do on the GPU (the suite's cases have <= 2000 uniques).  max_clust bounds the emulated rounds.
  python tools/emu_parity_big.py <n_uniques> <max_clust>        e.g. 70000 8 (two minutes), 150000 4 (five minutes)"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import numpy as np
import build_emu, dada2_b200.api as api
api._LIBPATH = build_emu.build(); api._LIB=None
import dada2_b200
from oracle import ref
from tests import cases
from tools import synth
n=int(sys.argv[1]); mc=int(sys.argv[2])
err=cases.tperr1()
t=time.time(); seqs, ab, q, _ = synth.illumina(n, seed=777); print("gen", round(time.time()-t,1), flush=True)
ref.set_threads(8)
t=time.time(); want=ref.dada_uniques(seqs, ab, None, err, q, homo_gap=-8, max_clust=mc, multithread=True); print("ref s", round(time.time()-t,1), "nclust", len(want["clustering"]["sequence"]), flush=True)
t=time.time(); got=dada2_b200.dada_uniques(seqs, ab, None, err, q, max_clust=mc); print("emu s", round(time.time()-t,1), {k:got["stats"][k] for k in ("gpu_launches","n_nw","n_gapless","n_rounds")}, flush=True)
cases.assert_same(got, want, rtol=1e-10, label="emu big")
print("IDENTICAL n=%d max_clust=%d" % (n, mc))
