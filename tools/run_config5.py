"""BASELINE configs[4] flavour on one GPU: PacBio-like ~1.5 kb uniques, BAND_SIZE=32, HOMOPOLYMER_GAP_PENALTY=-1 (scalar NW with
homopolymer gap costs) -- timing of the default path (warp-per-pair `nw_warp`) and of the restructured register kernel
(DADA2B_NWFWD_V2=1, the only register kernel that knows homopolymer gaps), each in its own subprocess, outputs diffed against
the CPU oracle on a bounded subsample.   python tools/run_config5.py [n_uniques=20000] [oracle_subsample=1500]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LEG = r'''
import sys, time, json
sys.path.insert(0, %r)
import numpy as np
from tools import synth
from tests import cases
import dada2_b200
n, nsub = int(sys.argv[1]), int(sys.argv[2])
seqs, ab, q = synth.pacbio(n, L=1500, nvar=30, seed=5)
err = synth.extend_err(cases.tperr1(), 94)
opts = dict(band_size=32, vectorized_alignment=False, homo_gap=-1)
res = dada2_b200.Resident(seqs, ab, None, q)
out = None
ts = []
for it in range(3):
    t0 = time.perf_counter(); out = res.run(err, **opts); ts.append((time.perf_counter() - t0) * 1e3)
st = out["stats"]
line = {"n_uniques": len(seqs), "maxlen": max(map(len, seqs)), "ms": [round(x, 1) for x in ts], "uniques_per_s": len(seqs) / (min(ts) / 1e3),
        "nclust": len(out["clustering"]["sequence"]), "stats": {k: (round(v, 2) if isinstance(v, float) else int(v)) for k, v in st.items()}}
if nsub:
    from oracle import port
    s2, a2 = seqs[:nsub], ab[:nsub]
    q2 = q[:nsub, :max(len(x) for x in s2)]          # the quality matrix must be exactly as wide as the longest read of the call
    got = dada2_b200.dada_uniques(s2, a2, None, err, q2, **opts)
    t0 = time.perf_counter(); want = port.dada_uniques(s2, a2, None, err, q2, **opts); line["oracle_s"] = round(time.perf_counter() - t0, 1)
    try:
        cases.assert_same(got, want, rtol=1e-10, label="config5"); line["parity_subsample"] = True
    except AssertionError as e:
        line["parity_subsample"] = "MISMATCH: " + str(e)[:200]
print("C5LEG " + json.dumps(line))
''' % ROOT


def main():
    n = sys.argv[1] if len(sys.argv) > 1 else "20000"
    nsub = sys.argv[2] if len(sys.argv) > 2 else "1500"
    for tag, env in (("default (nw_warp)", {}), ("DADA2B_NWFWD_V2=1", {"DADA2B_NWFWD_V2": "1"})):
        t0 = time.time()
        try:
            out = subprocess.run([sys.executable, "-c", LEG, n, nsub], env=dict(os.environ, **env), capture_output=True, text=True, timeout=1500)
            rows = [l for l in out.stdout.splitlines() if l.startswith("C5LEG ")]
            print(tag, "->", rows[-1][6:] if rows else ("FAILED: " + (out.stderr or out.stdout)[-500:]), "(%.0f s)" % (time.time() - t0), flush=True)
        except subprocess.TimeoutExpired:
            print(tag, "-> TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
