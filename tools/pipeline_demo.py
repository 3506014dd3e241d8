"""The accelerated steps chained on one GPU, host buffers between them (each step is its own C-ABI call, as the R package
would drive them): dereplication (8(f1), uniques left resident on the device) -> dada() core (8(a)-(e)) -> bimera detection on the resulting ASV table (8(f3)),
with per-step wall / device times.  Synthetic reads: `nreads` Illumina-like 250 nt reads of 100 variants plus bimeras of
them.   python tools/pipeline_demo.py [nreads=200000]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    nreads = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    import dada2_b200
    from dada2_b200 import bimera, derep
    from tests import cases
    rng = np.random.default_rng(5)
    L, nvar = 250, 100
    root = rng.integers(0, 4, L)
    var = []
    for k in range(nvar):
        v = root.copy(); p = rng.choice(L, int(rng.integers(1, 40)), replace=False); v[p] = (v[p] + rng.integers(1, 4, len(p))) % 4
        var.append(v)
    for k in range(nvar // 4):                                   # two-parent bimeras of abundant variants
        a, b = rng.choice(nvar // 2, 2, replace=False); bp = int(rng.integers(40, L - 40))
        var.append(np.concatenate([var[a][:bp], var[b][bp:]]))
    w = 1.0 / np.arange(1, len(var) + 1); w[nvar:] *= 0.2
    pick = rng.choice(len(var), nreads, p=w / w.sum())
    prof = np.clip(np.round(38 - 18 * (np.arange(L) / (L - 1)) ** 2), 2, 40)
    q = np.clip(np.round(prof[None, :] + rng.normal(0, 3, (nreads, L))), 2, 40).astype(np.uint8)
    reads = np.stack([var[v] for v in pick])
    err = rng.random((nreads, L)) < 10.0 ** (-q.astype(np.float64) / 10.0)
    reads = np.where(err, (reads + rng.integers(1, 4, (nreads, L))) % 4, reads)
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs = [bytes(nt[r]).decode() for r in reads]
    t0 = time.perf_counter(); d, res = derep.derep_reads(seqs, q.ravel(), return_stats=True, resident=True, want_quals=False); t1 = time.perf_counter()
    print("derep : %d reads -> %d uniques   %.1f ms wall (device %.1f ms, sort %.1f ms, %d launches)" % (
        nreads, len(d["uniques"]), (t1 - t0) * 1e3, d["stats"]["ms_device"], d["stats"]["ms_sort"], d["stats"]["gpu_launches"]))
    t0 = time.perf_counter(); r = res.run(cases.tperr1()); t1 = time.perf_counter()          # the uniques never left the device
    asv = r["clustering"]["sequence"]; ab = np.asarray(r["clustering"]["abundance"])
    print("dada  : %d uniques -> %d ASVs   %.1f ms wall (device %.1f ms, %d launches)" % (
        len(d["uniques"]), len(asv), (t1 - t0) * 1e3, r["stats"]["ms_device"], r["stats"]["gpu_launches"]))
    t0 = time.perf_counter(); b = bimera.C_table_bimera2(ab[None, :].astype(np.int32), asv, return_stats=True); t1 = time.perf_counter()
    flagged = bimera.isBimeraDenovoTable(ab[None, :].astype(np.int32), asv)
    true_bim = {bytes(nt[v]).decode() for v in var[nvar:]}
    print("bimera: %d ASVs, %d pairs -> %d flagged (%d of them planted bimeras)   %.1f ms wall (device %.1f ms)" % (
        len(asv), b["stats"]["n_pairs"], int(flagged.sum()), sum(asv[i] in true_bim for i in np.nonzero(flagged)[0]), (t1 - t0) * 1e3, b["stats"]["ms_device"]))


if __name__ == "__main__":
    main()
