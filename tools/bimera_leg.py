"""Bimera detection on one synthetic sequence table: the C-ABI call dada2b_table_bimera on host buffers
(pack + H2D + need / align / flag kernels + D2H inside the timed call) with the default alignment kernel (16-bit SIMD,
"simd16") and, pinned through the test hook DADA2B_BIMFWD, the register wavefront (=1) and the warp-per-pair traceback
kernel (=0) it falls back to, next to the reference's own C_table_bimera2 (oracle/_ref, all host threads, best of 3; the
CPU restatement when the compiled reference is absent), outputs diffed.  Prints one line `BIMLEG {json}`.  Run by bench.py
in a subprocess with a timeout after the measured region; never part of `value`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    nseq = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    nsample = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    from tools import synth
    from dada2_b200 import bimera
    seqs, mat = synth.bimera_table(nseq, nsample, seed=21)
    L = len(seqs[0])
    out = {"workload": "%d synthetic %d nt sequences x %d samples (tools/synth.py bimera_table seed 21), isBimeraDenovoTable defaults" % (len(seqs), L, nsample)}
    res = {}
    for tag, env in (("simd16", None), ("register", "1"), ("traceback", "0")):
        if env:
            os.environ["DADA2B_BIMFWD"] = env
        else:
            os.environ.pop("DADA2B_BIMFWD", None)
        try:
            bimera.C_table_bimera2(mat, seqs)                                   # warm-up (context, module load)
            ts, st = [], None
            for _ in range(3):
                t0 = time.perf_counter()
                r = bimera.C_table_bimera2(mat, seqs, return_stats=True)
                ts.append((time.perf_counter() - t0) * 1e3)
                st = r["stats"]
            res[tag] = r
            ms = float(np.median(ts))
            out[tag] = {"e2e_ms": round(ms, 2), "device_ms": round(st["ms_device"], 2), "k_align_ms": round(st["ms_k_align"], 2),
                        "pairs": int(st["n_pairs"]), "pairs_per_s_e2e": st["n_pairs"] / (ms / 1e3),
                        "gcups_align_kernel": st["n_cells"] / 1e9 / (st["ms_k_align"] / 1e3) if st["ms_k_align"] > 0 else None,
                        "gpu_launches": int(st["gpu_launches"]), "h2d_bytes": int(st["h2d_bytes"]), "d2h_bytes": int(st["d2h_bytes"])}
        except Exception as ex:
            out[tag] = {"failed": str(ex)[:200]}
    os.environ.pop("DADA2B_BIMFWD", None)
    try:
        from oracle import ref, port
        from bench import host_cpus
        ncores = host_cpus()
        if ref.available():
            ref.set_threads(ncores)
            dts = []
            for _ in range(3):                                                     # best of 3: a single pass swung 5x between two driver runs in round 1
                t0 = time.perf_counter(); want = ref.table_bimera(mat, seqs); dts.append(time.perf_counter() - t0)
            dt = min(dts)
            kind = "reference"
        else:
            t0 = time.perf_counter(); want = port.table_bimera(mat, seqs); dt = time.perf_counter() - t0
            kind, ncores = "port", 1
        npairs = next((out[t]["pairs"] for t in ("simd16", "register", "traceback") if "pairs" in out.get(t, {})), None)
        out["cpu_baseline"] = {"kind": kind, "cores": ncores, "s": round(dt, 3), "pairs_per_s": (npairs / dt) if npairs else None}
        for tag, r in res.items():
            ok = bool(np.array_equal(r["nflag"], want[0]) and np.array_equal(r["nsam"], want[1]))
            out[tag]["parity_vs_cpu"] = ok
            if ok and "e2e_ms" in out[tag]:
                out[tag]["speedup_e2e_vs_cpu"] = dt / (out[tag]["e2e_ms"] / 1e3)
        out["nflagged"] = int((want[0] > 0).sum())
    except Exception as ex:
        out["cpu_baseline"] = {"failed": str(ex)[:200]}
    print("BIMLEG " + json.dumps(out))


if __name__ == "__main__":
    main()
