"""One experimental variant (DADA2B_* switches taken from the environment) on a saved workload: resident steps timed,
outputs diffed against the default path's outputs saved by bench.py.  Prints one JSON object.  Run by bench.py in a
subprocess with a timeout, after the measured region, so that the round-end bench also says -- on hardware -- whether
each off-by-default kernel variant reproduces the default path's results and what it costs.  Never part of `value`."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def unflatten(z):
    out = {}
    for k in z.files:
        a, _, b = k.partition(".")
        v = z[k]
        if b:
            out.setdefault(a, {})[b] = v.tolist() if b in ("sequence", "ref", "sub") else v
        else:
            out[a] = v
    return out


def main():
    wl, ref_npz, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    warmup = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    import dada2_b200
    from tests import cases
    z = np.load(wl, allow_pickle=False)
    seqs = z["seqs"].tolist()
    res = dada2_b200.Resident(seqs, z["ab"], None, z["q"], device=0)
    err = z["err"]
    out = None
    for _ in range(warmup):
        out = res.run(err)
    ms, k = [], {}
    for _ in range(steps):
        t0 = time.perf_counter()
        out = res.run(err)
        ms.append((time.perf_counter() - t0) * 1e3)
    st = out["stats"]
    want = unflatten(np.load(ref_npz, allow_pickle=False))
    try:
        cases.assert_same(out, want, rtol=1e-10, label="ab")
        parity = True
    except AssertionError as e:
        parity = "MISMATCH: %s" % str(e)[:200]
    print("ABLEG " + json.dumps({
        "parity_vs_default": parity, "ms_per_step": round(float(np.median(ms)), 2), "step_ms": [round(x, 2) for x in ms],
        "device_ms": round(st["ms_device"], 2), "gpu_launches": int(st["gpu_launches"]),
        "kernel_ms": {k: round(st[k], 2) for k in ("ms_k_classify", "ms_k_align_nw", "ms_k_align_gl", "ms_k_align_final")},
        "host_ms": {k: round(st[k], 2) for k in ("ms_setup", "ms_loop", "ms_final")}}))
    res.close()


if __name__ == "__main__":
    main()
