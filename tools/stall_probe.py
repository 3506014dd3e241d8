"""Diagnostic (profiles/r2_host_stalls.md): N resident passes with DADA2B_STALLWATCH on -- the library reports every CUDA
runtime call of its driver, and every stretch of host code between two calls, that takes longer than 15 ms -- next to a
Python thread that only sleeps 2 ms at a time (is the whole process stalled?) and the cgroup's CPU-throttling counters."""
import sys, time, os, threading
sys.path.insert(0, '.')
os.environ.setdefault("DADA2B_STALLWATCH", "15")
import numpy as np
from tools import synth
from tests import cases
import dada2_b200


def cpu_stat():
    try:
        return {l.split()[0]: int(l.split()[1]) for l in open("/sys/fs/cgroup/cpu.stat")}
    except Exception:
        return {}


n = int(sys.argv[1]); reps = int(sys.argv[2])
seqs, ab, q, _ = synth.illumina(n, seed=12345)
err = cases.tperr1()
res = dada2_b200.Resident(seqs, ab, None, q)
res.run(err)
stop = False
T0 = time.perf_counter()


def watcher():
    while not stop:
        t = time.perf_counter()
        time.sleep(0.002)
        g = time.perf_counter() - t
        if g > 0.02:
            print("WATCHER gap %.1f ms at t=%.3f s" % (g * 1e3, t - T0), file=sys.stderr, flush=True)


th = threading.Thread(target=watcher, daemon=True); th.start()
c0 = cpu_stat()
for i in range(reps):
    t0 = time.perf_counter()
    r = res.run(err)
    s = r["stats"]
    print("PASS %d t=%.3f s wall %.1f ms setup %.1f loop %.1f final %.1f" % (i, t0 - T0, (time.perf_counter() - t0) * 1e3, s["ms_setup"], s["ms_loop"], s["ms_final"]),
          file=sys.stderr, flush=True)
stop = True
c1 = cpu_stat()
print("cgroup cpu.stat delta:", {k: c1[k] - c0.get(k, 0) for k in c1}, file=sys.stderr)
