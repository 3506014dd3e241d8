"""How many threads serve the reference's C++ best on this host?  (GPU boxes: 128 hardware threads under a 16-CPU cgroup quota)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import workload, host_cpus
from oracle import ref
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
seqs, ab, q, err = workload(n, 12345)
print("host_cpus()", host_cpus(), "os.cpu_count()", os.cpu_count(), flush=True)
for nt in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,32,128,16").split(",")]:
    ref.set_threads(nt)
    ref.dada_uniques(seqs, ab, None, err, q, multithread=True)
    print("threads %d: %.2f s" % (nt, ref.last_native_s), flush=True)
