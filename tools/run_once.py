import sys, time, os
sys.path.insert(0,'.')
import faulthandler; faulthandler.dump_traceback_later(int(os.environ.get("RUN_ONCE_WATCHDOG", "400")), exit=True)
import numpy as np
from tools import synth
from tests import cases
import dada2_b200
n = int(sys.argv[1])
seqs, ab, q, truth = synth.illumina(n, seed=12345)
err = cases.tperr1()
t=time.time()
r = dada2_b200.dada_uniques(seqs, ab, None, err, q)
print("time", time.time()-t, "nclust", len(r['clustering']['sequence']), r['stats'])
