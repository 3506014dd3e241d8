import sys, time
sys.path.insert(0, '.')
import numpy as np
from tools import synth
from tests import cases
import dada2_b200
n = int(sys.argv[1])
seqs, ab, q, truth = synth.illumina(n, seed=12345)
err = cases.tperr1()
call = dada2_b200.PackedCall(seqs, ab, None, err, q)
for i in range(4):
    r, ms = call.run(unpack=False)
    print("one-shot wall ms", round(ms, 2), {k: round(v, 2) if isinstance(v, float) else v for k, v in r["stats"].items() if k.startswith("ms_") or k.endswith("bytes")})
