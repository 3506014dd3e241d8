"""torchrun entry: one sample sharded over all ranks (dada2b_comm_init); checks parity against the
reference goldens on small cases, then times a synthetic workload of N uniques.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_sharded.py 100000"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from dada2_b200 import multi
from tests import cases
from tests.test_oracle import load_golden

ok = True
for name in ("syn800_default", "syn2000_default", "syn700_ragged", "syn800_nogreedy", "syn800_priors"):
    seqs, ab, pri, err, q, opts = cases.build_case(name)
    res = multi.sharded_resident(seqs, ab, pri, q, device=local)
    got = res.run(err, **opts)
    want = load_golden(name)
    pb = None
    if pri is not None:
        pb = np.zeros(len(want["clustering"]["sequence"]), dtype=bool)
        pb[1:] = want["clustering"]["birth_pval"][1:] >= opts.get("omegaA", 1e-40)
    try:
        cases.assert_same(got, want, rtol=1e-10, prior_born=pb, label=name)
        print("rank", rank, name, "PARITY OK", flush=True)
    except AssertionError as e:
        ok = False
        print("rank", rank, name, "MISMATCH", e, flush=True)
    res.close()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 0
if n:
    from tools import synth
    seqs, ab, q, _ = synth.illumina(n, seed=12345)
    err = cases.tperr1()
    res = multi.sharded_resident(seqs, ab, None, q, device=local)
    for it in range(4):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        r = res.run(err)
        torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) * 1e3
        if rank == 0:
            st = r["stats"]
            print("sharded world=%d n=%d: %.1f ms  (setup %.1f loop %.1f final %.1f; device %.1f; %s; launches %d) nclust %d" % (
                world, n, dt, st["ms_setup"], st["ms_loop"], st["ms_final"], st["ms_device"], " ".join("%s %.1f" % (k[5:], st[k]) for k in st if k.startswith("ms_k_")),
                st["gpu_launches"], len(r["clustering"]["sequence"])), flush=True)
    res.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
