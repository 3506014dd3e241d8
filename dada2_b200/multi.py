"""Multi-GPU host logic: one process per GPU (torch.distributed), samples of a run distributed over
ranks -- the reference's own per-sample loop (R/dada.R:266-365) made data-parallel.  No collective is
on the data path; results are gathered once at the end (gather_object).  Works with the `gloo` backend
on CPU for the logic tests (the per-sample runner is injectable)."""
from __future__ import annotations


def shard_samples(n_samples: int, rank: int, world: int):
    """Round-robin assignment (sample i -> rank i % world): balances the abundance-sorted sample lists
    dada() usually sees without any communication."""
    return list(range(rank, n_samples, world))


def dada_samples(samples, err, runner=None, group=None, **opts):
    """samples: list of (seqs, abundances, priors, quals).  Every rank passes the same list; rank r runs
    its shard with `runner` (default: dada2_b200.dada_uniques on this rank's GPU) and rank 0 returns the
    full list of per-sample results in input order (other ranks return None)."""
    import torch.distributed as dist
    if runner is None:
        from .api import dada_uniques as runner
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = shard_samples(len(samples), rank, world)
    local = {i: runner(samples[i][0], samples[i][1], samples[i][2], err, samples[i][3], **opts) for i in mine}
    if world == 1:
        return [local[i] for i in range(len(samples))]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0, group=group)
    if rank != 0:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    return [merged[i] for i in range(len(samples))]


def sharded_resident(seqs, abundances, priors, quals, device=None, group=None):
    """One sample sharded over all ranks of `group` (every rank passes the same uniques): returns a
    Resident whose run() aligns raw r on rank r % world and exchanges the new stored comparisons with one
    NCCL all-gather per split round (dada2b_comm_init).  Every rank gets the complete result."""
    import torch
    import torch.distributed as dist
    from . import api
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.cuda.current_device()
    res = api.Resident(seqs, abundances, priors, quals, device=device)
    box = [api.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    res.comm_init(rank, world, box[0])
    return res


def table_bimera_sharded(mat, seqs, device=None, group=None, runner=None, **opts):
    """C_table_bimera2 over all ranks of `group`: the queries shard with no data-path collective (rank r evaluates the
    sequences j with j % world == r against the replicated table, chimera.cpp:105 is an independent loop over j); one
    SUM all-reduce of the two int32 result vectors at the end.  Every rank gets the complete {"nflag", "nsam"}.
    `runner` is injectable for the CPU logic tests (default: dada2_b200.bimera.C_table_bimera2 on this rank's GPU)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if runner is None:
        from .bimera import C_table_bimera2 as runner
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    kw = dict(opts)
    if device is not None:
        kw["device"] = device
    r = runner(mat, seqs, shard_rank=rank, shard_world=world, **kw)
    if world == 1:
        return {"nflag": r["nflag"], "nsam": r["nsam"]}
    on_gpu = dist.get_backend(group) == "nccl"
    t = torch.from_numpy(np.stack([r["nflag"], r["nsam"]]).astype(np.int32))
    if on_gpu:
        t = t.cuda(device if device is not None else torch.cuda.current_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    t = t.cpu().numpy()
    return {"nflag": t[0].copy(), "nsam": t[1].copy()}
