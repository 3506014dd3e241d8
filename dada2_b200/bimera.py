"""Host-side mirror of the reference's bimera-detection entry points for the B200 path.

`C_table_bimera2` / `C_is_bimera` have the reference's names, argument order and meaning
(/root/reference/src/chimera.cpp:194-207, :18-59); `isBimeraDenovoTable` / `isBimeraDenovo` restate the thin R wrappers
around them (/root/reference/R/chimeras.R:220-250, :105-150).  All computation happens in libdada2b.so (hand-written
sm_100a CUDA behind include/dada2b_bimera.h); this module only marshals.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import api

ERRLEN = 256


class BimeraOpts(C.Structure):
    _fields_ = [("min_fold", C.c_double), ("min_abund", C.c_int32), ("allow_one_off", C.c_int32),
                ("min_one_off_par_dist", C.c_int32), ("match", C.c_int32), ("mismatch", C.c_int32), ("gap_p", C.c_int32),
                ("max_shift", C.c_int32), ("shard_rank", C.c_int32), ("shard_world", C.c_int32)]


class BimeraStats(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_cells", C.c_int64), ("gpu_launches", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("ms_device", C.c_double), ("ms_k_align", C.c_double), ("ms_total", C.c_double)]


_BOUND = False


def _lib():
    global _BOUND
    L = api.lib()
    if not _BOUND:
        P = C.POINTER
        L.dada2b_bimera_default_opts.argtypes = [P(BimeraOpts)]
        L.dada2b_table_bimera.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_char_p, C.c_void_p, P(BimeraOpts), C.c_int32,
                                          C.c_void_p, C.c_void_p, P(BimeraStats), C.c_char_p]
        L.dada2b_is_bimera.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                       P(BimeraOpts), C.c_int32, C.c_void_p, P(BimeraStats), C.c_char_p]
        L.dada2b_test_bimera_pairs.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                               P(BimeraOpts), C.c_int32, C.c_void_p, C.c_char_p]
        _BOUND = True
    return L


def _opts(**kw):
    o = BimeraOpts()
    _lib().dada2b_bimera_default_opts(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError("unknown bimera option %r" % k)
        setattr(o, k, float(v) if k == "min_fold" else int(v))
    return o


def _pack_seqs(seqs):
    if len(seqs) == 0:
        raise api.Dada2bError("Zero input sequences.")
    lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
    off = np.zeros(len(seqs) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    return "".join(seqs).encode(), off


def _stats(st):
    return {k: getattr(st, k) for k, _ in BimeraStats._fields_}


def C_table_bimera2(mat, seqs, min_fold=1.5, min_abund=2, allow_one_off=False, min_one_off_par_dist=4, match=5, mismatch=-4,
                    gap_p=-8, max_shift=16, device=0, shard_rank=0, shard_world=1, return_stats=False):
    """Drop-in for the reference's C_table_bimera2 (chimera.cpp:194-207).  mat: integer [nsample, nseq] (R's seqtab:
    rows = samples, columns = sequences); -> {"nflag": int32[nseq], "nsam": int32[nseq]}."""
    m = np.asfortranarray(np.asarray(mat, dtype=np.int32))       # column-major like R: mat(i, j) = vals[i + j * nrow]
    if m.ndim != 2 or m.shape[1] != len(seqs):
        raise api.Dada2bError("Input must be a valid sequence table.")
    buf, off = _pack_seqs(seqs)
    o = _opts(min_fold=min_fold, min_abund=min_abund, allow_one_off=allow_one_off, min_one_off_par_dist=min_one_off_par_dist,
              match=match, mismatch=mismatch, gap_p=gap_p, max_shift=max_shift, shard_rank=shard_rank, shard_world=shard_world)
    nflag, nsam = np.zeros(m.shape[1], np.int32), np.zeros(m.shape[1], np.int32)
    st = BimeraStats()
    eb = C.create_string_buffer(ERRLEN)
    rc = _lib().dada2b_table_bimera(m.shape[0], m.shape[1], m.ctypes.data, buf, off.ctypes.data, C.byref(o), int(device),
                                    nflag.ctypes.data, nsam.ctypes.data, C.byref(st), eb)
    if rc:
        raise api.Dada2bError(eb.value.decode())
    out = {"nflag": nflag, "nsam": nsam}
    if return_stats:
        out["stats"] = _stats(st)
    return out


def is_bimera_batch(seqs, query_idx, parent_lists, allow_one_off=False, min_one_off_par_dist=4, match=5, mismatch=-4, gap_p=-8,
                    max_shift=16, device=0, return_stats=False):
    """C_is_bimera (chimera.cpp:18-59) for many queries in one call: query q is seqs[query_idx[q]], its candidate parents
    are seqs[k] for k in parent_lists[q].  -> bool[nquery]"""
    buf, off = _pack_seqs(seqs)
    q = np.ascontiguousarray(query_idx, dtype=np.int32)
    po = np.zeros(len(q) + 1, np.int64)
    np.cumsum([len(p) for p in parent_lists], out=po[1:])
    pi = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.int32) for p in parent_lists]) if len(q) and po[-1] else np.zeros(0, np.int32))
    o = _opts(allow_one_off=allow_one_off, min_one_off_par_dist=min_one_off_par_dist, match=match, mismatch=mismatch, gap_p=gap_p,
              max_shift=max_shift)
    out = np.zeros(max(len(q), 1), np.uint8)
    st = BimeraStats()
    eb = C.create_string_buffer(ERRLEN)
    rc = _lib().dada2b_is_bimera(len(seqs), buf, off.ctypes.data, len(q), q.ctypes.data, po.ctypes.data, pi.ctypes.data,
                                 C.byref(o), int(device), out.ctypes.data, C.byref(st), eb)
    if rc:
        raise api.Dada2bError(eb.value.decode())
    res = out[:len(q)].astype(bool)
    return (res, _stats(st)) if return_stats else res


def C_is_bimera(sq, pars, allow_one_off=False, min_one_off_par_dist=4, match=5, mismatch=-4, gap_p=-8, max_shift=16, device=0):
    """Drop-in for the reference's C_is_bimera (chimera.cpp:18-59): one query against a list of parents."""
    if len(pars) == 0:
        return False
    seqs = [sq] + list(pars)
    return bool(is_bimera_batch(seqs, [0], [np.arange(1, len(seqs))], allow_one_off, min_one_off_par_dist, match, mismatch, gap_p,
                                max_shift, device)[0])


def isBimeraDenovoTable(seqtab, seqs, minSampleFraction=0.9, ignoreNNegatives=1, minFoldParentOverAbundance=1.5, minParentAbundance=2,
                        allowOneOff=False, minOneOffParentDistance=4, maxShift=16, device=0):
    """R/chimeras.R:220-250: per-sample flags from C_table_bimera2, then the consensus vote."""
    if len(set(seqs)) != len(seqs):
        raise api.Dada2bError("Duplicate sequences detected in input.")
    r = C_table_bimera2(seqtab, seqs, minFoldParentOverAbundance, minParentAbundance, allowOneOff, minOneOffParentDistance,
                        max_shift=maxShift, device=device)
    nflag, nsam = r["nflag"].astype(np.int64), r["nsam"].astype(np.int64)
    return (nflag >= nsam) | ((nflag > 0) & (nflag >= (nsam - ignoreNNegatives) * minSampleFraction))    # is.bim, :240-242


def isBimeraDenovo(seqs, abundances, minFoldParentOverAbundance=2, minParentAbundance=8, allowOneOff=False, minOneOffParentDistance=4,
                   maxShift=16, device=0):
    """R/chimeras.R:105-150 for distinct sequences: parents of i = sequences with abundance > fold * abund[i] and
    > minParentAbundance; fewer than two parents => FALSE; all queries go to the device in one batched call."""
    ab = np.asarray(abundances)
    queries, plists = [], []
    for i in range(len(seqs)):
        pars = np.nonzero((ab > minFoldParentOverAbundance * ab[i]) & (ab > minParentAbundance))[0]      # :127
        if len(pars) >= 2:                                                                                # :128-129
            queries.append(i); plists.append(pars)
    out = np.zeros(len(seqs), bool)
    if queries:
        out[np.asarray(queries)] = is_bimera_batch(seqs, queries, plists, allowOneOff, minOneOffParentDistance, max_shift=maxShift,
                                                   device=device)
    return out


def test_bimera_pairs(seqs, query, parent, allow_one_off=False, match=5, mismatch=-4, gap_p=-8, max_shift=16, device=0):
    """Kernel-level hook: -> int32[npairs, 5] = left, right, left_oo, right_oo, ham (chimera.cpp get_lr / get_ham_endsfree)."""
    buf, off = _pack_seqs(seqs)
    q = np.ascontiguousarray(query, dtype=np.int32)
    p = np.ascontiguousarray(parent, dtype=np.int32)
    o = _opts(allow_one_off=allow_one_off, match=match, mismatch=mismatch, gap_p=gap_p, max_shift=max_shift)
    out = np.zeros((len(q), 5), np.int32)
    eb = C.create_string_buffer(ERRLEN)
    rc = _lib().dada2b_test_bimera_pairs(len(seqs), buf, off.ctypes.data, len(q), q.ctypes.data, p.ctypes.data, C.byref(o), int(device),
                                         out.ctypes.data, eb)
    if rc:
        raise api.Dada2bError(eb.value.decode())
    return out


test_bimera_pairs.__test__ = False
