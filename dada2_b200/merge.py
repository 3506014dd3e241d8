"""Host-side mirror of mergePairs' native steps for the B200 path.

`merge_align` is the batched, fused form of the reference's per-pair chain C_nwalign -> C_eval_pair -> C_pair_consensus
(/root/reference/src/evaluate.cpp:18-174, driven by R/paired.R:150-168); `mergePairs_core` restates the few R lines around
it (prefer / accept / blanking of rejected sequences).  All computation happens in libdada2b.so behind
include/dada2b_merge.h; this module only marshals.  No CPU fallback.
"""
import ctypes as C

import numpy as np

from . import api
from .bimera import _pack_seqs

ERRLEN = 256


class MergeOpts(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_p", C.c_int32), ("homo_gap_p", C.c_int32), ("band", C.c_int32),
                ("trim_overhang", C.c_int32)]


class MergeOut(C.Structure):
    _fields_ = [("npairs", C.c_int32), ("nmatch", C.POINTER(C.c_int32)), ("nmismatch", C.POINTER(C.c_int32)), ("nindel", C.POINTER(C.c_int32)),
                ("cons_concat", C.POINTER(C.c_char)), ("cons_off", C.POINTER(C.c_int64)),
                ("n_cells", C.c_int64), ("gpu_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("ms_device", C.c_double), ("ms_k_merge", C.c_double), ("ms_total", C.c_double)]


_BOUND = False


def _lib():
    global _BOUND
    L = api.lib()
    if not _BOUND:
        P = C.POINTER
        L.dada2b_merge_default_opts.argtypes = [P(MergeOpts)]
        L.dada2b_merge_pairs.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, P(MergeOpts),
                                         C.c_int32, P(P(MergeOut)), C.c_char_p]
        L.dada2b_merge_free.argtypes = [P(MergeOut)]
        _BOUND = True
    return L


def merge_align(seqs, s1_idx, s2_idx, prefer, match=1, mismatch=-64, gap_p=-64, homo_gap_p=None, band=-1, trim_overhang=False, device=0,
                return_stats=False):
    """For every pairing x: nwalign(seqs[s1_idx[x]], seqs[s2_idx[x]], band, endsfree=TRUE) -> C_eval_pair -> C_pair_consensus
    (prefer[x], trim_overhang).  -> dict(nmatch, nmismatch, nindel: int32[npairs]; sequence: list[str])"""
    L = _lib()
    buf, off = _pack_seqs(seqs)
    a = np.ascontiguousarray(s1_idx, dtype=np.int32)
    b = np.ascontiguousarray(s2_idx, dtype=np.int32)
    p = np.ascontiguousarray(prefer, dtype=np.int32)
    if not (len(a) == len(b) == len(p)):
        raise api.Dada2bError("dada2b: s1_idx, s2_idx and prefer must have the same length.")
    o = MergeOpts(int(match), int(mismatch), int(gap_p), int(gap_p if homo_gap_p is None else homo_gap_p), int(band), int(bool(trim_overhang)))
    out = C.POINTER(MergeOut)()
    eb = C.create_string_buffer(ERRLEN)
    rc = L.dada2b_merge_pairs(len(seqs), buf, off.ctypes.data, len(a), a.ctypes.data, b.ctypes.data, p.ctypes.data, C.byref(o), int(device),
                              C.byref(out), eb)
    if rc:
        raise api.Dada2bError(eb.value.decode())
    try:
        r = out.contents
        n = r.npairs
        offs = np.ctypeslib.as_array(r.cons_off, shape=(n + 1,)).copy()
        raw = C.string_at(r.cons_concat, int(offs[-1])).decode() if n else ""
        res = {"nmatch": np.ctypeslib.as_array(r.nmatch, shape=(max(n, 1),))[:n].copy(),
               "nmismatch": np.ctypeslib.as_array(r.nmismatch, shape=(max(n, 1),))[:n].copy(),
               "nindel": np.ctypeslib.as_array(r.nindel, shape=(max(n, 1),))[:n].copy(),
               "sequence": [raw[offs[x]:offs[x + 1]] for x in range(n)]}
        if return_stats:
            res["stats"] = {k: getattr(r, k) for k in ("n_cells", "gpu_launches", "h2d_bytes", "d2h_bytes", "ms_device", "ms_k_merge", "ms_total")}
        return res
    finally:
        L.dada2b_merge_free(out)


def mergePairs_core(Fseqs, Rseqs_rc, forward, reverse, n0F, n0R, minOverlap=12, maxMismatch=0, trimOverhang=False, device=0):
    """R/paired.R:139-172 for the unique pairings (forward[x], reverse[x]) (0-based indices into Fseqs / Rseqs_rc, the
    latter already reverse-complemented): scores as set at :153-157, prefer = 1 + (n0R > n0F) (:162), accept rule (:163),
    rejected sequences blanked (:172)."""
    forward, reverse = np.asarray(forward), np.asarray(reverse)
    pool = list(Fseqs) + list(Rseqs_rc)
    mm = -64 if maxMismatch == 0 else -8
    prefer = 1 + (np.asarray(n0R)[reverse] > np.asarray(n0F)[forward]).astype(np.int32)
    r = merge_align(pool, forward, reverse + len(Fseqs), prefer, match=1, mismatch=mm, gap_p=mm, band=-1, trim_overhang=trimOverhang, device=device)
    accept = (r["nmatch"] >= minOverlap) & ((r["nmismatch"] + r["nindel"]) <= maxMismatch)
    r["prefer"] = prefer
    r["accept"] = accept
    r["sequence"] = [s if ok else "" for s, ok in zip(r["sequence"], accept)]
    return r
