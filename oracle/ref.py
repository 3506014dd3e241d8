"""TEST INFRASTRUCTURE ONLY -- ctypes front-end to oracle/_ref/libdada2ref.so, i.e. the
reference's own C++ (compiled unmodified, see oracle/Makefile) behind a flat C-ABI.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
import ctypes as C
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

DEFAULTS = dict(match=5, mismatch=-4, gap=-8, use_kmers=True, kdist_cutoff=0.42, band_size=16,
                omegaA=1e-40, omegaP=1e-4, omegaC=1e-40, detect_singletons=False, max_clust=0,
                min_fold=1.0, min_hamming=1, min_abund=1, use_quals=True, final_consensus=False,
                vectorized_alignment=True, homo_gap=-8, multithread=False, verbose=False, SSE=2,
                gapless=True, greedy=True)


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libdada2ref.so"))


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "libdada2ref.so"))
        L.ref_run.restype = C.c_void_p
        L.ref_error.restype = C.c_char_p
        L.ref_error.argtypes = [C.c_void_p]
        L.ref_free.argtypes = [C.c_void_p]
        for f in (L.ref_field_len,):
            f.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.ref_field_int.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p]
        L.ref_field_dbl.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_void_p]
        L.ref_field_str.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_field_str.restype = C.c_char_p
        L.ref_field_dims.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_ppois_upper.restype = C.c_double
        L.ref_ppois_upper.argtypes = [C.c_int, C.c_double]
        L.ref_calc_pA.restype = C.c_double
        L.ref_calc_pA.argtypes = [C.c_int, C.c_double, C.c_int]
        L.oracle_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def set_threads(n):
    lib().oracle_set_threads(int(n))


def _ivec(h, outer, inner=b""):
    L = lib()
    n = L.ref_field_len(h, outer, inner)
    out = np.zeros(max(n, 0), dtype=np.int32)
    if n > 0:
        L.ref_field_int(h, outer, inner, out.ctypes.data)
    return out


def _dvec(h, outer, inner=b""):
    L = lib()
    n = L.ref_field_len(h, outer, inner)
    out = np.zeros(max(n, 0), dtype=np.float64)
    if n > 0:
        L.ref_field_dbl(h, outer, inner, out.ctypes.data)
    return out


def _svec(h, outer, inner):
    L = lib()
    n = L.ref_field_len(h, outer, inner)
    return [L.ref_field_str(h, outer, inner, i).decode() for i in range(max(n, 0))]


last_native_s = 0.0


def dada_uniques(seqs, abundances, priors, err, quals, **opts):
    """Run the reference's dada_uniques (Rmain.cpp:30).  quals: float64 [nraw, maxlen] (NaN pad)
    or None; err: float64 [16, Q].  Returns a dict mirroring the R list."""
    o = dict(DEFAULTS)
    o.update(opts)
    L = lib()
    nraw = len(seqs)
    arr = (C.c_char_p * nraw)(*[s.encode() for s in seqs])
    ab = np.ascontiguousarray(abundances, dtype=np.int32)
    pr = np.ascontiguousarray(priors if priors is not None else np.zeros(nraw), dtype=np.uint8)
    err = np.asarray(err, dtype=np.float64)
    err_cm = np.asfortranarray(err)  # column-major 16 x Q
    if quals is not None:
        q = np.ascontiguousarray(quals, dtype=np.float64)  # [nraw, maxlen] row-major == maxlen x nraw col-major
        maxlen = q.shape[1]
        qp = q.ctypes.data_as(C.c_void_p)
    else:
        q = None
        maxlen = 0
        qp = None
    global last_native_s
    _t0 = time.perf_counter()
    h = L.ref_run(C.c_int(nraw), arr, ab.ctypes.data_as(C.c_void_p), pr.ctypes.data_as(C.c_void_p),
                  err_cm.ctypes.data_as(C.c_void_p), C.c_int(err.shape[1]), qp, C.c_int(maxlen),
                  C.c_int(o["match"]), C.c_int(o["mismatch"]), C.c_int(o["gap"]), C.c_int(o["use_kmers"]),
                  C.c_double(o["kdist_cutoff"]), C.c_int(o["band_size"]), C.c_double(o["omegaA"]),
                  C.c_double(o["omegaP"]), C.c_double(o["omegaC"]), C.c_int(o["detect_singletons"]),
                  C.c_int(o["max_clust"]), C.c_double(o["min_fold"]), C.c_int(o["min_hamming"]),
                  C.c_int(o["min_abund"]), C.c_int(o["use_quals"]), C.c_int(o["final_consensus"]),
                  C.c_int(o["vectorized_alignment"]), C.c_int(o["homo_gap"]), C.c_int(o["multithread"]),
                  C.c_int(o["verbose"]), C.c_int(o["SSE"]), C.c_int(o["gapless"]), C.c_int(o["greedy"]))
    last_native_s = time.perf_counter() - _t0          # the native call alone (what .Call('_dada2_dada_uniques') costs); marshalling of the
                                                       # Python lists above and of the result below is not the reference's time
    h = C.c_void_p(h)
    try:
        e = L.ref_error(h)
        if e:
            raise RuntimeError(e.decode())
        nr, nc = C.c_int(), C.c_int()
        res = {}
        cl = {}
        cl["sequence"] = _svec(h, b"clustering", b"sequence")
        for k in ("abundance", "n0", "n1", "nunq", "birth_from", "birth_ham"):
            cl[k] = _ivec(h, b"clustering", k.encode())
        for k in ("pval", "birth_pval", "birth_fold", "birth_qave"):
            cl[k] = _dvec(h, b"clustering", k.encode())
        res["clustering"] = cl
        bs = {}
        for k in ("pos", "clust"):
            bs[k] = _ivec(h, b"birth_subs", k.encode())
        for k in ("ref", "sub"):
            bs[k] = _svec(h, b"birth_subs", k.encode())
        bs["qual"] = _dvec(h, b"birth_subs", b"qual")
        res["birth_subs"] = bs
        L.ref_field_dims(h, b"subqual", C.byref(nr), C.byref(nc))
        res["subqual"] = _ivec(h, b"subqual").reshape((nc.value, nr.value)).T.copy()
        L.ref_field_dims(h, b"clusterquals", C.byref(nr), C.byref(nc))
        res["clusterquals"] = _dvec(h, b"clusterquals").reshape((nc.value, nr.value)).T.copy()
        res["map"] = _ivec(h, b"map")
        res["pval"] = _dvec(h, b"pval")
        return res
    finally:
        L.ref_free(h)


def pair(seq0, q0, seq1, q1, err, **opts):
    """Reference sub_new + compute_lambda_ts for (centre seq0, raw seq1)."""
    o = dict(DEFAULTS)
    o.update(opts)
    L = lib()
    l0, l1 = len(seq0), len(seq1)
    err_rm = np.ascontiguousarray(err, dtype=np.float64)
    lam = C.c_double()
    ns = C.c_int()
    kd, kod = C.c_double(-2), C.c_double(-2)
    mp = np.zeros(l0 + 1, dtype=np.uint16)
    pos = np.zeros(l0 + l1 + 1, dtype=np.uint16)
    nt0 = C.create_string_buffer(l0 + l1 + 1)
    nt1 = C.create_string_buffer(l0 + l1 + 1)
    sq0 = np.zeros(l0 + l1 + 1, dtype=np.uint8)
    sq1 = np.zeros(l0 + l1 + 1, dtype=np.uint8)
    al0 = C.create_string_buffer(l0 + l1 + 2)
    al1 = C.create_string_buffer(l0 + l1 + 2)
    eb = C.create_string_buffer(256)
    q0a = np.ascontiguousarray(q0, dtype=np.uint8) if q0 is not None else None
    q1a = np.ascontiguousarray(q1, dtype=np.uint8) if q1 is not None else None
    rc = L.ref_pair(seq0.encode(), q0a.ctypes.data_as(C.c_void_p) if q0a is not None else None,
                    seq1.encode(), q1a.ctypes.data_as(C.c_void_p) if q1a is not None else None,
                    err_rm.ctypes.data_as(C.c_void_p), C.c_int(err_rm.shape[1]),
                    C.c_int(o["match"]), C.c_int(o["mismatch"]), C.c_int(o["gap"]), C.c_int(o["homo_gap"]),
                    C.c_int(o["use_kmers"]), C.c_double(o["kdist_cutoff"]), C.c_int(o["band_size"]),
                    C.c_int(o["vectorized_alignment"]), C.c_int(o["SSE"]), C.c_int(o["gapless"]),
                    C.byref(lam), C.byref(ns), mp.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p),
                    nt0, nt1, sq0.ctypes.data_as(C.c_void_p), sq1.ctypes.data_as(C.c_void_p), al0, al1,
                    C.byref(kd), C.byref(kod), eb)
    if rc < 0:
        raise RuntimeError(eb.value.decode())
    n = max(ns.value, 0)
    return dict(shrouded=(rc == 1), lam=lam.value, nsubs=ns.value, map=mp[:l0].copy(), pos=pos[:n].copy(),
                nt0=bytes(nt0.raw[:n]), nt1=bytes(nt1.raw[:n]), q0=sq0[:n].copy(), q1=sq1[:n].copy(),
                al0=al0.value.decode(), al1=al1.value.decode(), kdist=kd.value, kodist=kod.value)


def align(seq0, seq1, match=5, mismatch=-4, gap=-8, homo_gap=-8, band=16, mode=0):
    L = lib()
    n = len(seq0) + len(seq1) + 2
    a0, a1 = C.create_string_buffer(n), C.create_string_buffer(n)
    rc = L.ref_align(seq0.encode(), seq1.encode(), C.c_int(match), C.c_int(mismatch), C.c_int(gap),
                     C.c_int(homo_gap), C.c_int(band), C.c_int(mode), a0, a1)
    if rc:
        raise RuntimeError("ref_align failed")
    return a0.value.decode(), a1.value.decode()


# ---------------- bimera detection (src/chimera.cpp) ----------------
BIMERA_DEFAULTS = dict(min_fold=1.5, min_abund=2, allow_one_off=False, min_one_off_par_dist=4,
                       match=5, mismatch=-4, gap_p=-8, max_shift=16)       # R/chimeras.R:220, R/dada.R:1-26


def _bimera_table(fn, mat, seqs, o):
    m = np.asfortranarray(np.asarray(mat, dtype=np.int32))          # nrow (samples) x ncol (sequences), column-major like R
    nrow, ncol = m.shape
    assert ncol == len(seqs)
    arr = (C.c_char_p * ncol)(*[s.encode() for s in seqs])
    nflag = np.zeros(ncol, np.int32)
    nsam = np.zeros(ncol, np.int32)
    return m, nrow, ncol, arr, nflag, nsam


def table_bimera(mat, seqs, **opts):
    """The reference's C_table_bimera2 (chimera.cpp:194-207).  mat: [nsample, nseq] counts.  -> (nflag, nsam)"""
    o = dict(BIMERA_DEFAULTS); o.update(opts)
    m, nrow, ncol, arr, nflag, nsam = _bimera_table(None, mat, seqs, o)
    eb = C.create_string_buffer(256)
    rc = lib().ref_table_bimera(C.c_int(nrow), C.c_int(ncol), m.ctypes.data_as(C.c_void_p), arr, C.c_double(o["min_fold"]),
                                C.c_int(o["min_abund"]), C.c_int(o["allow_one_off"]), C.c_int(o["min_one_off_par_dist"]),
                                C.c_int(o["match"]), C.c_int(o["mismatch"]), C.c_int(o["gap_p"]), C.c_int(o["max_shift"]),
                                nflag.ctypes.data_as(C.c_void_p), nsam.ctypes.data_as(C.c_void_p), eb)
    if rc:
        raise RuntimeError(eb.value.decode())
    return nflag, nsam


def is_bimera(sq, pars, **opts):
    """The reference's C_is_bimera (chimera.cpp:18-59)."""
    o = dict(BIMERA_DEFAULTS); o.update(opts)
    arr = (C.c_char_p * len(pars))(*[s.encode() for s in pars])
    rc = lib().ref_is_bimera(sq.encode(), C.c_int(len(pars)), arr, C.c_int(o["allow_one_off"]), C.c_int(o["min_one_off_par_dist"]),
                             C.c_int(o["match"]), C.c_int(o["mismatch"]), C.c_int(o["gap_p"]), C.c_int(o["max_shift"]))
    if rc < 0:
        raise RuntimeError("ref_is_bimera failed")
    return bool(rc)


def bimera_pair(sq, par, **opts):
    """nwalign_vectorized2 + get_lr + get_ham_endsfree for one (query, parent) pair -> dict."""
    o = dict(BIMERA_DEFAULTS); o.update(opts)
    out = np.zeros(5, np.int32)
    n = len(sq) + len(par) + 2
    a0, a1 = C.create_string_buffer(n), C.create_string_buffer(n)
    rc = lib().ref_bimera_pair(sq.encode(), par.encode(), C.c_int(o["allow_one_off"]), C.c_int(o["match"]), C.c_int(o["mismatch"]),
                               C.c_int(o["gap_p"]), C.c_int(o["max_shift"]), out.ctypes.data_as(C.c_void_p), a0, a1)
    if rc:
        raise RuntimeError("ref_bimera_pair failed")
    return dict(left=int(out[0]), right=int(out[1]), left_oo=int(out[2]), right_oo=int(out[3]), ham=int(out[4]),
                al0=a0.value.decode(), al1=a1.value.decode())


# ---------------- mergePairs' native steps (src/evaluate.cpp) ----------------
MERGE_DEFAULTS = dict(match=1, mismatch=-64, gap_p=-64, homo_gap_p=None, band=-1, prefer=1, trim_overhang=False)   # R/paired.R:153-157


def merge_pair(s1, s2, **opts):
    """C_nwalign(endsfree) -> C_eval_pair -> C_pair_consensus for one (forward, rc(reverse)) pair (R/paired.R:153-164)."""
    o = dict(MERGE_DEFAULTS); o.update(opts)
    hg = o["gap_p"] if o["homo_gap_p"] is None else o["homo_gap_p"]
    n = len(s1) + len(s2) + 2
    cnt = np.zeros(3, np.int32)
    cons, a0, a1 = C.create_string_buffer(n), C.create_string_buffer(n), C.create_string_buffer(n)
    rc = lib().ref_merge_pair(s1.encode(), s2.encode(), C.c_int(o["match"]), C.c_int(o["mismatch"]), C.c_int(o["gap_p"]), C.c_int(hg),
                              C.c_int(o["band"]), C.c_int(1), C.c_int(o["prefer"]), C.c_int(o["trim_overhang"]),
                              cnt.ctypes.data_as(C.c_void_p), cons, a0, a1)
    if rc:
        raise RuntimeError("ref_merge_pair failed")
    return dict(nmatch=int(cnt[0]), nmismatch=int(cnt[1]), nindel=int(cnt[2]), sequence=cons.value.decode(), al0=a0.value.decode(), al1=a1.value.decode())
