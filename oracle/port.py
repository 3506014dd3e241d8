"""TEST INFRASTRUCTURE ONLY -- ctypes front-end to oracle/_ref/liboracle_port.so (the CPU
restatement, oracle/port.cpp).  Same calling convention as dada2_b200.dada_uniques."""
import ctypes as C
import os
import numpy as np
from dada2_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "liboracle_port.so"))


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(os.path.join(_HERE, "_ref", "liboracle_port.so"))
        L.port_run.argtypes = [C.POINTER(_abi.In), C.POINTER(_abi.Opts), C.POINTER(C.POINTER(_abi.Out)), C.c_char_p]
        L.port_free.argtypes = [C.POINTER(_abi.Out)]
        L.port_calc_pA.restype = C.c_double
        L.port_calc_pA.argtypes = [C.c_int, C.c_double, C.c_int]
        L.port_ppois_upper.restype = C.c_double
        L.port_ppois_upper.argtypes = [C.c_int, C.c_double]
        _LIB = L
    return _LIB


def dada_uniques(seqs, abundances, priors, err, quals, **opts):
    L = lib()
    pin = _abi.PackedIn(seqs, abundances, priors, err, quals)
    o = _abi.make_opts(**opts)
    out = C.POINTER(_abi.Out)()
    eb = C.create_string_buffer(_abi.ERRLEN)
    rc = L.port_run(C.byref(pin.struct), C.byref(o), C.byref(out), eb)
    if rc:
        raise RuntimeError(eb.value.decode())
    try:
        return _abi.unpack_out(out.contents)
    finally:
        L.port_free(out)


def pair(seq0, q0, seq1, q1, err, use_kmers=True, kdist_cutoff=0.42, **opts):
    L = lib()
    o = _abi.make_opts(**opts)
    l0, l1 = len(seq0), len(seq1)
    err_rm = np.ascontiguousarray(err, dtype=np.float64)
    lam, ns = C.c_double(), C.c_int()
    mp = np.zeros(l0 + 1, dtype=np.uint16)
    pos = np.zeros(l0 + l1 + 1, dtype=np.uint16)
    nt0, nt1 = C.create_string_buffer(l0 + l1 + 1), C.create_string_buffer(l0 + l1 + 1)
    sq0, sq1 = np.zeros(l0 + l1 + 1, dtype=np.uint8), np.zeros(l0 + l1 + 1, dtype=np.uint8)
    al0, al1 = C.create_string_buffer(l0 + l1 + 2), C.create_string_buffer(l0 + l1 + 2)
    eb = C.create_string_buffer(256)
    q0a = np.ascontiguousarray(q0, dtype=np.uint8) if q0 is not None else None
    q1a = np.ascontiguousarray(q1, dtype=np.uint8) if q1 is not None else None
    kind = L.port_pair(seq0.encode(), q0a.ctypes.data_as(C.c_void_p) if q0a is not None else None,
                       seq1.encode(), q1a.ctypes.data_as(C.c_void_p) if q1a is not None else None,
                       err_rm.ctypes.data_as(C.c_void_p), C.c_int(err_rm.shape[1]), C.byref(o),
                       C.c_int(use_kmers), C.c_double(kdist_cutoff), C.byref(lam), C.byref(ns),
                       mp.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p), nt0, nt1,
                       sq0.ctypes.data_as(C.c_void_p), sq1.ctypes.data_as(C.c_void_p), al0, al1, eb)
    if kind < 0:
        raise RuntimeError(eb.value.decode())
    n = max(ns.value, 0)
    return dict(kind=kind, shrouded=(kind == 0), lam=lam.value, nsubs=ns.value, map=mp[:l0].copy(),
                pos=pos[:n].copy(), nt0=bytes(nt0.raw[:n]), nt1=bytes(nt1.raw[:n]), q0=sq0[:n].copy(),
                q1=sq1[:n].copy(), al0=al0.value.decode(), al1=al1.value.decode())


# ---------------- bimera detection (restatement of src/chimera.cpp in port.cpp) ----------------
BIMERA_DEFAULTS = dict(min_fold=1.5, min_abund=2, allow_one_off=False, min_one_off_par_dist=4,
                       match=5, mismatch=-4, gap_p=-8, max_shift=16)


def table_bimera(mat, seqs, **opts):
    o = dict(BIMERA_DEFAULTS); o.update(opts)
    m = np.asfortranarray(np.asarray(mat, dtype=np.int32))
    nrow, ncol = m.shape
    assert ncol == len(seqs)
    arr = (C.c_char_p * ncol)(*[s.encode() for s in seqs])
    nflag, nsam = np.zeros(ncol, np.int32), np.zeros(ncol, np.int32)
    lib().port_table_bimera(C.c_int(nrow), C.c_int(ncol), m.ctypes.data_as(C.c_void_p), arr, C.c_double(o["min_fold"]),
                            C.c_int(o["min_abund"]), C.c_int(o["allow_one_off"]), C.c_int(o["min_one_off_par_dist"]),
                            C.c_int(o["match"]), C.c_int(o["mismatch"]), C.c_int(o["gap_p"]), C.c_int(o["max_shift"]),
                            nflag.ctypes.data_as(C.c_void_p), nsam.ctypes.data_as(C.c_void_p))
    return nflag, nsam


def is_bimera(sq, pars, **opts):
    o = dict(BIMERA_DEFAULTS); o.update(opts)
    arr = (C.c_char_p * len(pars))(*[s.encode() for s in pars])
    return bool(lib().port_is_bimera(sq.encode(), C.c_int(len(pars)), arr, C.c_int(o["allow_one_off"]),
                                     C.c_int(o["min_one_off_par_dist"]), C.c_int(o["match"]), C.c_int(o["mismatch"]),
                                     C.c_int(o["gap_p"]), C.c_int(o["max_shift"])))


def bimera_pair(sq, par, **opts):
    o = dict(BIMERA_DEFAULTS); o.update(opts)
    out = np.zeros(5, np.int32)
    n = len(sq) + len(par) + 2
    a0, a1 = C.create_string_buffer(n), C.create_string_buffer(n)
    lib().port_bimera_pair(sq.encode(), par.encode(), C.c_int(o["allow_one_off"]), C.c_int(o["match"]), C.c_int(o["mismatch"]),
                           C.c_int(o["gap_p"]), C.c_int(o["max_shift"]), out.ctypes.data_as(C.c_void_p), a0, a1)
    return dict(left=int(out[0]), right=int(out[1]), left_oo=int(out[2]), right_oo=int(out[3]), ham=int(out[4]),
                al0=a0.value.decode(), al1=a1.value.decode())


MERGE_DEFAULTS = dict(match=1, mismatch=-64, gap_p=-64, homo_gap_p=None, band=-1, prefer=1, trim_overhang=False)


def merge_pair(s1, s2, **opts):
    """Restatement of C_nwalign(endsfree) -> C_eval_pair -> C_pair_consensus (evaluate.cpp:18-174)."""
    o = dict(MERGE_DEFAULTS); o.update(opts)
    hg = o["gap_p"] if o["homo_gap_p"] is None else o["homo_gap_p"]
    n = len(s1) + len(s2) + 2
    cnt = np.zeros(3, np.int32)
    cons, a0, a1 = C.create_string_buffer(n), C.create_string_buffer(n), C.create_string_buffer(n)
    rc = lib().port_merge_pair(s1.encode(), s2.encode(), C.c_int(o["match"]), C.c_int(o["mismatch"]), C.c_int(o["gap_p"]), C.c_int(hg),
                               C.c_int(o["band"]), C.c_int(o["prefer"]), C.c_int(o["trim_overhang"]), cnt.ctypes.data_as(C.c_void_p), cons, a0, a1)
    if rc:
        raise RuntimeError("port_merge_pair failed")
    return dict(nmatch=int(cnt[0]), nmismatch=int(cnt[1]), nindel=int(cnt[2]), sequence=cons.value.decode(), al0=a0.value.decode(), al1=a1.value.decode())
