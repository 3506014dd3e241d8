"""Arbitrary-precision Poisson upper tail P(X >= k | E) by direct summation (mpmath).

Independent ground truth for the ppois restatements (oracle/rmath_ppois.c and the
device function in dada2_b200/csrc/ppois.cuh).  Used by tools/make_golden.py.
"""
import mpmath as mp


def upper_tail(k, E, dps=60):
    """P(X >= k) for X ~ Poisson(E); k integer >= 0, E > 0 (mp.mpf)."""
    mp.mp.dps = dps
    E = mp.mpf(E)
    if k <= 0:
        return mp.mpf(1)
    eps = mp.mpf(10) ** (-(dps - 5))
    if E <= k:
        t = mp.exp(-E + k * mp.log(E) - mp.loggamma(k + 1))
        s = t
        j = k
        while True:
            j += 1
            t = t * E / j
            s += t
            if t < eps * s:
                break
        return s
    # complement: sum_{j<k}, running downwards from k-1
    j = k - 1
    t = mp.exp(-E + j * mp.log(E) - mp.loggamma(j + 1))
    s = t
    while j > 0:
        t = t * j / E
        j -= 1
        s += t
        if t < eps * s:
            break
    return 1 - s
