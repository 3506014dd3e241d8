"""Synthetic dereplicated amplicon data (SURVEY.md 8d / BASELINE.json configs 2-5).

Illumina-like (config 2/3/4): `nvar` true variants of length L derived from one random
root (variant k = root with U{1..max_subs} substitutions; a fraction carries a 1-3 nt
indel and is re-truncated to L, so NW has real gaps to find), Zipf(s=1) abundances,
per-base qualities clip(round(profile(pos)+N(0,3)),2,40) with a 38->20 profile and 2 %
random low-q bases, substitution errors with P=10^(-q/10).  Reads are drawn until
exactly `n_uniques` distinct sequences exist, then dereplicated the way derepFastq does
(/root/reference/R/sequenceIO.R:95-101,150-183): uniques sorted by decreasing abundance
(ties lexical), per-position mean quality as double.

PacBio-like (config 5): length-variable ~1450-1550 nt variants, CCS-like qualities
(mostly 93), errors dominated by homopolymer indels.

Everything is seeded (numpy default_rng) and deterministic.
"""
import numpy as np

NT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _qual_profile(L):
    pos = np.arange(L, dtype=np.float64)
    return 38.0 - 18.0 * (pos / max(L - 1, 1)) ** 2


def _derep(reads, quals, n_uniques=None):
    """reads uint8 [n, L] (codes 0-3), quals uint8 [n, L] -> uniques sorted by (-abund, seq)."""
    n, L = reads.shape
    view = np.ascontiguousarray(reads).view(np.dtype((np.void, L))).ravel()
    uniq, first, inv, counts = np.unique(view, return_index=True, return_inverse=True, return_counts=True)
    if n_uniques is not None and len(uniq) > n_uniques:
        cutoff = np.sort(first)[n_uniques]
        return _derep(reads[:cutoff], quals[:cutoff], None)
    order = np.argsort(inv, kind="stable")
    starts = np.zeros(len(uniq), dtype=np.int64)
    np.cumsum(counts[:-1], out=starts[1:])
    qsum = np.add.reduceat(quals[order].astype(np.float64), starts, axis=0)
    qmean = qsum / counts[:, None]
    rank = np.argsort(-counts, kind="stable")  # uniq is already lexically sorted
    useq = uniq.view(np.uint8).reshape(len(uniq), L)[rank]
    return useq, counts[rank].astype(np.int32), qmean[rank]


def illumina(n_uniques, L=250, nvar=100, max_subs=60, indel_frac=0.1, seed=12345, chunk0=400000,
             max_reads=None, as_strings=True, lowq_frac=0.02):
    """-> (seqs, abundances int32[n], quals float64[n, L], truth dict)"""
    rng = np.random.default_rng(seed)
    root = rng.integers(0, 4, size=L + 16, dtype=np.uint8)
    variants = []
    for k in range(nvar):
        v = root.copy()
        ns = 0 if k == 0 else int(rng.integers(1, max_subs + 1))
        pos = rng.choice(L, size=ns, replace=False)
        v[pos] = (v[pos] + rng.integers(1, 4, size=ns, dtype=np.uint8)) % 4
        if k > 0 and rng.random() < indel_frac:
            p = int(rng.integers(10, L - 10))
            n = int(rng.integers(1, 4))
            if rng.random() < 0.5:
                v = np.concatenate([v[:p], v[p + n:]])          # deletion
            else:
                v = np.concatenate([v[:p], rng.integers(0, 4, size=n, dtype=np.uint8), v[p:]])  # insertion
        variants.append(v[:L].copy())
    variants = np.stack(variants)
    w = 1.0 / np.arange(1, nvar + 1)
    w /= w.sum()
    prof = _qual_profile(L)
    reads_l, quals_l = [], []
    total, have = 0, 0
    seen = set()
    while True:
        chunk = int(min(chunk0, max(2000, 1.6 * (n_uniques - have) + 1000)))
        vi = rng.choice(nvar, size=chunk, p=w)
        q = np.clip(np.rint(prof[None, :] + rng.normal(0.0, 3.0, size=(chunk, L))), 2, 40)
        low = rng.random(size=(chunk, L)) < lowq_frac
        q[low] = rng.integers(2, 15, size=int(low.sum()))
        q = q.astype(np.uint8)
        perr = 10.0 ** (-q.astype(np.float64) / 10.0)
        is_err = rng.random(size=(chunk, L)) < perr
        r = variants[vi]
        shift = rng.integers(1, 4, size=int(is_err.sum()), dtype=np.uint8)
        r[is_err] = (r[is_err] + shift) % 4
        reads_l.append(r)
        quals_l.append(q)
        total += chunk
        allr = np.concatenate(reads_l)
        view = allr.view(np.dtype((np.void, L))).ravel()
        have = len(np.unique(view))
        if have > n_uniques or (max_reads and total >= max_reads):
            break
    allq = np.concatenate(quals_l)
    useq, ab, qm = _derep(allr, allq, n_uniques)
    truth = {"variants": ["".join("ACGT"[c] for c in v) for v in variants], "weights": w, "n_reads": int(ab.sum())}
    if as_strings:
        seqs = [bytes(NT[row]).decode() for row in useq]
    else:
        seqs = useq
    return seqs, ab, qm, truth


def pacbio(n_uniques, L=1500, nvar=30, seed=777, chunk=20000):
    """Length-variable CCS-like uniques; errors are mostly homopolymer indels.
    -> (seqs list[str], abundances, quals float64[n, maxlen] NaN-padded)"""
    rng = np.random.default_rng(seed)
    root = rng.integers(0, 4, size=L, dtype=np.uint8)
    # seed some homopolymers
    for _ in range(L // 25):
        p = int(rng.integers(0, L - 8)); n = int(rng.integers(3, 8)); root[p:p + n] = root[p]
    variants = []
    for k in range(nvar):
        v = list(root)
        for _ in range(0 if k == 0 else int(rng.integers(2, 40))):
            p = int(rng.integers(0, len(v))); v[p] = (v[p] + int(rng.integers(1, 4))) % 4
        dl = int(rng.integers(-50, 51))
        if dl > 0:
            p = int(rng.integers(50, len(v) - 50)); v[p:p] = list(rng.integers(0, 4, size=dl))
        elif dl < 0:
            p = int(rng.integers(50, len(v) - 50 + dl)); del v[p:p - dl]
        variants.append(np.array(v, dtype=np.uint8))
    w = 1.0 / np.arange(1, nvar + 1); w /= w.sum()
    table = {}
    while len(table) < n_uniques:
        for vi in rng.choice(nvar, size=chunk, p=w):
            v = list(variants[vi])
            q = np.where(rng.random(len(v) + 8) < 0.9, 93, rng.integers(20, 93, size=len(v) + 8)).astype(np.int64)
            ne = rng.poisson(0.8)
            for _ in range(ne):
                p = int(rng.integers(1, len(v) - 1))
                if rng.random() < 0.8:   # homopolymer indel
                    if rng.random() < 0.5: v.insert(p, v[p])
                    else: del v[p]
                else:
                    v[p] = (v[p] + int(rng.integers(1, 4))) % 4
            s = bytes(NT[np.array(v, dtype=np.uint8)]).decode()
            qq = q[:len(v)]
            e = table.get(s)
            if e is None:
                if len(table) >= n_uniques: continue
                table[s] = [1, qq.astype(np.float64)]
            else:
                e[0] += 1; e[1] += qq
    uniq = sorted(table.keys(), key=lambda s: (-table[s][0], s))
    maxlen = max(len(s) for s in uniq)
    ab = np.array([table[s][0] for s in uniq], dtype=np.int32)
    qm = np.full((len(uniq), maxlen), np.nan)
    for i, s in enumerate(uniq):
        qm[i, :len(s)] = table[s][1] / table[s][0]
    return uniq, ab, qm


def extend_err(err, ncol):
    """R/dada.R:303-313: repeat the last column until the matrix has `ncol` columns."""
    err = np.asarray(err, dtype=np.float64)
    if err.shape[1] >= ncol:
        return err
    return np.concatenate([err, np.repeat(err[:, -1:], ncol - err.shape[1], axis=1)], axis=1)


def bimera_table(nseq, nsample=8, L=250, npar=None, seed=1, frac_bimera=0.4, frac_oneoff=0.15, frac_indel=0.1, frac_shift=0.05,
                 lenvar=0):
    """Synthetic sequence table for bimera detection (input of the reference's isBimeraDenovoTable, R/chimeras.R:220-238):
    -> (seqs, mat[nsample, nseq] int32).  `npar` true variants (substitution variants of one root, like illumina());
    the rest are two-parent bimeras at a random breakpoint (some with one extra substitution = "one-off", some with a
    1-3 nt indel, some with shifted ends) plus plain substitution variants.  Abundances: Zipf over sequences, log-normal
    sample depths, ~60 % of the cells zero for the rarer sequences.  Sequences are distinct; column order is by
    decreasing total abundance (what makeSequenceTable + dada() produce is unordered; the algorithm does not care)."""
    rng = np.random.default_rng(seed)
    npar = npar or max(4, nseq // 8)
    root = rng.integers(0, 4, L)
    pars = []
    for _ in range(npar):
        s = root.copy()
        k = int(rng.integers(1, 40))
        pos = rng.choice(L, k, replace=False)
        s[pos] = (s[pos] + rng.integers(1, 4, k)) % 4
        if lenvar:
            s = s[:L - int(rng.integers(0, lenvar + 1))]
        pars.append(s)
    out = {}

    def add(s):
        key = "".join("ACGT"[x] for x in s)
        if len(key) >= 8 and key not in out:
            out[key] = len(out)

    for s in pars:
        add(s)
    guard = 0
    while len(out) < nseq and guard < 50 * nseq:
        guard += 1
        u = rng.random()
        a, b = pars[int(rng.integers(0, npar))], pars[int(rng.integers(0, npar))]
        if u < frac_bimera + frac_oneoff + frac_indel + frac_shift:
            bp = int(rng.integers(10, min(len(a), len(b)) - 10))
            s = np.concatenate([a[:bp], b[bp:]])
            v = u - frac_bimera
            if 0 <= v < frac_oneoff:
                p = int(rng.integers(0, len(s)))
                s = s.copy(); s[p] = (s[p] + int(rng.integers(1, 4))) % 4
            elif frac_oneoff <= v < frac_oneoff + frac_indel:
                p = int(rng.integers(5, len(s) - 5)); k = int(rng.integers(1, 4))
                s = np.concatenate([s[:p], rng.integers(0, 4, k), s[p:]]) if rng.random() < 0.5 else np.concatenate([s[:p], s[p + k:]])
            elif v >= frac_oneoff + frac_indel:
                k = int(rng.integers(1, 20))
                s = s[k:] if rng.random() < 0.5 else np.concatenate([rng.integers(0, 4, k), s])
        else:
            s = a.copy()
            k = int(rng.integers(1, 4))
            pos = rng.choice(len(s), k, replace=False)
            s[pos] = (s[pos] + rng.integers(1, 4, k)) % 4
        add(s)
    seqs = list(out)
    n = len(seqs)
    w = 1.0 / np.arange(1, n + 1)
    depth = rng.lognormal(9.0, 0.6, nsample)
    lam = depth[:, None] * (w / w.sum())[None, :]
    mat = rng.poisson(lam).astype(np.int32)
    drop = rng.random((nsample, n)) < np.minimum(0.6, np.arange(n)[None, :] / max(1, n) + 0.05)
    mat[drop] = 0
    mat[0, mat.sum(axis=0) == 0] = 1            # every sequence occurs somewhere
    return seqs, mat
