"""R-free readers for the reference's fixtures (used only to GENERATE tests/golden/*).

* read_err_rda(path): the 16xQ error matrix stored in data/{tperr1,errBalancedF,...}.rda
  (gzip'd R XDR serialisation; the REALSXP payload is found by scanning for the
  (16,41) double block whose columns sum to 1 per source nucleotide).
* derep_fastq(path): restatement of derepFastq()/qtables2() semantics for a single
  chunk (/root/reference/R/sequenceIO.R:45-124,150-183): uniques sorted by decreasing
  abundance, ties in lexical order; per-position mean quality (double) per unique.
"""
import gzip
import struct
import numpy as np


def read_err_rda(path, nrow=16, ncol=41):
    raw = gzip.open(path, "rb").read()
    n = nrow * ncol
    # REALSXP header: type word (0x0000??0E) then int32 length n, big-endian
    key = struct.pack(">i", n)
    off = 0
    while True:
        off = raw.find(key, off)
        if off < 0:
            raise ValueError("no %d-double vector in %s" % (n, path))
        start = off + 4
        if start + 8 * n <= len(raw):
            v = np.frombuffer(raw[start:start + 8 * n], dtype=">f8").astype(np.float64)
            m = v.reshape((ncol, nrow)).T  # column-major in R
            if np.all(np.isfinite(m)) and np.all(m >= 0) and np.all(m <= 1):
                s = m.reshape(4, 4, ncol).sum(axis=1)
                if np.allclose(s, 1.0, atol=1e-6):
                    return np.ascontiguousarray(m)
        off += 1


def read_fastq(path):
    op = gzip.open if path.endswith(".gz") else open
    seqs, quals = [], []
    with op(path, "rt") as f:
        while True:
            h = f.readline()
            if not h:
                break
            s = f.readline().strip()
            f.readline()
            q = f.readline().strip()
            seqs.append(s)
            quals.append(np.frombuffer(q.encode(), dtype=np.uint8).astype(np.int64) - 33)
    return seqs, quals


def derep_fastq(path):
    """-> (uniques list[str], abundances int32[n], quals float64[n, maxlen] NaN-padded, map int32[nreads])"""
    seqs, quals = read_fastq(path)
    table = {}
    for i, s in enumerate(seqs):
        table.setdefault(s, []).append(i)
    # sort: decreasing abundance, ties lexical (srsort order then stable order(decreasing))
    uniq = sorted(table.keys(), key=lambda s: (-len(table[s]), s))
    maxlen = max(len(s) for s in uniq)
    ab = np.array([len(table[s]) for s in uniq], dtype=np.int32)
    q = np.full((len(uniq), maxlen), np.nan, dtype=np.float64)
    rmap = np.zeros(len(seqs), dtype=np.int32)
    for u, s in enumerate(uniq):
        idx = table[s]
        acc = np.zeros(len(s), dtype=np.float64)
        for i in idx:
            acc += quals[i]
            rmap[i] = u
        q[u, :len(s)] = acc / len(idx)   # rowsum(...)/abundance
    return uniq, ab, q, rmap
