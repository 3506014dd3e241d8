"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's dereplication, derepFastq() / qtables2()
(/root/reference/R/sequenceIO.R:45-124, :150-183), SURVEY.md 8(f1).

PARITY UNPINNED against the reference itself: derepFastq is R code on top of ShortRead (FastqStreamer, srsort, srrank,
tables) and neither R nor ShortRead exists in the build image, so this restatement cannot be executed side by side with
it.  What pins it: (1) on the reference's own fixture inst/extdata/sam1F.fastq.gz it yields the 896 uniques / 1 500 reads
that the reference's dada() run on that file is known to see and reproduces the committed
config-1 input (tests/golden/config1_sam1F_input.npz) exactly; (2) the semantics below are restated line by line.

Semantics restated (numpy):
  * reads are taken in chunks of `n` (FastqStreamer(fl, n = n), :56-57, :62); zero-length reads are ignored, their map
    entry is NA (:152-157, :171-175);
  * within a chunk the uniques come in srsort's lexical order (A < C < G < T, a proper prefix sorts first, :159-166), with
    their abundances (tabulate of srrank, :161) and per-position quality SUMS (rowsum of the quality matrix, :177-180);
  * uniques first seen in a later chunk are appended after all earlier ones, in that chunk's lexical order (:76-91);
    abundances and quality sums of already-seen uniques are added (:82-86);
  * mean quality = sum / abundance (:95); then `order(derepCounts, decreasing=TRUE)` (:98) -- R's default radix order is
    stable, so ties keep the appended order (first chunk of appearance, then lexical);
  * map = unique index (1-based) of every read (:92, :101).
Outputs: quals is [nuniq, maxlen] float64 with NaN beyond a unique's length (what dada() passes on as t(derep$quals)).
"""
import numpy as np


def derep_reads(seqs, quals, n=1000000):
    """seqs: list[str]; quals: list of integer arrays (numeric quality per base, Phred offset already removed).
    -> dict(uniques list[str], abundances int32[nuniq], quals float64[nuniq, maxlen], map int32[nreads] 1-based, NA = INT32_MIN)"""
    nreads = len(seqs)
    if not any(len(s) > 0 for s in seqs):
        raise RuntimeError("Only zero-length sequences detected during dereplication.")        # :153
    order_of = {}                      # sequence -> appended position
    names, counts, qsums = [], [], []
    rmap = np.full(nreads, np.iinfo(np.int32).min, dtype=np.int64)
    for c0 in range(0, nreads, int(n)):
        idx = [i for i in range(c0, min(nreads, c0 + int(n))) if len(seqs[i]) > 0]
        chunk = {}
        for i in idx:
            chunk.setdefault(seqs[i], []).append(i)
        for s in sorted(chunk):        # str order: A < C < G < T, proper prefix first == srsort on A/C/G/T reads
            rd = chunk[s]
            qs = np.zeros(len(s), dtype=np.float64)
            for i in rd:
                qs += np.asarray(quals[i], dtype=np.float64)
            u = order_of.get(s)
            if u is None:
                u = order_of[s] = len(names)
                names.append(s); counts.append(len(rd)); qsums.append(qs)
            else:
                counts[u] += len(rd); qsums[u] += qs
            rmap[rd] = u
    counts = np.asarray(counts, dtype=np.int64)
    ord_ = np.argsort(-counts, kind="stable")                                                   # :98
    rank = np.empty(len(ord_), dtype=np.int64); rank[ord_] = np.arange(len(ord_))
    maxlen = max(len(s) for s in names)
    q = np.full((len(names), maxlen), np.nan)
    for r, u in enumerate(ord_):
        q[r, :len(names[u])] = qsums[u] / counts[u]                                             # :95
    m = np.where(rmap >= 0, rank[np.maximum(rmap, 0)] + 1, np.iinfo(np.int32).min).astype(np.int32)
    return {"uniques": [names[u] for u in ord_], "abundances": counts[ord_].astype(np.int32), "quals": q, "map": m}
