/* dada2b_merge.h -- C-ABI of the B200-native alignment / evaluation / consensus step of mergePairs (libdada2b.so).
 *
 * SURVEY.md 8(f4).  mergePairs (R/paired.R:150-168) runs, for every unique (forward ASV, reverse ASV) pairing of a sample,
 *     alvec <- nwalign(F, rc(R), band=-1)          -> C_nwalign(s1, s2, match, mismatch, gap_p, homo_gap_p, band, endsfree=TRUE)
 *                                                                                  /root/reference/src/evaluate.cpp:18-61
 *     C_eval_pair(alvec[1], alvec[2])              -> (match, mismatch, indel)     /root/reference/src/evaluate.cpp:73-120
 *     C_pair_consensus(alvec[1], alvec[2], prefer, trimOverhang)                   /root/reference/src/evaluate.cpp:131-174
 * one R-level call at a time (`mapply`); .Call stubs at src/RcppExports.cpp (C_nwalign, C_eval_pair, C_pair_consensus).
 * This boundary takes the whole batch of pairings in one call and fuses the three steps on the device; the aligned
 * strings themselves never leave it.  Everything else of mergePairs (pair tabulation, rc(), accept rule, sorting) stays in R.
 * Scores are the caller's (mergePairs sets MATCH 1, MISMATCH / GAP -64 for maxMismatch = 0, else -8: R/paired.R:153-157).
 * Sequences are A/C/G/T text (nwalign() rejects anything else, R/misc.R:186-188).  Returns 0 on success, non-zero with a
 * message in errbuf; nothing throws across the ABI.  There is no CPU path.
 */
#ifndef DADA2B_MERGE_H
#define DADA2B_MERGE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DADA2B_ERRLEN
#define DADA2B_ERRLEN 256
#endif

typedef struct {
  int32_t match, mismatch, gap_p;  /* C_nwalign's match / mismatch / gap_p                                        */
  int32_t homo_gap_p;              /* C_nwalign's homo_gap_p; == gap_p selects nwalign_endsfree, else ..._homo    */
  int32_t band;                    /* C_nwalign's band; -1 = unbanded (what mergePairs uses)                       */
  int32_t trim_overhang;           /* C_pair_consensus' trim_overhang (mergePairs' trimOverhang, default FALSE)    */
} dada2b_merge_opts;

typedef struct {
  int32_t npairs;
  int32_t *nmatch, *nmismatch, *nindel;   /* C_eval_pair per pairing                                              */
  char *cons_concat;                      /* C_pair_consensus per pairing, concatenated (no terminators) ...       */
  int64_t *cons_off;                      /* ... npairs + 1 offsets                                                */
  /* work counters and timings (information only) */
  int64_t n_cells, gpu_launches, h2d_bytes, d2h_bytes;
  double ms_device, ms_k_merge, ms_total;
} dada2b_merge_out;

void dada2b_merge_default_opts(dada2b_merge_opts *opts);   /* 1 / -64 / -64 / -64, band -1, trim_overhang 0 */

/* Pairing x aligns sequence s1_idx[x] (the forward ASV) with sequence s2_idx[x] (the reverse ASV, already
 * reverse-complemented by the caller as in R/paired.R:140) of one pool of nseq sequences; prefer[x] is 1 or 2
 * (R/paired.R:162); outputs are library-owned until dada2b_merge_free. */
int dada2b_merge_pairs(int32_t nseq, const char *seq_concat, const int64_t *seq_off, int32_t npairs, const int32_t *s1_idx,
                       const int32_t *s2_idx, const int32_t *prefer, const dada2b_merge_opts *opts, int32_t device,
                       dada2b_merge_out **out, char errbuf[DADA2B_ERRLEN]);
void dada2b_merge_free(dada2b_merge_out *out);

#ifdef __cplusplus
}
#endif
#endif
