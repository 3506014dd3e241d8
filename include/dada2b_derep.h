/* dada2b_derep.h -- C-ABI of the B200-native dereplication step (libdada2b.so).
 *
 * the step that feeds dada().  Drop-in boundary for what derepFastq() computes once ShortRead has parsed a
 * fastq file into reads and per-base qualities -- the work of qtables2() and of the chunk-merging loop around it
 * (/root/reference/R/sequenceIO.R:45-124, :150-183): unique sequences, their abundances, per-position mean qualities and the
 * read -> unique map, in derepFastq's output order.  Reading / decompressing the fastq stays on the host (ShortRead); the R
 * shim in INTEGRATION.md 2d hands the parsed reads over in one call.
 *
 * Output conventions are the reference's: uniques by decreasing abundance, ties in the order derepFastq leaves them (first
 * FastqStreamer chunk of `chunk_n` reads in which the sequence appears, then lexical A < C < G < T with a proper prefix
 * first); `quals` holds the per-position means as doubles, NA_real_ beyond a unique's length, laid out maxlen x nuniq
 * column-major (position fastest) -- i.e. t(derep$quals), the matrix dada() passes to dada_uniques (R/dada.R:339), so the
 * result can be fed to dada2b_run unchanged; `map` is 1-based, NA_integer_ for zero-length reads (sequenceIO.R:171-175).
 * Reads must be A/C/G/T (they are 2-bit packed on the device; dada() rejects anything else anyway, R/dada.R:269).
 * Returns 0 on success, non-zero with a message in errbuf; nothing throws across the ABI.  There is no CPU path.
 */
#ifndef DADA2B_DEREP_H
#define DADA2B_DEREP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DADA2B_ERRLEN
#define DADA2B_ERRLEN 256
#endif

typedef struct {
  int32_t nreads;
  const char *seq_concat;       /* concatenated A/C/G/T bytes of all reads                                         */
  const int64_t *seq_off;       /* nreads + 1 offsets into seq_concat (and into qual_concat)                       */
  const uint8_t *qual_concat;   /* numeric quality per base (Phred offset already removed, as(quality, "matrix"))  */
  int64_t chunk_n;              /* derepFastq's n (default 1e6): reads per FastqStreamer chunk; only decides the
                                   order of equally abundant uniques; <= 0 means one chunk                          */
} dada2b_derep_in;

typedef struct {
  int32_t nuniq, maxlen, nreads;
  char *seq_concat;             /* unique sequences, concatenated ...                                              */
  int64_t *seq_off;             /* ... nuniq + 1 offsets                                                           */
  int32_t *abund;               /* derep$uniques                                                                   */
  double *quals;                /* maxlen x nuniq column-major == t(derep$quals); NA_real_ padded                  */
  int32_t *map;                 /* derep$map, 1-based; NA_integer_ for zero-length reads                           */
  /* work counters and timings (information only) */
  int64_t gpu_launches, h2d_bytes, d2h_bytes;
  double ms_device, ms_sort, ms_total;
} dada2b_derep_out;

int dada2b_derep(const dada2b_derep_in *in, int32_t device, dada2b_derep_out **out, char errbuf[DADA2B_ERRLEN]);

/* Dereplicate and leave the uniques packed and RESIDENT on the device as a dada2b_ctx (include/dada2b.h) for
 * dada2b_run_resident: the quality means are rounded to dada()'s uint8 on the device ((uint8_t)round(q), containers.cpp:34)
 * and never travel to the host and back, the sequences are re-packed on the device.  `out` receives the same result as
 * dada2b_derep except that `quals` is NULL when want_quals == 0 (derep$quals is then never materialised on the host).
 * Free the context with dada2b_ctx_free.  A dada() run on the context equals dada2b_run on dada2b_derep's output. */
struct dada2b_ctx;
int dada2b_derep_resident(const dada2b_derep_in *in, int32_t device, int32_t want_quals, dada2b_derep_out **out,
                          struct dada2b_ctx **ctx, char errbuf[DADA2B_ERRLEN]);
void dada2b_derep_free(dada2b_derep_out *out);

#ifdef __cplusplus
}
#endif
#endif
