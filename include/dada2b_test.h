/* dada2b_test.h -- kernel-level entry points exported by libdada2b.so for the parity tests.
 * They run the SAME device kernels the dada2b_run() path uses (k_classify, k_align, calc_pA) on
 * caller-chosen pairs so that tests can diff individual stages against the oracle:
 *   raw_align / sub_new / compute_lambda_ts  (/root/reference/src/nwalign_endsfree.cpp:10-73,642-672; pval.cpp:144-197)
 *   calc_pA                                   (/root/reference/src/pval.cpp:44-64)
 */
#ifndef DADA2B_TEST_H
#define DADA2B_TEST_H
#include "dada2b.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Align npairs (centre[k], raw[k]) index pairs of an uploaded context.
 * kind: 0 shrouded (NULL sub), 1 gapless, 2 NW.  ops: opcap bytes per pair, alignment columns in
 * order, 1 = both bases, 2 = gap in the centre row, 3 = gap in the raw row.  Substitutions: subcap
 * entries per pair (pos0, nt0, nt1 as 0..3, q1). */
int dada2b_test_pairs(dada2b_ctx *ctx, int32_t npairs, const uint32_t *centre, const uint32_t *raw, const double *err,
                      int32_t Q, const dada2b_opts *opts, int32_t use_kmers, double kdist_cutoff, int32_t *kind,
                      double *lambda, int32_t *nsubs, uint8_t *ops, int32_t *nops, int32_t opcap, uint16_t *pos,
                      uint8_t *nt0, uint8_t *nt1, uint8_t *q1, int32_t subcap, char errbuf[DADA2B_ERRLEN]);

/* out[k] = calc_pA(reads[k], E[k], prior[k]) evaluated on the device. */
int dada2b_test_calc_pA(int32_t n, const int32_t *reads, const double *E, const int32_t *prior, double *out,
                        char errbuf[DADA2B_ERRLEN]);

/* Kernel-level hook for the LOOP aligners: every (centre[k], raw[k]) pair goes, alone, through one of the kernels that
 * produce b_compare's (lambda, hamming) without an alignment string (cluster.cpp:136-143):
 *   which = 0  k_nwrow<EXACT>   thread per pair, recorded moves + traceback        (dd_nwrow.cu)
 *   which = 1  k_nwlane         lane group per pair, software-pipelined rows        (dd_nwlane.cu)
 *   which = 2  k_nwfwd          lane-group anti-diagonal wavefront, carried lambda  (dd_nwfwd.cu)
 *   which = 3  k_nwrow<BOUND>   thread per pair, substitution count only (lambda[k] is left 0)
 * handled[k] = 0 when the kernel does not take the pair (lengths differ, band not instantiated, ...) and hands it on. */
int dada2b_test_loop_nw(dada2b_ctx *ctx, int32_t which, int32_t npairs, const uint32_t *centre, const uint32_t *raw, const double *err,
                        int32_t Q, const dada2b_opts *opts, double *lambda, int32_t *nsubs, int32_t *handled, char errbuf[DADA2B_ERRLEN]);

#ifdef __cplusplus
}
#endif
#endif
