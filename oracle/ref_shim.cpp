// TEST INFRASTRUCTURE ONLY (parity oracle) -- never linked into the product.
//
// C-ABI wrapper around the UNMODIFIED reference translation units compiled from
// /root/reference/src (see oracle/Makefile).  Exposes
//   * ref_run(): the reference's dada_uniques() (Rmain.cpp:30) end to end;
//   * pair-level hooks: raw_align/sub_new/compute_lambda_ts, the three NW variants,
//     the k-mer distances -- for differential tests of individual CUDA kernels.
// Outputs are read back through ref_field_* getters (flat arrays).
#include "dada.h"
#include <atomic>

Rcpp::List dada_uniques(std::vector<std::string> seqs, std::vector<int> abundances, std::vector<bool> priors,
                        Rcpp::NumericMatrix err, Rcpp::NumericMatrix quals, int match, int mismatch, int gap,
                        bool use_kmers, double kdist_cutoff, int band_size, double omegaA, double omegaP,
                        double omegaC, bool detect_singletons, int max_clust, double min_fold, int min_hamming,
                        int min_abund, bool use_quals, bool final_consensus, bool vectorized_alignment,
                        int homo_gap, bool multithread, bool verbose, int SSE, bool gapless, bool greedy);

Rcpp::DataFrame C_table_bimera2(Rcpp::IntegerMatrix mat, std::vector<std::string> seqs, double min_fold, int min_abund,
                                bool allow_one_off, int min_one_off_par_dist, int match, int mismatch, int gap_p, int max_shift);
bool C_is_bimera(std::string sq, std::vector<std::string> pars, bool allow_one_off, int min_one_off_par_dist, int match,
                 int mismatch, int gap_p, int max_shift);
Rcpp::CharacterVector C_nwalign(std::string s1, std::string s2, int match, int mismatch, int gap_p, int homo_gap_p, int band, bool endsfree);
Rcpp::IntegerVector C_eval_pair(std::string s1, std::string s2);
Rcpp::CharacterVector C_pair_consensus(std::string s1, std::string s2, int prefer, bool trim_overhang);
int get_ham_endsfree(const char *seq1, const char *seq2);
void get_lr(char **al, int &left, int &right, int &left_oo, int &right_oo, bool allow_one_off, int max_shift);

static std::atomic<int> g_threads(1);
extern "C" int oracle_get_threads(void) { return g_threads.load(); }
extern "C" void oracle_set_threads(int n) { g_threads.store(n < 1 ? 1 : n); }

struct RefResult {
  Rcpp::List out;
  std::string err;
};

static const Rcpp::Any *find(RefResult *r, const char *outer, const char *inner) {
  const Rcpp::Any *a = r->out.get(outer);
  if (!a) return nullptr;
  if (inner && inner[0]) {
    if (!a->list) return nullptr;
    return a->list->get(inner);
  }
  return a;
}

extern "C" {

// seqs: nraw NUL-terminated ACGT strings; quals: maxlen x nraw column-major doubles
// (NULL => no quals, nrow 0); err: 16 x ncol_err column-major.
void *ref_run(int nraw, const char **seqs, const int *abund, const unsigned char *prior, const double *err,
              int ncol_err, const double *quals, int maxlen, int match, int mismatch, int gap, int use_kmers,
              double kdist_cutoff, int band_size, double omegaA, double omegaP, double omegaC,
              int detect_singletons, int max_clust, double min_fold, int min_hamming, int min_abund,
              int use_quals, int final_consensus, int vectorized_alignment, int homo_gap, int multithread,
              int verbose, int SSE, int gapless, int greedy) {
  RefResult *res = new RefResult();
  try {
    std::vector<std::string> s(nraw);
    std::vector<int> ab(nraw);
    std::vector<bool> pr(nraw);
    for (int i = 0; i < nraw; i++) { s[i] = seqs[i]; ab[i] = abund[i]; pr[i] = prior ? prior[i] != 0 : false; }
    Rcpp::NumericMatrix E(16, ncol_err);
    for (int i = 0; i < 16 * ncol_err; i++) (*E.d)[i] = err[i];
    Rcpp::NumericMatrix Q = quals ? Rcpp::NumericMatrix(maxlen, nraw) : Rcpp::NumericMatrix();
    if (quals) for (size_t i = 0; i < (size_t)maxlen * nraw; i++) (*Q.d)[i] = quals[i];
    res->out = dada_uniques(s, ab, pr, E, Q, match, mismatch, gap, use_kmers != 0, kdist_cutoff, band_size,
                            omegaA, omegaP, omegaC, detect_singletons != 0, max_clust, min_fold, min_hamming,
                            min_abund, use_quals != 0, final_consensus != 0, vectorized_alignment != 0, homo_gap,
                            multithread != 0, verbose != 0, SSE, gapless != 0, greedy != 0);
  } catch (std::exception &e) {
    res->err = e.what();
    if (res->err.empty()) res->err = "error";
  }
  return res;
}
const char *ref_error(void *h) { RefResult *r = (RefResult *)h; return r->err.empty() ? nullptr : r->err.c_str(); }
void ref_free(void *h) { delete (RefResult *)h; }

int ref_field_len(void *h, const char *outer, const char *inner) {
  const Rcpp::Any *a = find((RefResult *)h, outer, inner);
  if (!a) return -1;
  if (a->iv) return (int)a->iv->size();
  if (a->nv) return (int)a->nv->size();
  if (a->sv) return (int)a->sv->size();
  return -1;
}
int ref_field_dims(void *h, const char *outer, int *nr, int *nc) {
  const Rcpp::Any *a = find((RefResult *)h, outer, nullptr);
  if (!a) return -1;
  *nr = a->nr; *nc = a->nc; return 0;
}
int ref_field_int(void *h, const char *outer, const char *inner, int *out) {
  const Rcpp::Any *a = find((RefResult *)h, outer, inner);
  if (!a || !a->iv) return -1;
  memcpy(out, a->iv->data(), a->iv->size() * sizeof(int));
  return (int)a->iv->size();
}
int ref_field_dbl(void *h, const char *outer, const char *inner, double *out) {
  const Rcpp::Any *a = find((RefResult *)h, outer, inner);
  if (!a || !a->nv) return -1;
  memcpy(out, a->nv->data(), a->nv->size() * sizeof(double));
  return (int)a->nv->size();
}
const char *ref_field_str(void *h, const char *outer, const char *inner, int i) {
  const Rcpp::Any *a = find((RefResult *)h, outer, inner);
  if (!a || !a->sv || i < 0 || i >= (int)a->sv->size()) return nullptr;
  return (*a->sv)[i].c_str();
}

// ---------------- pair-level hooks ----------------
// Sequences come in as ACGT text; converted with the reference's nt2int.
static Raw *mk_raw(const char *seq, const unsigned char *qual, int use_kmers, unsigned maxlen,
                   std::vector<void *> &frees) {
  size_t n = strlen(seq);
  char *buf = (char *)malloc(n + 1);
  strcpy(buf, seq);
  nt2int(buf, buf);
  std::vector<double> q(n);
  if (qual) for (size_t i = 0; i < n; i++) q[i] = qual[i];
  Raw *raw = raw_new(buf, qual ? q.data() : NULL, 1, false);
  free(buf);
  raw->kmer8 = NULL; raw->kmer = NULL; raw->kord = NULL;
  if (use_kmers) {
    size_t n_kmer = 1 << (2 * KMER_SIZE);
    raw->kmer8 = (uint8_t *)malloc(n_kmer);
    raw->kmer = (uint16_t *)malloc(n_kmer * 2);
    raw->kord = (uint16_t *)calloc(maxlen, 2);
    frees.push_back(raw->kmer8); frees.push_back(raw->kmer); frees.push_back(raw->kord);
    assign_kmer8(raw->kmer8, raw->seq, KMER_SIZE);
    assign_kmer(raw->kmer, raw->seq, KMER_SIZE);
    assign_kmer_order(raw->kord, raw->seq, KMER_SIZE);
  }
  return raw;
}

// Full reference sub_new + compute_lambda_ts for one (centre=seq0, raw=seq1) pair.
// Returns 0 ok, 1 = NULL sub (shrouded), <0 error.  al0/al1 (capacity len0+len1+1) receive
// the alignment actually used (ACGT/-), map has len0 entries.
int ref_pair(const char *seq0, const unsigned char *q0, const char *seq1, const unsigned char *q1,
             const double *err_rowmajor, int ncol, int match, int mismatch, int gap, int homo_gap,
             int use_kmers, double kdist_cutoff, int band, int vectorized, int SSE, int gapless,
             double *lambda, int *nsubs, unsigned short *map, unsigned short *pos, char *nt0, char *nt1,
             unsigned char *sq0, unsigned char *sq1, char *al0, char *al1, double *kdist, double *kodist,
             char *errbuf) {
  std::vector<void *> frees;
  int rc = 0;
  try {
    unsigned maxlen = (unsigned)std::max(strlen(seq0), strlen(seq1));
    Raw *r0 = mk_raw(seq0, q0, use_kmers, maxlen, frees);
    Raw *r1 = mk_raw(seq1, q1, use_kmers, maxlen, frees);
    if (use_kmers) {
      double kd = kmer_dist_SSEi_8(r0->kmer8, r0->length, r1->kmer8, r1->length, KMER_SIZE);
      if (kd < 0) kd = kmer_dist_SSEi(r0->kmer, r0->length, r1->kmer, r1->length, KMER_SIZE);
      if (kdist) *kdist = kd;
      if (kodist) *kodist = kord_dist_SSEi(r0->kord, r0->length, r1->kord, r1->length, KMER_SIZE);
    }
    char **al = raw_align(r0, r1, match, mismatch, gap, homo_gap, use_kmers != 0, kdist_cutoff, band,
                          vectorized != 0, SSE, gapless != 0);
    if (al) {
      if (al0) { strcpy(al0, al[0]); int2nt(al0, al0); }
      if (al1) { strcpy(al1, al[1]); int2nt(al1, al1); }
      free(al[0]); free(al[1]); free(al);
    }
    Sub *sub = sub_new(r0, r1, match, mismatch, gap, homo_gap, use_kmers != 0, kdist_cutoff, band,
                       vectorized != 0, SSE, gapless != 0);
    *lambda = compute_lambda_ts(r1, sub, ncol, (double *)err_rowmajor, q1 != NULL);
    if (!sub) { rc = 1; *nsubs = -1; }
    else {
      *nsubs = sub->nsubs;
      if (map) memcpy(map, sub->map, sub->len0 * sizeof(uint16_t));
      for (unsigned s = 0; s < sub->nsubs; s++) {
        if (pos) pos[s] = sub->pos[s];
        if (nt0) nt0[s] = sub->nt0[s];
        if (nt1) nt1[s] = sub->nt1[s];
        if (sq0 && sub->q0) sq0[s] = sub->q0[s];
        if (sq1 && sub->q1) sq1[s] = sub->q1[s];
      }
      sub_free(sub);
    }
    raw_free(r0); raw_free(r1);
  } catch (std::exception &e) {
    if (errbuf) { strncpy(errbuf, e.what(), 255); errbuf[255] = 0; }
    rc = -1;
  }
  for (void *p : frees) free(p);
  return rc;
}

// Direct aligners.  mode: 0 = nwalign_vectorized2 (end_gap 0), 1 = nwalign_endsfree,
// 2 = nwalign_endsfree_homo, 3 = nwalign_gapless, 4 = nwalign (global), 5 = vectorized2 non-endsfree.
int ref_align(const char *seq0, const char *seq1, int match, int mismatch, int gap, int homo_gap, int band,
              int mode, char *al0, char *al1) {
  size_t l0 = strlen(seq0), l1 = strlen(seq1);
  char *s0 = (char *)malloc(l0 + 1), *s1 = (char *)malloc(l1 + 1);
  strcpy(s0, seq0); strcpy(s1, seq1);
  nt2int(s0, s0); nt2int(s1, s1);
  int score[4][4];
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) score[i][j] = i == j ? match : mismatch;
  char **al = NULL;
  int rc = 0;
  try {
    switch (mode) {
      case 0: al = nwalign_vectorized2(s0, l0, s1, l1, (int16_t)match, (int16_t)mismatch, (int16_t)gap, 0, band); break;
      case 1: al = nwalign_endsfree(s0, l0, s1, l1, score, gap, band); break;
      case 2: al = nwalign_endsfree_homo(s0, l0, s1, l1, score, gap, homo_gap, band); break;
      case 3: al = nwalign_gapless(s0, l0, s1, l1); break;
      case 4: al = nwalign(s0, l0, s1, l1, score, gap, band); break;
      case 5: al = nwalign_vectorized2(s0, l0, s1, l1, (int16_t)match, (int16_t)mismatch, (int16_t)gap, (int16_t)gap, band); break;
      default: rc = -2;
    }
  } catch (std::exception &e) { rc = -1; }
  if (al) {
    strcpy(al0, al[0]); int2nt(al0, al0);
    strcpy(al1, al[1]); int2nt(al1, al1);
    free(al[0]); free(al[1]); free(al);
  }
  free(s0); free(s1);
  return rc;
}

// ---------------- bimera detection (src/chimera.cpp) ----------------
// C_table_bimera2 (chimera.cpp:194-207): mat is nrow (samples) x ncol (sequences), column-major.
int ref_table_bimera(int nrow, int ncol, const int *mat, const char **seqs, double min_fold, int min_abund,
                     int allow_one_off, int min_one_off_par_dist, int match, int mismatch, int gap_p, int max_shift,
                     int *nflag, int *nsam, char *errbuf) {
  try {
    Rcpp::IntegerMatrix M(nrow, ncol);
    for (size_t i = 0; i < (size_t)nrow * ncol; i++) (*M.d)[i] = mat[i];
    std::vector<std::string> s(ncol);
    for (int i = 0; i < ncol; i++) s[i] = seqs[i];
    Rcpp::DataFrame df = C_table_bimera2(M, s, min_fold, min_abund, allow_one_off != 0, min_one_off_par_dist, match,
                                         mismatch, gap_p, max_shift);
    const Rcpp::Any *f = df.get("nflag"), *n = df.get("nsam");
    memcpy(nflag, f->iv->data(), ncol * sizeof(int));
    memcpy(nsam, n->iv->data(), ncol * sizeof(int));
  } catch (std::exception &e) {
    if (errbuf) { strncpy(errbuf, e.what(), 255); errbuf[255] = 0; }
    return -1;
  }
  return 0;
}
// C_is_bimera (chimera.cpp:18-59)
int ref_is_bimera(const char *sq, int npar, const char **pars, int allow_one_off, int min_one_off_par_dist, int match,
                  int mismatch, int gap_p, int max_shift) {
  std::vector<std::string> p(npar);
  for (int i = 0; i < npar; i++) p[i] = pars[i];
  try {
    return C_is_bimera(std::string(sq), p, allow_one_off != 0, min_one_off_par_dist, match, mismatch, gap_p, max_shift) ? 1 : 0;
  } catch (std::exception &e) { return -1; }
}
// One (query, parent) pair: the alignment of chimera.cpp:122 followed by get_lr (:239-269) and get_ham_endsfree (:210-236).
int ref_bimera_pair(const char *sq, const char *par, int allow_one_off, int match, int mismatch, int gap_p, int max_shift,
                    int *out5, char *al0, char *al1) {
  try {
    char **al = nwalign_vectorized2(sq, strlen(sq), par, strlen(par), (int16_t)match, (int16_t)mismatch, (int16_t)gap_p, 0, max_shift);
    int left = 0, right = 0, left_oo = 0, right_oo = 0;
    get_lr(al, left, right, left_oo, right_oo, allow_one_off != 0, max_shift);
    out5[0] = left; out5[1] = right; out5[2] = left_oo; out5[3] = right_oo; out5[4] = get_ham_endsfree(al[0], al[1]);
    if (al0) strcpy(al0, al[0]);
    if (al1) strcpy(al1, al[1]);
    free(al[0]); free(al[1]); free(al);
  } catch (std::exception &e) { return -1; }
  return 0;
}

// ---------------- mergePairs' native steps (src/evaluate.cpp) ----------------
// R/paired.R:153-164 for one pair: C_nwalign(s1, s2, ...) -> C_eval_pair(al0, al1) -> C_pair_consensus(al0, al1, prefer, trim).
// counts3 = match, mismatch, indel; cons / al0 / al1 need capacity len1 + len2 + 1.
int ref_merge_pair(const char *s1, const char *s2, int match, int mismatch, int gap_p, int homo_gap_p, int band, int endsfree,
                   int prefer, int trim_overhang, int *counts3, char *cons, char *al0, char *al1) {
  try {
    Rcpp::CharacterVector al = C_nwalign(s1, s2, match, mismatch, gap_p, homo_gap_p, band, endsfree != 0);
    Rcpp::IntegerVector ev = C_eval_pair(al[0], al[1]);
    counts3[0] = ev[0]; counts3[1] = ev[1]; counts3[2] = ev[2];
    Rcpp::CharacterVector c = C_pair_consensus(al[0], al[1], prefer, trim_overhang != 0);
    strcpy(cons, c[0].c_str());
    if (al0) strcpy(al0, al[0].c_str());
    if (al1) strcpy(al1, al[1].c_str());
  } catch (std::exception &e) { return -1; }
  return 0;
}

double ref_ppois_upper(int reads_minus_1, double E) { return oracle_ppois((double)reads_minus_1, E, 0, 0); }
double ref_calc_pA(int reads, double E_reads, int prior) { return calc_pA(reads, E_reads, prior != 0); }

}  // extern "C"
