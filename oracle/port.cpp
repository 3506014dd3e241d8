// TEST INFRASTRUCTURE ONLY (parity oracle) -- never linked into or called by the product.
//
// CPU restatement ("port") of the reference's dada_uniques()/run_dada() hot path in
// plain C++ with flat containers.  Every function cites the reference lines it
// restates (paths under /root/reference/src).  It is pinned against the reference's
// own code executed here (oracle/_ref/libdada2ref.so) by tests/test_oracle_*.py and
// against the committed goldens in tests/golden/.  The one piece of arithmetic that
// lives outside the reference, R's ppois, comes from oracle/rmath_ppois.c.
//
// Uses the product's public struct definitions (include/dada2b.h) so that tests can
// diff product output against port output field by field.
#include "../include/dada2b.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <string>
#include <vector>
#include <unordered_map>
#include <stdexcept>
#include <algorithm>

extern "C" double oracle_ppois(double x, double lambda, int lower_tail, int log_p);

namespace {

const int K = 5;                 // KMER_SIZE        dada.h:27
const int NKMER = 1 << (2 * K);  // 4^5 bins
const unsigned GAP_GLYPH = 9999; // dada.h:31
const int SEQLEN = 9999;         // dada.h:24
const int MAX_SHUFFLE = 10;      // dada.h:30
const double TAIL_APPROX_CUTOFF = 1e-7;  // dada.h:25

struct Stop : std::runtime_error { using std::runtime_error::runtime_error; };

double na_real() { union { double d; uint64_t u; } v; v.u = 0x7FF00000000007A2ULL; return v.d; }

struct Cmp { uint32_t i, index; double lambda; uint32_t hamming; };   // dada.h:42-47

struct Sub {                                                          // dada.h:53-62
  bool null = true;
  unsigned nsubs = 0, len0 = 0;
  std::vector<uint16_t> map, pos;
  std::vector<char> nt0, nt1;
  std::vector<uint8_t> q0, q1;
  bool has_q = false;
};

struct Raw {                                                          // dada.h:65-80
  std::string seq;              // codes 1..4
  std::vector<uint8_t> qual;    // empty when no quals
  bool prior = false;
  std::vector<uint16_t> kmer;   // 1024 counts
  std::vector<uint16_t> kord;   // ordered 5-mers
  unsigned reads = 0, index = 0;
  double p = 0.0, E_minmax = -999.0;                                  // containers.cpp:38-39
  Cmp comp{0, 0, 0.0, 0};
  bool lock = false, correct = true;
};

struct Bi {                                                           // dada.h:85-107
  int center = -1;
  std::vector<int> raw;         // member indices in slot order
  unsigned reads = 0;
  bool update_e = true, check_locks = true;                           // containers.cpp:61-63
  double self = 0;
  char birth_type = 'I';
  unsigned birth_from = 0;
  double birth_pval = 0, birth_fold = 1, birth_e = 0;
  Cmp birth_comp{0, 0, 0.0, 0};
  std::vector<Cmp> comp;
};

struct Params {
  int match, mismatch, gap, homo_gap, band;
  bool use_kmers, vectorized, gapless, greedy, use_quals;
  double kdist_cutoff;
  int SSE;
};

// ---- misc.cpp:38-68 nt2int -------------------------------------------------------
std::string nt2int(const char *s, size_t n) {
  std::string o(n, 0);
  for (size_t i = 0; i < n; i++) {
    switch (s[i]) {
      case 'A': o[i] = 1; break;
      case 'C': o[i] = 2; break;
      case 'G': o[i] = 3; break;
      case 'T': o[i] = 4; break;
      case 'N': o[i] = 5; break;
      default: throw Stop("Unexpected nucleotide.");
    }
  }
  return o;
}
char int2nt(char c) {  // misc.cpp:71-99
  switch (c) { case 1: return 'A'; case 2: return 'C'; case 3: return 'G'; case 4: return 'T';
               case 5: return 'N'; default: return c; }
}

// ---- kmers.cpp:207-279 assign_kmer / assign_kmer_order ------------------------------
void assign_kmers(Raw &r) {
  int len = (int)r.seq.size();
  r.kmer.assign(NKMER, 0);
  r.kord.assign(len - K + 1, 0);
  for (int i = 0; i + K <= len; i++) {
    unsigned km = 0;
    for (int j = i; j < i + K; j++) {
      int nti = r.seq[j] - 1;
      if (nti < 0 || nti > 3) throw Stop("Unexpected nucleotide.");
      km = 4 * km + nti;
    }
    r.kmer[km]++;
    r.kord[i] = (uint16_t)km;
  }
}

// ---- kmers.cpp:13-93: all three kmer_dist variants reduce to the exact min-sum (the
// 8-bit SSE path returns -1 on any saturated lane and is then replaced by the 16-bit
// one, raw_align nwalign_endsfree.cpp:22-26) ------------------------------------------
unsigned kmer_minsum(const Raw &a, const Raw &b) {
  unsigned s = 0;
  for (int k = 0; k < NKMER; k++) s += std::min(a.kmer[k], b.kmer[k]);
  return (uint16_t)s;  // dotsum is uint16_t in the reference
}
// kmers.cpp:121-150 kord_dist_SSEi (x86: no length check); :102-116 scalar returns -1 if
// lengths differ.
int kord_match(const Raw &a, const Raw &b) {
  int klen = (int)std::min(a.seq.size(), b.seq.size()) - K + 1, m = 0;
  for (int i = 0; i < klen; i++) m += a.kord[i] == b.kord[i];
  return m;
}

// ---- nwalign_endsfree.cpp:76-216 (homo_gap == gap) and :220-396 (homopolymer gaps):
// banded ends-free NW, int32, full-matrix semantics restated over the band only.
// nwalign_vectorized.cpp:71-318 computes the same alignment (int16 anti-diagonal layout,
// shorter sequence first with flipped precedence) -- verified pairwise against _ref. ---
void nw_endsfree(const std::string &s1, const std::string &s2, int match, int mismatch, int gap_p,
                 int homo_gap_p, bool homo, int band, std::string &al0, std::string &al1) {
  int len1 = (int)s1.size(), len2 = (int)s2.size();
  std::vector<unsigned char> homo1(len1, 0), homo2(len2, 0);
  if (homo) {  // :230-255 runs of length >= 3
    for (int i = 0, j = 0; j < len1; j++)
      if (j == len1 - 1 || s1[j] != s1[j + 1]) { for (int k = i; k <= j; k++) homo1[k] = (j - i >= 2); i = j + 1; }
    for (int i = 0, j = 0; j < len2; j++)
      if (j == len2 - 1 || s2[j] != s2[j + 1]) { for (int k = i; k <= j; k++) homo2[k] = (j - i >= 2); i = j + 1; }
  }
  size_t ncol = (size_t)len2 + 1;
  std::vector<int> d((size_t)(len1 + 1) * ncol), p((size_t)(len1 + 1) * ncol, 0);
  for (int i = 0; i <= len1; i++) { d[i * ncol] = 0; p[i * ncol] = 3; }   // :91-95
  for (int j = 0; j <= len2; j++) { d[j] = 0; p[j] = 2; }                  // :97-101
  int lband, rband;                                                         // :103-113
  if (len2 > len1) { lband = band; rband = band + len2 - len1; }
  else if (len1 > len2) { lband = band + len1 - len2; rband = band; }
  else { lband = band; rband = band; }
  if (band >= 0 && (band < len1 || band < len2)) {                          // :115-121
    for (int i = 0; i <= len1; i++) {
      if (i - lband - 1 >= 0) d[i * ncol + i - lband - 1] = -9999;
      if (i + rband + 1 <= len2) d[i * ncol + i + rband + 1] = -9999;
    }
  }
  for (int i = 1; i <= len1; i++) {                                         // :124-160
    int l, r;
    if (band >= 0) { l = std::max(1, i - lband); r = std::min(len2, i + rband); }
    else { l = 1; r = len2; }
    for (int j = l; j <= r; j++) {
      int left, up, diag;
      if (i == len1) left = d[i * ncol + j - 1];
      else if (homo && homo2[j - 1]) left = d[i * ncol + j - 1] + homo_gap_p;
      else left = d[i * ncol + j - 1] + gap_p;
      if (j == len2) up = d[(i - 1) * ncol + j];
      else if (homo && homo1[i - 1]) up = d[(i - 1) * ncol + j] + homo_gap_p;
      else up = d[(i - 1) * ncol + j] + gap_p;
      diag = d[(i - 1) * ncol + j - 1] + (s1[i - 1] == s2[j - 1] ? match : mismatch);
      if (up >= diag && up >= left) { d[i * ncol + j] = up; p[i * ncol + j] = 3; }
      else if (left >= diag) { d[i * ncol + j] = left; p[i * ncol + j] = 2; }
      else { d[i * ncol + j] = diag; p[i * ncol + j] = 1; }
    }
  }
  std::string r0, r1;                                                       // :166-190 traceback
  int i = len1, j = len2;
  while (i > 0 || j > 0) {
    switch (p[i * ncol + j]) {
      case 1: r0.push_back(s1[--i]); r1.push_back(s2[--j]); break;
      case 2: r0.push_back('-'); r1.push_back(s2[--j]); break;
      case 3: r0.push_back(s1[--i]); r1.push_back('-'); break;
      default: throw Stop("N-W Align out of range.");
    }
  }
  al0.assign(r0.rbegin(), r0.rend());
  al1.assign(r1.rbegin(), r1.rend());
}

// ---- nwalign_endsfree.cpp:539-555 nwalign_gapless ----------------------------------
void nw_gapless(const std::string &s1, const std::string &s2, std::string &al0, std::string &al1) {
  size_t n = std::max(s1.size(), s2.size());
  al0.resize(n); al1.resize(n);
  for (size_t i = 0; i < n; i++) { al0[i] = i < s1.size() ? s1[i] : '-'; al1[i] = i < s2.size() ? s2[i] : '-'; }
}

// ---- nwalign_endsfree.cpp:10-73 raw_align: returns false for a NULL alignment ------
// kind: 0 shrouded, 1 gapless, 2 NW
int raw_align(const Raw &r1, const Raw &r2, const Params &P, bool use_kmers, double kdist_cutoff,
              std::string &al0, std::string &al1) {
  double kdist = 0.0, kodist = -1.0;
  int minlen = (int)std::min(r1.seq.size(), r2.seq.size());
  if (use_kmers) kdist = 1. - ((double)kmer_minsum(r1, r2)) / (minlen - K + 1.);
  if (use_kmers && P.gapless) {
    if (P.SSE == 0 && r1.seq.size() != r2.seq.size()) kodist = -1.0;          // kmers.cpp:107
    else kodist = 1. - ((double)(uint16_t)kord_match(r1, r2)) / (minlen - K + 1.);
  }
  if (use_kmers && kdist > kdist_cutoff) return 0;
  if (P.band == 0 || (P.gapless && kodist == kdist)) { nw_gapless(r1.seq, r2.seq, al0, al1); return 1; }
  if (P.vectorized) nw_endsfree(r1.seq, r2.seq, P.match, P.mismatch, P.gap, P.gap, false,
                                P.band < 0 ? -1 : P.band, al0, al1);
  else if (P.homo_gap != P.gap && P.homo_gap <= 0)
    nw_endsfree(r1.seq, r2.seq, P.match, P.mismatch, P.gap, P.homo_gap, true, P.band, al0, al1);
  else nw_endsfree(r1.seq, r2.seq, P.match, P.mismatch, P.gap, P.gap, false, P.band, al0, al1);
  return 2;
}

// ---- nwalign_endsfree.cpp:570-639 al2subs + :642-672 sub_new ------------------------
Sub sub_new(const Raw &r0, const Raw &r1, const Params &P, bool use_kmers, double kdist_cutoff, int *kind = nullptr) {
  Sub sub;
  std::string al0, al1;
  int k = raw_align(r0, r1, P, use_kmers, kdist_cutoff, al0, al1);
  if (kind) *kind = k;
  if (k == 0) return sub;
  sub.null = false;
  int i0 = -1, i1 = -1;
  for (size_t i = 0; i < al0.size(); i++) {
    bool is0 = al0[i] >= 1 && al0[i] <= 5, is1 = al1[i] >= 1 && al1[i] <= 5;
    if (is0) i0++;
    if (is1) i1++;
    if (is0) sub.map.push_back(is1 ? (uint16_t)i1 : (uint16_t)GAP_GLYPH);
    if (is0 && is1 && al0[i] != al1[i] && al0[i] != 5 && al1[i] != 5) {
      sub.pos.push_back((uint16_t)i0); sub.nt0.push_back(al0[i]); sub.nt1.push_back(al1[i]);
    }
  }
  sub.len0 = (unsigned)sub.map.size();
  sub.nsubs = (unsigned)sub.pos.size();
  if (!r0.qual.empty() && !r1.qual.empty()) {
    sub.has_q = true;
    for (unsigned s = 0; s < sub.nsubs; s++) {
      sub.q0.push_back(r0.qual[sub.pos[s]]);
      sub.q1.push_back(r1.qual[sub.map[sub.pos[s]]]);
    }
  }
  return sub;
}

// ---- pval.cpp:144-197 compute_lambda_ts (== :92-141 compute_lambda) -----------------
// err is row-major 16 x ncol like the C array of cluster.cpp:162-170.
double compute_lambda(const Raw &raw, const Sub &sub, unsigned ncol, const double *err, bool use_quals) {
  if (sub.null) return 0.0;
  int len1 = (int)raw.seq.size();
  std::vector<unsigned> tvec(len1), qind(len1);
  for (int pos1 = 0; pos1 < len1; pos1++) {
    int nti1 = raw.seq[pos1] - 1;
    if (nti1 < 0 || nti1 > 3) throw Stop("Non-ACGT sequences in compute_lambda.");
    tvec[pos1] = nti1 * 4 + nti1;
    qind[pos1] = use_quals ? raw.qual[pos1] : 0;
    if (qind[pos1] > ncol - 1) throw Stop("Rounded quality exceeded range of err lookup table.");
  }
  for (unsigned s = 0; s < sub.nsubs; s++) {
    int pos1 = sub.map[sub.pos[s]];
    if (pos1 < 0 || pos1 >= len1) throw Stop("CL: Bad pos1.");
    tvec[pos1] = (sub.nt0[s] - 1) * 4 + (sub.nt1[s] - 1);
  }
  double lambda = 1.0;
  for (int pos1 = 0; pos1 < len1; pos1++) lambda = lambda * err[tvec[pos1] * ncol + qind[pos1]];
  if (lambda < 0 || lambda > 1) throw Stop("Bad lambda.");
  return lambda;
}

// ---- pval.cpp:44-64 calc_pA, :67-89 get_pA ------------------------------------------
double calc_pA(int reads, double E_reads, bool prior) {
  double pval = oracle_ppois((double)(reads - 1), E_reads, 0, 0);
  if (!prior) {
    double norm = (1.0 - exp(-E_reads));
    if (norm < TAIL_APPROX_CUTOFF) norm = E_reads - 0.5 * E_reads * E_reads;
    pval = pval / norm;
  }
  return pval;
}
double get_pA(const Raw &raw, const Bi &bi, bool detect_singletons) {
  if (raw.reads == 1 && !raw.prior && !detect_singletons) return 1.;
  if (raw.comp.hamming == 0) return 1.;
  if (raw.comp.lambda == 0) return 0.;
  return calc_pA(raw.reads, raw.comp.lambda * bi.reads, raw.prior || detect_singletons);
}

struct B {                                                           // dada.h:110-123
  std::vector<Raw> &raw;
  std::vector<Bi> bi;
  unsigned reads = 0, nalign = 0, nshroud = 0;
  double omegaA, omegaP;
  explicit B(std::vector<Raw> &r) : raw(r) {}
};

void bi_add_raw(Bi &bi, const Raw &r) { bi.raw.push_back(r.index); bi.reads += r.reads; bi.update_e = true; }  // containers.cpp:150-162
int bi_pop_raw(Bi &bi, std::vector<Raw> &raws, unsigned r) {          // containers.cpp:183-197 (swap-with-last)
  if (r >= bi.raw.size()) throw Stop("Container Error (Bi): Tried to pop out-of-range raw.");
  int pop = bi.raw[r];
  bi.raw[r] = bi.raw.back();
  bi.raw.pop_back();
  bi.reads -= raws[pop].reads;
  bi.update_e = true;
  return pop;
}
void bi_assign_center(Bi &bi, std::vector<Raw> &raws) {               // cluster.cpp:371-386
  unsigned max_reads = 0;
  bi.center = -1;
  for (int idx : bi.raw) {
    raws[idx].lock = false;
    if (raws[idx].reads > max_reads) { bi.center = idx; max_reads = raws[idx].reads; }
  }
  bi.check_locks = true;
}

// ---- cluster.cpp:152-204 b_compare_parallel (== :13-88 b_compare where that is defined)
static FILE *g_trace = nullptr;   // optional: PORT_TRACE=<file> dumps one record per real alignment (analysis tooling)
void b_compare(B &b, unsigned i, const Params &P, double kdist_cutoff, const double *err, unsigned ncol) {
  const Raw &center = b.raw[b.bi[i].center];
  for (unsigned index = 0; index < b.raw.size(); index++) {
    Raw &raw = b.raw[index];
    Cmp comp{i, index, 0.0, (uint32_t)-1};
    if (!(P.greedy && raw.reads > center.reads) && !(P.greedy && raw.lock)) {
      int kind = 0;
      Sub sub = sub_new(center, raw, P, P.use_kmers, kdist_cutoff, &kind);
      b.nalign++;
      if (sub.null) b.nshroud++;
      comp.lambda = compute_lambda(raw, sub, ncol, err, P.use_quals);
      if (!sub.null) comp.hamming = sub.nsubs;
      if (g_trace && kind == 2) {
        struct { uint32_t i, index, nsubs, stored; double lambda, emm, breads; } rec{i, index, sub.nsubs,
            (uint32_t)(comp.lambda * b.reads > raw.E_minmax), comp.lambda, raw.E_minmax, (double)b.reads};
        fwrite(&rec, sizeof rec, 1, g_trace);
      }
    }
    double lambda = comp.lambda;
    if (index == (unsigned)b.bi[i].center) b.bi[i].self = lambda;
    if (lambda * b.reads > raw.E_minmax) {                            // cluster.cpp:192
      if (lambda * center.reads > raw.E_minmax) raw.E_minmax = lambda * center.reads;
      b.bi[i].comp.push_back(comp);
      if (i == 0 || (int)index == b.bi[i].center) raw.comp = comp;
    }
  }
}

// ---- cluster.cpp:210-266 b_shuffle2 -------------------------------------------------
bool b_shuffle2(B &b) {
  size_t nraw = b.raw.size();
  std::vector<double> emax(nraw);
  std::vector<const Cmp *> compmax(nraw);
  for (size_t index = 0; index < nraw; index++) {
    compmax[index] = &b.bi[0].comp[index];
    emax[index] = compmax[index]->lambda * b.bi[0].reads;
  }
  for (size_t i = 1; i < b.bi.size(); i++)
    for (const Cmp &c : b.bi[i].comp) {
      double e = c.lambda * b.bi[i].reads;
      if (e > emax[c.index]) { compmax[c.index] = &c; emax[c.index] = e; }
    }
  bool shuffled = false;
  for (size_t i = 0; i < b.bi.size(); i++)
    for (int r = (int)b.bi[i].raw.size() - 1; r >= 0; r--) {
      Raw &raw = b.raw[b.bi[i].raw[r]];
      if (compmax[raw.index]->i != i) {
        if ((int)raw.index == b.bi[i].center) continue;
        bi_pop_raw(b.bi[i], b.raw, r);
        bi_add_raw(b.bi[compmax[raw.index]->i], raw);
        raw.comp = *compmax[raw.index];
        shuffled = true;
      }
    }
  return shuffled;
}

// ---- pval.cpp:14-40 b_p_update ------------------------------------------------------
void b_p_update(B &b, bool greedy, bool detect_singletons) {
  for (Bi &bi : b.bi) {
    if (bi.update_e) {
      for (int idx : bi.raw) b.raw[idx].p = get_pA(b.raw[idx], bi, detect_singletons);
      bi.update_e = false;
    }
    if (greedy && bi.check_locks) {
      for (int idx : bi.raw) {
        Raw &raw = b.raw[idx];
        double E_reads_center = b.raw[bi.center].reads * raw.comp.lambda;
        if (E_reads_center > raw.reads) raw.lock = true;
        if (idx == bi.center) raw.lock = true;
      }
      bi.check_locks = false;
    }
  }
}

// ---- cluster.cpp:274-350 b_bud ------------------------------------------------------
int b_bud(B &b, double min_fold, int min_hamming, int min_abund) {
  int mini = -1, minr = -1, mini_prior = -1, minr_prior = -1;
  const Raw *minraw = &b.raw[b.bi[0].center], *minraw_prior = minraw;
  for (int i = 0; i < (int)b.bi.size(); i++)
    for (int r = 1; r < (int)b.bi[i].raw.size(); r++) {
      const Raw &raw = b.raw[b.bi[i].raw[r]];
      if ((int)raw.reads < min_abund) continue;
      int hamming = (int)raw.comp.hamming;
      double lambda = raw.comp.lambda;
      if (hamming >= min_hamming) {
        if (min_fold <= 1 || ((double)raw.reads) >= min_fold * lambda * b.bi[i].reads) {
          if ((raw.p < minraw->p) || (raw.p == minraw->p && raw.reads > minraw->reads)) { mini = i; minr = r; minraw = &raw; }
          if (raw.prior && ((raw.p < minraw_prior->p) || (raw.p == minraw_prior->p && raw.reads > minraw_prior->reads))) {
            mini_prior = i; minr_prior = r; minraw_prior = &raw;
          }
        }
      }
    }
  double pA = minraw->p * b.raw.size(), pP = minraw_prior->p;
  auto birth = [&](int from, int r, const Raw *mr, char type, double pv) {
    double expected = mr->comp.lambda * b.bi[from].reads;
    Cmp bc = mr->comp;
    int idx = bi_pop_raw(b.bi[from], b.raw, r);
    b.bi.emplace_back();
    int i = (int)b.bi.size() - 1;
    Bi &nb = b.bi[i];
    nb.birth_type = type;
    nb.birth_from = (type == 'A') ? (unsigned)from : bc.i;   // 'P' leaves birth_from uninitialised in the
    nb.birth_pval = pv;                                      // reference (cluster.cpp:334-339); we emit comp.i
    nb.birth_fold = b.raw[idx].reads / expected;
    nb.birth_e = expected;
    nb.birth_comp = bc;
    bi_add_raw(nb, b.raw[idx]);
    bi_assign_center(nb, b.raw);
    return i;
  };
  if (pA < b.omegaA && mini >= 0) return birth(mini, minr, minraw, 'A', pA);
  if (pP < b.omegaP && mini_prior >= 0) return birth(mini_prior, minr_prior, minraw_prior, 'P', pP);
  return 0;
}

template <typename T> T *dup(const std::vector<T> &v) {
  T *p = (T *)malloc(std::max<size_t>(1, v.size()) * sizeof(T));
  if (!v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

// ---- Rmain.cpp:30-295 dada_uniques --------------------------------------------------
dada2b_out *run(const dada2b_in *in, const dada2b_opts *o) {
  if (const char *tp = getenv("PORT_TRACE")) g_trace = fopen(tp, "wb");
  struct TraceCloser { ~TraceCloser() { if (g_trace) { fclose(g_trace); g_trace = nullptr; } } } trace_closer;
  unsigned nraw = in->nraw;
  if (nraw == 0) throw Stop("Zero input sequences.");
  unsigned maxlen = 0, minlen = SEQLEN;
  for (unsigned i = 0; i < nraw; i++) {
    unsigned l = (unsigned)(in->seq_off[i + 1] - in->seq_off[i]);
    maxlen = std::max(maxlen, l); minlen = std::min(minlen, l);
  }
  if (maxlen >= (unsigned)SEQLEN) throw Stop("Input sequences exceed the maximum allowed string length.");
  if (minlen <= (unsigned)K) throw Stop("Input sequences must all be longer than the kmer-size (5).");
  bool has_quals = in->maxlen > 0;
  if (has_quals && (unsigned)in->maxlen != maxlen) throw Stop("Sequence must have associated qualities for each nucleotide position.");
  unsigned ncol = in->Q;
  std::vector<double> err((size_t)16 * ncol);                        // cluster.cpp:162-170 row-major copy
  for (unsigned r = 0; r < 16; r++) for (unsigned c = 0; c < ncol; c++) err[r * ncol + c] = in->err[r + 16 * c];

  Params P;
  P.match = o->match; P.mismatch = o->mismatch; P.gap = o->gap; P.homo_gap = o->homo_gap; P.band = o->band_size;
  P.use_kmers = o->use_kmers; P.vectorized = o->vectorized_alignment; P.gapless = o->gapless; P.greedy = o->greedy;
  P.use_quals = o->use_quals; P.kdist_cutoff = o->kdist_cutoff; P.SSE = o->SSE;

  std::vector<Raw> raws(nraw);                                        // Rmain.cpp:103-120, containers.cpp:19-43
  for (unsigned i = 0; i < nraw; i++) {
    Raw &r = raws[i];
    size_t l = in->seq_off[i + 1] - in->seq_off[i];
    r.seq = nt2int(in->seq_concat + in->seq_off[i], l);
    if (has_quals) {
      r.qual.resize(l);
      for (size_t pos = 0; pos < l; pos++) r.qual[pos] = (uint8_t)round(in->quals[pos + (size_t)in->maxlen * i]);
    }
    r.reads = in->abund[i]; r.prior = in->prior ? in->prior[i] != 0 : false; r.index = i;
    if (o->use_kmers) assign_kmers(r);
  }

  // ---- run_dada Rmain.cpp:297-336 ----
  B b(raws);
  b.omegaA = o->omegaA; b.omegaP = o->omegaP;
  for (auto &r : raws) b.reads += r.reads;
  b.bi.emplace_back();                                                // b_init containers.cpp:111-137
  b.bi[0].birth_e = b.reads;
  for (auto &r : raws) bi_add_raw(b.bi[0], r);
  bi_assign_center(b.bi[0], raws);
  b_compare(b, 0, P, 1.0, err.data(), ncol);
  b_p_update(b, P.greedy, o->detect_singletons);
  int max_clust = o->max_clust < 1 ? (int)nraw : o->max_clust;
  int newi, nrounds = 0, nshuf_total = 0;
  while ((int)b.bi.size() < max_clust && (newi = b_bud(b, o->min_fold, o->min_hamming, o->min_abund))) {
    b_compare(b, newi, P, P.kdist_cutoff, err.data(), ncol);
    int nshuffle = 0; bool shuffled;
    do { shuffled = b_shuffle2(b); nshuf_total++; } while (shuffled && ++nshuffle < MAX_SHUFFLE);
    b_p_update(b, P.greedy, o->detect_singletons);
    nrounds++;
  }
  unsigned nclust = (unsigned)b.bi.size();

  // ---- final subs Rmain.cpp:179-236 ----
  std::vector<Sub> subs(nraw), birth_subs(nclust);
  for (unsigned i = 0; i < nclust; i++) {
    for (int idx : b.bi[i].raw) subs[idx] = sub_new(raws[b.bi[i].center], raws[idx], P, false, 1.0);
    if (i > 0) birth_subs[i] = sub_new(raws[b.bi[b.bi[i].birth_comp.i].center], raws[b.bi[i].center], P, P.use_kmers, 1.0);
  }
  // ---- final p Rmain.cpp:239-252 ----
  std::vector<double> pval(nraw);
  for (unsigned i = 0; i < nclust; i++)
    for (int idx : b.bi[i].raw) {
      Raw &raw = raws[idx];
      if (b.bi[i].center == idx) raw.p = 1.0;
      else { raw.p = calc_pA(raw.reads, raw.comp.lambda * b.bi[i].reads, true); if (raw.p < o->omegaC) raw.correct = false; }
      pval[idx] = raw.p;
    }

  dada2b_out *out = (dada2b_out *)calloc(1, sizeof(dada2b_out));
  out->nclust = nclust; out->nraw = nraw; out->maxlen = maxlen; out->Q = ncol;
  out->n_align = b.nalign; out->n_shroud = b.nshroud; out->n_rounds = nrounds; out->n_shuffles = nshuf_total;

  // ---- error.cpp:9-127 b_make_clustering_df ----
  std::string cseq; std::vector<int64_t> coff(1, 0);
  std::vector<int32_t> ab(nclust, 0), n0(nclust, 0), n1(nclust, 0), nunq(nclust, 0), bfrom(nclust), bham(nclust);
  std::vector<double> cpval(nclust), bpval(nclust), bfold(nclust), bqave(nclust);
  for (unsigned i = 0; i < nclust; i++) {
    unsigned max_reads = 0; int max_raw = -1;
    for (int idx : b.bi[i].raw) if (raws[idx].reads > max_reads) { max_raw = idx; max_reads = raws[idx].reads; }
    if (max_raw >= 0) for (char c : raws[max_raw].seq) cseq.push_back(int2nt(c));
    coff.push_back((int64_t)cseq.size());
    for (int idx : b.bi[i].raw) {
      const Raw &raw = raws[idx];
      if (raw.correct) {
        ab[i] += raw.reads; nunq[i]++;
        if (!subs[idx].null) { if (subs[idx].nsubs == 0) n0[i] += raw.reads; if (subs[idx].nsubs == 1) n1[i] += raw.reads; }
      }
    }
    if (i == 0) { bpval[i] = na_real(); bfrom[i] = INT_MIN; bfold[i] = na_real(); bham[i] = INT_MIN; bqave[i] = na_real(); }
    else {
      bfrom[i] = b.bi[i].birth_from + 1; bpval[i] = b.bi[i].birth_pval; bfold[i] = b.bi[i].birth_fold;
      bham[i] = (int)b.bi[i].birth_comp.hamming;
      if (has_quals) {
        double q_ave = 0.0; const Sub &s = birth_subs[i];
        if (!s.null && s.has_q) { for (unsigned k = 0; k < s.nsubs; k++) q_ave += s.q1[k]; q_ave = q_ave / ((double)s.nsubs); }
        bqave[i] = q_ave;
      } else bqave[i] = na_real();
    }
  }
  {  // post-hoc pval error.cpp:99-119
    std::unordered_map<unsigned, unsigned> center_of;
    for (unsigned i = 0; i < nclust; i++) center_of[b.bi[i].center] = i;
    std::vector<double> tot_e(nclust, 0.0);
    for (unsigned i = 0; i < nclust; i++)
      for (const Cmp &c : b.bi[i].comp) {
        auto it = center_of.find(c.index);
        if (it != center_of.end() && it->second != i) tot_e[it->second] += c.lambda * b.bi[i].reads;
      }
    for (unsigned i = 0; i < nclust; i++) cpval[i] = calc_pA(raws[b.bi[i].center].reads, tot_e[i], true);
  }
  out->cl_seq_concat = (char *)malloc(cseq.size() + 1); memcpy(out->cl_seq_concat, cseq.data(), cseq.size()); out->cl_seq_concat[cseq.size()] = 0;
  out->cl_seq_off = dup(coff);
  out->cl_abundance = dup(ab); out->cl_n0 = dup(n0); out->cl_n1 = dup(n1); out->cl_nunq = dup(nunq);
  out->cl_pval = dup(cpval); out->cl_birth_from = dup(bfrom); out->cl_birth_pval = dup(bpval);
  out->cl_birth_fold = dup(bfold); out->cl_birth_ham = dup(bham); out->cl_birth_qave = dup(bqave);

  // ---- error.cpp:131-172 transition matrix ----
  int tcol = has_quals ? (int)ncol : 1;
  std::vector<int32_t> trans((size_t)16 * tcol, 0);
  for (unsigned i = 0; i < nclust; i++) {
    const Raw &center = raws[b.bi[i].center];
    for (int idx : b.bi[i].raw) {
      const Raw &raw = raws[idx];
      if (!raw.correct || subs[idx].null) continue;
      for (unsigned pos0 = 0; pos0 < center.seq.size(); pos0++) {
        unsigned pos1 = subs[idx].map[pos0];
        if (pos1 == GAP_GLYPH) continue;
        unsigned t = 4 * (center.seq[pos0] - 1) + (raw.seq[pos1] - 1);
        unsigned q = has_quals ? raw.qual[pos1] : 0;
        trans[t + 16 * (size_t)q] += raw.reads;
      }
    }
  }
  out->subqual = dup(trans); out->subqual_ncol = tcol;

  // ---- error.cpp:225-258 cluster quality matrix ----
  std::vector<double> cq((size_t)maxlen * nclust, 0.0);
  if (has_quals)
    for (unsigned i = 0; i < nclust; i++) {
      unsigned seqlen = (unsigned)raws[b.bi[i].center].seq.size();
      std::vector<unsigned> nreads(maxlen, 0);
      for (int idx : b.bi[i].raw) {
        const Raw &raw = raws[idx];
        if (!raw.correct || subs[idx].null) continue;
        for (unsigned pos0 = 0; pos0 < seqlen; pos0++) {
          unsigned pos1 = subs[idx].map[pos0];
          if (pos1 == GAP_GLYPH) continue;
          nreads[pos0] += raw.reads;
          cq[pos0 + (size_t)maxlen * i] += (raw.qual[pos1] * raw.reads);
        }
      }
      for (unsigned pos0 = 0; pos0 < seqlen; pos0++) cq[pos0 + (size_t)maxlen * i] = cq[pos0 + (size_t)maxlen * i] / nreads[pos0];
      for (unsigned pos0 = seqlen; pos0 < maxlen; pos0++) cq[pos0 + (size_t)maxlen * i] = na_real();
    }
  out->clusterquals = dup(cq);

  // ---- error.cpp:261-300 birth subs ----
  std::vector<int32_t> bs_pos, bs_clust; std::vector<char> bs_ref, bs_sub; std::vector<double> bs_qual;
  for (unsigned i = 0; i < nclust; i++) {
    const Sub &s = birth_subs[i];
    if (s.null) continue;
    for (unsigned k = 0; k < s.nsubs; k++) {
      bs_pos.push_back(s.pos[k] + 1); bs_ref.push_back(int2nt(s.nt0[k])); bs_sub.push_back(int2nt(s.nt1[k]));
      bs_qual.push_back(has_quals ? (double)s.q1[k] : na_real()); bs_clust.push_back(i + 1);
    }
  }
  out->n_birth_subs = (int32_t)bs_pos.size();
  out->bs_pos = dup(bs_pos); out->bs_clust = dup(bs_clust); out->bs_ref = dup(bs_ref); out->bs_sub = dup(bs_sub); out->bs_qual = dup(bs_qual);

  // ---- map Rmain.cpp:269-279 ----
  std::vector<int32_t> map(nraw);
  for (unsigned i = 0; i < nclust; i++) for (int idx : b.bi[i].raw) map[idx] = raws[idx].correct ? (int)i + 1 : INT_MIN;
  out->map = dup(map); out->pval = dup(pval);
  return out;
}


// =====================================================================================
// Bimera detection: restatement of /root/reference/src/chimera.cpp.
// The alignment there is nwalign_vectorized2(query, parent, match, mismatch, gap_p, end_gap 0, band = max_shift)
// (chimera.cpp:27, :122), which computes the same alignment as the banded ends-free NW restated above.
// =====================================================================================
struct BimPair { int left, right, left_oo, right_oo, ham; };

// get_lr, chimera.cpp:239-269.  The integer conversions of the original are kept: `pos < len` and
// `pos > +(len - max_shift)` compare an int with a size_t (unsigned 64-bit arithmetic; the second wraps when
// len < max_shift), `pos < max_shift` and `pos >= 0` are int comparisons.  The one-off credit looks at the column
// AFTER the first mismatch (:254-256, :266-268).  Reads at al[.][len] see the terminating NUL, like the original.
void get_lr(const std::string &a0, const std::string &a1, int &left, int &right, int &left_oo, int &right_oo,
            bool allow_one_off, int max_shift) {
  const size_t len = a0.size();
  const char *al0 = a0.c_str(), *al1 = a1.c_str();
  int pos = 0; left = 0;
  while (al0[pos] == '-' && (size_t)pos < len) pos++;                                  // :242-244
  while (al1[pos] == '-' && pos < max_shift) { pos++; left++; }                        // :245-247
  while ((size_t)pos < len && al0[pos] == al1[pos]) { pos++; left++; }                 // :248-250
  if (allow_one_off) {                                                                 // :251-258
    left_oo = left; pos++;
    if ((size_t)pos < len && al0[pos] != '-') left_oo++;
    while ((size_t)pos < len && al0[pos] == al1[pos]) { pos++; left_oo++; }
  }
  pos = (int)len - 1; right = 0;
  while (al0[pos] == '-' && pos >= 0) pos--;                                           // :261-263
  while (al1[pos] == '-' && (size_t)pos > +(len - (size_t)max_shift)) { pos--; right++; }   // :264-266
  while (pos >= 0 && al0[pos] == al1[pos]) { pos--; right++; }                         // :267-269
  if (allow_one_off) {
    right_oo = right; pos--;
    if (pos >= 0 && al0[pos] != '-') right_oo++;
    while (pos >= 0 && al0[pos] == al1[pos]) { pos--; right_oo++; }
  }
}

// get_ham_endsfree, chimera.cpp:210-236: mismatching columns (internal gaps included) between the end-gap runs.
int get_ham_endsfree(const std::string &s1, const std::string &s2) {
  int i = 0, j = (int)s2.size() - 1;
  bool g1 = s1[i] == '-', g2 = s2[i] == '-';
  while (g1 || g2) { i++; g1 = g1 && s1[i] == '-'; g2 = g2 && s2[i] == '-'; }
  g1 = s1[j] == '-'; g2 = s2[j] == '-';
  while (g1 || g2) { j--; g1 = g1 && s1[j] == '-'; g2 = g2 && s2[j] == '-'; }
  int ham = 0;
  for (int pos = i; pos <= j; pos++) if (s1[pos] != s2[pos]) ham++;
  return ham;
}

BimPair bimera_pair(const std::string &sq, const std::string &par, bool allow_one_off, int match, int mismatch, int gap_p,
                    int max_shift, std::string *o0 = nullptr, std::string *o1 = nullptr) {
  std::string a0, a1;
  nw_endsfree(sq, par, match, mismatch, gap_p, gap_p, false, max_shift, a0, a1);
  BimPair r{0, 0, 0, 0, 0};
  get_lr(a0, a1, r.left, r.right, r.left_oo, r.right_oo, allow_one_off, max_shift);
  r.ham = get_ham_endsfree(a0, a1);
  if (o0) *o0 = a0;
  if (o1) *o1 = a1;
  return r;
}

// =====================================================================================
// mergePairs' native steps: restatement of /root/reference/src/evaluate.cpp:18-174.
// =====================================================================================
// C_eval_pair, evaluate.cpp:73-120: matches / mismatches / indels between the end-gap runs.
void eval_pair(const std::string &s1, const std::string &s2, int &match, int &mismatch, int &indel) {
  bool s1gap = true, s2gap = true;
  int start = -1, end;
  do {                                                                     // :85-91
    start++;
    s1gap = s1gap && (s1.c_str()[start] == '-');
    s2gap = s2gap && (s2.c_str()[start] == '-');
  } while ((s1gap || s2gap) && (size_t)start < s1.size());
  s1gap = s2gap = true;
  end = (int)s1.size();
  do {                                                                     // :94-101
    end--;
    if (end < 0) break;                                                    // the original would read s1[-1] (empty input only)
    s1gap = s1gap && (s1[end] == '-');
    s2gap = s2gap && (s2[end] == '-');
  } while ((s1gap || s2gap) && end >= start);
  match = mismatch = indel = 0;
  for (int i = start; i <= end; i++) {                                     // :104-113
    if (s1[i] == '-' || s2[i] == '-') indel++;
    else if (s1[i] == s2[i]) match++;
    else mismatch++;
  }
}

// C_pair_consensus, evaluate.cpp:131-174
std::string pair_consensus(const std::string &s1, const std::string &s2, int prefer, bool trim_overhang) {
  std::string o(s1.size(), '-');
  for (size_t i = 0; i < s1.size(); i++) {
    if (s1[i] == s2[i]) o[i] = s1[i];
    else if (s2[i] == '-') o[i] = s1[i];
    else if (s1[i] == '-') o[i] = s2[i];
    else o[i] = prefer == 1 ? s1[i] : (prefer == 2 ? s2[i] : 'N');
  }
  if (trim_overhang) {                                                     // :152-160
    for (size_t i = 0; i < s1.size(); i++) { if (s1[i] != '-') break; o[i] = '-'; }
    for (int i = (int)s1.size() - 1; i >= 0; i--) { if (s2[i] != '-') break; o[i] = '-'; }
  }
  std::string r;
  for (char c : o) if (c != '-') r.push_back(c);                           // :162-168
  return r;
}

}  // namespace

extern "C" {

int port_run(const dada2b_in *in, const dada2b_opts *opts, dada2b_out **out, char errbuf[DADA2B_ERRLEN]) {
  try { *out = run(in, opts); return 0; }
  catch (std::exception &e) { snprintf(errbuf, DADA2B_ERRLEN, "%s", e.what()); *out = nullptr; return 1; }
}

void port_free(dada2b_out *o) {
  if (!o) return;
  free(o->cl_seq_concat); free(o->cl_seq_off); free(o->cl_abundance); free(o->cl_n0); free(o->cl_n1); free(o->cl_nunq);
  free(o->cl_pval); free(o->cl_birth_from); free(o->cl_birth_pval); free(o->cl_birth_fold); free(o->cl_birth_ham);
  free(o->cl_birth_qave); free(o->bs_pos); free(o->bs_ref); free(o->bs_sub); free(o->bs_qual); free(o->bs_clust);
  free(o->subqual); free(o->clusterquals); free(o->map); free(o->pval); free(o);
}

// Pair-level hook: centre seq0 vs raw seq1 (ACGT text), err row-major 16 x ncol.
// Returns kind (0 shrouded, 1 gapless, 2 NW) or -1 on error.
int port_pair(const char *seq0, const uint8_t *q0, const char *seq1, const uint8_t *q1, const double *err_rowmajor,
              int ncol, const dada2b_opts *o, int use_kmers, double kdist_cutoff, double *lambda, int *nsubs,
              uint16_t *map, uint16_t *pos, char *nt0, char *nt1, uint8_t *sq0, uint8_t *sq1, char *al0, char *al1,
              char *errbuf) {
  try {
    Params P;
    P.match = o->match; P.mismatch = o->mismatch; P.gap = o->gap; P.homo_gap = o->homo_gap; P.band = o->band_size;
    P.use_kmers = use_kmers; P.vectorized = o->vectorized_alignment; P.gapless = o->gapless; P.greedy = o->greedy;
    P.use_quals = q1 != nullptr; P.kdist_cutoff = kdist_cutoff; P.SSE = o->SSE;
    Raw r0, r1;
    r0.seq = nt2int(seq0, strlen(seq0)); r1.seq = nt2int(seq1, strlen(seq1));
    if (q0) r0.qual.assign(q0, q0 + r0.seq.size());
    if (q1) r1.qual.assign(q1, q1 + r1.seq.size());
    if (use_kmers) { assign_kmers(r0); assign_kmers(r1); }
    std::string a0, a1;
    int kind = raw_align(r0, r1, P, use_kmers, kdist_cutoff, a0, a1);
    if (al0) { for (size_t i = 0; i < a0.size(); i++) al0[i] = int2nt(a0[i]); al0[a0.size()] = 0; }
    if (al1) { for (size_t i = 0; i < a1.size(); i++) al1[i] = int2nt(a1[i]); al1[a1.size()] = 0; }
    Sub sub = sub_new(r0, r1, P, use_kmers, kdist_cutoff);
    *lambda = compute_lambda(r1, sub, ncol, err_rowmajor, P.use_quals);
    *nsubs = sub.null ? -1 : (int)sub.nsubs;
    if (!sub.null) {
      if (map) memcpy(map, sub.map.data(), sub.len0 * 2);
      for (unsigned s = 0; s < sub.nsubs; s++) {
        if (pos) pos[s] = sub.pos[s];
        if (nt0) nt0[s] = sub.nt0[s];
        if (nt1) nt1[s] = sub.nt1[s];
        if (sq0 && sub.has_q) sq0[s] = sub.q0[s];
        if (sq1 && sub.has_q) sq1[s] = sub.q1[s];
      }
    }
    return kind;
  } catch (std::exception &e) { if (errbuf) snprintf(errbuf, 256, "%s", e.what()); return -1; }
}

// C_table_bimera2 / BimeraTableParallel, chimera.cpp:61-207.  mat: nrow (samples) x ncol (sequences), column-major.
int port_table_bimera(int nrow, int ncol, const int *vals, const char **seqs, double min_fold, int min_abund,
                      int allow_one_off, int min_one_off_par_dist, int match, int mismatch, int gap_p, int max_shift,
                      int *nflag_out, int *nsam_out) {
  std::vector<std::string> sq(ncol);
  for (int j = 0; j < ncol; j++) sq[j] = seqs[j];
  std::vector<int> lefts(ncol), rights(ncol), lefts_oo(ncol), rights_oo(ncol);
  std::vector<char> allowed(ncol);
  for (int j = 0; j < ncol; j++) {                                                       // :105
    int nsam = 0, nflag = 0;
    const int sqlen = (int)sq[j].size();
    std::fill(lefts.begin(), lefts.end(), -1); std::fill(rights.begin(), rights.end(), -1);
    std::fill(lefts_oo.begin(), lefts_oo.end(), -1); std::fill(rights_oo.begin(), rights_oo.end(), -1);
    std::fill(allowed.begin(), allowed.end(), 0);
    for (int i = 0; i < nrow; i++) {                                                     // :116
      if (vals[i + (size_t)j * nrow] <= 0) continue;
      nsam++;
      int max_left = 0, max_right = 0, oo_max_left = 0, oo_max_right = 0, oo_max_left_oo = 0, oo_max_right_oo = 0;
      for (int k = 0; k < ncol; k++) {                                                   // :121
        if (vals[i + (size_t)k * nrow] > (min_fold * vals[i + (size_t)j * nrow]) && vals[i + (size_t)k * nrow] >= min_abund) {
          if (lefts[k] < 0) {                                                            // :123-147
            BimPair r = bimera_pair(sq[j], sq[k], allow_one_off != 0, match, mismatch, gap_p, max_shift);
            if (allow_one_off && r.ham >= min_one_off_par_dist) allowed[k] = 1;
            if (r.left + r.right < sqlen) { lefts[k] = r.left; rights[k] = r.right; if (allow_one_off) { lefts_oo[k] = r.left_oo; rights_oo[k] = r.right_oo; } }
            else { lefts[k] = 0; rights[k] = 0; if (allow_one_off) { lefts_oo[k] = 0; rights_oo[k] = 0; } }
          }
          if (lefts[k] > max_left) max_left = lefts[k];                                  // :149-157
          if (rights[k] > max_right) max_right = rights[k];
          if (allow_one_off && allowed[k]) {
            if (lefts[k] > oo_max_left) oo_max_left = lefts[k];
            if (rights[k] > oo_max_right) oo_max_right = rights[k];
            if (lefts_oo[k] > oo_max_left_oo) oo_max_left_oo = lefts_oo[k];
            if (rights_oo[k] > oo_max_right_oo) oo_max_right_oo = rights_oo[k];
          }
        }
      }
      if (max_right + max_left >= sqlen) nflag++;                                        // :162-169
      else if (allow_one_off) {
        if (oo_max_left + oo_max_right_oo >= sqlen || oo_max_left_oo + oo_max_right >= sqlen) nflag++;
      }
    }
    nflag_out[j] = nflag; nsam_out[j] = nsam;                                            // :172-173
  }
  return 0;
}

// C_is_bimera, chimera.cpp:18-59 (the early exit once a model is found does not change the result: the maxima only grow).
int port_is_bimera(const char *sq_c, int npar, const char **pars, int allow_one_off, int min_one_off_par_dist, int match,
                   int mismatch, int gap_p, int max_shift) {
  const std::string sq(sq_c);
  int max_left = 0, max_right = 0, oo_max_left = 0, oo_max_right = 0, oo_max_left_oo = 0, oo_max_right_oo = 0;
  bool rval = false;
  for (int i = 0; i < npar && !rval; i++) {
    BimPair r = bimera_pair(sq, pars[i], allow_one_off != 0, match, mismatch, gap_p, max_shift);
    if ((size_t)(r.left + r.right) >= sq.size()) continue;                               // :30-32
    if (r.left > max_left) max_left = r.left;
    if (r.right > max_right) max_right = r.right;
    if (allow_one_off && r.ham >= min_one_off_par_dist) {                                // :37-42
      if (r.left > oo_max_left) oo_max_left = r.left;
      if (r.right > oo_max_right) oo_max_right = r.right;
      if (r.left_oo > oo_max_left_oo) oo_max_left_oo = r.left_oo;
      if (r.right_oo > oo_max_right_oo) oo_max_right_oo = r.right_oo;
    }
    if ((size_t)(max_right + max_left) >= sq.size()) rval = true;                        // :45-52
    if (allow_one_off && ((size_t)(oo_max_left + oo_max_right_oo) >= sq.size() || (size_t)(oo_max_left_oo + oo_max_right) >= sq.size())) rval = true;
  }
  return rval ? 1 : 0;
}

int port_bimera_pair(const char *sq, const char *par, int allow_one_off, int match, int mismatch, int gap_p, int max_shift,
                     int *out5, char *al0, char *al1) {
  std::string a0, a1;
  BimPair r = bimera_pair(sq, par, allow_one_off != 0, match, mismatch, gap_p, max_shift, &a0, &a1);
  out5[0] = r.left; out5[1] = r.right; out5[2] = r.left_oo; out5[3] = r.right_oo; out5[4] = r.ham;
  if (al0) strcpy(al0, a0.c_str());
  if (al1) strcpy(al1, a1.c_str());
  return 0;
}

// C_nwalign (endsfree) + C_eval_pair + C_pair_consensus for one pair, as chained by R/paired.R:153-164.
int port_merge_pair(const char *s1, const char *s2, int match, int mismatch, int gap_p, int homo_gap_p, int band, int prefer,
                    int trim_overhang, int *counts3, char *cons, char *al0, char *al1) {
  try {
    std::string a0, a1;
    nw_endsfree(s1, s2, match, mismatch, gap_p, homo_gap_p, gap_p != homo_gap_p, band, a0, a1);      // evaluate.cpp:40-46
    eval_pair(a0, a1, counts3[0], counts3[1], counts3[2]);
    strcpy(cons, pair_consensus(a0, a1, prefer, trim_overhang != 0).c_str());
    if (al0) strcpy(al0, a0.c_str());
    if (al1) strcpy(al1, a1.c_str());
    return 0;
  } catch (std::exception &e) { return -1; }
}

double port_calc_pA(int reads, double E_reads, int prior) { return calc_pA(reads, E_reads, prior != 0); }
double port_ppois_upper(int reads_minus_1, double E) { return oracle_ppois((double)reads_minus_1, E, 0, 0); }

}  // extern "C"
