// Shared declarations of the bimera kernels (dd_bimera.cu: traceback kernel, table kernels, host driver;
// dd_bimfwd.cu: register-resident alignment kernel).  Product code.
#pragma once
#include "dd_common.h"
#include "dd_kernels.h"

namespace dd2 {

// Packed per-(query, parent) record: bits 0-14 left, 15-29 right, 30-44 left_oo, 45-59 right_oo, 60 allowed
// (get_ham_endsfree >= min_one_off_par_dist), 63 valid.  Alignments have < 2 * 9999 columns, so 15 bits suffice.
constexpr unsigned long long BIM_VALID = 1ull << 63, BIM_ALLOWED = 1ull << 60;
__host__ __device__ inline unsigned long long bim_pack(int l, int r, int lo, int ro, bool allowed) {
  return BIM_VALID | (allowed ? BIM_ALLOWED : 0ull) | (unsigned long long)(l & 0x7FFF) | ((unsigned long long)(r & 0x7FFF) << 15) |
         ((unsigned long long)(lo & 0x7FFF) << 30) | ((unsigned long long)(ro & 0x7FFF) << 45);
}

struct BimSeqs {
  int n, maxlen, minlen, SW;
  const uint32_t *seq2;     // [n][SW] 2-bit packed, like DevIn::seq2
  const uint16_t *len;      // [n]
};

struct BimAlignArgs {
  BimSeqs sq;
  AlnParams P;
  int allow_one_off, min_one_off_par_dist, max_shift;
  const uint32_t *jq, *jk;                  // job -> (query sequence, parent sequence)
  const unsigned long long *njobs_ptr;      // device-side count (NULL => njobs_fixed)
  unsigned long long njobs_fixed;
  int dst_mode;                             // 0: record index = job;  1: slot(jq) * ncol + jk  (table batches)
  uint32_t j0; int ncol;
  int q_mul, q_add;                         //    batch slot of query q = (q - q_add) / q_mul - j0  (sharded table calls)
  unsigned long long *rec;                  // packed records (may be NULL)
  int32_t *raw5;                            // [job][5] unpacked get_lr / ham values (test hook; may be NULL)
  unsigned long long *ctr;                  // [0] jobs (k_bim_need), [1] cells, [2] error flag
  int warp_words, seq_bytes, H_words, ops_words, mask_words, ptr_in_smem;
  uint32_t *ptr_scratch; unsigned long long ptr_words;
  const uint32_t *job_list;                 // optional indirection: job id = job_list[x], x < njobs (fallback pass)
  uint32_t *fb_list; unsigned long long *fb_count;   // k_bimfwd: jobs it could not hold in registers
  int fast_ok;                              // k_bimfwd: interior fast path allowed (scores cannot approach the sentinel)
};

// run of set bits of mask M starting at pos and going up, never past position n - 1
__device__ __forceinline__ int run_fwd(const uint32_t *M, int pos, int n) {
  int cnt = 0;
  while (pos < n) {
    const int sh = pos & 31, avail = 32 - sh;
    const uint32_t w = ~(M[pos >> 5] >> sh);
    const int t = w ? __ffs((int)w) - 1 : 32;
    const int r = min(min(t, avail), n - pos);
    cnt += r; pos += r;
    if (r < avail) break;
  }
  return cnt;
}
// run of set bits starting at pos and going down to 0
__device__ __forceinline__ int run_bwd(const uint32_t *M, int pos) {
  int cnt = 0;
  while (pos >= 0) {
    const int avail = (pos & 31) + 1;
    const uint32_t w = ~(M[pos >> 5] << (31 - (pos & 31)));
    const int t = __clz((int)w);
    const int r = min(t, avail);
    cnt += r; pos -= r;
    if (r < avail) break;
  }
  return cnt;
}
__device__ __forceinline__ int bit_at(const uint32_t *M, int pos) { return (M[pos >> 5] >> (pos & 31)) & 1; }

// get_lr (chimera.cpp:239-269) and get_ham_endsfree (:210-236) on the column masks of an alignment of n columns:
// Q = query row has '-', Pm = parent row has '-', E = both rows hold the same base.  Warp-uniform (every lane computes
// the same scalars from shared memory).  Integer conversions of the original are kept: `pos > +(len - max_shift)` is an
// unsigned 64-bit comparison (false for every pos when len < max_shift); the one-off credit inspects the column after
// the first mismatch.
__device__ inline void bim_scan(const uint32_t *Q, const uint32_t *Pm, const uint32_t *E, int n, int neq, bool one_off, int max_shift, int out[5]) {
  int pos = run_fwd(Q, 0, n);                                                   // :242-244
  int left = run_fwd(Pm, pos, min(n, max(max_shift, 0)));                       // :245-247 (pos < max_shift)
  pos += left;
  { const int r = run_fwd(E, pos, n); left += r; pos += r; }                    // :248-250
  int left_oo = 0, right_oo = 0;
  if (one_off) {                                                                // :251-258
    left_oo = left; pos++;
    if (pos < n && !bit_at(Q, pos)) left_oo++;
    if (pos < n) left_oo += run_fwd(E, pos, n);
  }
  pos = n - 1;
  pos -= run_bwd(Q, pos);                                                       // :261-263
  int right = 0;
  {                                                                             // :264-266
    const unsigned long long thr = (unsigned long long)n - (unsigned long long)(long long)max_shift;
    if (pos >= 0 && (unsigned long long)pos > thr) {
      const int r = min(run_bwd(Pm, pos), (int)((unsigned long long)pos - thr));
      right += r; pos -= r;
    }
  }
  { const int r = run_bwd(E, pos); right += r; pos -= r; }                      // :267-269
  if (one_off) {
    right_oo = right; pos--;
    if (pos >= 0 && !bit_at(Q, pos)) right_oo++;
    if (pos >= 0) right_oo += run_bwd(E, pos);
  }
  // get_ham_endsfree: the end-gap run on either side belongs to whichever row starts (ends) with a gap
  const int i = bit_at(Q, 0) ? run_fwd(Q, 0, n) : run_fwd(Pm, 0, n);
  const int j = n - 1 - (bit_at(Q, n - 1) ? run_bwd(Q, n - 1) : run_bwd(Pm, n - 1));
  out[0] = left; out[1] = right; out[2] = left_oo; out[3] = right_oo; out[4] = (j - i + 1) - neq;
}


// dd_bimfwd.cu: register-resident wavefront NW that stores 2-bit moves per step and traces back per pair (EXPERIMENTAL,
// DADA2B_BIMFWD=1).  Pairs whose band does not fit an instantiation are appended to fb_list for k_bim_align.
bool launch_bimfwd(const BimAlignArgs &a, int slots_needed, unsigned long long njobs_upper, int num_sms, cudaStream_t s, int *grid_out);
size_t bimfwd_scratch_words(int slots_needed, int maxlen, int num_sms);
// dd_bimfwd16.cu: two jobs per lane group on the 16-bit SIMD datapath (EXPERIMENTAL, DADA2B_BIMFWD=2); jobs whose neighbour has
// another query / parent length come back in uneq_list for k_bimfwd.
bool launch_bimfwd16(const BimAlignArgs &a, uint32_t *uneq_list, unsigned long long *uneq_count, int slots_needed,
                     unsigned long long njobs_upper, int num_sms, cudaStream_t s);

}  // namespace dd2
