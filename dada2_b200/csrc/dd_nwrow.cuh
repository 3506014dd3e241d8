// Shared pieces of the thread-per-pair / lane-group row-sweep NW kernels (dd_nwrow.cu, dd_nwlane.cu).  Product code.
// See dd_nwrow.cu for the word layout (score << 16 | prec << 14 | nsubs) and the reference lines it restates.
#pragma once
#include "dd_common.h"
#include "dd_kernels.h"

namespace dd2 {
namespace {

constexpr int NW_CLR = (int)0xFFFF3FFFu;          // clears the prec field
constexpr int NW_NMASK = 0x3FFF;                  // nsubs field
constexpr int NW_SENT_H = -20000;                 // out-of-band score (biased space), far below any real score
constexpr int NW_PREC_LEFT = 1 << 14, NW_PREC_UP = 2 << 14;

struct RowConsts {
  int cU, cU0, cL, cL0, delta, matchS;
};

__host__ __device__ inline RowConsts row_consts(const AlnParams &P) {
  RowConsts c;
  c.matchS = P.match * 65536;
  c.cU = (P.gap - P.match) * 65536 + NW_PREC_UP;       // up: previous row -> one more unit of bias
  c.cU0 = (0 - P.match) * 65536 + NW_PREC_UP;          // free end gap in column len2
  c.cL = P.gap * 65536 + NW_PREC_LEFT;                 // left: same row
  c.cL0 = NW_PREC_LEFT;                                // free end gap in row len1
  c.delta = (P.mismatch - P.match) * 65536 + 1;        // mismatching diagonal move: score difference, one substitution
  return c;
}

// mismatch bit of cell d: bit 2*(d % 16) of word d / 16 of the row's mask (odd bits: don't care), as 0 / 1 through the FMA pipe
template <int NWW> __device__ __forceinline__ int mis_bit(const uint32_t (&mm)[NWW], int d) {
  return (int)__umulhi(mm[d >> 4] << (31 - 2 * (d & 15)), 2u);
}

// One DP row, in place.  CHECKED = false: interior row (no boundary cell).  MOVES: also return the row's 2-bit moves.
template <int B, bool CHECKED, bool MOVES>
__device__ __forceinline__ int nw_row(int (&S)[2 * B + 1], const uint32_t (&mm)[(2 * (2 * B + 1) + 31) / 32], const RowConsts &c, int cLrow,
                                       int dpin, int pinval, int dfree, uint32_t (&mv)[(2 * (2 * B + 1) + 31) / 32]) {
  constexpr int W = 2 * B + 1;
  constexpr int SENT = NW_SENT_H * 65536;
  int left = SENT, mB = 0;
#pragma unroll
  for (int w = 0; w < (2 * W + 31) / 32; w++) if (MOVES) mv[w] = 0u;
#pragma unroll
  for (int d = 0; d < W; d++) {
    const int diag = S[d] + mis_bit(mm, d) * c.delta;
    const int up = (d + 1 < W) ? S[d + 1] : SENT;
    int cu = c.cU;
    if (CHECKED) cu = (d == dfree) ? c.cU0 : c.cU;
    const int t = __viaddmax_s32(up, cu, diag);
    int m = __viaddmax_s32(left, cLrow, t);
    if (MOVES) mv[d >> 4] |= (((uint32_t)m >> 14) & 3u) << (2 * (d & 15));
    if (d == B) mB = m;                       // the main-diagonal cell with its prec field still in place
    m &= NW_CLR;
    if (CHECKED) m = (d == dpin) ? pinval : m;
    S[d] = m; left = m;
  }
  return mB;
}


// warp-aggregated append of r (for lanes with flag set) to list / count
__device__ __forceinline__ void warp_append(bool flag, uint32_t r, uint32_t *list, unsigned long long *count) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = r;
}


// ---- the tail of an exact alignment, shared by k_nwrow<EXACT> and k_nwlane: traceback, lambda, store rule ----------------
// Moves were recorded as one word per (row, word slot): word `w` of row i (1-based) lives at mv[(i - 1) * row_stride +
// w * word_stride]; a word holds CPW cells, cell d of the band sits in word d / CPW at bits 2 * (d % CPW) (TOP: counted from the top).
// The ROW sequence of a traceback is known in advance (every diag / up move goes to row i - 1, left moves stay), so the
// words are streamed through registers RB rows at a time: one memory latency per RB rows instead of one per move
// (measured: ~300 ns per dependent L2 access made the naive walk 3x longer than the DP of a small round).
// Substituted raw positions are written to `sub` ([k * sub_stride]: position | centre base << 14), last one first.
// Returns nsubs of the traced path (al2subs, nwalign_endsfree.cpp:570-639; walk order of :169-188).
template <int NWORDS, int CPW, int RB, bool TOP = false>
__device__ __forceinline__ int trace_moves(const uint32_t *mv, size_t row_stride, size_t word_stride, int L, int B, const uint8_t *s_cen,
                                           const uint32_t *raw2 /* packed raw row (shared or global) */, uint16_t *sub, size_t sub_stride) {
  int i = L, j = L, nsub = 0;
  bool done = false;
  while (!done && i > 0) {
    uint32_t buf[RB][NWORDS];
#pragma unroll
    for (int rr = 0; rr < RB; rr++) {
      const int row = i - rr;
#pragma unroll
      for (int w = 0; w < NWORDS; w++) buf[rr][w] = row >= 1 ? mv[(size_t)(row - 1) * row_stride + (size_t)w * word_stride] : 0u;
    }
#pragma unroll
    for (int rr = 0; rr < RB; rr++) {
      if (done || i < 1) { done = true; continue; }
      for (;;) {                                        // moves inside this row (left moves), then one move that leaves it
        if (j == 0) { done = true; break; }             // left column: only up moves remain, no raw base involved (p = 3, :89-92)
        const int d = j - i + B, wi = d / CPW;
        uint32_t word = buf[rr][0];
#pragma unroll
        for (int w = 1; w < NWORDS; w++) word = (wi == w) ? buf[rr][w] : word;
        // TOP: the lane kernel shifts moves in from the top of the word (cell k of a lane ends at bits 32 - 2 (CPW - k))
        const uint32_t mvv = (word >> (TOP ? 32 - 2 * (CPW - (d - wi * CPW)) : 2 * (d - wi * CPW))) & 3u;
        if (mvv == 1u) { j--; continue; }               // left: raw base against a gap (self transition)
        if (mvv == 0u) {                                // diag: centre base i-1 against raw base j-1
          const uint32_t b1 = s_cen[i - 1], b2 = (raw2[(j - 1) >> 4] >> (2 * ((j - 1) & 15))) & 3u;
          if (b1 != b2) { sub[(size_t)nsub * sub_stride] = (uint16_t)((uint32_t)(j - 1) | (b1 << 14)); nsub++; }
          j--;
        }
        i--;                                            // diag or up: on to the row above
        break;
      }
    }
  }
  return nsub;
}

// lambda = product over the raw's positions, in order, of err[transition][quality] (compute_lambda_ts, pval.cpp:158-193): the
// self transition everywhere except at the substituted positions found by the traceback.  Qualities are fetched 16 bytes at a
// time, one block ahead of their use.  *errflag gets ERR_QUAL when a rounded quality exceeds the table (pval.cpp:169-171).
__device__ __forceinline__ double lambda_from_subs(const uint32_t *raw2, const uint8_t *qrow, int L, int ncol, int use_quals, const double *s_err,
                                                   const uint16_t *sub, size_t sub_stride, int nsub, int *errflag) {
  double lam = 1.0;
  int k = nsub - 1;
  uint32_t nxt = k >= 0 ? sub[(size_t)k * sub_stride] : 0xFFFFu;
  uint4 qn = *(const uint4 *)qrow;                      // QS is a multiple of 16: whole blocks are always readable
  for (int p0 = 0; p0 < L; p0 += 16) {
    const uint4 qv = qn;
    if (p0 + 16 < L) qn = *(const uint4 *)(qrow + p0 + 16);
    uint32_t bw = raw2[p0 >> 4];
    const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int p = p0 + u;
      if (p < L) {
        const uint32_t b = bw & 3u;
        int q = use_quals ? (int)((qq[u >> 2] >> (8 * (u & 3))) & 0xFFu) : 0;
        if (q > ncol - 1) { *errflag = ERR_QUAL; q = ncol - 1; }
        uint32_t tt = 5u * b;
        if ((nxt & 0x3FFFu) == (uint32_t)p) {
          tt = 4u * (nxt >> 14) + b;
          k--;
          nxt = k >= 0 ? sub[(size_t)k * sub_stride] : 0xFFFFu;
        }
        lam = lam * s_err[tt * ncol + q];
      }
      bw >>= 2;
    }
  }
  return lam;
}

// lambda of the gapless alignment of two equally long sequences (the traced path is the main diagonal): position p of the raw against
// position p of the centre, in raw-position order -- the same product compute_lambda_ts forms from that alignment's subs
__device__ __forceinline__ double lambda_diag(const uint32_t *raw2, const uint8_t *s_cen, const uint8_t *qrow, int L, int ncol, int use_quals,
                                              const double *s_err, int *errflag) {
  double lam = 1.0;
  uint4 qn = *(const uint4 *)qrow;
  for (int p0 = 0; p0 < L; p0 += 16) {
    const uint4 qv = qn;
    if (p0 + 16 < L) qn = *(const uint4 *)(qrow + p0 + 16);
    uint32_t bw = raw2[p0 >> 4];
    const uint32_t qq[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
    for (int u = 0; u < 16; u++) {
      const int p = p0 + u;
      if (p < L) {
        const uint32_t b = bw & 3u, cb = s_cen[p];
        int q = use_quals ? (int)((qq[u >> 2] >> (8 * (u & 3))) & 0xFFu) : 0;
        if (q > ncol - 1) { *errflag = ERR_QUAL; q = ncol - 1; }
        lam = lam * s_err[(4u * cb + b) * ncol + q];
      }
      bw >>= 2;
    }
  }
  return lam;
}

// the "selectively store" step of b_compare (cluster.cpp:179-201) for one exact comparison
__device__ __forceinline__ void store_comparison(const FwdArgs &a, uint32_t r, double lam, int ns) {
  const double emm = a.st.E_minmax[r];
  if (lam * (double)a.total_reads > emm) {
    const double ec = lam * (double)a.centre_reads;
    if (ec > emm) a.st.E_minmax[r] = ec;
    {
      const unsigned long long slot = a.cluster_i == 0 ? (unsigned long long)r : atomicAdd(&a.st.ctr[CTR_CS_COUNT], 1ull);
      if (slot < a.st.cs_cap) {
        a.st.cs_index[slot] = r; a.st.cs_i[slot] = a.cluster_i; a.st.cs_lambda[slot] = lam; a.st.cs_ham[slot] = (uint32_t)ns;
      }
      if (a.cluster_i == 0 || r == a.centre_idx) { a.st.comp_lambda[r] = lam; a.st.comp_ham[r] = (uint32_t)ns; }
    }
  }
}

}  // namespace
bool nwrow_applicable(const FwdArgs &f, int len1);
}  // namespace dd2
