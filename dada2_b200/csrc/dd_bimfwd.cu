// k_bimfwd<G, ND>: register-resident banded ends-free NW for bimera detection, G lanes per (query, parent) pair, ND
// diagonals per lane (same wavefront as k_nwfwd, dd_nwfwd.cu).  Product code (sm_100a).  EXPERIMENTAL: DADA2B_BIMFWD=1.
//
// chimera.cpp needs the alignment itself only through get_lr / get_ham_endsfree (:210-269), i.e. through the shape of the
// path near its two ends.  The DP keeps scores in registers (no shared-memory score array, several pairs per warp) and
// writes the 2-bit move of every cell to a per-pair scratch row (one 32-bit word per lane and step; L2-resident for the
// batch in flight); lane 0 of each group then walks the path back from (len1, len2) -- the right-hand scan of get_lr runs
// in exactly that direction -- records the three column masks (query gap / parent gap / equal bases) and evaluates
// bim_scan on them.  Alignments are the reference's (precedence up > left > diag, nwalign_endsfree.cpp:147-156; free end
// gaps on the last row / column), so every number is identical to the warp-per-pair traceback kernel k_bim_align.
#include "dd_bimera.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace dd2 {


template <int G, int ND>
__global__ void __launch_bounds__(128) k_bimfwd(BimAlignArgs a) {
  constexpr int NSL = ND / 2;             // cells per lane per step
  constexpr int PPW = 32 / G;             // pairs per warp
  static_assert(ND % 2 == 0 && NSL <= 16, "base windows / move words are one 32-bit register each");
  extern __shared__ uint32_t smem[];
  const AlnParams &P = a.P;
  const int nwarps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane / G, gl = lane % G;
  // per pair: query bases [seq_bytes] | parent bases [seq_bytes] | 6 x mask_words (reversed + forward column masks)
  const int pair_words = 2 * (a.seq_bytes >> 2) + 6 * a.mask_words;
  uint32_t *s_pair = smem + (size_t)(wid * PPW + gid) * pair_words;
  uint8_t *s_q = (uint8_t *)s_pair;
  uint8_t *s_p = s_q + a.seq_bytes;
  uint32_t *rQ = s_pair + 2 * (a.seq_bytes >> 2), *rP = rQ + a.mask_words, *rE = rP + a.mask_words;
  uint32_t *mQ = rE + a.mask_words, *mP = mQ + a.mask_words, *mE = mP + a.mask_words;

  const unsigned long long njobs = a.njobs_ptr ? *a.njobs_ptr : a.njobs_fixed;
  if ((unsigned long long)blockIdx.x * nwarps * PPW >= njobs) return;
  const int SENT = P.sentinel, match = P.match, mismatch = P.mismatch, gap = P.gap;
  // scratch row of this pair slot: [step][lane of the group]
  uint32_t *moves = a.ptr_scratch + ((size_t)(blockIdx.x * nwarps + wid) * PPW + gid) * a.ptr_words;
  int errflag = 0;
  long long cells_lane = 0;

  for (unsigned long long base = (unsigned long long)(blockIdx.x * nwarps + wid) * PPW; base < njobs;
       base += (unsigned long long)gridDim.x * nwarps * PPW) {
    const unsigned long long jx = base + gid;
    bool act = jx < njobs;
    const unsigned long long jb = act ? (a.job_list ? (unsigned long long)a.job_list[jx] : jx) : 0;   // job_list: hand-over from k_bimfwd16
    const uint32_t q = act ? a.jq[jb] : 0, par = act ? a.jk[jb] : 0;
    const int len1 = act ? (int)a.sq.len[q] : 16;
    const int len2 = act ? (int)a.sq.len[par] : len1;   // idle groups run a benign geometry (their lanes still execute)
    if (act) {
      const uint32_t *qrow = a.sq.seq2 + (size_t)q * a.sq.SW, *prow = a.sq.seq2 + (size_t)par * a.sq.SW;
      for (int p = gl; p < len1; p += G) s_q[p] = (uint8_t)((qrow[p >> 4] >> (2 * (p & 15))) & 3u);
      for (int p = gl; p < len2; p += G) s_p[p] = (uint8_t)((prow[p >> 4] >> (2 * (p & 15))) & 3u);
    }
    __syncwarp();
    // ---- band geometry (nwalign_endsfree.cpp:101-111) ----
    int lband, rband;
    if (len2 > len1) { lband = P.band; rband = P.band + len2 - len1; }
    else if (len1 > len2) { lband = P.band + len1 - len2; rband = P.band; }
    else { lband = P.band; rband = P.band; }
    const int lb = min(lband, len1), rb = min(rband, len2);
    const int LB = (lb + 1) & ~1;
    const int lo = LB - lb, hi = LB + rb;          // in-band dd range [lo, hi]
    if (act && (P.band < 0 || hi >= G * ND)) {     // does not fit this instantiation: hand over to k_bim_align
      if (gl == 0) { unsigned long long s = atomicAdd(a.fb_count, 1ull); a.fb_list[s] = (uint32_t)jb; }
      act = false;
    }
    const int tlo = lo - gl * ND, thi = hi - gl * ND;   // in-band local t range for this lane
    const int D = gl * ND - LB;                          // delta of local t = 0 (even)
    const int nsteps = act ? len1 + len2 : 0;
    int maxsteps = nsteps;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxsteps = max(maxsteps, __shfl_xor_sync(0xffffffffu, maxsteps, o));

    int H[ND];
#pragma unroll
    for (int t = 0; t < ND; t++) H[t] = SENT;
    int I = -(D / 2), J = D / 2;      // windows for step k = 0; slot c: s1[I-1-c], s2[J-1+c]
    uint32_t A = 0, B = 0;
#pragma unroll
    for (int cc = 0; cc < NSL; cc++) {
      const int i1 = I - 1 - cc, j1 = J - 1 + cc;
      const uint32_t b1 = (act && i1 >= 0 && i1 < len1) ? s_q[i1] : 0u;
      const uint32_t b2 = (act && j1 >= 0 && j1 < len2) ? s_p[j1] : 0u;
      A |= b1 << (2 * cc); B |= b2 << (2 * cc);
    }
    // interior steps (no boundary cell, no free end gap, every pair of the warp still running): branch-free path with
    // out-of-band slots held far below any real score by an additive penalty
    int kf_lo = act ? max(lb, rb) + 2 : 0, kf_hi = act ? min(2 * len1 - lb, 2 * len2 - rb) - 1 : 0x3fffffff;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      kf_lo = max(kf_lo, __shfl_xor_sync(0xffffffffu, kf_lo, o));
      kf_hi = min(kf_hi, __shfl_xor_sync(0xffffffffu, kf_hi, o));
    }
    if (!a.fast_ok) kf_hi = -1;
    constexpr int BIGPEN = 1 << 20;
    int PEN[ND];
#pragma unroll
    for (int t = 0; t < ND; t++) PEN[t] = (t >= tlo && t <= thi) ? 0 : -BIGPEN;

    for (int kk = 0; kk <= maxsteps; kk += 2) {
      if (kk >= kf_lo && kk + 1 <= kf_hi) {
        // ---- interior: branch-free, ~11 instructions per cell (3 adds, VIMNMX3, 2 compares + 2 selects for the move, shift/or, penalty)
#pragma unroll
        for (int PAR = 0; PAR < 2; PAR++) {
          int Hn;
          if (PAR == 0) { Hn = __shfl_up_sync(0xffffffffu, H[ND - 1], 1, G); if (gl == 0) Hn = -BIGPEN; }
          else { Hn = __shfl_down_sync(0xffffffffu, H[0], 1, G); if (gl == G - 1) Hn = -BIGPEN; }
          const uint32_t X = A ^ B;
          const uint32_t EQ = ~(X | (X >> 1));                                  // bit 2cc set <=> bases of slot cc are equal
          int Hnew[NSL];
          uint32_t mv = 0;
#pragma unroll
          for (int cc = 0; cc < NSL; cc++) {
            const int t = 2 * cc + PAR;
            const int hl = (PAR == 0 && cc == 0) ? Hn : H[t - 1 < 0 ? 0 : t - 1];
            const int hu = (PAR == 1 && cc == NSL - 1) ? Hn : H[t + 1 >= ND ? ND - 1 : t + 1];
            const int left = hl + gap, up = hu + gap, diag = H[t] + (((EQ >> (2 * cc)) & 1u) ? match : mismatch);
            const int m = __vimax3_s32(left, up, diag);
            const uint32_t pm = (up == m) ? 3u : ((left == m) ? 2u : 1u);       // precedence up > left > diag (nwalign_endsfree.cpp:147-156)
            mv |= pm << (2 * cc);
            Hnew[cc] = m + PEN[t];
          }
#pragma unroll
          for (int cc = 0; cc < NSL; cc++) H[2 * cc + PAR] = Hnew[cc];
          moves[(size_t)(kk + PAR) * G + gl] = mv;                             // every pair of the warp is inside its interior here
          if (PAR == 0) {
            uint32_t nbB = __shfl_down_sync(0xffffffffu, B, 1, G);
            uint32_t newb = nbB & 3u;
            if (gl == G - 1) { const int jn = J + NSL - 1; newb = (jn >= 0 && jn < len2) ? s_p[jn] : 0u; }
            B = (B >> 2) | (newb << (2 * (NSL - 1)));
          } else {
            uint32_t nbA = __shfl_up_sync(0xffffffffu, A, 1, G);
            uint32_t newa = (nbA >> (2 * (NSL - 1))) & 3u;
            if (gl == 0) newa = (I >= 0 && I < len1) ? s_q[I] : 0u;
            A = ((A << 2) | newa) & (NSL == 16 ? 0xffffffffu : ((1u << (2 * NSL)) - 1u));
            I += 1; J += 1;
          }
        }
        continue;
      }
#pragma unroll
      for (int PAR = 0; PAR < 2; PAR++) {
        const int k = kk + PAR;
        int Hn;
        if (PAR == 0) { Hn = __shfl_up_sync(0xffffffffu, H[ND - 1], 1, G); if (gl == 0) Hn = SENT; }
        else { Hn = __shfl_down_sync(0xffffffffu, H[0], 1, G); if (gl == G - 1) Hn = SENT; }
        const uint32_t X = A ^ B;
        const int Jp = J + PAR;
        int Hnew[NSL];
        uint32_t mv = 0;
#pragma unroll
        for (int cc = 0; cc < NSL; cc++) {
          const int t = 2 * cc + PAR;
          const int hl = (PAR == 0 && cc == 0) ? Hn : H[t - 1 < 0 ? 0 : t - 1];
          const int hu = (PAR == 1 && cc == NSL - 1) ? Hn : H[t + 1 >= ND ? ND - 1 : t + 1];
          const bool eq = ((X >> (2 * cc)) & 3u) == 0u;
          const int diag = H[t] + (eq ? match : mismatch);
          const int i = I - cc, j = Jp + cc;
          const bool valid = (t >= tlo) && (t <= thi) && i >= 0 && j >= 0 && i <= len1 && j <= len2 && k <= nsteps;
          const int left = hl + ((i == len1) ? 0 : gap);                        // free end gaps, nwalign_endsfree.cpp:130-141
          const int up = hu + ((j == len2) ? 0 : gap);
          const int m = max(max(left, up), diag);
          const uint32_t pm = (up == m) ? 3u : ((left == m) ? 2u : 1u);
          int val = m;
          if (i == 0 || j == 0) val = 0;                                        // first row / column: ends-free (:91-101)
          mv |= pm << (2 * cc);
          Hnew[cc] = valid ? val : H[t];
        }
#pragma unroll
        for (int cc = 0; cc < NSL; cc++) H[2 * cc + PAR] = Hnew[cc];
        if (act && k <= nsteps) moves[(size_t)k * G + gl] = mv;
        if (PAR == 0) {           // even -> odd: parent window moves one base
          uint32_t nbB = __shfl_down_sync(0xffffffffu, B, 1, G);
          uint32_t newb = nbB & 3u;
          if (gl == G - 1) { const int jn = J + NSL - 1; newb = (act && jn >= 0 && jn < len2) ? s_p[jn] : 0u; }
          B = (B >> 2) | (newb << (2 * (NSL - 1)));
        } else {                  // odd -> even: query window moves one base
          uint32_t nbA = __shfl_up_sync(0xffffffffu, A, 1, G);
          uint32_t newa = (nbA >> (2 * (NSL - 1))) & 3u;
          if (gl == 0) newa = (act && I >= 0 && I < len1) ? s_q[I] : 0u;
          A = ((A << 2) | newa) & (NSL == 16 ? 0xffffffffu : ((1u << (2 * NSL)) - 1u));
          I += 1; J += 1;
        }
      }
    }
    __syncwarp();                 // the group's move words are visible to its lane 0
    // ---- traceback (nwalign_endsfree.cpp:166-190) by lane 0 of the group: reversed column masks, then bim_scan ----
    if (act && gl == 0) {
      const int MW = a.mask_words;
      for (int w = 0; w < MW; w++) { rQ[w] = 0; rP[w] = 0; rE[w] = 0; }
      int i = len1, j = len2, n = 0, neq = 0;
      while (i > 0 || j > 0) {
        int p;
        if (i == 0) p = 2;
        else if (j == 0) p = 3;
        else {
          const int dd = (j - i) + LB;
          const int ow = dd / ND, t = dd - ow * ND;
          p = (moves[(size_t)(i + j) * G + ow] >> (2 * (t >> 1))) & 3u;
        }
        const uint32_t bit = 1u << (n & 31);
        if (p == 1) {
          i--; j--;
          if (((s_q[i] ^ s_p[j]) & 3) == 0) { rE[n >> 5] |= bit; neq++; }
        } else if (p == 2) { j--; rQ[n >> 5] |= bit; }       // gap in the query row
        else if (p == 3) { i--; rP[n >> 5] |= bit; }         // gap in the parent row
        else { errflag = ERR_TRACE; break; }
        n++;
      }
      // forward masks: bit c of m = bit n-1-c of r
      const int nw = (n + 31) >> 5;
      for (int w = 0; w <= nw && w < MW; w++) {
        const int s = n - 32 * w - 32;                        // lowest reversed bit of this forward word
        uint32_t q0, p0, e0;
        if (s >= 0) {
          const int wi = s >> 5, sh = s & 31;
          q0 = __funnelshift_r(rQ[wi], wi + 1 < MW ? rQ[wi + 1] : 0u, sh);
          p0 = __funnelshift_r(rP[wi], wi + 1 < MW ? rP[wi + 1] : 0u, sh);
          e0 = __funnelshift_r(rE[wi], wi + 1 < MW ? rE[wi + 1] : 0u, sh);
        } else if (s > -32) { q0 = rQ[0] << (-s); p0 = rP[0] << (-s); e0 = rE[0] << (-s); }
        else { q0 = p0 = e0 = 0u; }
        mQ[w] = __brev(q0); mP[w] = __brev(p0); mE[w] = __brev(e0);
      }
      if (!errflag) {
        int v[5];
        bim_scan(mQ, mP, mE, n, neq, a.allow_one_off != 0, a.max_shift, v);
        if (a.raw5) { int32_t *o = a.raw5 + (size_t)jb * 5; o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3]; o[4] = v[4]; }
        if (a.rec) {
          const bool allowed = a.allow_one_off && v[4] >= a.min_one_off_par_dist;      // chimera.cpp:125-127
          const bool keep = v[0] + v[1] < len1;                                         // :129-142
          const size_t dst = a.dst_mode ? (size_t)((q - (uint32_t)a.q_add) / (uint32_t)a.q_mul - a.j0) * a.ncol + par : (size_t)jb;
          a.rec[dst] = keep ? bim_pack(v[0], v[1], a.allow_one_off ? v[2] : 0, a.allow_one_off ? v[3] : 0, allowed) : bim_pack(0, 0, 0, 0, allowed);
        }
      }
      cells_lane += band_cells_cf(len1, len2, lband, rband);
    }
    __syncwarp();
  }
  if (errflag) atomicMax(&a.ctr[2], (unsigned long long)errflag);
#pragma unroll
  for (int o = 16; o; o >>= 1) cells_lane += __shfl_xor_sync(0xffffffffu, cells_lane, o);
  if (lane == 0 && cells_lane) atomicAdd(&a.ctr[1], (unsigned long long)cells_lane);
}

static void pick(int slots_needed, int &G, int &ND) {
  const char *force = getenv("DADA2B_NWFWD");          // tuning override shared with k_nwfwd, e.g. "8x6"
  int fg = 0, fnd = 0;
  if (force && sscanf(force, "%dx%d", &fg, &fnd) == 2) { G = fg; ND = fnd; return; }
  if (slots_needed <= 40) { G = 4; ND = 10; }
  else if (slots_needed <= 48) { G = 8; ND = 6; }
  else if (slots_needed <= 64) { G = 8; ND = 8; }
  else if (slots_needed <= 128) { G = 16; ND = 8; }
  else if (slots_needed <= 256) { G = 32; ND = 8; }
  else { G = 0; ND = 0; }
}
static int bimfwd_grid(int num_sms) { return num_sms * 4; }

size_t bimfwd_scratch_words(int slots_needed, int maxlen, int num_sms) {
  int G, ND;
  pick(slots_needed, G, ND);
  if (!G) return 0;
  return (size_t)bimfwd_grid(num_sms) * 4 * (32 / G) * (size_t)(2 * maxlen + 2) * G;
}

template <int G, int ND> static void launch_one(const BimAlignArgs &a, int grid, size_t smem, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_bimfwd<G, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr_set = true; }
  k_bimfwd<G, ND><<<grid, 128, smem, s>>>(a);
}

// a.ptr_scratch must hold bimfwd_scratch_words(); a.ptr_words is set here (words per pair slot).
bool launch_bimfwd(const BimAlignArgs &a0, int slots_needed, unsigned long long njobs_upper, int num_sms, cudaStream_t s, int *grid_out) {
  int G, ND;
  pick(slots_needed, G, ND);
  if (!G || G * ND < slots_needed || a0.P.band < 0) return false;
  BimAlignArgs a = a0;
  const int PPW = 32 / G;
  a.ptr_words = (unsigned long long)(2 * a.sq.maxlen + 2) * G;
  const size_t smem = (size_t)4 * PPW * (2 * (a.seq_bytes >> 2) + 6 * a.mask_words) * 4;
  if (smem > 160 * 1024) return false;
  const unsigned long long warps = (njobs_upper + PPW - 1) / PPW;
  int grid = (int)std::min<unsigned long long>((warps + 3) / 4, (unsigned long long)bimfwd_grid(num_sms));
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = grid;
  if (G == 4 && ND == 10) launch_one<4, 10>(a, grid, smem, s);
  else if (G == 8 && ND == 6) launch_one<8, 6>(a, grid, smem, s);
  else if (G == 8 && ND == 8) launch_one<8, 8>(a, grid, smem, s);
  else if (G == 16 && ND == 8) launch_one<16, 8>(a, grid, smem, s);
  else if (G == 32 && ND == 8) launch_one<32, 8>(a, grid, smem, s);
  else return false;
  return true;
}

}  // namespace dd2
