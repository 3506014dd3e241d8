// k_bimfwd16<G, ND>: k_bimfwd (dd_bimfwd.cu) on the 16-bit SIMD datapath of sm_100a -- TWO (query, parent) jobs per lane
// group.  Product code.  EXPERIMENTAL: DADA2B_BIMFWD=2.
//
// k_bim_need emits the jobs of one query in runs, and in amplicon tables most parents have the same length, so two
// consecutive jobs usually share the query AND the band geometry: they ride in the two halves of every score register
// (VIADD.16x2, VIMNMX.S16x2 with one predicate per half = the up > left > diag tie order), and their 2-bit moves share one
// move word per lane and step (low half / high half; 2 * NSL <= 16 bits each).  Lanes 0 and 1 of the group then walk the two
// paths back concurrently and evaluate bim_scan as in k_bimfwd.  Pairs of jobs that do not match (different query or
// parent length) go to k_bimfwd through a list; bands that do not fit go to k_bim_align.
#include "dd_bimera.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace dd2 {

__device__ __forceinline__ uint32_t b16_rep(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }
__device__ __forceinline__ uint32_t b16_sel2(bool p_hi, bool p_lo, uint32_t a, uint32_t b) {   // half h = p_h ? a_h : b_h
  uint32_t r = b;
  if (p_lo) r = __byte_perm(r, a, 0x3254);
  if (p_hi) r = __byte_perm(r, a, 0x7610);
  return r;
}
constexpr int B16_NEG = -16000;

struct B16Args {
  BimAlignArgs a;
  uint32_t *uneq_list;                 // jobs whose neighbour in the list has another query / parent length -> k_bimfwd
  unsigned long long *uneq_count;
};

template <int G, int ND>
__global__ void __launch_bounds__(128) k_bimfwd16(B16Args ba) {
  constexpr int NSL = ND / 2;
  constexpr int PPW = 32 / G;
  static_assert(ND % 2 == 0 && NSL <= 8 && G >= 2, "two 2*NSL-bit move fields per word; two traceback lanes per group");
  const BimAlignArgs &a = ba.a;
  extern __shared__ uint32_t smem[];
  const AlnParams &P = a.P;
  const int nwarps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane / G, gl = lane % G;
  // per group: query bases | parent 0 bases | parent 1 bases | 2 x 6 x mask_words
  const int grp_words = 3 * (a.seq_bytes >> 2) + 12 * a.mask_words;
  uint32_t *s_grp = smem + (size_t)(wid * PPW + gid) * grp_words;
  uint8_t *s_q = (uint8_t *)s_grp, *s_p0 = s_q + a.seq_bytes, *s_p1 = s_p0 + a.seq_bytes;
  uint32_t *s_masks = s_grp + 3 * (a.seq_bytes >> 2);
  const unsigned long long njobs = a.njobs_ptr ? *a.njobs_ptr : a.njobs_fixed, npairs = (njobs + 1) >> 1;
  if ((unsigned long long)blockIdx.x * nwarps * PPW >= npairs) return;
  const int match = P.match, mismatch = P.mismatch, gap = P.gap;
  const uint32_t NEG2 = b16_rep(B16_NEG), gap2 = b16_rep(gap), match2 = b16_rep(match), dmm = (uint32_t)(mismatch - match) & 0xFFFFu;
  const uint32_t ONE2 = 0x00010001u, TWO2 = 0x00020002u, THREE2 = 0x00030003u;
  uint32_t *moves = a.ptr_scratch + ((size_t)(blockIdx.x * nwarps + wid) * PPW + gid) * a.ptr_words;
  int errflag = 0;
  long long cells_lane = 0;

  for (unsigned long long base = (unsigned long long)(blockIdx.x * nwarps + wid) * PPW; base < npairs;
       base += (unsigned long long)gridDim.x * nwarps * PPW) {
    const unsigned long long pj = base + gid;
    bool act = pj < npairs;
    const bool two = act && 2 * pj + 1 < njobs;
    const unsigned long long j0 = 2 * pj, j1 = two ? 2 * pj + 1 : 2 * pj;
    const uint32_t q = act ? a.jq[j0] : 0, par0 = act ? a.jk[j0] : 0, par1 = act ? a.jk[j1] : 0;
    const int len1 = act ? (int)a.sq.len[q] : 16;
    const int len2 = act ? (int)a.sq.len[par0] : len1;
    if (two && (a.jq[j1] != q || (int)a.sq.len[par1] != len2)) {          // different query or geometry: scalar register kernel
      if (gl == 0) { const unsigned long long s = atomicAdd(ba.uneq_count, 2ull); ba.uneq_list[s] = (uint32_t)j0; ba.uneq_list[s + 1] = (uint32_t)j1; }
      act = false;
    }
    if (act) {
      const uint32_t *qrow = a.sq.seq2 + (size_t)q * a.sq.SW, *r0 = a.sq.seq2 + (size_t)par0 * a.sq.SW, *r1 = a.sq.seq2 + (size_t)par1 * a.sq.SW;
      for (int p = gl; p < len1; p += G) s_q[p] = (uint8_t)((qrow[p >> 4] >> (2 * (p & 15))) & 3u);
      for (int p = gl; p < len2; p += G) {
        s_p0[p] = (uint8_t)((r0[p >> 4] >> (2 * (p & 15))) & 3u);
        s_p1[p] = (uint8_t)((r1[p >> 4] >> (2 * (p & 15))) & 3u);
      }
    }
    __syncwarp();
    int lband, rband;                               // nwalign_endsfree.cpp:101-111
    if (len2 > len1) { lband = P.band; rband = P.band + len2 - len1; }
    else if (len1 > len2) { lband = P.band + len1 - len2; rband = P.band; }
    else { lband = P.band; rband = P.band; }
    const int lb = min(lband, len1), rb = min(rband, len2);
    const int LB = (lb + 1) & ~1;
    const int lo = LB - lb, hi = LB + rb;
    if (act && (P.band < 0 || hi >= G * ND)) {      // does not fit: warp-per-pair traceback kernel
      if (gl == 0) { const unsigned long long s = atomicAdd(a.fb_count, two ? 2ull : 1ull); a.fb_list[s] = (uint32_t)j0; if (two) a.fb_list[s + 1] = (uint32_t)j1; }
      act = false;
    }
    const int tlo = lo - gl * ND, thi = hi - gl * ND;
    const int D = gl * ND - LB;
    const int nsteps = act ? len1 + len2 : 0;
    int maxsteps = nsteps;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxsteps = max(maxsteps, __shfl_xor_sync(0xffffffffu, maxsteps, o));
    uint32_t H[ND], INB[ND];
#pragma unroll
    for (int t = 0; t < ND; t++) { H[t] = NEG2; INB[t] = ((t >= tlo) && (t <= thi)) ? 0xFFFFFFFFu : 0u; }
    int I = -(D / 2), J = D / 2;
    uint32_t A = 0, B0 = 0, B1 = 0;
#pragma unroll
    for (int cc = 0; cc < NSL; cc++) {
      const int i1 = I - 1 - cc, j1x = J - 1 + cc;
      const uint32_t b1 = (act && i1 >= 0 && i1 < len1) ? s_q[i1] : 0u;
      const bool ok = act && j1x >= 0 && j1x < len2;
      A |= b1 << (2 * cc); B0 |= (ok ? (uint32_t)s_p0[j1x] : 0u) << (2 * cc); B1 |= (ok ? (uint32_t)s_p1[j1x] : 0u) << (2 * cc);
    }
    int kf_lo = act ? max(lb, rb) + 2 : 0, kf_hi = act ? min(2 * len1 - lb, 2 * len2 - rb) - 1 : 0x3fffffff;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      kf_lo = max(kf_lo, __shfl_xor_sync(0xffffffffu, kf_lo, o));
      kf_hi = min(kf_hi, __shfl_xor_sync(0xffffffffu, kf_hi, o));
    }
    if (!a.fast_ok) kf_hi = -1;

    for (int kk = 0; kk <= maxsteps; kk += 2) {
      const bool fast = kk >= kf_lo && kk + 1 <= kf_hi;
#pragma unroll
      for (int PAR = 0; PAR < 2; PAR++) {
        const int k = kk + PAR;
        uint32_t Hn;
        if (PAR == 0) { Hn = __shfl_up_sync(0xffffffffu, H[ND - 1], 1, G); if (gl == 0) Hn = NEG2; }
        else { Hn = __shfl_down_sync(0xffffffffu, H[0], 1, G); if (gl == G - 1) Hn = NEG2; }
        const uint32_t X0 = A ^ B0, X1 = A ^ B1;
        const uint32_t W = ((X0 | (X0 >> 1)) & 0x55555555u) | (((X1 | (X1 >> 1)) & 0x55555555u) << 1);   // bit 2cc: job 0 mismatches in slot cc, bit 2cc+1: job 1
        const int Jp = J + PAR;
        uint32_t mv = 0;
        if (fast) {
#pragma unroll
          for (int cc = 0; cc < NSL; cc++) {
            const int t = 2 * cc + PAR;
            const uint32_t hl = (PAR == 0 && cc == 0) ? Hn : H[t - 1 < 0 ? 0 : t - 1];
            const uint32_t hu = (PAR == 1 && cc == NSL - 1) ? Hn : H[t + 1 >= ND ? ND - 1 : t + 1];
            const uint32_t neq2 = ((((W >> (2 * cc)) & 3u)) * 0x8001u) & 0x10001u;
            const uint32_t diag = __vadd2(__vadd2(H[t], match2), neq2 * dmm);
            const uint32_t left = __vadd2(hl, gap2), up = __vadd2(hu, gap2);
            bool pl_hi, pl_lo, pu_hi, pu_lo;
            const uint32_t t2 = __vibmax_s16x2(left, diag, &pl_hi, &pl_lo);      // left wins the tie against diag
            const uint32_t m = __vibmax_s16x2(up, t2, &pu_hi, &pu_lo);           // up wins every tie (nwalign_endsfree.cpp:147-156)
            mv |= b16_sel2(pu_hi, pu_lo, THREE2, b16_sel2(pl_hi, pl_lo, TWO2, ONE2)) << (2 * cc);
            H[t] = (m & INB[t]) | (NEG2 & ~INB[t]);
          }
          moves[(size_t)k * G + gl] = mv;
        } else {
#pragma unroll
          for (int cc = 0; cc < NSL; cc++) {
            const int t = 2 * cc + PAR;
            const uint32_t hl = (PAR == 0 && cc == 0) ? Hn : H[t - 1 < 0 ? 0 : t - 1];
            const uint32_t hu = (PAR == 1 && cc == NSL - 1) ? Hn : H[t + 1 >= ND ? ND - 1 : t + 1];
            const uint32_t neq2 = ((((W >> (2 * cc)) & 3u)) * 0x8001u) & 0x10001u;
            const uint32_t diag = __vadd2(__vadd2(H[t], match2), neq2 * dmm);
            const int i = I - cc, j = Jp + cc;
            const bool valid = (t >= tlo) && (t <= thi) && i >= 0 && j >= 0 && i <= len1 && j <= len2 && k <= nsteps;
            const uint32_t left = __vadd2(hl, (i == len1) ? 0u : gap2), up = __vadd2(hu, (j == len2) ? 0u : gap2);   // free end gaps (:130-141)
            bool pl_hi, pl_lo, pu_hi, pu_lo;
            const uint32_t t2 = __vibmax_s16x2(left, diag, &pl_hi, &pl_lo);
            const uint32_t m = __vibmax_s16x2(up, t2, &pu_hi, &pu_lo);
            mv |= b16_sel2(pu_hi, pu_lo, THREE2, b16_sel2(pl_hi, pl_lo, TWO2, ONE2)) << (2 * cc);
            H[t] = valid ? ((i == 0 || j == 0) ? 0u : m) : H[t];                 // first row / column: ends-free (:91-101)
          }
          if (act && k <= nsteps) moves[(size_t)k * G + gl] = mv;
        }
        if (PAR == 0) {
          uint32_t n0 = __shfl_down_sync(0xffffffffu, B0, 1, G) & 3u, n1 = __shfl_down_sync(0xffffffffu, B1, 1, G) & 3u;
          if (gl == G - 1) { const int jn = J + NSL - 1; const bool ok = act && jn >= 0 && jn < len2; n0 = ok ? s_p0[jn] : 0u; n1 = ok ? s_p1[jn] : 0u; }
          B0 = (B0 >> 2) | (n0 << (2 * (NSL - 1))); B1 = (B1 >> 2) | (n1 << (2 * (NSL - 1)));
        } else {
          uint32_t na = (__shfl_up_sync(0xffffffffu, A, 1, G) >> (2 * (NSL - 1))) & 3u;
          if (gl == 0) na = (act && I >= 0 && I < len1) ? s_q[I] : 0u;
          A = ((A << 2) | na) & ((1u << (2 * NSL)) - 1u);
          I += 1; J += 1;
        }
      }
    }
    __syncwarp();
    // ---- traceback: lane h of the group walks job h (nwalign_endsfree.cpp:166-190), then bim_scan on its column masks ----
    if (act && gl < (two ? 2 : 1)) {
      const int h = gl;
      const uint8_t *s_p = h ? s_p1 : s_p0;
      const unsigned long long jb = h ? j1 : j0;
      const uint32_t par = h ? par1 : par0;
      const int MW = a.mask_words;
      uint32_t *rQ = s_masks + (size_t)h * 6 * MW, *rP = rQ + MW, *rE = rP + MW, *mQ = rE + MW, *mP = mQ + MW, *mE = mP + MW;
      for (int w = 0; w < MW; w++) { rQ[w] = 0; rP[w] = 0; rE[w] = 0; }
      int i = len1, j = len2, n = 0, neq = 0;
      bool bad = false;
      while (i > 0 || j > 0) {
        int p;
        if (i == 0) p = 2;
        else if (j == 0) p = 3;
        else {
          const int dd = (j - i) + LB;
          const int ow = dd / ND, t = dd - ow * ND;
          p = (moves[(size_t)(i + j) * G + ow] >> (16 * h + 2 * (t >> 1))) & 3u;
        }
        const uint32_t bit = 1u << (n & 31);
        if (p == 1) { i--; j--; if (((s_q[i] ^ s_p[j]) & 3) == 0) { rE[n >> 5] |= bit; neq++; } }
        else if (p == 2) { j--; rQ[n >> 5] |= bit; }
        else if (p == 3) { i--; rP[n >> 5] |= bit; }
        else { bad = true; break; }
        n++;
      }
      if (bad) errflag = ERR_TRACE;
      else {
        const int nw = (n + 31) >> 5;
        for (int w = 0; w <= nw && w < MW; w++) {                // forward masks: bit c = reversed bit n-1-c
          const int s = n - 32 * w - 32;
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
        int v[5];
        bim_scan(mQ, mP, mE, n, neq, a.allow_one_off != 0, a.max_shift, v);
        if (a.raw5) { int32_t *o = a.raw5 + (size_t)jb * 5; o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3]; o[4] = v[4]; }
        if (a.rec) {
          const bool allowed = a.allow_one_off && v[4] >= a.min_one_off_par_dist;      // chimera.cpp:125-127
          const bool keep = v[0] + v[1] < len1;                                         // :129-142
          const size_t dst = a.dst_mode ? (size_t)((q - (uint32_t)a.q_add) / (uint32_t)a.q_mul - a.j0) * a.ncol + par : (size_t)jb;
          a.rec[dst] = keep ? bim_pack(v[0], v[1], a.allow_one_off ? v[2] : 0, a.allow_one_off ? v[3] : 0, allowed) : bim_pack(0, 0, 0, 0, allowed);
        }
        cells_lane += band_cells_cf(len1, len2, lband, rband);
      }
    }
    __syncwarp();
  }
  if (errflag) atomicMax(&a.ctr[2], (unsigned long long)errflag);
#pragma unroll
  for (int o = 16; o; o >>= 1) cells_lane += __shfl_xor_sync(0xffffffffu, cells_lane, o);
  if (lane == 0 && cells_lane) atomicAdd(&a.ctr[1], (unsigned long long)cells_lane);
}

static bool pick16(int slots_needed, int &G, int &ND) {
  if (slots_needed <= 40) { G = 4; ND = 10; }
  else if (slots_needed <= 48) { G = 8; ND = 6; }
  else if (slots_needed <= 64) { G = 8; ND = 8; }
  else if (slots_needed <= 128) { G = 16; ND = 8; }
  else if (slots_needed <= 256) { G = 32; ND = 8; }
  else return false;
  return true;
}
template <int G, int ND> static void launch_one16(const B16Args &a, int grid, size_t smem, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_bimfwd16<G, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr_set = true; }
  k_bimfwd16<G, ND><<<grid, 128, smem, s>>>(a);
}

// a0.ptr_scratch must hold bimfwd_scratch_words() (same layout as k_bimfwd: one move word per lane and step and group).
bool launch_bimfwd16(const BimAlignArgs &a0, uint32_t *uneq_list, unsigned long long *uneq_count, int slots_needed,
                     unsigned long long njobs_upper, int num_sms, cudaStream_t s) {
  int G, ND;
  if (!pick16(slots_needed, G, ND) || G * ND < slots_needed || a0.P.band < 0) return false;
  const long worst = (long)a0.sq.maxlen * std::max(std::abs(a0.P.mismatch), std::abs(a0.P.match)) + std::abs(a0.P.gap) + 16;
  if (worst >= -B16_NEG / 2) return false;                       // scores must stay inside the int16 headroom
  B16Args a{a0, uneq_list, uneq_count};
  const int PPW = 32 / G;
  a.a.ptr_words = (unsigned long long)(2 * a0.sq.maxlen + 2) * G;
  const size_t smem = (size_t)4 * PPW * (3 * (a0.seq_bytes >> 2) + 12 * a0.mask_words) * 4;
  if (smem > 160 * 1024) return false;
  const unsigned long long warps = ((njobs_upper + 1) / 2 + PPW - 1) / PPW;
  int grid = (int)std::min<unsigned long long>((warps + 3) / 4, (unsigned long long)num_sms * 4);
  if (grid < 1) grid = 1;
  if (G == 4 && ND == 10) launch_one16<4, 10>(a, grid, smem, s);
  else if (G == 8 && ND == 6) launch_one16<8, 6>(a, grid, smem, s);
  else if (G == 8 && ND == 8) launch_one16<8, 8>(a, grid, smem, s);
  else if (G == 16 && ND == 8) launch_one16<16, 8>(a, grid, smem, s);
  else launch_one16<32, 8>(a, grid, smem, s);
  return true;
}

}  // namespace dd2
