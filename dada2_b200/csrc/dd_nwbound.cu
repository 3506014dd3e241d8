// k_nwbound16<G, ND>: the bound pass of the two-phase loop NW on the 16-bit SIMD datapath of sm_100a --
// TWO raws per lane group, scores and substitution counts as s16x2 halves.  Product code.  EXPERIMENTAL: DADA2B_BOUND16=1
// on top of DADA2B_TWOPHASE=1; checked on the host SIMT emulator only.
//
// The bound pass needs, for every (centre, raw) pair of a round, only the number of substitutions on the path the
// reference's traceback would take (precedence up > left > diag, nwalign_endsfree.cpp:147-156): lambda <= S_r * rho_r^nsubs
// then proves most pairs unstorable (cluster.cpp:192) and only the survivors go through the exact fp64 kernel.  All pairs
// of a launch share the centre, and raws of equal length share the whole band geometry, so two raws ride in the two
// halves of every register: per step one VIADD.16x2 each for left / up / diag, two VIMNMX.S16x2 (max with per-half
// predicate outputs: exactly the tie order needed) and per-half selects of the carried counts -- about 9 SASS
// instructions per cell and raw instead of ~24 in the scalar bound pass of k_nwfwd2.  Scores stay within int16 by
// construction (|score| <= maxlen * max|match, mismatch| + |gap|, checked on the host; out-of-band slots are pinned
// to a constant instead of accumulating a penalty).  Pairs of consecutive jobs with different lengths are handed to the
// scalar bound pass through a list; bands that do not fit go to the exact traceback kernel like everywhere else.
#include "dd_common.h"
#include "dd_kernels.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace dd2 {

struct BoundArgs {
  FwdArgs f;
  uint32_t *uneq_list;                 // raws whose partner has a different length -> scalar bound pass
  unsigned long long *uneq_count;
};

__device__ __forceinline__ uint32_t rep16(int v) { return ((uint32_t)v & 0xFFFFu) * 0x10001u; }
// per-half select: half h of the result = p_h ? a_h : b_h   (two predicated byte merges)
__device__ __forceinline__ uint32_t sel2(bool p_hi, bool p_lo, uint32_t a, uint32_t b) {
  uint32_t r = b;
  if (p_lo) r = __byte_perm(r, a, 0x3254);
  if (p_hi) r = __byte_perm(r, a, 0x7610);
  return r;
}

constexpr int NEG16 = -16000;          // out-of-band / unreached slots: far below any real score, far above INT16_MIN

struct BStep { int gl, len1, len2, tlo, thi, nsteps; uint32_t gap2, match2, dmm; };

// One anti-diagonal step (cells of parity PAR), in place, both raws at once.  FAST: interior step, branch-free.
template <int G, int ND, int PAR, bool FAST>
__device__ __forceinline__ void nb_step(uint32_t (&H)[ND], uint32_t (&N)[ND], const uint32_t (&INB)[ND], uint32_t A, uint32_t B0, uint32_t B1,
                                        int I, int J, int k, const BStep &c) {
  constexpr int NSL = ND / 2;
  const uint32_t NEG2 = rep16(NEG16);
  uint32_t Hn, Nn;
  if (PAR == 0) {
    Hn = __shfl_up_sync(0xffffffffu, H[ND - 1], 1, G); Nn = __shfl_up_sync(0xffffffffu, N[ND - 1], 1, G);
    if (c.gl == 0) { Hn = NEG2; Nn = 0; }
  } else {
    Hn = __shfl_down_sync(0xffffffffu, H[0], 1, G); Nn = __shfl_down_sync(0xffffffffu, N[0], 1, G);
    if (c.gl == G - 1) { Hn = NEG2; Nn = 0; }
  }
  const uint32_t X0 = A ^ B0, X1 = A ^ B1;
  // W: bit 2cc = raw 0 MISmatches in slot cc, bit 2cc+1 = raw 1 mismatches
  const uint32_t W = ((X0 | (X0 >> 1)) & 0x55555555u) | (((X1 | (X1 >> 1)) & 0x55555555u) << 1);
  const int Jp = J + PAR;
#pragma unroll
  for (int cc = 0; cc < NSL; cc++) {
    const int t = 2 * cc + PAR;
    const uint32_t hl = (PAR == 0 && cc == 0) ? Hn : H[t - 1 < 0 ? 0 : t - 1];
    const uint32_t nl = (PAR == 0 && cc == 0) ? Nn : N[t - 1 < 0 ? 0 : t - 1];
    const uint32_t hu = (PAR == 1 && cc == NSL - 1) ? Hn : H[t + 1 >= ND ? ND - 1 : t + 1];
    const uint32_t nu = (PAR == 1 && cc == NSL - 1) ? Nn : N[t + 1 >= ND ? ND - 1 : t + 1];
    const uint32_t x = (W >> (2 * cc)) & 3u;
    const uint32_t neq2 = (x * 0x8001u) & 0x10001u;                          // 1 in the half of every raw that mismatches here
    const uint32_t diag = __vadd2(__vadd2(H[t], c.match2), neq2 * c.dmm);    // + match, or + match + (mismatch - match)
    bool pl_hi, pl_lo, pu_hi, pu_lo;
    if (FAST) {
      const uint32_t left = __vadd2(hl, c.gap2), up = __vadd2(hu, c.gap2);
      const uint32_t t2 = __vibmax_s16x2(left, diag, &pl_hi, &pl_lo);       // left >= diag: left wins the tie
      const uint32_t m = __vibmax_s16x2(up, t2, &pu_hi, &pu_lo);            // up >= max(left, diag): up wins the tie
      N[t] = sel2(pu_hi, pu_lo, nu, sel2(pl_hi, pl_lo, nl, __vadd2(N[t], neq2)));
      H[t] = (m & INB[t]) | (NEG2 & ~INB[t]);                                 // one LOP3: out-of-band slots pinned to NEG16
    } else {
      const int i = I - cc, j = Jp + cc;
      const bool valid = (t >= c.tlo) && (t <= c.thi) && i >= 0 && j >= 0 && i <= c.len1 && j <= c.len2 && k <= c.nsteps;
      const uint32_t left = __vadd2(hl, (i == c.len1) ? 0u : c.gap2);       // free end gaps, nwalign_endsfree.cpp:130-141
      const uint32_t up = __vadd2(hu, (j == c.len2) ? 0u : c.gap2);
      const uint32_t t2 = __vibmax_s16x2(left, diag, &pl_hi, &pl_lo);
      const uint32_t m = __vibmax_s16x2(up, t2, &pu_hi, &pu_lo);
      const bool edge = (i == 0) || (j == 0);                                // first row / column: score 0, no substitutions yet (:91-101)
      const uint32_t np = edge ? 0u : sel2(pu_hi, pu_lo, nu, sel2(pl_hi, pl_lo, nl, __vadd2(N[t], neq2)));
      H[t] = valid ? (edge ? 0u : m) : H[t];
      N[t] = valid ? np : N[t];
    }
  }
}

template <int G, int ND>
__global__ void __launch_bounds__(128) k_nwbound16(BoundArgs ba) {
  constexpr int NSL = ND / 2;
  constexpr int PPW = 32 / G;             // lane groups per warp; every group carries two raws
  static_assert(ND % 2 == 0 && NSL <= 16, "base windows are one 32-bit register each");
  const FwdArgs &a = ba.f;
  extern __shared__ uint32_t smem[];
  const AlnParams &P = a.P;
  uint8_t *s_cen = (uint8_t *)smem;                                       // centre bases (one centre per launch)
  const int nwarps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gid = lane / G, gl = lane % G;
  uint8_t *s_raw0 = s_cen + a.seq_bytes + (size_t)(wid * PPW + gid) * 2 * a.seq_bytes, *s_raw1 = s_raw0 + a.seq_bytes;
  const unsigned long long njobs = *a.njobs_ptr, npairs = (njobs + 1) >> 1;
  if ((unsigned long long)blockIdx.x * nwarps * PPW >= npairs) return;
  const int len1 = a.in.len[a.centre_idx];
  {
    const uint32_t *crow = a.in.seq2 + (size_t)a.centre_idx * a.in.SW;
    for (int p = threadIdx.x; p < len1; p += blockDim.x) s_cen[p] = (uint8_t)((crow[p >> 4] >> (2 * (p & 15))) & 3u);
  }
  __syncthreads();
  long long cells_lane = 0;
  const uint32_t NEG2 = rep16(NEG16);

  for (unsigned long long base = (unsigned long long)(blockIdx.x * nwarps + wid) * PPW; base < npairs;
       base += (unsigned long long)gridDim.x * nwarps * PPW) {
    const unsigned long long pj = base + gid;
    bool act = pj < npairs;
    const bool two = act && 2 * pj + 1 < njobs;
    const uint32_t r0 = act ? a.jobs[2 * pj] : 0, r1 = two ? a.jobs[2 * pj + 1] : r0;
    const int len2 = act ? (int)a.in.len[r0] : len1;
    if (two && (int)a.in.len[r1] != len2) {          // different band geometry: both go to the scalar bound pass
      if (gl == 0) { const unsigned long long s = atomicAdd(ba.uneq_count, 2ull); ba.uneq_list[s] = r0; ba.uneq_list[s + 1] = r1; }
      act = false;
    }
    if (act) {
      const uint32_t *row0 = a.in.seq2 + (size_t)r0 * a.in.SW, *row1 = a.in.seq2 + (size_t)r1 * a.in.SW;
      for (int p = gl; p < len2; p += G) {
        s_raw0[p] = (uint8_t)((row0[p >> 4] >> (2 * (p & 15))) & 3u);
        s_raw1[p] = (uint8_t)((row1[p >> 4] >> (2 * (p & 15))) & 3u);
      }
    }
    __syncwarp();
    // ---- band geometry (nwalign_endsfree.cpp:101-111), shared by both raws ----
    int lband, rband;
    if (len2 > len1) { lband = P.band; rband = P.band + len2 - len1; }
    else if (len1 > len2) { lband = P.band + len1 - len2; rband = P.band; }
    else { lband = P.band; rband = P.band; }
    const int lb = min(lband, len1), rb = min(rband, len2);
    const int LB = (lb + 1) & ~1;
    const int lo = LB - lb, hi = LB + rb;
    if (act && (P.band < 0 || hi >= G * ND)) {       // does not fit this instantiation: exact traceback kernel (k_align)
      if (gl == 0) {
        const unsigned long long s = atomicAdd(a.fb_count, two ? 2ull : 1ull);
        a.fb_list[s] = r0; if (two) a.fb_list[s + 1] = r1;
      }
      act = false;
    }
    const int tlo = lo - gl * ND, thi = hi - gl * ND;
    const int D = gl * ND - LB;
    const int nsteps = act ? len1 + len2 : 0;
    int maxsteps = nsteps;
#pragma unroll
    for (int o = 16; o; o >>= 1) maxsteps = max(maxsteps, __shfl_xor_sync(0xffffffffu, maxsteps, o));

    uint32_t H[ND], N[ND];
    uint32_t INB[ND];
#pragma unroll
    for (int t = 0; t < ND; t++) { H[t] = NEG2; N[t] = 0; INB[t] = ((t >= tlo) && (t <= thi)) ? 0xFFFFFFFFu : 0u; }
    int I = -(D / 2), J = D / 2;
    uint32_t A = 0, B0 = 0, B1 = 0;
#pragma unroll
    for (int cc = 0; cc < NSL; cc++) {
      const int i1 = I - 1 - cc, j1 = J - 1 + cc;
      const uint32_t b1 = (i1 >= 0 && i1 < len1) ? s_cen[i1] : 0u;
      const bool ok = act && j1 >= 0 && j1 < len2;
      A |= b1 << (2 * cc); B0 |= (ok ? (uint32_t)s_raw0[j1] : 0u) << (2 * cc); B1 |= (ok ? (uint32_t)s_raw1[j1] : 0u) << (2 * cc);
    }
    int kf_lo = act ? max(lb, rb) + 2 : 0, kf_hi = act ? min(2 * len1 - lb, 2 * len2 - rb) - 1 : 0x3fffffff;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      kf_lo = max(kf_lo, __shfl_xor_sync(0xffffffffu, kf_lo, o));
      kf_hi = min(kf_hi, __shfl_xor_sync(0xffffffffu, kf_hi, o));
    }
    if (!a.fast_ok) kf_hi = -1;
    auto advB = [&]() {
      uint32_t n0 = __shfl_down_sync(0xffffffffu, B0, 1, G) & 3u, n1 = __shfl_down_sync(0xffffffffu, B1, 1, G) & 3u;
      if (gl == G - 1) { const int jn = J + NSL - 1; const bool ok = act && jn >= 0 && jn < len2; n0 = ok ? s_raw0[jn] : 0u; n1 = ok ? s_raw1[jn] : 0u; }
      B0 = (B0 >> 2) | (n0 << (2 * (NSL - 1))); B1 = (B1 >> 2) | (n1 << (2 * (NSL - 1)));
    };
    auto advA = [&]() {
      uint32_t na = (__shfl_up_sync(0xffffffffu, A, 1, G) >> (2 * (NSL - 1))) & 3u;
      if (gl == 0) na = (I >= 0 && I < len1) ? s_cen[I] : 0u;
      A = ((A << 2) | na) & (NSL == 16 ? 0xffffffffu : ((1u << (2 * NSL)) - 1u));
      I += 1; J += 1;
    };
    const BStep cx{gl, len1, len2, tlo, thi, nsteps, rep16(P.gap), rep16(P.match), (uint32_t)(P.mismatch - P.match) & 0xFFFFu};
    int kk = 0;
    const int kfa = (kf_lo + 1) & ~1;
    for (; kk < kfa && kk <= maxsteps; kk += 2) {
      nb_step<G, ND, 0, false>(H, N, INB, A, B0, B1, I, J, kk, cx); advB();
      nb_step<G, ND, 1, false>(H, N, INB, A, B0, B1, I, J, kk + 1, cx); advA();
    }
    for (; kk + 1 <= kf_hi && kk <= maxsteps; kk += 2) {
      nb_step<G, ND, 0, true>(H, N, INB, A, B0, B1, I, J, kk, cx); advB();
      nb_step<G, ND, 1, true>(H, N, INB, A, B0, B1, I, J, kk + 1, cx); advA();
    }
    for (; kk <= maxsteps; kk += 2) {
      nb_step<G, ND, 0, false>(H, N, INB, A, B0, B1, I, J, kk, cx); advB();
      nb_step<G, ND, 1, false>(H, N, INB, A, B0, B1, I, J, kk + 1, cx); advA();
    }
    // ---- result: cell (len1, len2); one bound test per raw (same arithmetic as the scalar bound pass, dd_nwfwd.cu) ----
    const int tf = (len2 - len1 + LB) - gl * ND;
    uint32_t nsw = 0;
#pragma unroll
    for (int t = 0; t < ND; t++) if (t == tf) nsw = N[t];
    const bool owner = act && tf >= 0 && tf < ND;
    bool s0 = false, s1 = false;
    if (owner) {
      cells_lane += band_cells_cf(len1, len2, lband, rband) * (two ? 2 : 1);
      const double b0 = a.raw_S[r0] * pow(a.raw_rho[r0], (double)(nsw & 0xFFFFu)) * (double)a.total_reads * (1.0 + 1e-9);
      s0 = !(b0 <= a.st.E_minmax[r0]) || b0 < 1e-280;
      if (two) {
        const double b1 = a.raw_S[r1] * pow(a.raw_rho[r1], (double)(nsw >> 16)) * (double)a.total_reads * (1.0 + 1e-9);
        s1 = !(b1 <= a.st.E_minmax[r1]) || b1 < 1e-280;
      }
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, s0), m1 = __ballot_sync(0xffffffffu, s1);
    unsigned long long bs = 0;
    if (lane == 0 && (m0 | m1)) bs = atomicAdd(a.surv_count, (unsigned long long)(__popc(m0) + __popc(m1)));
    bs = __shfl_sync(0xffffffffu, bs, 0);
    const unsigned lt = (1u << lane) - 1u;
    if (s0) a.surv_list[bs + __popc(m0 & lt)] = r0;
    if (s1) a.surv_list[bs + __popc(m0) + __popc(m1 & lt)] = r1;
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) cells_lane += __shfl_xor_sync(0xffffffffu, cells_lane, o);
  if (lane == 0 && cells_lane) atomicAdd(&a.st.ctr[CTR_CELLS], (unsigned long long)cells_lane);
}

template <int G, int ND> static void launch_one(const BoundArgs &a, int grid, size_t smem, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_nwbound16<G, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr_set = true; }
  k_nwbound16<G, ND><<<grid, 128, smem, s>>>(a);
}

// Bound pass over a.jobs two raws at a time; raws whose neighbour in the job list has a different length come back in
// uneq_list (count in *uneq_count, to be zeroed by the caller) for the scalar bound pass.  false: nothing launched.
bool launch_nwbound16(const FwdArgs &f, uint32_t *uneq_list, unsigned long long *uneq_count, int slots_needed, unsigned long long njobs_upper,
                      unsigned long long njobs_hint, int num_sms, cudaStream_t s) {
  // int16 headroom: every real score must stay well inside (NEG16, -NEG16)
  const long worst = (long)f.in.maxlen * std::max(std::abs(f.P.mismatch), std::abs(f.P.match)) + std::abs(f.P.gap) + 16;
  if (worst >= -NEG16 / 2 || f.P.homo || f.P.band < 0) return false;
  int G, ND;
  const bool big = njobs_hint > (unsigned long long)num_sms * 1024;
  if (slots_needed <= 40) { G = big ? 4 : 8; ND = big ? 10 : 6; }
  else if (slots_needed <= 48) { G = 8; ND = 6; }
  else if (slots_needed <= 64) { G = 8; ND = 8; }
  else if (slots_needed <= 128) { G = 16; ND = 8; }
  else if (slots_needed <= 256) { G = 32; ND = 8; }
  else return false;
  if (G * ND < slots_needed) return false;
  BoundArgs a{f, uneq_list, uneq_count};
  const int PPW = 32 / G;
  const size_t smem = (size_t)f.seq_bytes + (size_t)4 * PPW * 2 * f.seq_bytes;
  if (smem > 160 * 1024) return false;
  const unsigned long long warps = ((njobs_upper + 1) / 2 + PPW - 1) / PPW;
  int grid = (int)std::min<unsigned long long>((warps + 3) / 4, (unsigned long long)num_sms * 16);
  if (grid < 1) grid = 1;
  count_launch(1);
  if (G == 4 && ND == 10) launch_one<4, 10>(a, grid, smem, s);
  else if (G == 8 && ND == 6) launch_one<8, 6>(a, grid, smem, s);
  else if (G == 8 && ND == 8) launch_one<8, 8>(a, grid, smem, s);
  else if (G == 16 && ND == 8) launch_one<16, 8>(a, grid, smem, s);
  else launch_one<32, 8>(a, grid, smem, s);
  return true;
}

}  // namespace dd2
