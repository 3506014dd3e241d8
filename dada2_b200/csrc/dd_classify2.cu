// k_classify2: the k-mer screen of the loop with a pivot pre-filter (product code, sm_100a) -- EXPERIMENTAL, selected
// with DADA2B_PIVOT=1; checked on the host SIMT emulator of tests/emu, not yet on hardware.  k_classify stays the
// default (and keeps serving the pair mode).
//
// 88 % of the (raw, seed) pairs of a round are shrouded by the k-mer screen (raw_align, nwalign_endsfree.cpp:51), and
// k_classify reads the raw's packed row and extracts its 5-mers to find that out.  Two thirds of all pairs can be
// proven shrouded from 8 bytes instead: sum_k min(c_a[k], c_b[k]) = (n_a + n_b - L1(a, b)) / 2 over 5-mer count
// vectors, and L1 obeys the triangle inequality.  Every raw remembers a pivot -- the earlier centre p it shares most
// 5-mers with, and that exact min-sum.  For a new seed s
//     minsum(r, s) <= floor((n_r + n_s - |L1(r, p) - L1(p, s)|) / 2)
// where L1(p, s) needs one min-sum per existing centre and round (k_seed_dists).  kdist is monotone in the min-sum, so
// "1 - bound/denom > cutoff" proves the pair shrouded exactly -- the same argument as k_classify's presence-bitmap
// tier.  Lanes test 32 raws at a time; the survivors go through the unchanged warp-per-pair screen one by one, which
// also refreshes the pivot whenever it computes a larger exact min-sum.
#include "dd_common.h"
#include "dd_kernels.h"

namespace dd2 {

namespace {
__device__ __forceinline__ unsigned c2_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned c2_kmer_at(const uint32_t *row, int p) {
  const uint32_t w0 = row[p >> 4], w1 = row[(p + 4) >> 4];
  return __funnelshift_r(w0, w1, 2 * (p & 15)) & 0x3FFu;
}
__device__ __forceinline__ int c2_warp_sum(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
}  // namespace

// seed_ms[c] = sum_k min(count_c[k], count_seed[k]) for every existing centre c < nclust (one warp per centre).
// shared: seed counts (512 words) + per-warp scratch counts (512 words each).
__global__ void __launch_bounds__(256) k_seed_dists(DevIn in, const uint32_t *cl_center, int nclust, uint32_t seed, uint16_t *seed_ms) {
  extern __shared__ uint32_t smem[];
  const int nwarps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = c2_lane();
  uint32_t *cen_cnt = smem, *wtab = smem + 512 + wid * 512;
  for (int x = threadIdx.x; x < 512 + nwarps * 512; x += blockDim.x) smem[x] = 0;
  __syncthreads();
  const uint32_t *srow = in.seq2 + (size_t)seed * in.SW;
  const int slen = in.len[seed];
  for (int p = threadIdx.x; p + KMER <= slen; p += blockDim.x) { const unsigned km = c2_kmer_at(srow, p); atomicAdd(&cen_cnt[km >> 1], 1u << (16 * (km & 1))); }
  __syncthreads();
  for (int c = blockIdx.x * nwarps + wid; c < nclust; c += gridDim.x * nwarps) {
    const uint32_t cr = cl_center[c];
    const uint32_t *row = in.seq2 + (size_t)cr * in.SW;
    const int len = in.len[cr];
    int ms = 0;
    for (int p = lane; p + KMER <= len; p += 32) {
      const unsigned km = c2_kmer_at(row, p), sh = 16 * (km & 1);
      const unsigned old = atomicAdd(&wtab[km >> 1], 1u << sh);
      ms += ((old >> sh) & 0xFFFFu) < ((cen_cnt[km >> 1] >> sh) & 0xFFFFu);
    }
    ms = c2_warp_sum(ms);
    __syncwarp();
    for (int p = lane; p + KMER <= len; p += 32) wtab[c2_kmer_at(row, p) >> 1] = 0;
    __syncwarp();
    if (lane == 0) seed_ms[c] = (uint16_t)ms;
  }
}

// Same shared-memory layout and per-pair arithmetic as k_classify (dd_kernels.cu); loop mode only.
__global__ void __launch_bounds__(256) k_classify2(ClassifyArgs a, PivotArgs pv) {
  extern __shared__ uint32_t smem[];
  const int nwarps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = c2_lane();
  const int SW = a.in.SW;
  uint32_t *cen_cnt = smem;
  uint16_t *cen_kord = (uint16_t *)(smem + 512);
  uint32_t *wtab = smem + 512 + a.kord_words + wid * 512;
  uint32_t *wseq = smem + 512 + a.kord_words + nwarps * 512 + wid * SW;
  const AlnParams &P = a.P;
  const uint32_t c = a.centre_idx;
  const int len1 = a.in.len[c];
  __shared__ uint32_t cen_bits_s[32];
  uint32_t cen_bits = 0;
  {
    if (threadIdx.x < 32) cen_bits_s[threadIdx.x] = 0;
    for (int x = threadIdx.x; x < 512; x += blockDim.x) cen_cnt[x] = 0;
    for (int x = threadIdx.x; x < nwarps * 512; x += blockDim.x) smem[512 + a.kord_words + x] = 0;
    __syncthreads();
    const uint32_t *crow = a.in.seq2 + (size_t)c * SW;
    for (int p = threadIdx.x; p + KMER <= len1; p += blockDim.x) {
      const unsigned km = c2_kmer_at(crow, p);
      cen_kord[p] = (uint16_t)km;
      atomicAdd(&cen_cnt[km >> 1], 1u << (16 * (km & 1)));
      atomicOr(&cen_bits_s[km >> 5], 1u << (km & 31));
    }
    __syncthreads();
    cen_bits = cen_bits_s[lane];
  }
  const int gw = blockIdx.x * nwarps + wid, tw = gridDim.x * nwarps;
  const int sw = max(a.shard_world, 1), sr = a.shard_rank;
  const int rend = (a.in.nraw - sr + sw - 1) / sw;
  const int n_s = len1 - KMER + 1;                       // number of 5-mers of the seed
  uint32_t st_nw = 0, st_gl = 0;
  int n_nw = 0, n_gl = 0, c_align = 0, c_shroud = 0, c_pivot = 0;     // c_* are per lane here (summed at the end)
  auto flush = [&](uint32_t *list, unsigned long long *counter, uint32_t &stage, int &n) {
    if (n == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(counter, (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < n) list[base + lane] = stage;
    n = 0;
  };
  for (int it0 = gw * 32; it0 < rend; it0 += tw * 32) {
    // ---- 32 raws at a time: greedy skips and the pivot bound, one raw per lane ----
    const int it = it0 + lane;
    uint32_t my_r = 0;
    bool need = it < rend;
    if (need) {
      my_r = (uint32_t)(it * sw + sr);
      if (a.greedy && (a.in.reads[my_r] > a.centre_reads || a.lock[my_r])) need = false;      // cluster.cpp:127-131
    }
    if (need && pv.cluster_i > 0) {
      const uint32_t pc = pv.pv_cluster[my_r];
      if (pc != 0xFFFFFFFFu) {
        const int len2 = a.in.len[my_r], n_r = len2 - KMER + 1, n_p = (int)a.in.len[pv.cl_center[pc]] - KMER + 1;
        const int l1_rp = n_r + n_p - 2 * (int)pv.pv_ms[my_r], l1_ps = n_p + n_s - 2 * (int)pv.seed_ms[pc];
        const int gap = l1_rp > l1_ps ? l1_rp - l1_ps : l1_ps - l1_rp;
        const int ub = (n_r + n_s - gap) >> 1;                                               // >= the exact min-sum (>= 0)
        const double denom = (double)(min(len1, len2) - KMER) + 1.;
        if (1. - ((double)ub) / denom > P.kdist_cutoff) { c_align++; c_shroud++; c_pivot++; need = false; }
      }
    }
    unsigned todo = __ballot_sync(0xffffffffu, need);
    // ---- the survivors, warp-cooperatively (same arithmetic as k_classify) ----
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      const uint32_t r = __shfl_sync(0xffffffffu, my_r, src);
      const int len2 = a.in.len[r];
      const uint32_t *rrow = a.in.seq2 + (size_t)r * SW;
      for (int x = lane; x < SW; x += 32) wseq[x] = rrow[x];
      __syncwarp();
      const int minlen = min(len1, len2), nko = minlen - KMER + 1;
      const double denom = (double)(minlen - KMER) + 1.;
      int U = 0;
      for (int p0 = 0; p0 + KMER <= len2; p0 += 32) {
        const int p = p0 + lane;
        const bool ok = p + KMER <= len2;
        const unsigned km = ok ? c2_kmer_at(wseq, p) : 0u;
        const uint32_t w = __shfl_sync(0xffffffffu, cen_bits, km >> 5);
        U += ok ? (int)((w >> (km & 31)) & 1u) : 0;
      }
      U = c2_warp_sum(U);
      if (1. - ((double)(U & 0xFFFF)) / denom > P.kdist_cutoff) {
        if (lane == 0) { c_align++; c_shroud++; }
        __syncwarp();
        continue;
      }
      int ms = 0, om = 0;
      for (int p = lane; p + KMER <= len2; p += 32) {
        const unsigned km = c2_kmer_at(wseq, p), sh = 16 * (km & 1);
        const unsigned old = atomicAdd(&wtab[km >> 1], 1u << sh);
        ms += ((old >> sh) & 0xFFFFu) < ((cen_cnt[km >> 1] >> sh) & 0xFFFFu);     // kmers.cpp:13-26
        if (p < nko) om += (km == cen_kord[p]);                                    // kmers.cpp:102-116
      }
      ms = c2_warp_sum(ms);
      om = c2_warp_sum(om);
      __syncwarp();
      for (int p = lane; p + KMER <= len2; p += 32) wtab[c2_kmer_at(wseq, p) >> 1] = 0;
      __syncwarp();
      if (lane == 0) {                                   // keep the closest centre seen so far as the pivot
        const uint32_t pc = pv.pv_cluster[r];
        if (pc == 0xFFFFFFFFu || (unsigned)ms > (unsigned)pv.pv_ms[r]) { pv.pv_cluster[r] = pv.cluster_i; pv.pv_ms[r] = (uint16_t)ms; }
      }
      const double kdist = 1. - ((double)(ms & 0xFFFF)) / denom;
      const bool ko_valid = P.gapless && !(P.sse == 0 && len1 != len2);
      const double kodist = ko_valid ? 1. - ((double)(om & 0xFFFF)) / denom : -1.0;
      int kind;
      if (kdist > P.kdist_cutoff) kind = KIND_SHROUD;
      else if (P.band == 0 || (P.gapless && kodist == kdist)) kind = KIND_GAPLESS;
      else kind = KIND_NW;
      if (lane == 0) { c_align++; if (kind == KIND_SHROUD) c_shroud++; }
      if (kind == KIND_GAPLESS) { if (lane == n_gl) st_gl = r; if (++n_gl == 32) flush(a.gl_list, &a.ctr[CTR_GL], st_gl, n_gl); }
      else if (kind == KIND_NW) { if (lane == n_nw) st_nw = r; if (++n_nw == 32) flush(a.nw_list, &a.ctr[CTR_NW], st_nw, n_nw); }
    }
  }
  flush(a.gl_list, &a.ctr[CTR_GL], st_gl, n_gl);
  flush(a.nw_list, &a.ctr[CTR_NW], st_nw, n_nw);
  c_align = c2_warp_sum(c_align);
  c_shroud = c2_warp_sum(c_shroud);
  c_pivot = c2_warp_sum(c_pivot);
  if (lane == 0 && c_pivot) atomicAdd(&a.ctr[CTR_GLTOT], (unsigned long long)c_pivot);      // diagnostic: pairs settled by the pivot bound
  if (lane == 0 && c_align) { atomicAdd(&a.ctr[CTR_ALIGN], (unsigned long long)c_align); if (c_shroud) atomicAdd(&a.ctr[CTR_SHROUD], (unsigned long long)c_shroud); }
}

void launch_seed_dists(const DevIn &in, const uint32_t *cl_center, int nclust, uint32_t seed, uint16_t *seed_ms, int num_sms, cudaStream_t s) {
  if (nclust <= 0) return;
  count_launch(1);
  const int g = std::max(1, std::min((nclust + 7) / 8, num_sms));
  k_seed_dists<<<g, 256, (size_t)(512 + 8 * 512) * 4, s>>>(in, cl_center, nclust, seed, seed_ms);
}
void launch_classify2(const ClassifyArgs &a, const PivotArgs &pv, int grid, int block, size_t smem, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_classify2, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr_set = true; }
  count_launch(1);
  k_classify2<<<grid, block, smem, s>>>(a, pv);
}

}  // namespace dd2
