// k_prescreen: the streaming first tier of the k-mer screen (product code, sm_100a).
//
// raw_align shrouds a pair when kmer_dist > cutoff (/root/reference/src/nwalign_endsfree.cpp:44-52; kmers.cpp:13-93:
// kdist = 1 - sum_k min(c_raw[k], c_centre[k]) / (min(len) - 5 + 1)).  ~88 % of all (centre, raw) pairs of a run end there.
// With B_x the 1024-bit presence bitmap of x's 5-mers,
//     sum_k min(c_r[k], c_c[k])  <=  popc(B_r & B_c) + (n_r - popc(B_r))          n_r = len_r - 4 five-mers of the raw
// (each shared 5-mer counts once, plus every repeated occurrence in the raw), and kdist is monotone in the min-sum, so
//     1 - U / denom > cutoff   with U = that bound
// proves the pair shrouded EXACTLY as the reference would decide it, from 128 bytes per raw and 32 AND + POPC.  Pairs the
// bound cannot settle go to k_classify (dd_kernels.cu), which computes the exact integer min-sum and the gapless test.
//
// This is the DRAM-streaming kernel of the path (the screen is "the part that can approach the HBM roof"):
// per round every active raw's bitmap row (128 B) + 8 B of metadata are read once.  The rows are staged to shared
// memory by the TMA engine -- 1-D bulk copies (cp.async.bulk, dd_tma.cuh) of 128-row tiles into a 4-stage ring,
// completion tracked by one mbarrier per stage -- and consumed one raw per thread (eight rotated 16-byte shared loads,
// conflict-free), with the thread's 13 bytes of per-raw metadata requested before it waits on the tile.
//
// k_kmer_bits builds the bitmap rows of this rank's raws (row `it` <-> raw it * world + rank) once per run.
#include "dd_common.h"
#include "dd_kernels.h"
#include "dd_tma.cuh"
#include <algorithm>
#include <cstring>

namespace dd2 {

namespace {
constexpr int PS_TILE = 128;         // raws per TMA tile (16 KB): one raw per thread
constexpr int PS_STAGES = 4;         // 64 KB of tiles in flight per CTA, three CTAs per SM
constexpr int PS_BLOCK = PS_TILE;

__device__ __forceinline__ unsigned kmer10(const uint32_t *row, int p) {      // same labelling as dd_kernels.cu:kmer_at
  const uint32_t w0 = row[p >> 4], w1 = row[(p + 4) >> 4];
  return __funnelshift_r(w0, w1, 2 * (p & 15)) & 0x3FFu;
}
}  // namespace

// one warp per owned raw: bitmap row + meta word (slack << 16 | len)
__global__ void __launch_bounds__(256) k_kmer_bits(DevIn in, int rank, int world, int nown, uint32_t *kbits, uint32_t *kmeta) {
  __shared__ uint32_t s_bits[8][32];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int it = blockIdx.x * 8 + wid; it < nown; it += gridDim.x * 8) {
    const uint32_t r = (uint32_t)it * (uint32_t)world + (uint32_t)rank;
    const uint32_t *row = in.seq2 + (size_t)r * in.SW;
    const int len = in.len[r];
    s_bits[wid][lane] = 0u;
    __syncwarp();
    for (int p = lane; p + KMER <= len; p += 32) { const unsigned km = kmer10(row, p); atomicOr(&s_bits[wid][km >> 5], 1u << (km & 31)); }
    __syncwarp();
    const uint32_t w = s_bits[wid][lane];
    kbits[(size_t)it * 32 + lane] = w;
    int pc = __popc(w);
#pragma unroll
    for (int o = 16; o; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
    if (lane == 0) kmeta[it] = ((uint32_t)(len - KMER + 1 - pc) << 16) | (uint32_t)len;
    __syncwarp();
  }
}

struct PrescreenArgs {
  DevIn in;
  const uint32_t *kbits, *kmeta;
  int nown, rank, world;
  uint32_t centre_idx, centre_reads;
  int greedy;
  const uint8_t *lock;
  double kdist_cutoff;
  uint32_t *cand_list;
  unsigned long long *cand_count;
  unsigned long long *ctr;
};

__global__ void __launch_bounds__(PS_BLOCK) k_prescreen(PrescreenArgs a) {
  extern __shared__ __align__(128) uint32_t s_dyn[];         // PS_STAGES tiles of PS_TILE bitmap rows
  __align__(8) __shared__ uint64_t s_full[PS_STAGES];
  __align__(16) __shared__ uint32_t s_cen[32];
  const int tid = threadIdx.x, lane = tid & 31;
  const int ntiles = (a.nown + PS_TILE - 1) / PS_TILE;
  const int len1 = a.in.len[a.centre_idx];
  if (tid < 32) s_cen[tid] = 0u;
  if (tid == 0) {
    for (int s = 0; s < PS_STAGES; s++) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  {
    const uint32_t *crow = a.in.seq2 + (size_t)a.centre_idx * a.in.SW;
    for (int p = tid; p + KMER <= len1; p += blockDim.x) { const unsigned km = kmer10(crow, p); atomicOr(&s_cen[km >> 5], 1u << (km & 31)); }
  }
  auto issue = [&](int tile, int stage) {                     // elected thread: arm the stage's barrier, start the bulk copy
    const int rows = min(PS_TILE, a.nown - tile * PS_TILE);
    const uint32_t bytes = (uint32_t)rows * 128u;
    tma_fence_generic_before_async();                         // the stage was last read through the generic proxy
    mbar_expect_tx(&s_full[stage], bytes);
    tma_load_1d(s_dyn + (size_t)stage * PS_TILE * 32, a.kbits + (size_t)tile * PS_TILE * 32, bytes, &s_full[stage]);
  };
  if (tid == 0)
    for (int s = 0; s < PS_STAGES; s++) { const int t = blockIdx.x + s * gridDim.x; if (t < ntiles) issue(t, s); }
  __syncthreads();
  int c_align = 0, c_shroud = 0;
  int k = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
    const int stage = k % PS_STAGES;
    // one raw per thread; its metadata is requested before the wait on the tile so that both latencies overlap
    const int it = tile * PS_TILE + tid;
    const bool valid = it < a.nown;
    const uint32_t r = (uint32_t)it * (uint32_t)a.world + (uint32_t)a.rank;
    const uint32_t meta = valid ? a.kmeta[it] : 0u;
    const bool skip = !valid || (a.greedy && (a.in.reads[r] > a.centre_reads || a.lock[r]));      // cluster.cpp:127-131
    mbar_wait(&s_full[stage], (uint32_t)((k / PS_STAGES) & 1));
    const uint32_t *row = s_dyn + (size_t)stage * PS_TILE * 32 + (size_t)tid * 32;
    int pc = 0;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      const int ch = (q + tid) & 7;                           // rotate the 16-byte chunks over the lanes: conflict-free 128-bit shared loads
      const uint4 v = *(const uint4 *)(row + 4 * ch);
      const uint4 cb = *(const uint4 *)(s_cen + 4 * ch);
      pc += __popc(v.x & cb.x) + __popc(v.y & cb.y) + __popc(v.z & cb.z) + __popc(v.w & cb.w);
    }
    bool cand = false;
    if (!skip) {
      const int len2 = (int)(meta & 0xFFFFu), U = pc + (int)(meta >> 16);
      const double denom = (double)(min(len1, len2) - KMER) + 1.;
      const double kd_lb = 1. - ((double)(U & 0xFFFF)) / denom;           // kmers.cpp:24 / :91 with the bound in place of the min-sum
      if (kd_lb > a.kdist_cutoff) { c_align++; c_shroud++; }
      else cand = true;
    }
    const unsigned m = __ballot_sync(0xffffffffu, cand);
    if (m) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.cand_count, (unsigned long long)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (cand) a.cand_list[base + __popc(m & ((1u << lane) - 1u))] = r;
    }
    __syncthreads();                                          // every thread is done with this stage
    if (tid == 0) { const int t = tile + PS_STAGES * gridDim.x; if (t < ntiles) issue(t, stage); }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) { c_align += __shfl_xor_sync(0xffffffffu, c_align, o); c_shroud += __shfl_xor_sync(0xffffffffu, c_shroud, o); }
  if (lane == 0 && c_align) { atomicAdd(&a.ctr[CTR_ALIGN], (unsigned long long)c_align); atomicAdd(&a.ctr[CTR_SHROUD], (unsigned long long)c_shroud); }
}

void launch_kmer_bits(const DevIn &in, int rank, int world, int nown, uint32_t *kbits, uint32_t *kmeta, int num_sms, cudaStream_t s) {
  count_launch(1);
  const int grid = std::max(1, std::min((nown + 7) / 8, num_sms * 8));
  k_kmer_bits<<<grid, 256, 0, s>>>(in, rank, world, nown, kbits, kmeta);
}

// Streams this rank's bitmap rows against centre `c`; pairs not proven shrouded are appended to cand_list (count zeroed
// by the caller).  Proven pairs are counted into CTR_ALIGN / CTR_SHROUD exactly as k_classify would have.
void launch_prescreen(const DevIn &in, const uint32_t *kbits, const uint32_t *kmeta, int nown, int rank, int world, uint32_t centre_idx,
                      uint32_t centre_reads, int greedy, const uint8_t *lock, double kdist_cutoff, uint32_t *cand_list, unsigned long long *cand_count,
                      unsigned long long *ctr, int num_sms, cudaStream_t s) {
  PrescreenArgs a{in, kbits, kmeta, nown, rank, world, centre_idx, centre_reads, greedy, lock, kdist_cutoff, cand_list, cand_count, ctr};
  const int ntiles = (nown + PS_TILE - 1) / PS_TILE;
  const int grid = std::max(1, std::min(ntiles, num_sms * 3));
  const size_t smem = (size_t)PS_STAGES * PS_TILE * 128;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_prescreen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
  count_launch(1);
  k_prescreen<<<grid, PS_BLOCK, smem, s>>>(a);
}

}  // namespace dd2
