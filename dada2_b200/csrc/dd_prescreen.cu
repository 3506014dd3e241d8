// k_prescreen: the streaming first tier of the k-mer screen (product code, sm_100a).
//
// raw_align shrouds a pair when kmer_dist > cutoff (/root/reference/src/nwalign_endsfree.cpp:44-52; kmers.cpp:13-93:
// kdist = 1 - sum_k min(c_raw[k], c_centre[k]) / (min(len) - 5 + 1)).  ~88 % of all (centre, raw) pairs of a run end there.
// With B_x the 1024-bit presence bitmap of x's 5-mers,
//     sum_k min(c_r[k], c_c[k])  <=  popc(B_r & B_c) + (n_r - popc(B_r))          n_r = len_r - 4 five-mers of the raw
// (each shared 5-mer counts once, plus every repeated occurrence in the raw), and kdist is monotone in the min-sum, so
//     1 - U / denom > cutoff   with U = that bound
// proves the pair shrouded EXACTLY as the reference would decide it, from 128 bytes per raw and 32 AND + POPC.
// Where the bound is inconclusive the SAME thread finishes the exact integer min-sum: only 5-mers that occur at least
// twice in the raw can contribute more than their presence bit,
//     sum_k min(c_r[k], c_c[k])  =  popc(B_r & B_c) + sum over the raw's repeated 5-mers k present in the centre of (min(c_r[k], c_c[k]) - 1),
// and a 250-nt read has about 25 of those: they are kept as a list of up to 48 entries per raw (k_kmer_bits) and looked up in a
// count table of the centre in shared memory.  So every pair leaves this kernel decided exactly (shrouded or not); the
// pairs that are not shrouded go to k_kord (one thread per pair: ordered 5-mer matches from the XOR of the packed rows ->
// gapless or NW, raw_align :53-56).  Raws whose list overflows take the warp-per-pair k_classify (dd_kernels.cu) instead.
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

__device__ __forceinline__ void warp_append_u32(bool flag, uint32_t v, uint32_t *list, unsigned long long *count) {
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (!m) return;
  const int lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(m));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (flag) list[base + __popc(m & ((1u << lane) - 1u))] = v;
}
__device__ __forceinline__ unsigned kmer10(const uint32_t *row, int p) {      // same labelling as dd_kernels.cu:kmer_at
  const uint32_t w0 = row[p >> 4], w1 = row[(p + 4) >> 4];
  return __funnelshift_r(w0, w1, 2 * (p & 15)) & 0x3FFu;
}
}  // namespace

constexpr int KREP = 48;              // repeated-5-mer list entries per raw (u16: k-mer | (count - 1) << 10, packed from the front, 0 = padding); a 250-nt read has ~25
constexpr uint32_t META_OVF = 1u << 31;   // kmeta flag: the list does not hold every repeated 5-mer of this raw (or a count above 64)

// one warp per owned raw: bitmap row, repeated-5-mer list, meta word (overflow flag | slack << 16 | len)
__global__ void __launch_bounds__(256) k_kmer_bits(DevIn in, int rank, int world, int nown, uint32_t *kbits, uint32_t *kmeta, uint16_t *krep, unsigned *n_overflow) {
  __shared__ uint32_t s_bits[8][32];
  __shared__ uint32_t s_cnt[8][512];            // 1024 u16 counters per warp
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int x = lane; x < 512; x += 32) s_cnt[wid][x] = 0u;
  __syncwarp();
  for (int it = blockIdx.x * 8 + wid; it < nown; it += gridDim.x * 8) {
    const uint32_t r = (uint32_t)it * (uint32_t)world + (uint32_t)rank;
    const uint32_t *row = in.seq2 + (size_t)r * in.SW;
    const int len = in.len[r];
    s_bits[wid][lane] = 0u;
    __syncwarp();
    for (int p = lane; p + KMER <= len; p += 32) {
      const unsigned km = kmer10(row, p);
      atomicOr(&s_bits[wid][km >> 5], 1u << (km & 31));
      atomicAdd(&s_cnt[wid][km >> 1], 1u << (16 * (km & 1)));
    }
    __syncwarp();
    const uint32_t w = s_bits[wid][lane];
    kbits[(size_t)it * 32 + lane] = w;
    int pc = __popc(w);
#pragma unroll
    for (int o = 16; o; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
    // repeated 5-mers: lane l scans k-mers 32 l .. 32 l + 31 (the bits of its bitmap word), in k-mer order
    uint16_t *lst = krep + (size_t)it * KREP;
    for (int x = lane; x < KREP; x += 32) lst[x] = 0u;                 // padding: count field 0 (a real entry has count - 1 >= 1)
    __syncwarp();
    int nrep = 0; bool ovf = false;
    for (int b = 0; b < 32; b++) {
      const unsigned km = (unsigned)lane * 32u + (unsigned)b;
      const unsigned cnt = (s_cnt[wid][km >> 1] >> (16 * (km & 1))) & 0xFFFFu;
      const bool rep = cnt >= 2;
      const unsigned m = __ballot_sync(0xffffffffu, rep);
      if (rep) {
        const int slot = nrep + __popc(m & ((1u << lane) - 1u));
        if (slot < KREP && cnt <= 64) lst[slot] = (uint16_t)(km | ((cnt - 1) << 10));
        else ovf = true;
      }
      nrep += __popc(m);
    }
    ovf = __any_sync(0xffffffffu, ovf);
    if (lane == 0) kmeta[it] = (ovf ? META_OVF : 0u) | ((uint32_t)(len - KMER + 1 - pc) << 16) | (uint32_t)len;
    if (lane == 0 && ovf) atomicAdd(n_overflow, 1u);
    __syncwarp();
    for (int p = lane; p + KMER <= len; p += 32) s_cnt[wid][kmer10(row, p) >> 1] = 0u;       // clear what was touched
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
  const uint16_t *krep;                // [nown][KREP] repeated 5-mers of every raw
  uint32_t *cand_list;                 // pairs that are NOT shrouded (exact), or whose list overflowed ...
  uint16_t *cand_ms;                   // ... with their exact min-sum (0xFFFF: unknown, overflowed list)
  unsigned long long *cand_count;
  unsigned long long *ctr;
};

__global__ void __launch_bounds__(PS_BLOCK) k_prescreen(PrescreenArgs a) {
  extern __shared__ __align__(128) uint32_t s_dyn[];         // PS_STAGES tiles of PS_TILE bitmap rows
  __align__(8) __shared__ uint64_t s_full[PS_STAGES];
  __align__(16) __shared__ uint32_t s_cen[32];
  __shared__ uint32_t s_ccnt[512];                           // the centre's 5-mer counts, 1024 x u16
  __shared__ uint8_t s_t1[1024];                             // max(count - 1, 0): what a repeated 5-mer of the raw can add beyond its presence bit
  const int tid = threadIdx.x, lane = tid & 31;
  const int ntiles = (a.nown + PS_TILE - 1) / PS_TILE;
  const int len1 = a.in.len[a.centre_idx];
  if (tid < 32) s_cen[tid] = 0u;
  for (int x = tid; x < 512; x += blockDim.x) s_ccnt[x] = 0u;
  if (tid == 0) {
    for (int s = 0; s < PS_STAGES; s++) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  {
    const uint32_t *crow = a.in.seq2 + (size_t)a.centre_idx * a.in.SW;
    for (int p = tid; p + KMER <= len1; p += blockDim.x) {
      const unsigned km = kmer10(crow, p);
      atomicOr(&s_cen[km >> 5], 1u << (km & 31));
      atomicAdd(&s_ccnt[km >> 1], 1u << (16 * (km & 1)));
    }
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
  for (int x = tid; x < 1024; x += blockDim.x) {
    const uint32_t cc = (s_ccnt[x >> 1] >> (16 * (x & 1))) & 0xFFFFu;
    s_t1[x] = (uint8_t)(cc ? min(cc - 1u, 255u) : 0u);       // list counts stop at 63: the cap never binds
  }
  __syncthreads();
  int c_align = 0, c_shroud = 0;
  int k = 0;
  // One raw per thread.  No global load of the loop is waited for in the iteration that issues it: the 13 bytes of per-raw
  // metadata are requested TWO tiles ahead, and a raw whose presence bound is inconclusive requests its whole repeated-5-mer
  // list (six independent 16-byte loads) and is only resolved in the NEXT iteration, after that iteration's tile has been
  // waited for and scanned -- a warp otherwise pays the slowest lane's chain of dependent DRAM round trips on every tile.
  auto fetch_meta = [&](int tile, uint32_t &meta, bool &skip) {
    const int it = tile * PS_TILE + tid;
    const bool valid = tile < ntiles && it < a.nown;
    const uint32_t r = (uint32_t)it * (uint32_t)a.world + (uint32_t)a.rank;
    meta = valid ? a.kmeta[it] : 0u;
    skip = !valid || (a.greedy && (a.in.reads[r] > a.centre_reads || a.lock[r]));      // cluster.cpp:127-131
  };
  auto append = [&](bool cand, uint32_t r, uint32_t msv) {
    const unsigned m = __ballot_sync(0xffffffffu, cand);
    if (m) {
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(a.cand_count, (unsigned long long)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (cand) { const unsigned long long at = base + __popc(m & ((1u << lane) - 1u)); a.cand_list[at] = r; a.cand_ms[at] = (uint16_t)msv; }
    }
  };
  // the raw of the previous tile that still waits for its exact min-sum: its list (in registers), presence count and denominator
  bool pend = false;
  uint32_t pend_r = 0;
  int pend_pc = 0;
  double pend_denom = 1.;
  uint4 L[KREP / 8];
  auto resolve = [&]() {
    bool cand = false;
    uint32_t msv = 0;
    if (pend) {
      // exact min-sum: presence bits + what the raw's repeated 5-mers add beyond their presence bit: for an entry (k, count - 1),
      // min(count, c_centre[k]) - 1 = min(count - 1, max(c_centre[k] - 1, 0)) when the 5-mer is in the centre, 0 when not -- the
      // same expression with the table holding max(c - 1, 0); padding entries (count field 0) add nothing
      int ms = pend_pc;
#pragma unroll
      for (int blk = 0; blk < KREP / 8; blk++) {
        const uint32_t lw[4] = {L[blk].x, L[blk].y, L[blk].z, L[blk].w};
#pragma unroll
        for (int e = 0; e < 8; e++) {
          const uint32_t ent = (lw[e >> 1] >> (16 * (e & 1))) & 0xFFFFu;
          ms += (int)min(ent >> 10, (uint32_t)s_t1[ent & 0x3FFu]);
        }
        if ((lw[3] >> 26) == 0u) break;                                       // the block's last entry is padding: the list ended here
      }
      const double kdist = 1. - ((double)(ms & 0xFFFF)) / pend_denom;         // exactly raw_align's kdist (N1: integer min-sum)
      if (kdist > a.kdist_cutoff) { c_align++; c_shroud++; }
      else { cand = true; msv = (uint32_t)ms; }
    }
    append(cand, pend_r, msv);
    pend = false;
  };
  uint32_t meta_n1, meta_n2; bool skip_n1, skip_n2;
  fetch_meta(blockIdx.x, meta_n1, skip_n1);
  fetch_meta(blockIdx.x + gridDim.x, meta_n2, skip_n2);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, k++) {
    const int stage = k % PS_STAGES;
    const uint32_t meta = meta_n1; const bool skip = skip_n1;
    meta_n1 = meta_n2; skip_n1 = skip_n2;
    const uint32_t r = (uint32_t)(tile * PS_TILE + tid) * (uint32_t)a.world + (uint32_t)a.rank;
    fetch_meta(tile + 2 * gridDim.x, meta_n2, skip_n2);
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
    resolve();                                                // the previous tile's inconclusive raw: its list has landed by now
    bool cand = false;
    if (!skip) {
      const int len2 = (int)(meta & 0xFFFFu), U = pc + (int)((meta >> 16) & 0x3FFFu);
      const double denom = (double)(min(len1, len2) - KMER) + 1.;
      const double kd_lb = 1. - ((double)(U & 0xFFFF)) / denom;           // kmers.cpp:24 / :91 with the bound in place of the min-sum
      if (kd_lb > a.kdist_cutoff) { c_align++; c_shroud++; }
      else if (meta & META_OVF) cand = true;                                // list overflow: the warp-per-pair screen decides
      else {
        const uint4 *lp = (const uint4 *)(a.krep + (size_t)(tile * PS_TILE + tid) * KREP);
#pragma unroll
        for (int blk = 0; blk < KREP / 8; blk++) L[blk] = lp[blk];
        pend = true; pend_r = r; pend_pc = pc; pend_denom = denom;
      }
    }
    append(cand, r, 0xFFFFu);
    __syncthreads();                                          // every thread is done with this stage
    if (tid == 0) { const int t = tile + PS_STAGES * gridDim.x; if (t < ntiles) issue(t, stage); }
  }
  resolve();
#pragma unroll
  for (int o = 16; o; o >>= 1) { c_align += __shfl_xor_sync(0xffffffffu, c_align, o); c_shroud += __shfl_xor_sync(0xffffffffu, c_shroud, o); }
  if (lane == 0 && c_align) { atomicAdd(&a.ctr[CTR_ALIGN], (unsigned long long)c_align); atomicAdd(&a.ctr[CTR_SHROUD], (unsigned long long)c_shroud); }
}

// ---- the pairs that are not shrouded: gapless or NW?  raw_align :53-56 compares kodist (ordered 5-mer matches, kmers.cpp:121-150)
// with kdist; equal denominators make it the integer test om == ms.  One thread per pair: the 5-mer at position p matches iff
// bases p .. p+4 all match, i.e. five consecutive set bits in the equality mask of the two packed rows.
struct KordArgs {
  DevIn in;
  AlnParams P;
  uint32_t centre_idx;
  const uint32_t *cand_list;
  const uint16_t *cand_ms;
  const unsigned long long *cand_count;
  uint32_t *nw_list, *gl_list, *old_list;        // old_list: min-sum unknown -> k_classify
  unsigned long long *ctr, *old_count;
};

__global__ void __launch_bounds__(128) k_kord(KordArgs a) {
  extern __shared__ uint32_t s_crow[];             // the centre's packed row + one zero word
  const unsigned long long n = *a.cand_count;
  if ((unsigned long long)blockIdx.x * blockDim.x >= n) return;
  const int SW = a.in.SW, len1 = a.in.len[a.centre_idx];
  for (int x = threadIdx.x; x <= SW; x += blockDim.x) s_crow[x] = x < SW ? a.in.seq2[(size_t)a.centre_idx * SW + x] : 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int c_align = 0;
  for (unsigned long long wb = (unsigned long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); wb < n; wb += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long k = wb + lane;
    const bool act = k < n;
    const uint32_t r = act ? a.cand_list[k] : 0u;
    const uint32_t ms = act ? a.cand_ms[k] : 0u;
    bool to_old = act && ms == 0xFFFFu, to_gl = false, to_nw = false;
    if (act && !to_old) {
      c_align++;
      const int len2 = a.in.len[r];
      bool gapless = a.P.band == 0;
      const bool ko_valid = a.P.gapless && !(a.P.sse == 0 && len1 != len2);        // scalar kord_dist: -1 for unequal lengths (kmers.cpp:107)
      if (!gapless && ko_valid) {
        const uint32_t *rrow = a.in.seq2 + (size_t)r * SW;
        const int nko = min(len1, len2) - KMER + 1;                                // ordered positions compared (kmers.cpp:121-150)
        int om = 0;
        uint32_t e_cur;
        { const uint32_t x = rrow[0] ^ s_crow[0]; e_cur = ~(x | (x >> 1)) & 0x55555555u; }
        for (int w = 0; w * 16 < nko; w++) {
          uint32_t e_nxt = 0u;
          if (w + 1 < SW) { const uint32_t x = rrow[w + 1] ^ s_crow[w + 1]; e_nxt = ~(x | (x >> 1)) & 0x55555555u; }
          uint32_t run = e_cur & __funnelshift_r(e_cur, e_nxt, 2) & __funnelshift_r(e_cur, e_nxt, 4) & __funnelshift_r(e_cur, e_nxt, 6) &
                         __funnelshift_r(e_cur, e_nxt, 8);                         // bit 2b: the 5-mers starting at base 16 w + b are equal
          const int left = nko - w * 16;                                           // positions of this word below nko
          if (left < 16) run &= (1u << (2 * left)) - 1u;
          om += __popc(run);
          e_cur = e_nxt;
        }
        gapless = (uint32_t)om == ms;                                              // kodist == kdist
      }
      to_gl = gapless; to_nw = !gapless;
    }
    warp_append_u32(to_gl, r, a.gl_list, &a.ctr[CTR_GL]);
    warp_append_u32(to_nw, r, a.nw_list, &a.ctr[CTR_NW]);
    warp_append_u32(to_old, r, a.old_list, a.old_count);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) c_align += __shfl_xor_sync(0xffffffffu, c_align, o);
  if (lane == 0 && c_align) atomicAdd(&a.ctr[CTR_ALIGN], (unsigned long long)c_align);
}

void launch_kmer_bits(const DevIn &in, int rank, int world, int nown, uint32_t *kbits, uint32_t *kmeta, uint16_t *krep, unsigned *n_overflow, int num_sms,
                      cudaStream_t s) {
  count_launch(1);
  const int grid = std::max(1, std::min((nown + 7) / 8, num_sms * 8));
  k_kmer_bits<<<grid, 256, 0, s>>>(in, rank, world, nown, kbits, kmeta, krep, n_overflow);
}

// Streams this rank's bitmap rows against centre `c`: every pair is decided exactly (shrouded or not), except raws whose
// repeated-5-mer list overflowed.  The pairs that are not shrouded land in cand_list / cand_ms (count zeroed by the
// caller); shrouded ones are counted into CTR_ALIGN / CTR_SHROUD exactly as k_classify would have.
void launch_prescreen(const DevIn &in, const uint32_t *kbits, const uint32_t *kmeta, const uint16_t *krep, int nown, int rank, int world,
                      uint32_t centre_idx, uint32_t centre_reads, int greedy, const uint8_t *lock, double kdist_cutoff, uint32_t *cand_list,
                      uint16_t *cand_ms, unsigned long long *cand_count, unsigned long long *ctr, int num_sms, cudaStream_t s) {
  PrescreenArgs a{in, kbits, kmeta, nown, rank, world, centre_idx, centre_reads, greedy, lock, kdist_cutoff, krep, cand_list, cand_ms, cand_count, ctr};
  const int ntiles = (nown + PS_TILE - 1) / PS_TILE;
  const int grid = std::max(1, std::min(ntiles, num_sms * 3));
  const size_t smem = (size_t)PS_STAGES * PS_TILE * 128;
  static bool attr_set = false;
  if (!attr_set) { cudaFuncSetAttribute(k_prescreen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
  count_launch(1);
  k_prescreen<<<grid, PS_BLOCK, smem, s>>>(a);
}

// gapless / NW decision for the candidates; raws with an unknown min-sum are forwarded to old_list (count zeroed by the caller)
void launch_kord(const DevIn &in, const AlnParams &P, uint32_t centre_idx, const uint32_t *cand_list, const uint16_t *cand_ms,
                 const unsigned long long *cand_count, uint32_t *nw_list, uint32_t *gl_list, uint32_t *old_list, unsigned long long *old_count,
                 unsigned long long *ctr, unsigned long long upper, int num_sms, cudaStream_t s) {
  KordArgs a{in, P, centre_idx, cand_list, cand_ms, cand_count, nw_list, gl_list, old_list, ctr, old_count};
  const int grid = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((upper + 127) / 128, (unsigned long long)num_sms * 16));
  count_launch(1);
  k_kord<<<grid, 128, (size_t)(in.SW + 1) * 4, s>>>(a);
}

}  // namespace dd2
