// Per-round control kernels of the divisive loop (product code, sm_100a): everything between two
// bud decisions runs on the device without host round trips.
//
//   k_round_begin     b_bud's container updates for the new cluster      cluster.cpp:315-330, containers.cpp:150-197
//   k_shuffle_*       b_shuffle2                                        cluster.cpp:210-266
//   k_p_update        b_p_update / get_pA / calc_pA + greedy locks      pval.cpp:14-89
//   k_bud_*           b_bud's arg-min scan                              cluster.cpp:274-308
//   k_report          snapshot for the host (mapped pinned memory)
//
// The host only replays the (rare) membership moves on its slot-ordered member arrays -- which the
// reference's tie-breaks depend on -- and picks the bud winner among exact ties.
#include "dd_common.h"
#include "dd_kernels.h"
#include "ppois.cuh"

namespace dd2 {

void count_launch(int n);

// ---- new cluster + per-round counter reset (one thread) --------------------------------
__global__ void k_round_begin(DevState st, int apply, uint32_t r, uint32_t from, uint32_t newi, uint32_t reads_r) {
  if (threadIdx.x == 0) {
    if (apply) {   // bi_pop_raw(from) ; b_add_bi ; bi_add_raw ; bi_assign_center
      st.cluster_of[r] = newi; st.is_center[r] = 1; st.slot0[r] = 1; st.lock[r] = 0;
      st.cl_center[newi] = r;
      st.cl_reads[from] -= reads_r; st.cl_reads_next[from] = st.cl_reads[from];
      st.cl_reads[newi] = reads_r; st.cl_reads_next[newi] = reads_r;
      st.cl_update_e[from] = 1; st.cl_update_e[newi] = 1; st.cl_check_locks[newi] = 1;
    }
    st.ctr[CTR_NW] = 0; st.ctr[CTR_GL] = 0; st.ctr[CTR_FB] = 0; st.ctr[CTR_NMOVE] = 0;
    st.ctr[CTR_CAND] = 0; st.ctr[CTR_OLD] = 0; st.ctr[CTR_SURV] = 0; st.ctr[CTR_UNEQ_B] = 0; st.ctr[CTR_UNEQ_X] = 0;   // (no memsets in the round)
  }
  if (threadIdx.x < MAX_PASS + 2) st.pinfo[threadIdx.x] = 0;
}

// ---- b_shuffle2 ---------------------------------------------------------------------------
//   max : emax[raw] = max over stored comparisons of lambda * reads(cluster)   (order-preserving
//         u64 image of the non-negative double; emax is kept zeroed between passes)
//   arg : best[raw] = lowest entry id attaining emax (entries are appended cluster by cluster, so
//         lowest id == lowest cluster index == the reference's strict '>' scan, cluster.cpp:229-239)
//   move: raws whose best cluster differs from the current one move; centres stay (:248-251).
//         Cluster reads are updated in cl_reads_next so that pass-wide e values use frozen reads.
__device__ __forceinline__ bool pass_skipped(const DevState &st, int pass) {
  return pass > 0 && st.pinfo[pass] == st.pinfo[pass - 1];      // previous pass moved nothing
}
__global__ void k_shuffle_max(DevState st, int pass) {
  if (pass_skipped(st, pass)) return;
  const unsigned long long n = st.ctr[CTR_CS_COUNT];
  for (unsigned long long x = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; x < n;
       x += (unsigned long long)gridDim.x * blockDim.x) {
    const double e = st.cs_lambda[x] * (double)st.cl_reads[st.cs_i[x]];
    atomicMax(&st.emax_bits[st.cs_index[x]], (unsigned long long)__double_as_longlong(e));
  }
}
__global__ void k_shuffle_arg(DevState st, int pass) {
  if (pass_skipped(st, pass)) return;
  const unsigned long long n = st.ctr[CTR_CS_COUNT];
  for (unsigned long long x = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; x < n;
       x += (unsigned long long)gridDim.x * blockDim.x) {
    const double e = st.cs_lambda[x] * (double)st.cl_reads[st.cs_i[x]];
    const uint32_t r = st.cs_index[x];
    if ((unsigned long long)__double_as_longlong(e) == st.emax_bits[r]) atomicMin(&st.best_entry[r], (uint32_t)x);
  }
}
__global__ void k_shuffle_move(DevState st, DevIn in, int pass) {
  if (pass_skipped(st, pass)) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= in.nraw) return;
  const uint32_t be = st.best_entry[r];
  st.emax_bits[r] = 0ull; st.best_entry[r] = 0xFFFFFFFFu;        // leave the scratch clean for the next pass
  if (be == 0xFFFFFFFFu) return;
  const uint32_t to = st.cs_i[be], from = st.cluster_of[r];
  if (to != from && !st.is_center[r]) {
    const unsigned long long s = atomicAdd(&st.ctr[CTR_NMOVE], 1ull);
    if (s < st.move_cap) { st.moves[2 * s] = (uint32_t)r; st.moves[2 * s + 1] = to; }
    st.cluster_of[r] = to;
    st.comp_lambda[r] = st.cs_lambda[be];                         // raw->comp = *compmax (:257)
    st.comp_ham[r] = st.cs_ham[be];
    const uint32_t rd = in.reads[r];
    atomicSub(&st.cl_reads_next[from], rd);
    atomicAdd(&st.cl_reads_next[to], rd);
    st.cl_update_e[from] = 1; st.cl_update_e[to] = 1;
  }
}
__global__ void k_shuffle_commit(DevState st, int nclust, int pass) {
  for (int x = threadIdx.x; x < nclust; x += blockDim.x) st.cl_reads[x] = st.cl_reads_next[x];
  if (threadIdx.x == 0) st.pinfo[pass + 1] = (uint32_t)st.ctr[CTR_NMOVE];
}

// ---- b_p_update ---------------------------------------------------------------------------
__device__ __forceinline__ bool converged_after(const DevState &st, int last_pass) {
  return st.pinfo[last_pass + 1] == st.pinfo[last_pass];
}
__global__ void k_p_update(DevState st, DevIn in, int greedy, int detect_singletons, int last_pass) {
  if (last_pass >= 0 && !converged_after(st, last_pass)) return;   // more shuffling needed: host takes over
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r == 0) {                                                   // initialise the bud scan
    st.ctr[CTR_PMIN] = ~0ull; st.ctr[CTR_RMAX] = 0ull; st.ctr[CTR_NTIE] = 0ull;
    st.ctr[CTR_PMIN_PR] = ~0ull; st.ctr[CTR_RMAX_PR] = 0ull; st.ctr[CTR_NTIE_PR] = 0ull;
  }
  if (r >= in.nraw) return;
  const uint32_t ci = st.cluster_of[r];
  const uint32_t reads = in.reads[r];
  const double lambda = st.comp_lambda[r];
  if (st.cl_update_e[ci]) {                                        // get_pA pval.cpp:67-89
    const bool prior = in.prior[r] != 0;
    double pval;
    if (reads == 1 && !prior && !detect_singletons) pval = 1.;
    else if (st.comp_ham[r] == 0) pval = 1.;
    else if (lambda == 0) pval = 0.;
    else pval = calc_pA((int)reads, lambda * (double)st.cl_reads[ci], prior || detect_singletons);
    st.p[r] = pval;
  }
  if (greedy && st.cl_check_locks[ci]) {                           // pval.cpp:29-38
    const uint32_t cen = st.cl_center[ci];
    const double E_reads_center = (double)in.reads[cen] * lambda;
    if (E_reads_center > (double)reads) st.lock[r] = 1;
    if ((uint32_t)r == cen) st.lock[r] = 1;
  }
}

// ---- b_bud scan: lexicographic minimum of (p asc, reads desc) over eligible raws; every raw
// attaining it is reported so the host can apply the (cluster, slot) scan-order tie-break.
__device__ __forceinline__ bool bud_eligible(const DevState &st, const DevIn &in, int r, const BudParams &bp) {
  if (st.slot0[r]) return false;                                   // r starts at 1 (cluster.cpp:285)
  const uint32_t reads = in.reads[r];
  if ((int)reads < bp.min_abund) return false;
  if ((int)st.comp_ham[r] < bp.min_hamming) return false;
  if (!(bp.min_fold <= 1 || ((double)reads) >= bp.min_fold * st.comp_lambda[r] * (double)st.cl_reads[st.cluster_of[r]])) return false;
  return true;
}
__global__ void k_bud_pmin(DevState st, DevIn in, BudParams bp, int nclust, int last_pass) {
  if (last_pass >= 0 && !converged_after(st, last_pass)) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < nclust) { st.cl_update_e[r] = 0; st.cl_check_locks[r] = 0; }   // flags consumed by k_p_update
  unsigned long long pb = ~0ull, pbp = ~0ull;
  if (r < in.nraw && bud_eligible(st, in, r, bp)) {
    pb = (unsigned long long)__double_as_longlong(st.p[r]);
    if (in.prior[r]) pbp = pb;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, pb, o); pb = t < pb ? t : pb;
    t = __shfl_xor_sync(0xffffffffu, pbp, o); pbp = t < pbp ? t : pbp;
  }
  if ((threadIdx.x & 31) == 0) {
    if (pb != ~0ull) atomicMin(&st.ctr[CTR_PMIN], pb);
    if (pbp != ~0ull) atomicMin(&st.ctr[CTR_PMIN_PR], pbp);
  }
}
__global__ void k_bud_rmax(DevState st, DevIn in, BudParams bp, int last_pass) {
  if (last_pass >= 0 && !converged_after(st, last_pass)) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= in.nraw) return;
  const unsigned long long pb = (unsigned long long)__double_as_longlong(st.p[r]);
  const bool a = pb == st.ctr[CTR_PMIN], b = in.prior[r] && pb == st.ctr[CTR_PMIN_PR];
  if (!(a || b) || !bud_eligible(st, in, r, bp)) return;
  if (a) atomicMax(&st.ctr[CTR_RMAX], (unsigned long long)in.reads[r]);
  if (b) atomicMax(&st.ctr[CTR_RMAX_PR], (unsigned long long)in.reads[r]);
}
__global__ void k_bud_collect(DevState st, DevIn in, BudParams bp, int last_pass, uint32_t *big_ties, uint32_t *big_ties_pr,
                              unsigned big_cap) {
  if (last_pass >= 0 && !converged_after(st, last_pass)) return;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= in.nraw) return;
  const unsigned long long pb = (unsigned long long)__double_as_longlong(st.p[r]);
  const unsigned long long rd = in.reads[r];
  const bool a = pb == st.ctr[CTR_PMIN] && rd == st.ctr[CTR_RMAX];
  const bool b = in.prior[r] && pb == st.ctr[CTR_PMIN_PR] && rd == st.ctr[CTR_RMAX_PR];
  if (!(a || b) || !bud_eligible(st, in, r, bp)) return;
  if (a) {
    const unsigned long long s = atomicAdd(&st.ctr[CTR_NTIE], 1ull);
    if (s < TIE_MAX) { st.report->tie_r[s] = (uint32_t)r; st.report->tie_lam[s] = st.comp_lambda[r]; st.report->tie_ham[s] = st.comp_ham[r]; }
    if (big_ties && s < big_cap) big_ties[s] = (uint32_t)r;
  }
  if (b) {
    const unsigned long long s = atomicAdd(&st.ctr[CTR_NTIE_PR], 1ull);
    if (s < TIE_MAX) { st.report->tiep_r[s] = (uint32_t)r; st.report->tiep_lam[s] = st.comp_lambda[r]; st.report->tiep_ham[s] = st.comp_ham[r]; }
    if (big_ties_pr && s < big_cap) big_ties_pr[s] = (uint32_t)r;
  }
}
__global__ void k_report(DevState st, int last_pass) {
  if (threadIdx.x < CTR_N) st.report->ctr[threadIdx.x] = st.ctr[threadIdx.x];
  if (threadIdx.x < MAX_PASS + 2) st.report->pinfo[threadIdx.x] = st.pinfo[threadIdx.x];
  if (threadIdx.x == 0) st.report->converged = (last_pass < 0 || converged_after(st, last_pass)) ? 1u : 0u;
}

// Two-phase loop NW: per-raw bound factors.  S_r = product over the raw's positions of its self-transition factor
// err[5*nt][q]; rho_r = max over positions and nt0 != nt of err[4*nt0+nt][q] / err[5*nt][q].
__global__ void k_raw_bounds(DevIn in, const double *err, int ncol, int use_quals, double *S, double *rho, int rank, int world) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) * world + rank;        // this rank's raws only (their quality rows may be the only ones uploaded)
  if (r >= in.nraw) return;
  const uint32_t *row = in.seq2 + (size_t)r * in.SW;
  const uint8_t *q = in.qual + (size_t)r * in.QS;
  double s = 1.0, rh = 0.0;
  const int L = in.len[r];
  for (int p = 0; p < L; p++) {
    const int b = (row[p >> 4] >> (2 * (p & 15))) & 3;
    const int qq = use_quals ? min((int)q[p], ncol - 1) : 0;
    const double self = err[(5 * b) * ncol + qq];
    s *= self;
    for (int a0 = 0; a0 < 4; a0++) if (a0 != b) rh = fmax(rh, err[(4 * a0 + b) * ncol + qq] / self);
  }
  S[r] = s; rho[r] = rh;
}

__global__ void k_fill_f64(double *p, double v, size_t n) {
  for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < n; x += (size_t)gridDim.x * blockDim.x) p[x] = v;
}
__global__ void k_center_cluster(int *cc, const uint32_t *cl_center, int nclust) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nclust) cc[cl_center[i]] = i;
}

// Quality rows of a list of raws <-> a dense buffer (16-byte units).  Sharded uploads keep a raw's quality row on its owner only
// (dd_driver.cu:do_upload); the rows of the cluster centres are exchanged before the birth subs need them.
__global__ void k_qrows_gather(const uint8_t *qual, int QS, const uint32_t *rows, int nrows, int only_rank, int world, uint8_t *dense) {
  const int u = QS / 16, x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nrows * u) return;
  const uint32_t r = rows[x / u];
  uint4 v = make_uint4(0, 0, 0, 0);
  if (only_rank < 0 || (int)(r % (uint32_t)world) == only_rank) v = ((const uint4 *)(qual + (size_t)r * QS))[x % u];
  ((uint4 *)dense)[x] = v;
}
__global__ void k_qrows_scatter(uint8_t *qual, int QS, const uint32_t *rows, int nrows, int rank, int world, const uint8_t *dense) {
  const int u = QS / 16, x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nrows * u) return;
  const uint32_t r = rows ? rows[x / u] : (uint32_t)(x / u) * (uint32_t)world + (uint32_t)rank;     // rows == NULL: the rank's own raws in order
  ((uint4 *)(qual + (size_t)r * QS))[x % u] = ((const uint4 *)dense)[x];
}

// ------------------------------- launch wrappers --------------------------------------
void launch_qrows_gather(const uint8_t *qual, int QS, const uint32_t *rows, int nrows, int only_rank, int world, uint8_t *dense, cudaStream_t s) {
  if (nrows <= 0) return;
  count_launch(1);
  k_qrows_gather<<<(nrows * (QS / 16) + 255) / 256, 256, 0, s>>>(qual, QS, rows, nrows, only_rank, world, dense);
}
void launch_qrows_scatter(uint8_t *qual, int QS, const uint32_t *rows, int nrows, int rank, int world, const uint8_t *dense, cudaStream_t s) {
  if (nrows <= 0) return;
  count_launch(1);
  k_qrows_scatter<<<(nrows * (QS / 16) + 255) / 256, 256, 0, s>>>(qual, QS, rows, nrows, rank, world, dense);
}
void launch_raw_bounds(const DevIn &in, const double *err_rowmajor, int ncol, int use_quals, double *S, double *rho, int rank, int world, cudaStream_t s) {
  count_launch(1);
  const int nown = (in.nraw - rank + world - 1) / world;
  k_raw_bounds<<<(nown + 127) / 128, 128, 0, s>>>(in, err_rowmajor, ncol, use_quals, S, rho, rank, world);
}
void launch_fill_f64(double *p, double v, size_t n, cudaStream_t s) {
  count_launch(1);
  k_fill_f64<<<(unsigned)std::min<size_t>((n + 255) / 256, 2048), 256, 0, s>>>(p, v, n);
}
void launch_center_cluster(int *cc, const uint32_t *cl_center, int nclust, cudaStream_t s) {
  count_launch(1);
  k_center_cluster<<<(nclust + 127) / 128, 128, 0, s>>>(cc, cl_center, nclust);
}
void launch_round_begin(const DevState &st, int apply, uint32_t r, uint32_t from, uint32_t newi, uint32_t reads_r, cudaStream_t s) {
  count_launch(1);
  k_round_begin<<<1, 32, 0, s>>>(st, apply, r, from, newi, reads_r);
}
void launch_shuffle_pass(const DevState &st, const DevIn &in, unsigned long long n_entries_upper, int nclust, int pass, cudaStream_t s) {
  const int B = 256;
  const unsigned g = (unsigned)std::min<unsigned long long>((n_entries_upper + B - 1) / B, 148ull * 16);
  count_launch(4);
  k_shuffle_max<<<g, B, 0, s>>>(st, pass);
  k_shuffle_arg<<<g, B, 0, s>>>(st, pass);
  k_shuffle_move<<<(in.nraw + B - 1) / B, B, 0, s>>>(st, in, pass);
  k_shuffle_commit<<<1, 256, 0, s>>>(st, nclust, pass);
}
void launch_p_update(const DevState &st, const DevIn &in, int greedy, int detect_singletons, int last_pass, cudaStream_t s) {
  count_launch(1);
  k_p_update<<<(in.nraw + 127) / 128, 128, 0, s>>>(st, in, greedy, detect_singletons, last_pass);
}
void launch_bud_scan(const DevState &st, const DevIn &in, const BudParams &bp, int nclust, int last_pass, cudaStream_t s) {
  const int B = 256, G = (std::max(in.nraw, nclust) + B - 1) / B;
  count_launch(3);
  k_bud_pmin<<<G, B, 0, s>>>(st, in, bp, nclust, last_pass);
  k_bud_rmax<<<G, B, 0, s>>>(st, in, bp, last_pass);
  k_bud_collect<<<G, B, 0, s>>>(st, in, bp, last_pass, nullptr, nullptr, 0);
}
void launch_bud_collect_big(const DevState &st, const DevIn &in, const BudParams &bp, uint32_t *ties, uint32_t *ties_pr, unsigned cap,
                            cudaStream_t s) {
  const int B = 256, G = (in.nraw + B - 1) / B;
  count_launch(1);
  k_bud_collect<<<G, B, 0, s>>>(st, in, bp, -1, ties, ties_pr, cap);
}
void launch_report(const DevState &st, int last_pass, cudaStream_t s) {
  count_launch(1);
  k_report<<<1, 64, 0, s>>>(st, last_pass);
}

}  // namespace dd2
