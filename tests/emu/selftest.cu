// Self-test of the host SIMT emulator (tests/emu): collectives, early exits, shared memory, atomics -- and a kernel with
// a missing __syncwarp() whose result must depend on CUEMU_SCHED (that is what the schedule knob is for).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void k_collectives(int *out) {
  extern __shared__ int sm[];
  __shared__ int total;
  const int t = threadIdx.x;
  if (t == 0) total = 0;
  sm[t] = t;
  __syncthreads();
  const int v = sm[(t + 1) % blockDim.x];
  const int lane = t & 31;
  int x = 1;                                               // inclusive warp scan
  for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
  const unsigned b = __ballot_sync(0xffffffffu, t & 1);
  const double d = __shfl_xor_sync(0xffffffffu, (double)t * 0.5, 3);
  if (t >= 100) return;                                    // early exit: must not block the collectives below
  const int s8 = __shfl_down_sync(0xffffffffu, t, 1, 8);
  const int bc = __shfl_sync(0xffffffffu, t, 2, 4);
  atomicAdd(&total, 1);
  __syncthreads();
  out[blockIdx.x * blockDim.x + t] = v * 1000000 + x * 10000 + (b == 0xaaaaaaaau) * 1000 + s8 + (total == 100) * 500 + (d == (t ^ 3) * 0.5) * 100000000 + bc * 0;
  out[2 * 128 + blockIdx.x * blockDim.x + t] = bc;
}

// Lane groups of 8 that diverged from each other: every group runs its own number of group-wide collectives (partial masks), as
// k_nwlane's group-wide lambda does; groups 1 and 3 skip the block entirely.
__global__ void k_group_collectives(int *out) {
  const int lane = threadIdx.x & 31, g = lane & 7, grp = lane >> 3;
  const unsigned gmask = 0xFFu << (lane & ~7);
  int acc = 0;
  if (grp == 0 || grp == 2) {
    const int rounds = grp == 0 ? 3 : 5;
    for (int k = 0; k < rounds; k++) {
      int v = lane * 10 + k;
      for (int o = 4; o; o >>= 1) v += __shfl_xor_sync(gmask, v, o);       // sum over the group
      acc += __shfl_sync(gmask, v, (lane & ~7) + ((g + 1) & 7));            // every lane reads a neighbour's copy of the sum
    }
  }
  __syncwarp();
  out[threadIdx.x] = acc;
}

// Each lane writes its slot, then reads its neighbour's WITHOUT a __syncwarp(): undefined on the GPU, order dependent here.
__global__ void k_missing_syncwarp(int *out) {
  __shared__ int slot[32];
  const int lane = threadIdx.x & 31;
  slot[lane] = -1;
  __syncwarp();
  slot[lane] = lane;
  out[lane] = slot[(lane + 1) & 31];
}

int main() {
  int *d;
  cudaMalloc(&d, 4 * 128 * 4);
  k_collectives<<<2, 128, 128 * 4>>>(d);
  int bad = 0;
  for (int b = 0; b < 2; b++)
    for (int t = 0; t < 99; t++) {                         // t == 99 reads lane 4 of the last warp, which has exited (poison)
      const int lane = t & 31, s8 = ((t % 8) + 1 >= 8) ? t : t + 1, bc = (t / 4) * 4 + 2;
      const int exp = ((t + 1) % 128) * 1000000 + (lane + 1) * 10000 + 1000 + s8 + 500 + 100000000;
      if (d[b * 128 + t] != exp) { bad++; std::printf("collectives: block %d thread %d got %d expected %d\n", b, t, d[b * 128 + t], exp); }
      if (t < 96 && d[2 * 128 + b * 128 + t] != bc) { bad++; std::printf("shfl width 4: thread %d got %d expected %d\n", t, d[2 * 128 + b * 128 + t], bc); }
    }
  int *gq;
  cudaMalloc(&gq, 64 * 4);
  k_group_collectives<<<1, 64>>>(gq);
  for (int t = 0; t < 64; t++) {
    const int lane = t & 31, grp = lane >> 3, base = lane & ~7;
    int exp = 0;
    if (grp == 0 || grp == 2) for (int k = 0; k < (grp == 0 ? 3 : 5); k++) { int sum = 0; for (int j = 0; j < 8; j++) sum += (base + j) * 10 + k; exp += sum; }
    if (gq[t] != exp) { bad++; std::printf("group collectives: thread %d got %d expected %d\n", t, gq[t], exp); }
  }
  std::printf("collectives bad=%d\n", bad);
  int *e;
  cudaMalloc(&e, 32 * 4);
  k_missing_syncwarp<<<1, 32>>>(e);
  int stale = 0;
  for (int l = 0; l < 32; l++) stale += e[l] != ((l + 1) & 31);
  std::printf("missing_syncwarp stale=%d\n", stale);
  return bad != 0;
}
