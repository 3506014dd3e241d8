// Micro-benchmarks of the sm_100a integer instructions the NW row kernel is built from (throughput per SM per clock and
// dependent-issue latency).  Measurement tool, not product code:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu alu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { VIADDMNMX, VIMNMX, VIMNMX3, IADD, LOP, IMAD, IMADHI, SHF, CELL, CELL_MNMX3, NOPS };
const char *names[] = {"VIADDMNMX", "VIMNMX", "VIMNMX3", "IADD3", "LOP3", "IMAD", "IMAD.HI", "SHF", "cell(2xVIADDMNMX+LOP3+3xIMAD*)", "cell(VIMNMX3+LOP3+4 fma-pipe)"};

template <int OP, int CH>
__global__ void k(int *out, int iters, int a0, int b0, int c0) {
  int x[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) x[i] = threadIdx.x + i * a0;
  int b = b0, c = c0;
  unsigned mm = (unsigned)a0 * 2654435761u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      if (OP == VIADDMNMX) x[i] = __viaddmax_s32(x[i], b, c);
      if (OP == VIMNMX) x[i] = max(x[i], b) ^ 0;   // plain max (b varies below)
      if (OP == VIMNMX3) x[i] = __vimax3_s32(x[i], b, c);
      if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(x[i]) : "r"(b));
      if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(b), "r"(c));
      if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
      if (OP == IMADHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(c));
      if (OP == SHF) x[i] = __funnelshift_r(x[i], b, 2);
      if (OP == CELL) {          // one DP cell of dd_nwrow.cu: x[i] plays S[d], the chain goes through `c`
        const int bit = (int)__umulhi(mm << (i & 31), 2u);
        const int diag = x[i] + bit * b;
        const int t = __viaddmax_s32(x[(i + 1) % CH], b0, diag);
        int m = __viaddmax_s32(c, a0, t);
        m &= 0xFFFF3FFF;
        x[i] = m; c = m;
      }
      if (OP == CELL_MNMX3) {
        const int bit = (int)__umulhi(mm << (i & 31), 2u);
        int diag, up, left;
        asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(diag) : "r"(bit), "r"(b), "r"(x[i]));
        asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(up) : "r"(x[(i + 1) % CH]), "r"(1), "r"(b0));
        asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(left) : "r"(c), "r"(1), "r"(a0));
        int m = __vimax3_s32(up, left, diag);
        m &= 0xFFFF3FFF;
        x[i] = m; c = m;
      }
    }
    b += 1; mm = mm * 1664525u + 1013904223u;
  }
  int s = c;
#pragma unroll
  for (int i = 0; i < CH; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP, int CH> void run(int *d, int sms, int warps_per_sm, const char *tag) {
  const int iters = 4096;
  const int block = 128, grid = sms * (warps_per_sm * 32 / block);
  k<OP, CH><<<grid, block>>>(d, 16, 3, 5, 7);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP, CH><<<grid, block>>>(d, iters, 3, 5, 7);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double ops = (double)grid * block * iters * CH;
  const double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-34s CH=%2d warps/SM=%2d  %8.3f ms  %7.2f thread-ops/clk/SM  (%.2f warp-ops/clk/SMSP)  %s\n", names[OP], CH, warps_per_sm, ms,
         ops / cyc / sms, ops / cyc / sms / 128.0, tag);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int *d; cudaMalloc(&d, sizeof(int) * sms * 2048 * 4);
  printf("SMs %d\n", sms);
  // throughput: 8 independent chains, 32 warps per SM
  run<VIADDMNMX, 8>(d, sms, 32, "tput"); run<VIMNMX, 8>(d, sms, 32, "tput"); run<VIMNMX3, 8>(d, sms, 32, "tput");
  run<IADD, 8>(d, sms, 32, "tput"); run<LOP, 8>(d, sms, 32, "tput"); run<IMAD, 8>(d, sms, 32, "tput");
  run<IMADHI, 8>(d, sms, 32, "tput"); run<SHF, 8>(d, sms, 32, "tput");
  // dependent latency: one chain, one warp per SMSP
  run<VIADDMNMX, 1>(d, sms, 4, "latency: cycles/op = 1/(warp-ops/clk/SMSP)"); run<LOP, 1>(d, sms, 4, "latency"); run<IMAD, 1>(d, sms, 4, "latency");
  run<VIMNMX3, 1>(d, sms, 4, "latency");
  // the cell: per 'op' = one DP cell
  for (int w : {4, 8, 16, 24, 32}) {
    if (w == 4) { run<CELL, 33>(d, sms, 4, "cells"); run<CELL_MNMX3, 33>(d, sms, 4, "cells"); }
    if (w == 8) { run<CELL, 33>(d, sms, 8, "cells"); run<CELL_MNMX3, 33>(d, sms, 8, "cells"); }
    if (w == 16) { run<CELL, 33>(d, sms, 16, "cells"); run<CELL_MNMX3, 33>(d, sms, 16, "cells"); }
    if (w == 24) { run<CELL, 33>(d, sms, 24, "cells"); run<CELL_MNMX3, 33>(d, sms, 24, "cells"); }
    if (w == 32) { run<CELL, 33>(d, sms, 32, "cells"); run<CELL_MNMX3, 33>(d, sms, 32, "cells"); }
  }
  return 0;
}
