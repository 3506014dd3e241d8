// Where do the 0.3-1 s step outliers on the GPU boxes come from?  (profiles/r2_host_stalls.md)
// Four observers run side by side for SECS seconds and log every gap above a threshold with its wall-clock time:
//   cpu    a host thread that only reads the clock                      -> the process is descheduled / throttled
//   gpu    one device thread that only reads %globaltimer               -> the GPU context is not running
//   poll   launch of a tiny kernel that writes a mapped host flag, host spins on the flag (no driver call in the wait)
//   sync   launch of a tiny kernel + cudaStreamSynchronize
// build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/stall tools/ubench/stall.cu -lpthread
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Gap { double at, gap_ms; };

__global__ void k_watch(unsigned long long dur_ns, unsigned long long thr_ns, unsigned long long *out, int cap, int *n) {
  unsigned long long t0, t, prev;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  prev = t0;
  int k = 0;
  for (;;) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - prev > thr_ns && k < cap) { out[2 * k] = prev - t0; out[2 * k + 1] = t - prev; k++; }
    prev = t;
    if (t - t0 > dur_ns) break;
  }
  *n = k;
}
__global__ void k_flag(volatile unsigned *flag, unsigned v) { *flag = v; }
__global__ void k_nop() {}

static std::string slurp(const char *p) { std::ifstream f(p); std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>()); return s; }

int main(int argc, char **argv) {
  const double secs = argc > 1 ? atof(argv[1]) : 8.0;
  cudaSetDevice(0);
  cudaFree(0);
  printf("cpu.max: %s", slurp("/sys/fs/cgroup/cpu.max").c_str());
  printf("cpu.stat before:\n%s", slurp("/sys/fs/cgroup/cpu.stat").c_str());
  printf("hardware threads: %u\n", std::thread::hardware_concurrency());
  unsigned long long *d_out; int *d_n;
  cudaMalloc(&d_out, 2 * 4096 * 8); cudaMalloc(&d_n, 4); cudaMemset(d_n, 0, 4);
  cudaStream_t sw, sp, ss;
  cudaStreamCreateWithFlags(&sw, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&sp, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&ss, cudaStreamNonBlocking);
  unsigned *h_flag, *d_flag;
  cudaHostAlloc(&h_flag, 64, cudaHostAllocMapped); *h_flag = 0; cudaHostGetDevicePointer(&d_flag, h_flag, 0);
  k_nop<<<1, 1, 0, ss>>>(); k_flag<<<1, 1, 0, sp>>>(d_flag, 0); cudaDeviceSynchronize();
  const double T0 = now_s();
  std::atomic<bool> stop(false);
  std::vector<Gap> g_cpu, g_poll, g_sync;
  unsigned long long n_poll = 0, n_sync = 0; double sum_poll = 0, sum_sync = 0;
  k_watch<<<1, 1, 0, sw>>>((unsigned long long)(secs * 1e9), 200000ull, d_out, 4096, d_n);
  std::thread tc([&] { double p = now_s(); while (!stop) { double t = now_s(); if (t - p > 1e-3) g_cpu.push_back({p - T0, (t - p) * 1e3}); p = t; } });
  std::thread tp([&] {
    unsigned v = 0;
    while (!stop) {
      const double a = now_s(); v++;
      k_flag<<<1, 1, 0, sp>>>(d_flag, v);
      const double b = now_s();
      while (*(volatile unsigned *)h_flag != v) {}
      const double c = now_s();
      n_poll++; sum_poll += c - a;
      if (c - a > 1e-3) { g_poll.push_back({a - T0, (c - a) * 1e3}); g_poll.push_back({-1.0, (b - a) * 1e3}); }
    }
  });
  std::thread tsy([&] {
    while (!stop) {
      const double a = now_s();
      k_nop<<<1, 1, 0, ss>>>();
      const double b = now_s();
      cudaStreamSynchronize(ss);
      const double c = now_s();
      n_sync++; sum_sync += c - a;
      if (c - a > 1e-3) { g_sync.push_back({a - T0, (c - a) * 1e3}); g_sync.push_back({-1.0, (b - a) * 1e3}); }
    }
  });
  std::this_thread::sleep_for(std::chrono::duration<double>(secs));
  stop = true; tc.join(); tp.join(); tsy.join();
  cudaDeviceSynchronize();
  int n = 0; cudaMemcpy(&n, d_n, 4, cudaMemcpyDeviceToHost);
  std::vector<unsigned long long> o(2 * 4096); cudaMemcpy(o.data(), d_out, o.size() * 8, cudaMemcpyDeviceToHost);
  printf("cpu.stat after:\n%s", slurp("/sys/fs/cgroup/cpu.stat").c_str());
  printf("poll: %llu iterations, mean %.1f us; sync: %llu iterations, mean %.1f us\n", n_poll, sum_poll / (n_poll ? n_poll : 1) * 1e6, n_sync, sum_sync / (n_sync ? n_sync : 1) * 1e6);
  printf("cpu gaps > 1 ms: %zu\n", g_cpu.size());
  for (auto &g : g_cpu) printf("  cpu  t=%.3f s  %.1f ms\n", g.at, g.gap_ms);
  printf("gpu gaps > 0.2 ms: %d\n", n);
  for (int k = 0; k < n && k < 200; k++) printf("  gpu  t=%.3f s  %.2f ms\n", o[2 * k] * 1e-9, o[2 * k + 1] * 1e-6);
  printf("poll round trips > 1 ms: %zu\n", g_poll.size() / 2);
  for (size_t k = 0; k + 1 < g_poll.size() && k < 200; k += 2) printf("  poll t=%.3f s  %.1f ms (launch call %.1f ms)\n", g_poll[k].at, g_poll[k].gap_ms, g_poll[k + 1].gap_ms);
  printf("sync round trips > 1 ms: %zu\n", g_sync.size() / 2);
  for (size_t k = 0; k + 1 < g_sync.size() && k < 200; k += 2) printf("  sync t=%.3f s  %.1f ms (launch call %.1f ms)\n", g_sync[k].at, g_sync[k].gap_ms, g_sync[k + 1].gap_ms);
  return 0;
}
