// 1-D bulk asynchronous copies (TMA engine: cp.async.bulk, SASS UBLKCP) and the mbarrier objects that track them.
// Product code (sm_100a).  One elected thread arms a shared-memory mbarrier with the byte count it expects and issues the
// bulk copy global -> shared; the consumers spin on the barrier's phase parity (try_wait suspends the warp in hardware).
//
// Under the host SIMT emulator (tests/emu, -DDADA2B_EMU: test infrastructure) the copy is a memcpy by the issuing fiber and
// the wait is a block barrier: every kernel that uses these helpers waits on a stage with ALL its threads, uniformly.
#pragma once
#include <cstdint>

namespace dd2 {

#ifndef DADA2B_EMU
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_fence_generic_before_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  }
}
#else
__device__ __forceinline__ void mbar_init(uint64_t *bar, int) { *bar = 0; }
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *, uint32_t) {}
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *) { memcpy(smem_dst, gmem_src, bytes); }
__device__ __forceinline__ void tma_fence_generic_before_async() {}
__device__ __forceinline__ void mbar_wait(uint64_t *, uint32_t) { __syncthreads(); }
#endif

}  // namespace dd2
