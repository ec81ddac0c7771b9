// tma.cuh — mbarrier + 1-D TMA bulk-copy helpers (sm_90+/sm_100a PTX) used by the streaming
// kernels. A "stage" of shared memory is filled by ONE elected thread with cp.async.bulk; the copy
// engine signals completion on an mbarrier with the byte count, so no thread spends registers or
// issue slots on global loads and the bytes in flight do not depend on what the warps are doing.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
// Make barrier initialisation visible to the async proxy (TMA complete_tx) and to other threads.
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// Order earlier generic-proxy accesses to shared memory before later async-proxy (TMA) accesses.
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
                   smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
// Suspend-time hint of try_wait: without it a waiting warp re-issues the try every few cycles and
// takes issue slots from the warps it is waiting for (26 % of all issued instructions in the first
// profile of the filter kernel).
constexpr uint32_t kMbarSuspendNs = 2000;
// Blocks until the phase with the given parity has completed (try_wait suspends in hardware for a
// bounded time, so this loop does not burn issue slots).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_addr(bar)),
      "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
}
// L2 eviction-priority policy for data that is read exactly once.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// Global -> shared bulk copy. dst/src 16 B aligned, bytes a multiple of 16. Completion is
// reported to `bar` as `bytes` transaction bytes.
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
          "r"(smem_addr(dst)),
      "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
      : "memory");
}
// Barrier among a subset of the CTA's warps (id 1..15; id 0 is __syncthreads).
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Non-blocking arrival at a named barrier (the producer side of PTX's bar.arrive / bar.sync
// producer-consumer pattern): the waiting side sits in bar.sync, descheduled by the hardware,
// instead of polling an mbarrier.
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#endif  // __CUDACC__
