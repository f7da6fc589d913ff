// Bulk asynchronous copies (the 1-D form of the Tensor Memory Accelerator,
// cp.async.bulk, SASS UBLKCP) and the mbarriers that track them, as inline PTX
// for sm_100a.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace pb {

__device__ __forceinline__ std::uint32_t SharedAddress(void const* const p) {
  return static_cast<std::uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void MbarrierInit(std::uint64_t* const barrier,
                                             int const arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
               :
               : "r"(SharedAddress(barrier)), "r"(arrivals)
               : "memory");
}

// Makes freshly initialised barriers visible to the asynchronous proxy.
__device__ __forceinline__ void MbarrierInitFence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Orders this thread's earlier generic-proxy accesses to shared memory before
// later asynchronous-proxy ones (a bulk copy that overwrites a buffer the
// warp has just read).
__device__ __forceinline__ void AsyncProxyFence() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One arrival that also announces `bytes` of bulk-copy traffic.
__device__ __forceinline__ void MbarrierArriveExpectTx(
    std::uint64_t* const barrier, std::uint32_t const bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :
               : "r"(SharedAddress(barrier)), "r"(bytes)
               : "memory");
}

// Global -> shared bulk copy of `bytes` (a multiple of 16, both addresses
// 16-byte aligned) that completes on `barrier`.
__device__ __forceinline__ void BulkCopyGlobalToShared(
    void* const destination, void const* const source,
    std::uint32_t const bytes, std::uint64_t* const barrier) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];"
      :
      : "r"(SharedAddress(destination)), "l"(source), "r"(bytes),
        "r"(SharedAddress(barrier))
      : "memory");
}

// Blocks until the phase of `barrier` with the given parity has completed.
__device__ __forceinline__ void MbarrierWait(std::uint64_t* const barrier,
                                             std::uint32_t const parity) {
  std::uint32_t const address = SharedAddress(barrier);
  std::uint32_t done = 0;
  do {
    asm volatile(
        "{\n"
        "  .reg .pred p;\n"
        "  mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "  selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(address), "r"(parity)
        : "memory");
  } while (done == 0);
}

}  // namespace pb
