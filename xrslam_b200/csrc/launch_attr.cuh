// Opt-in to > 48 KB of dynamic shared memory.  The attribute belongs to the (kernel, device) pair, not to the process:
// a context created on a second GPU of the same process needs its own opt-in, and lanes launch from several host threads.
// A kernel whose shared-memory size depends on its arguments asks again whenever a launch needs more than any before it.
#pragma once
#include <atomic>
#include <cuda_runtime.h>

namespace xv {
struct SmemOptIn { std::atomic<int> bytes[64]; SmemOptIn() { for (auto &b : bytes) b.store(0); } };

template <class K> inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, SmemOptIn &done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::atomic<int> &have = done.bytes[dev & 63];
    if (have.load(std::memory_order_acquire) >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);      // idempotent: a race only repeats it
    if (e == cudaSuccess) {
        int cur = have.load(std::memory_order_relaxed);
        while (cur < bytes && !have.compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
    }
    return e;
}
}  // namespace xv
