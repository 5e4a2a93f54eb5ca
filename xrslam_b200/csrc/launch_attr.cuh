// Opt-in to > 48 KB of dynamic shared memory.  The attribute belongs to the (kernel, device) pair, not to the process:
// a context created on a second GPU of the same process needs its own opt-in, and lanes launch from several host threads.
#pragma once
#include <atomic>
#include <cuda_runtime.h>

namespace xv {
template <class K> inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, std::atomic<unsigned long long> &done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done_mask.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);      // idempotent: a race only repeats it
    if (e == cudaSuccess) done_mask.fetch_or(bit, std::memory_order_release);
    return e;
}
}  // namespace xv
