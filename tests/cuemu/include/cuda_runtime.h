// TEST INFRASTRUCTURE ONLY -- a stand-in for <cuda_runtime.h> that lets the product's .cu sources be compiled by g++
// and executed on the CPU by a small SIMT emulator (tests/cuemu/cuemu.cpp): every CUDA block runs as a set of
// cooperatively scheduled fibers, __syncthreads / warp shuffles are scheduler barriers, device memory is host memory.
// It exists so that kernel changes can be checked against the oracle on machines without a GPU (tests/test_emulated.py).
// It is never built into or loaded by xrslam_b200: the product library is compiled by nvcc only and refuses to run
// without a CUDA device.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <functional>

#define CUEMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__
#ifndef __restrict__
#define __restrict__ __restrict
#endif

// ---- vector types ----
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(4) short2 { short x, y; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline short2 make_short2(short x, short y) { return short2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }

// ---- runtime API (host memory stands in for device memory; everything is synchronous) ----
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 11, cudaErrorNotReady = 600 };
enum { cudaEventDefault = 0, cudaEventBlockingSync = 1, cudaEventDisableTiming = 2 };
typedef void *cudaStream_t;
typedef struct cuemu_event *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

cudaError_t cudaMalloc(void **p, size_t n);
cudaError_t cudaFree(void *p);
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned flags);
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaMemset(void *p, int v, size_t n);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *st, unsigned flags);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaStreamDestroy(cudaStream_t st);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceCount(int *n);
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st = nullptr);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaEventDestroy(cudaEvent_t e);
// the emulator executes every launch synchronously: an event is complete as soon as it is recorded
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *st, unsigned flags, int) { return cudaStreamCreateWithFlags(st, flags); }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyFromSymbol(void *dst, const T &sym, size_t n) { std::memcpy(dst, &sym, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMemcpyToSymbol(T &sym, const void *src, size_t n) { std::memcpy(&sym, src, n); return cudaSuccess; }

// ---- SIMT emulation ----
namespace cuemu {
struct ThreadCtx { uint3 tid, bid; dim3 bdim, gdim; int linear, lane, warp; };
ThreadCtx &cur();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
void *dyn_smem();
void sync_block();
int sync_block_reduce(int pred, int mode);       // mode 0: and, 1: or
void sync_warp(unsigned mask);
unsigned long long exchange(unsigned mask, unsigned long long v, int src_lane, bool valid_src);   // value of src_lane (own value if invalid)
unsigned ballot(unsigned mask, int pred);
long long clock();
}  // namespace cuemu

#define threadIdx (cuemu::cur().tid)
#define blockIdx (cuemu::cur().bid)
#define blockDim (cuemu::cur().bdim)
#define gridDim (cuemu::cur().gdim)
constexpr int warpSize = 32;

inline void __syncthreads() { cuemu::sync_block(); }
inline int __syncthreads_and(int p) { return cuemu::sync_block_reduce(p, 0); }
inline int __syncthreads_or(int p) { return cuemu::sync_block_reduce(p, 1); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { cuemu::sync_warp(mask); }
inline long long clock64() { return cuemu::clock(); }

namespace cuemu {
template <class T> inline unsigned long long to_bits(T v) { static_assert(sizeof(T) <= 8, "shuffle of > 8 bytes"); unsigned long long b = 0; std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(unsigned long long b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }
}
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = cuemu::cur().lane, base = lane - lane % width;
    return cuemu::from_bits<T>(cuemu::exchange(mask, cuemu::to_bits(v), base + ((src % width) + width) % width, true));
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
    const int lane = cuemu::cur().lane, src = lane - (int)d;
    return cuemu::from_bits<T>(cuemu::exchange(mask, cuemu::to_bits(v), src, src >= lane - lane % width));
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
    const int lane = cuemu::cur().lane, src = lane + (int)d;
    return cuemu::from_bits<T>(cuemu::exchange(mask, cuemu::to_bits(v), src, src < lane - lane % width + width));
}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
    const int lane = cuemu::cur().lane, src = lane ^ m;
    return cuemu::from_bits<T>(cuemu::exchange(mask, cuemu::to_bits(v), src, src < lane - lane % width + width && src >= 0));
}
inline unsigned __ballot_sync(unsigned mask, int pred) { return cuemu::ballot(mask, pred); }
inline int __any_sync(unsigned mask, int pred) { return cuemu::ballot(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return cuemu::ballot(mask, !pred) == 0; }

// atomics: fibers of one block are scheduled cooperatively on one OS thread and blocks run one after the other
template <class T, class U> inline T atomicAdd(T *p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> inline T atomicMax(T *p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T, class U> inline T atomicMin(T *p, U v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }

// ---- math / conversion intrinsics ----
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
inline int __float2int_rn(float x) { return (int)std::nearbyintf(x); }
inline int __double2int_rn(double x) { return (int)std::nearbyint(x); }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
using std::min; using std::max; using std::isfinite; using std::isnan;
inline int min(int a, unsigned b) { return (int)std::min<long long>(a, b); }
inline int min(unsigned a, int b) { return (int)std::min<long long>(a, b); }
inline int max(int a, unsigned b) { return (int)std::max<long long>(a, b); }
inline int max(unsigned a, int b) { return (int)std::max<long long>(a, b); }
inline long long min(long long a, int b) { return std::min<long long>(a, b); }
inline long long min(int a, long long b) { return std::min<long long>(a, b); }
inline size_t min(size_t a, int b) { return std::min<size_t>(a, (size_t)b); }
inline size_t min(int a, size_t b) { return std::min<size_t>((size_t)a, b); }
