// Micro-benchmark for the back-end roofline: FP64 FMA and FP64 tensor (DMMA m8n8k4) issue rate per SM, FP32 FMA for
// scale, CTA barrier cost with 16 warps.  The round-1 section profile of solve_kernel (tools/phase_probe.py) is
// consistent with ~2 DFMA / clk / SM; this tool measures that number directly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_probe tools/fp64_probe.cu && tools/fp64_probe
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <typename T> __global__ void fma_kernel(T *out, long long *cycles, T a, T b) {
    T x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < ITER; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    __syncthreads();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void dmma_kernel(double *out, long long *cycles, double a, double b) {
    double c0[2] = {0.0, 0.0}, c1[2] = {1.0, 1.0}, c2[2] = {2.0, 2.0}, c3[2] = {3.0, 3.0};
    const double fa = a + threadIdx.x * 1e-9, fb = b;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITER; ++i) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[0]), "+d"(c0[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c1[0]), "+d"(c1[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c2[0]), "+d"(c2[1]) : "d"(fa), "d"(fb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c3[0]), "+d"(c3[1]) : "d"(fa), "d"(fb));
    }
    __syncthreads();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c0[1] + c1[0] + c1[1] + c2[0] + c2[1] + c3[0] + c3[1];
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void barrier_kernel(long long *cycles) {
    __shared__ double v[512];
    v[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < ITER; ++i) {
        v[(threadIdx.x + i) & 511] += 1.0;           // one shared-memory store in flight at each barrier, like the solver's column loop
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double mean_cycles(const long long *h, int n) { double s = 0; for (int i = 0; i < n; ++i) s += (double)h[i]; return s / n; }

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { std::printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    const int sms = p.multiProcessorCount, T = 512;
    double *out; long long *cyc, *h = new long long[sms];
    cudaMalloc(&out, sizeof(double) * sms * T); cudaMalloc(&cyc, sizeof(long long) * sms);
    auto run = [&](auto launch) { launch(); launch(); cudaDeviceSynchronize(); cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost); return mean_cycles(h, sms); };
    const double c64 = run([&] { fma_kernel<double><<<sms, T>>>(out, cyc, 1.0000001, 1e-9); });
    const double c32 = run([&] { fma_kernel<float><<<sms, T>>>((float *)out, cyc, 1.0000001f, 1e-9f); });
    const double cmm = run([&] { dmma_kernel<<<sms, T>>>(out, cyc, 1.0000001, 1e-9); });
    const double cbar = run([&] { barrier_kernel<<<sms, T>>>(cyc); });
    const double fma_per_cta = (double)T * 8 * ITER;
    const double dmma_fma_per_cta = (double)(T / 32) * 4 * ITER * (8 * 8 * 4);      // m8n8k4 = 256 FMA per warp instruction
    std::printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %d, \"dfma_per_clk_per_sm\": %.2f, \"ffma_per_clk_per_sm\": %.2f, "
                "\"dmma_fma_per_clk_per_sm\": %.2f, \"barrier_cycles_16_warps\": %.1f}\n",
                p.name, sms, p.clockRate / 1000, fma_per_cta / c64, fma_per_cta / c32, dmma_fma_per_cta / cmm, cbar / ITER);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) std::printf("cuda error: %s\n", cudaGetErrorString(e));
    return 0;
}
