// Stand-alone probe of the TMA tile load used by klt_kernel (diagnostic; built and run on the GPU box, one variant per process
// because a faulting variant poisons the context):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu -ldl && for v in 0 1 2 3 4 5 6 7 8; do /tmp/tma_probe $v; done
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK, bool FENCE> __device__ void load_tile(unsigned char *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z, unsigned bytes) {
    if ((threadIdx.x & 31) == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        if (FENCE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
}

template <int RANK, bool FENCE> __global__ void k_global(const CUtensorMap *map, int x, int y, int z, unsigned bytes, unsigned char *out) {
    __shared__ alignas(128) unsigned char tile[24 * 64];
    __shared__ unsigned long long bar;
    load_tile<RANK, FENCE>(tile, map, &bar, x, y, z, bytes);
    for (int i = threadIdx.x; i < (int)bytes; i += 32) out[i] = tile[i];
}
__global__ void k_param(const __grid_constant__ CUtensorMap map, int x, int y, int z, unsigned bytes, unsigned char *out) {
    __shared__ alignas(128) unsigned char tile[24 * 64];
    __shared__ unsigned long long bar;
    load_tile<3, true>(tile, &map, &bar, x, y, z, bytes);
    for (int i = threadIdx.x; i < (int)bytes; i += 32) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int W = 640, H = 480, N = 6;
    const size_t total = 408000;
    std::vector<unsigned char> h(total * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d, *out;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 24 * 64);
    void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    if (variant == 7) {
        void *lib = dlopen("libcuda.so.1", RTLD_NOW);
        fn = lib ? dlsym(lib, "cuTensorMapEncodeTiled") : nullptr;
    } else cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    if (!fn) { printf("variant %d: no entry point\n", variant); return 1; }
    int x = 17, y = 9, z = 3, rank = 3, bw = 32;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
    CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE;
    int esize = 1;
    const char *what = "3D u8 box 32x24 x=17 (as in klt_kernel)";
    switch (variant) {
    case 1: x = 16; what = "3D u8 box 32x24 x=16 (16-byte aligned start)"; break;
    case 2: rank = 2; what = "2D u8 box 32x24 x=17"; break;
    case 3: rank = 2; x = 16; what = "2D u8 box 32x24 x=16"; break;
    case 4: bw = 64; what = "3D u8 box 64x24 x=17"; break;
    case 5: l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B; what = "3D u8 box 32x24 x=17, L2 promotion 128B"; break;
    case 6: dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; esize = 4; x = 4; bw = 8; what = "3D u32 view box 8x24 (32 bytes) x=4 words"; break;
    case 7: what = "3D u8 box 32x24 x=17, encode via dlsym(libcuda)"; break;
    case 8: what = "3D u8 box 32x24 x=17, descriptor as __grid_constant__ parameter"; break;
    case 9: what = "3D u8 box 32x24 x=17, no fence.proxy.async"; break;
    default: break;
    }
    alignas(64) CUtensorMap m;
    memset(&m, 0, sizeof(m));
    cuuint64_t dims[3] = {(cuuint64_t)(W / esize), (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)total};
    cuuint32_t box[3] = {(cuuint32_t)bw, 24, 1}, es[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&m, dt, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("variant %d (%s): encode failed %d\n", variant, what, (int)r); return 1; }
    printf("variant %d descriptor words:", variant);
    for (int i = 0; i < 16; ++i) printf(" %016llx", ((unsigned long long *)&m)[i]);
    printf("\n");
    CUtensorMap *dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
    const unsigned bytes = bw * esize * 24;
    if (rank == 2) { z = 0; k_global<2, true><<<1, 32>>>(dm, x, y, 0, bytes, out); }
    else if (variant == 8) k_param<<<1, 32>>>(m, x, y, z, bytes, out);
    else if (variant == 9) k_global<3, false><<<1, 32>>>(dm, x, y, z, bytes, out);
    else k_global<3, true><<<1, 32>>>(dm, x, y, z, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d (%s): %s\n", variant, what, cudaGetErrorString(e)); return 2; }
    std::vector<unsigned char> got(bytes);
    cudaMemcpy(got.data(), out, bytes, cudaMemcpyDeviceToHost);
    int bad = 0;
    const int rowb = bw * esize;
    for (int rr = 0; rr < 24; ++rr) for (int c = 0; c < rowb; ++c) {
        const int gx = x * esize + c, gy = y + rr;
        const unsigned char want = (gx < W && gy < H) ? h[(size_t)z * total + (size_t)gy * W + gx] : 0;
        bad += got[rr * rowb + c] != want;
    }
    printf("variant %d (%s): OK, %d mismatches\n", variant, what, bad);
    return 0;
}
