// Stand-alone probe of the TMA tile load used by klt_kernel (diagnostic; built and run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_probe tools/tma_probe.cu && /tmp/tma_probe
// Variants: descriptor in kernel-parameter space (__grid_constant__, static / dynamically indexed) or in global memory; 3-D map.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Maps { CUtensorMap m[4]; unsigned ok; };

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ void load_tile(unsigned char *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z, unsigned bytes) {
    if ((threadIdx.x & 31) == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
}

__global__ void k_param(const __grid_constant__ Maps maps, int level, int x, int y, int z, unsigned char *out) {
    __shared__ alignas(128) unsigned char tile[24][32];
    __shared__ unsigned long long bar;
    load_tile(&tile[0][0], &maps.m[level], &bar, x, y, z, 768);
    for (int i = threadIdx.x; i < 768; i += 32) out[i] = (&tile[0][0])[i];
}
__global__ void k_global(const Maps *maps, int level, int x, int y, int z, unsigned char *out) {
    __shared__ alignas(128) unsigned char tile[24][32];
    __shared__ unsigned long long bar;
    load_tile(&tile[0][0], &maps->m[level], &bar, x, y, z, 768);
    for (int i = threadIdx.x; i < 768; i += 32) out[i] = (&tile[0][0])[i];
}

int main() {
    const int W[4] = {640, 320, 160, 80}, H[4] = {480, 240, 120, 60};
    size_t off[4], total = 0;
    for (int l = 0; l < 4; ++l) { off[l] = total; total += (size_t)W[l] * H[l]; }
    const int N = 6;
    std::vector<unsigned char> h(total * N);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d, *out;
    cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 768);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    printf("entry point: %s qr=%d fn=%p\n", cudaGetErrorString(ge), (int)qr, fn);
    if (!fn) return 1;
    Maps m; memset(&m, 0, sizeof(m));
    for (int l = 0; l < 4; ++l) {
        cuuint64_t dims[3] = {(cuuint64_t)W[l], (cuuint64_t)H[l], (cuuint64_t)N};
        cuuint64_t strides[2] = {(cuuint64_t)W[l], (cuuint64_t)total};
        cuuint32_t box[3] = {32, 24, 1}, es[3] = {1, 1, 1};
        CUresult r = ((EncodeFn)fn)(&m.m[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d + off[l], dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode level %d: %d\n", l, (int)r);
        if (r == CUDA_SUCCESS) m.ok |= 1u << l;
    }
    Maps *dm; cudaMalloc(&dm, sizeof(Maps)); cudaMemcpy(dm, &m, sizeof(Maps), cudaMemcpyHostToDevice);
    std::vector<unsigned char> got(768);
    for (int variant = 0; variant < 2; ++variant)
        for (int l = 0; l < 4; ++l) {
            const int x = 17, y = 9, z = 3;
            if (variant == 0) k_global<<<1, 32>>>(dm, l, x, y, z, out); else k_param<<<1, 32>>>(m, l, x, y, z, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %s level %d: %s\n", variant ? "param" : "global", l, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(got.data(), out, 768, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int r = 0; r < 24; ++r) for (int c = 0; c < 32; ++c) {
                const int gx = x + c, gy = y + r;
                const unsigned char want = (gx < W[l] && gy < H[l]) ? h[(size_t)z * total + off[l] + (size_t)gy * W[l] + gx] : 0;
                bad += got[r * 32 + c] != want;
            }
            printf("variant %s level %d: %d mismatches\n", variant ? "param" : "global", l, bad);
        }
    return 0;
}
