// Front-end kernels for sm_100a.  See frontend.cuh.  Compiled with -fmad=false: the float
// arithmetic of CLAHE interpolation, Harris and LK mirrors OpenCV's scalar code, which is not
// FMA-contracted in its SSE3 baseline build.
#include "frontend.cuh"
#include "launch_attr.cuh"
#ifndef CUEMU
#include <cuda.h>                 // CUtensorMap, cuTensorMapEncodeTiled (resolved at run time through cudaGetDriverEntryPoint)
#endif
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace xvfe {

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}
// single reflection, valid for -len < p < 2*len - 1 (all KLT window offsets); clamped for safety
__device__ __forceinline__ int reflect1(int p, int len) {
    p = p < 0 ? -p : p;
    p = p >= len ? 2 * len - 2 - p : p;
    return min(max(p, 0), len - 1);
}
__device__ __forceinline__ int cv_round(float v) { return __float2int_rn(v); }   // round half to even
__device__ __forceinline__ uint8_t sat_u8(int v) { return (uint8_t)min(max(v, 0), 255); }

// ------------------------------------------------------------------------------------------------
// K1 CLAHE.  OpenCV clahe.cpp: per-tile histogram -> clip -> redistribute -> cdf LUT; then bilinear
// blend of the four neighbouring tile LUTs (float).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clahe_lut_kernel(const uint8_t *__restrict__ src, const int *__restrict__ sidx,
                                                         size_t src_stride, uint8_t *__restrict__ lut,
                                                         int w, int h, int tw, int th, int tiles_x, int clip,
                                                         float lut_scale) {
    __shared__ int hist[256];
    __shared__ int wsum[8];
    __shared__ int scan_base[8];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, img = blockIdx.y;
    const int tx = tile % tiles_x, ty = tile / tiles_x;
    const uint8_t *I = src + (size_t)(sidx ? sidx[img] : img) * src_stride;
    hist[tid] = 0;
    __syncthreads();
    const int npx = tw * th;
    const int x0 = tx * tw, y0 = ty * th;
    for (int i = tid; i < npx; i += 256) {
        int ly = i / tw, lx = i - ly * tw;
        int x = x0 + lx, y = y0 + ly;
        if (x >= w) x = reflect101(x, w);   // copyMakeBorder(..., BORDER_REFLECT_101) on right/bottom
        if (y >= h) y = reflect101(y, h);
        atomicAdd(&hist[I[(size_t)y * w + x]], 1);
    }
    __syncthreads();
    int v = hist[tid];
    int excess = 0;
    if (clip > 0 && v > clip) { excess = v - clip; v = clip; }
    // block sum of excess
    int s = excess;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) wsum[tid >> 5] = s;
    __syncthreads();
    int clipped = 0;
    for (int k = 0; k < 8; ++k) clipped += wsum[k];
    if (clip > 0) {
        int batch = clipped / 256;
        int residual = clipped - batch * 256;
        v += batch;
        if (residual != 0) {
            int step = max(256 / residual, 1);
            if (tid % step == 0 && tid / step < residual) v++;
        }
    }
    // inclusive scan over 256 bins
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += t;
    }
    __syncthreads();
    if ((tid & 31) == 31) wsum[tid >> 5] = inc;
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int k = 0; k < 8; ++k) { scan_base[k] = acc; acc += wsum[k]; }
    }
    __syncthreads();
    int sum = inc + scan_base[tid >> 5];
    float f = (float)sum * lut_scale;
    lut[((size_t)img * gridDim.x + tile) * 256 + tid] = sat_u8(cv_round(f));
}

__global__ void __launch_bounds__(256) clahe_apply_kernel(const uint8_t *__restrict__ src, const int *__restrict__ sidx,
                                                           size_t src_stride, uint8_t *__restrict__ dst,
                                                           const int *__restrict__ didx, size_t dst_stride,
                                                           const uint8_t *__restrict__ lut, int w, int h, int tiles_x,
                                                           int tiles_y, float inv_tw, float inv_th) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int img = blockIdx.z;
    if (x >= w) return;
    const uint8_t *L = lut + (size_t)img * tiles_x * tiles_y * 256;
    float txf = (float)x * inv_tw - 0.5f;
    int tx1 = (int)floorf(txf);
    int tx2 = tx1 + 1;
    float xa = txf - (float)tx1, xa1 = 1.0f - xa;
    tx1 = max(tx1, 0); tx2 = min(tx2, tiles_x - 1);
    float tyf = (float)y * inv_th - 0.5f;
    int ty1 = (int)floorf(tyf);
    int ty2 = ty1 + 1;
    float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
    ty1 = max(ty1, 0); ty2 = min(ty2, tiles_y - 1);
    const size_t p = (size_t)y * w + x;
    int v = src[(size_t)(sidx ? sidx[img] : img) * src_stride + p];
    const uint8_t *P1 = L + (size_t)ty1 * tiles_x * 256;
    const uint8_t *P2 = L + (size_t)ty2 * tiles_x * 256;
    float l11 = (float)P1[tx1 * 256 + v], l12 = (float)P1[tx2 * 256 + v];
    float l21 = (float)P2[tx1 * 256 + v], l22 = (float)P2[tx2 * 256 + v];
    float res = (l11 * xa1 + l12 * xa) * ya1 + (l21 * xa1 + l22 * xa) * ya;
    dst[(size_t)(didx ? didx[img] : img) * dst_stride + p] = sat_u8(cv_round(res));
}

void clahe(const uint8_t *src, const int *src_idx, size_t src_stride, uint8_t *dst, const int *dst_idx, size_t dst_stride,
           uint8_t *lut, int n, int w, int h, double clip_limit, int tiles_x, int tiles_y, cudaStream_t st) {
    int wext = (w % tiles_x == 0) ? w : w + (tiles_x - w % tiles_x);
    int hext = (h % tiles_y == 0) ? h : h + (tiles_y - h % tiles_y);
    if (w % tiles_x != 0 || h % tiles_y != 0) { wext = w + (tiles_x - w % tiles_x); hext = h + (tiles_y - h % tiles_y); }
    int tw = wext / tiles_x, th = hext / tiles_y;
    int area = tw * th;
    float lut_scale = 255.0f / (float)area;
    int clip = 0;
    if (clip_limit > 0.0) { clip = (int)(clip_limit * area / 256); clip = clip > 1 ? clip : 1; }
    clahe_lut_kernel<<<dim3(tiles_x * tiles_y, n), 256, 0, st>>>(src, src_idx, src_stride, lut, w, h, tw, th, tiles_x, clip, lut_scale);
    float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    clahe_apply_kernel<<<dim3((w + 255) / 256, h, n), 256, 0, st>>>(src, src_idx, src_stride, dst, dst_idx, dst_stride, lut, w, h, tiles_x, tiles_y, inv_tw, inv_th);
}

// ------------------------------------------------------------------------------------------------
// K2 pyrDown (5x5 Gaussian [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, rounding (s + 128) >> 8)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pyrdown_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                                       const int *__restrict__ idx, int sw, int sh, int dw, int dh,
                                                       size_t img_stride) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= dw || y >= dh) return;
    const size_t item = idx ? idx[blockIdx.z] : blockIdx.z;
    const uint8_t *S = src + item * img_stride;
    uint8_t *D = dst + item * img_stride;
    const int wts[5] = {1, 4, 6, 4, 1};
    int acc = 0;
#pragma unroll
    for (int j = -2; j <= 2; ++j) {
        int yy = reflect101(2 * y + j, sh);
        const uint8_t *row = S + (size_t)yy * sw;
        int r = 0;
#pragma unroll
        for (int i = -2; i <= 2; ++i) {
            int xx = reflect101(2 * x + i, sw);
            r += wts[i + 2] * row[xx];
        }
        acc += wts[j + 2] * r;
    }
    D[(size_t)y * dw + x] = (uint8_t)((acc + 128) >> 8);
}

void pyramid(uint8_t *pyr, const int *idx, int n, const PyrLayout &L, cudaStream_t st) {
    for (int l = 1; l < kPyrLevels; ++l) {
        dim3 grid((L.w[l] + 31) / 32, (L.h[l] + 7) / 8, n);
        pyrdown_kernel<<<grid, 256, 0, st>>>(pyr + L.off[l - 1], pyr + L.off[l], idx, L.w[l - 1], L.h[l - 1], L.w[l], L.h[l], L.total);
    }
}

// Scharr derivatives as calcSharrDeriv (lkpyramid.cpp): interleaved (dx, dy) int16.
__global__ void scharr_kernel(const uint8_t *__restrict__ img, int16_t *__restrict__ deriv, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w) return;
    const uint8_t *I = img + (size_t)blockIdx.z * w * h;
    int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
    int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    auto P = [&](int xx, int yy) { return (int)I[(size_t)yy * w + xx]; };
    int t0m = (P(xm, ym) + P(xm, yp)) * 3 + P(xm, y) * 10;
    int t0p = (P(xp, ym) + P(xp, yp)) * 3 + P(xp, y) * 10;
    int t1m = P(xm, yp) - P(xm, ym);
    int t1c = P(x, yp) - P(x, ym);
    int t1p = P(xp, yp) - P(xp, ym);
    size_t o = ((size_t)blockIdx.z * w * h + (size_t)y * w + x) * 2;
    deriv[o] = (int16_t)(t0p - t0m);
    deriv[o + 1] = (int16_t)((t1p + t1m) * 3 + t1c * 10);
}
void scharr(const uint8_t *img, int16_t *deriv, int n, int w, int h, cudaStream_t st) {
    scharr_kernel<<<dim3((w + 127) / 128, h, n), 128, 0, st>>>(img, deriv, w, h);
}

// ------------------------------------------------------------------------------------------------
// K3 GFTT with Harris response (goodFeaturesToTrack(useHarris) = cornerHarris(block 3, Sobel 3,
// k 0.04) -> threshold q*max -> 3x3 NMS -> sort -> min-distance grid -> first max_corners).
// ------------------------------------------------------------------------------------------------
constexpr int HT_W = 32, HT_H = 16;       // response tile
template <int HALO>                        // HALO = 0: max pass, 1: NMS pass
__device__ __forceinline__ void harris_tile(const uint8_t *__restrict__ I, int w, int h, int bx, int by,
                                            float (*resp)[HT_W + 2 * HALO]) {
    // image tile with halo HALO + 2, derivative tiles with halo HALO + 1
    constexpr int IW = HT_W + 2 * (HALO + 2), IH = HT_H + 2 * (HALO + 2);
    constexpr int DW = HT_W + 2 * (HALO + 1), DH = HT_H + 2 * (HALO + 1);
    constexpr int RW = HT_W + 2 * HALO, RH = HT_H + 2 * HALO;
    __shared__ uint8_t simg[IH][IW];
    __shared__ float sdx[DH][DW], sdy[DH][DW];
    const int tid = threadIdx.x;
    const int ix0 = bx - (HALO + 2), iy0 = by - (HALO + 2);
    for (int i = tid; i < IW * IH; i += blockDim.x) {
        int ly = i / IW, lx = i - ly * IW;
        simg[ly][lx] = I[(size_t)reflect101(iy0 + ly, h) * w + reflect101(ix0 + lx, w)];
    }
    __syncthreads();
    const float s1 = (float)(1.0 / 3060.0), s2 = (float)(2.0 / 3060.0);   // scale = 1/(4*3*255)
    const int dx0 = bx - (HALO + 1), dy0 = by - (HALO + 1);
    for (int i = tid; i < DW * DH; i += blockDim.x) {
        int ly = i / DW, lx = i - ly * DW;
        int gx = dx0 + lx, gy = dy0 + ly;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
            int cx = lx + 1, cy = ly + 1;   // position in simg
            // Dx: row filter [-1 0 1] (exact), column filter [s 2s s]
            float rm = (float)((int)simg[cy - 1][cx + 1] - (int)simg[cy - 1][cx - 1]);
            float r0 = (float)((int)simg[cy][cx + 1] - (int)simg[cy][cx - 1]);
            float rp = (float)((int)simg[cy + 1][cx + 1] - (int)simg[cy + 1][cx - 1]);
            sdx[ly][lx] = s2 * r0 + s1 * (rm + rp);
            // Dy: row filter [s 2s s], column filter [-1 0 1]
            float um = s2 * (float)simg[cy - 1][cx] + s1 * (float)((int)simg[cy - 1][cx - 1] + (int)simg[cy - 1][cx + 1]);
            float up = s2 * (float)simg[cy + 1][cx] + s1 * (float)((int)simg[cy + 1][cx - 1] + (int)simg[cy + 1][cx + 1]);
            sdy[ly][lx] = up - um;
        }
    }
    __syncthreads();
    // out-of-image derivative positions take the value at the reflected position (BORDER_REFLECT_101 of boxFilter)
    for (int i = tid; i < DW * DH; i += blockDim.x) {
        int ly = i / DW, lx = i - ly * DW;
        int gx = dx0 + lx, gy = dy0 + ly;
        if (!(gx >= 0 && gx < w && gy >= 0 && gy < h)) {
            int rx = reflect101(gx, w) - dx0, ry = reflect101(gy, h) - dy0;
            if (rx >= 0 && rx < DW && ry >= 0 && ry < DH) { sdx[ly][lx] = sdx[ry][rx]; sdy[ly][lx] = sdy[ry][rx]; }
            else { sdx[ly][lx] = 0.f; sdy[ly][lx] = 0.f; }
        }
    }
    __syncthreads();
    for (int i = tid; i < RW * RH; i += blockDim.x) {
        int ly = i / RW, lx = i - ly * RW;
        int gx = bx - HALO + lx, gy = by - HALO + ly;
        float val = -FLT_MAX;
        if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
            float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float ra = 0.f, rb = 0.f, rc = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float dx = sdx[ly + j][lx + k], dy = sdy[ly + j][lx + k];
                    ra += dx * dx; rb += dx * dy; rc += dy * dy;
                }
                a += ra; b += rb; c += rc;
            }
            double da = a, db = b, dc = c;
            val = (float)(da * dc - db * db - 0.04 * (da + dc) * (da + dc));
        }
        resp[ly][lx] = val;
    }
    __syncthreads();
}

// Single pass over the image: Harris response on a tile with a 1-px halo, 3x3 non-maximum suppression
// and the per-image maximum.  Every positive local maximum becomes a candidate; the quality threshold
// (q * max, which needs the global maximum) is applied by corner_select_kernel.
__global__ void __launch_bounds__(256) harris_nms_kernel(const uint8_t *__restrict__ img, const int *__restrict__ idx,
                                                          size_t img_stride, int w, int h,
                                                          unsigned int *__restrict__ maxval,
                                                          unsigned long long *__restrict__ cand,
                                                          int *__restrict__ cand_count) {
    __shared__ float resp[HT_H + 2][HT_W + 2];
    __shared__ float wmax[8];
    const uint8_t *I = img + (size_t)(idx ? idx[blockIdx.z] : blockIdx.z) * img_stride;
    const int bx = blockIdx.x * HT_W, by = blockIdx.y * HT_H;
    harris_tile<1>(I, w, h, bx, by, resp);
    float m = 0.f;
    for (int i = threadIdx.x; i < HT_W * HT_H; i += blockDim.x) {
        int ly = i / HT_W, lx = i - ly * HT_W;
        int gx = bx + lx, gy = by + ly;
        if (gx >= w || gy >= h) continue;
        float v = resp[ly + 1][lx + 1];
        m = fmaxf(m, v);
        if (gx < 1 || gx >= w - 1 || gy < 1 || gy >= h - 1) continue;
        if (!(v > 0.f)) continue;
        bool is_max = true;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) is_max = is_max && (resp[ly + j][lx + k] <= v);
        if (is_max) {
            int slot = atomicAdd(&cand_count[blockIdx.z], 1);
            if (slot < kMaxCand)
                cand[(size_t)blockIdx.z * kMaxCand + slot] =
                    ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)(gy * w + gx);
        }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 8; ++k) m = fmaxf(m, wmax[k]);
        if (m > 0.f) atomicMax(&maxval[blockIdx.z], __float_as_uint(m));   // positive floats order as uints
    }
}

// Candidate selection, one block per image: (1) a histogram over the top bits of the response picks
// a cut so that only the strongest ~kSortTarget candidates are compacted and bitonic-sorted
// (descending by response, then by descending address == descending y*w+x, as greaterThanPtr);
// (2) greedy min-distance selection on a cell grid (goodFeaturesToTrack).  If the greedy pass runs out
// of sorted candidates before max_corners are accepted, the whole list is sorted and the pass repeated.
constexpr int kSortTarget = 2048;
constexpr int kHistBins = 2048;
constexpr int kSortSmem = 4096;        // candidates that can be sorted inside shared memory (32 KB of keys)
__device__ void bitonic_desc(unsigned long long *C, int np2) {
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) {
                int l = i ^ j;
                if (l > i) {
                    unsigned long long a = C[i], b = C[l];
                    bool desc = ((i & k) == 0);
                    if (desc ? (a < b) : (a > b)) { C[i] = b; C[l] = a; }
                }
            }
            __syncthreads();
        }
    }
}
__global__ void __launch_bounds__(1024) corner_select_kernel(unsigned long long *__restrict__ cand,
                                                              const int *__restrict__ cand_count,
                                                              const unsigned int *__restrict__ maxval, float quality,
                                                              int w, int h, int max_corners, float min_distance,
                                                              float *__restrict__ corners, int *__restrict__ count) {
    const int img = blockIdx.x;
    unsigned long long *C = cand + (size_t)img * kMaxCand;
    // threshold(eig, maxVal * qualityLevel, THRESH_TOZERO): drop candidates that are not above the cut
    __shared__ int s_n;
    {
        const int n_all = min(cand_count[img], kMaxCand);
        const float thr = (float)((double)__uint_as_float(maxval[img]) * (double)quality);
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        // stable-order-free compaction in place: survivors are appended to the front region through a counter; a
        // survivor at position i is only ever written to a slot <= i because slots are handed out in index order per pass
        for (int base = 0; base < n_all; base += blockDim.x) {
            const int i = base + threadIdx.x;
            unsigned long long key = 0ull;
            bool keep = false;
            if (i < n_all) { key = C[i]; keep = __uint_as_float((unsigned int)(key >> 32)) > thr; }
            __syncthreads();
            // warp-aggregated append: one shared-memory atomic per warp instead of one per survivor (tens of thousands of
            // candidates pass through this counter; per-thread atomics on it were the kernel's largest single cost)
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (bal) {
                const int lane = threadIdx.x & 31, leader = __ffs((int)bal) - 1;
                int slot0 = 0;
                if (lane == leader) slot0 = atomicAdd(&s_n, __popc(bal));
                slot0 = __shfl_sync(0xffffffffu, slot0, leader);
                if (keep) C[slot0 + __popc(bal & ((1u << lane) - 1u))] = key;
            }
            __syncthreads();
        }
    }
    const int n = s_n;
    extern __shared__ unsigned char smem_raw[];
    constexpr int CELL_CAP = 8;
    const int cell = __float2int_rn(min_distance);
    const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
    short *cell_cnt = reinterpret_cast<short *>(smem_raw);                 // gw*gh
    short2 *cell_pts = reinterpret_cast<short2 *>(cell_cnt + ((gw * gh + 1) & ~1));   // gw*gh*CELL_CAP
    // sort buffer behind the grid (8-byte aligned): the strongest <= kSortSmem candidates are sorted here instead of in global memory
    unsigned long long *S = reinterpret_cast<unsigned long long *>(smem_raw + (((size_t)((gw * gh + 1) & ~1) * sizeof(short) + (size_t)gw * gh * CELL_CAP * sizeof(short2) + 7) & ~(size_t)7));
    __shared__ int hist[kHistBins];
    __shared__ int s_cut, s_m, s_done, s_ncorners;
    // ---- (1) choose the cut: bin = top 11 bits below the sign of the float response (exponent + 3 mantissa bits) ----
    int m_sorted = n;
    if (n > 2 * kSortTarget) {
        for (int i = threadIdx.x; i < kHistBins; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(&hist[(unsigned)(C[i] >> 32) >> 20], 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, b = kHistBins - 1;
            for (; b > 0; --b) { acc += hist[b]; if (acc >= kSortTarget) break; }
            s_cut = b; s_m = 0;
        }
        __syncthreads();
        // partition: strong candidates to the front (order inside the partitions is irrelevant before the sort)
        const int cut = s_cut;
        // two-pointer style compaction through a shared counter: swap-based in-place partition
        // pass 1: count strong
        int strong = 0;
        for (int b = cut + threadIdx.x; b < kHistBins; b += blockDim.x) strong += hist[b];
        for (int o = 16; o > 0; o >>= 1) strong += __shfl_xor_sync(0xffffffffu, strong, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_m, strong);
        __syncthreads();
        const int m = s_m;
        // elements in [0,m) that are weak must swap with elements in [m,n) that are strong
        __shared__ int s_wi, s_si;
        if (threadIdx.x == 0) { s_wi = 0; s_si = 0; }
        __syncthreads();
        // collect indices of misplaced elements into the tail of the candidate buffer of this image (scratch beyond n)
        // kMaxCand slots exist; use the region [kMaxCand - 2*m_mis, kMaxCand) only if free -> simpler: do it in two sweeps with atomics
        unsigned long long *W = C;   // alias
        // sweep A: weak elements in the front block record their positions
        int *pos_w = reinterpret_cast<int *>(cell_pts);            // reuse the (not yet used) grid storage: gw*gh*CELL_CAP*4 bytes
        const int pos_cap = gw * gh * CELL_CAP / 2;                  // ints available per list
        int *pos_s = pos_w + pos_cap;
        bool overflow = false;
        for (int i = threadIdx.x; i < m; i += blockDim.x)
            if ((int)((unsigned)(W[i] >> 32) >> 20) < cut) { int k = atomicAdd(&s_wi, 1); if (k < pos_cap) pos_w[k] = i; else overflow = true; }
        for (int i = m + threadIdx.x; i < n; i += blockDim.x)
            if ((int)((unsigned)(W[i] >> 32) >> 20) >= cut) { int k = atomicAdd(&s_si, 1); if (k < pos_cap) pos_s[k] = i; else overflow = true; }
        const int any_over = __syncthreads_or(overflow ? 1 : 0);
        if (!any_over) {
            const int ns = min(s_wi, s_si);
            for (int k = threadIdx.x; k < ns; k += blockDim.x) {
                const unsigned long long t = W[pos_w[k]]; W[pos_w[k]] = W[pos_s[k]]; W[pos_s[k]] = t;
            }
            m_sorted = m;
        }
        __syncthreads();
    }
    for (int attempt = 0; attempt < 2; ++attempt) {
        const int cnt = attempt == 0 ? m_sorted : n;
        int np2 = 1;
        while (np2 < cnt) np2 <<= 1;
        const unsigned long long *sorted = C;
        // the usual case: the strong candidates fit the shared-memory buffer -> ~80 compare-exchange passes at shared-memory
        // latency with a CTA barrier each, instead of the same passes on global memory (the kernel's former critical path)
        if (np2 <= kSortSmem) {
            for (int i = threadIdx.x; i < np2; i += blockDim.x) S[i] = i < cnt ? C[i] : 0ull;
            __syncthreads();
            bitonic_desc(S, np2);
            sorted = S;
        } else
        // pad [cnt, np2) with zeros: when only the strong part is sorted, save and restore the overwritten tail
        if (attempt == 0 && np2 > cnt && cnt < n) {
            // the padding would destroy weak candidates in [cnt, np2): sort the larger power of two instead (they are
            // all weaker than the cut, so the strong prefix still comes out in exact order)
            np2 = min(np2, 1 << 30);
            const int lim = min(np2, n);
            for (int i = lim + threadIdx.x; i < np2; i += blockDim.x) C[i] = 0ull;
        } else {
            for (int i = cnt + threadIdx.x; i < np2; i += blockDim.x) C[i] = 0ull;
        }
        __syncthreads();
        if (sorted == C) bitonic_desc(C, np2);
        const int usable = attempt == 0 ? m_sorted : n;      // candidates guaranteed to be in global order
        for (int i = threadIdx.x; i < gw * gh; i += blockDim.x) cell_cnt[i] = 0;
        if (threadIdx.x == 0) { s_done = 0; s_ncorners = 0; }
        __syncthreads();
        if (threadIdx.x < 32) {
            // Greedy min-distance selection in response order (goodFeaturesToTrack).  Most candidates fall next to an already accepted
            // corner, so 32 candidates are tested against the grid AT ONCE (one lane each); the few that pass are then accepted
            // strictly in order, each acceptance striking the later passers of the same batch that lie within the distance --
            // exactly the sequential rule, at a fraction of its ~3000 serial iterations per image.
            const int lane = threadIdx.x;
            const float md2 = min_distance * min_distance;
            int ncorners = 0;
            int i = 0;
            for (; i < usable && ncorners < max_corners; i += 32) {
                const bool valid = i + lane < usable;
                const unsigned long long key = valid ? sorted[i + lane] : 0ull;
                const int idx = (int)(key & 0xffffffffu);
                const int y = idx / w, x = idx - y * w;
                const int xc = x / cell, yc = y / cell;
                bool bad = !valid;
                if (valid)
                    for (int cy = max(yc - 1, 0); cy <= min(yc + 1, gh - 1) && !bad; ++cy)
                        for (int cx = max(xc - 1, 0); cx <= min(xc + 1, gw - 1) && !bad; ++cx) {
                            const int c = cy * gw + cx;
                            const int mm = cell_cnt[c];
                            for (int r = 0; r < mm; ++r) {
                                const short2 p = cell_pts[c * CELL_CAP + r];
                                const float dx = (float)(x - p.x), dy = (float)(y - p.y);
                                if (dx * dx + dy * dy < md2) { bad = true; break; }
                            }
                        }
                unsigned pass = __ballot_sync(0xffffffffu, !bad);
                while (pass && ncorners < max_corners) {
                    const int q = __ffs((int)pass) - 1;
                    pass &= ~(1u << q);
                    const int xq = __shfl_sync(0xffffffffu, x, q), yq = __shfl_sync(0xffffffffu, y, q);
                    if (lane == q) {
                        const int c = yc * gw + xc;
                        const int mm = cell_cnt[c];
                        if (mm < CELL_CAP) { cell_pts[c * CELL_CAP + mm] = make_short2((short)x, (short)y); cell_cnt[c] = mm + 1; }
                        float *o = corners + ((size_t)img * kMaxCorners + ncorners) * 3;
                        o[0] = (float)x; o[1] = (float)y; o[2] = __uint_as_float((unsigned int)(key >> 32));
                    }
                    ncorners++;
                    const float dx = (float)(x - xq), dy = (float)(y - yq);
                    pass &= ~__ballot_sync(0xffffffffu, dx * dx + dy * dy < md2);       // later passers too close to the new corner
                    __syncwarp();
                }
            }
            if (lane == 0) { s_ncorners = ncorners; s_done = (ncorners >= max_corners || usable >= n) ? 1 : 0; }
        }
        __syncthreads();
        if (s_done) break;
    }
    if (threadIdx.x == 0) count[img] = s_ncorners;
}

void gftt_harris(const uint8_t *img, const int *idx, size_t img_stride, int n, int w, int h, int max_corners, float quality,
                 float min_distance, float *corners, int *count, unsigned long long *cand, int *cand_count,
                 unsigned int *maxval, cudaStream_t st) {
    cudaMemsetAsync(maxval, 0, sizeof(unsigned int) * n, st);
    cudaMemsetAsync(cand_count, 0, sizeof(int) * n, st);
    dim3 grid((w + HT_W - 1) / HT_W, (h + HT_H - 1) / HT_H, n);
    harris_nms_kernel<<<grid, 256, 0, st>>>(img, idx, img_stride, w, h, maxval, cand, cand_count);
    int cell = (int)lrintf(min_distance);
    int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
    size_t smem = (((size_t)((gw * gh + 1) & ~1) * sizeof(short) + (size_t)gw * gh * 8 * sizeof(short2) + 7) & ~(size_t)7) + (size_t)kSortSmem * sizeof(unsigned long long);
    if (max_corners > kMaxCorners) max_corners = kMaxCorners;
    static xv::SmemOptIn attr_done;
    xv::ensure_dynamic_smem(corner_select_kernel, (int)smem, attr_done);
    corner_select_kernel<<<n, 1024, smem, st>>>(cand, cand_count, maxval, quality, w, h, max_corners, min_distance, corners, count);
}

// ------------------------------------------------------------------------------------------------
// K4 pyramidal Lucas-Kanade (calcOpticalFlowPyrLK, winSize 21x21, maxLevel 3, 30 iterations, eps
// 0.01, OPTFLOW_USE_INITIAL_FLOW, minEigThreshold 1e-4) -- one warp per keypoint, forward pass,
// the reference's gates, backward pass and the 0.5 px round-trip check fused.
// ------------------------------------------------------------------------------------------------
constexpr int WIN = 21;
constexpr int WINP = WIN + 1;     // 22: window + 1 for bilinear taps
constexpr int RAWP = WIN + 3;     // 24: + 1 px ring for the Scharr stencil
constexpr int KLT_WARPS = 4;

constexpr int JREG = 32;          // side of the staged search region (window 22 + 10 px of slack)
// Row pitch of both staged tiles = inner box of their TMA loads.  A tile load must START on a 16-byte boundary of global
// memory (measured: tools/tma_probe.cu, a u8 box starting at x = 17 raises an illegal-instruction fault on sm_100, x = 16 is
// fine), so the box starts at x & ~15 and is 48 wide: the wanted 24 / 32 columns begin at offset x & 15 <= 15 inside it.
constexpr int RAWPITCH = 48;
struct alignas(128) KltSmem {
    union {                           // both are destinations of TMA tile loads: 128-byte aligned, pitch = box width
        uint8_t raw[RAWP][RAWPITCH];  // level image around the template window (reflect-101 border)
        uint8_t J[JREG][RAWPITCH];    // search region in the other image (staged once per level, re-staged on exit)
    };
    short dx[WINP][WINP];             // Scharr derivatives at the 22x22 positions (0 outside the image)
    short dy[WINP][WINP];
    short I[WIN * WIN];               // template patch  (<< 5)
    short Ix[WIN * WIN], Iy[WIN * WIN];
    unsigned long long bar;           // mbarrier of this warp's TMA loads
};

// ---- TMA staging (north_star: "image patches ... staged through TMA into shared memory") --------------------------------
// One tensor map per pyramid level over the whole pyramid buffer: dims (x = level width, y = level height, z = pyramid
// index), u8, box 48 x 24 (template neighbourhood) or 48 x 32 (search region), see RAWPITCH.  A tile that lies completely inside the level
// is fetched by ONE cp.async.bulk.tensor issued by lane 0 and awaited on the warp's mbarrier, instead of 18-32 predicated
// byte loads per lane with reflect-101 index arithmetic; tiles that touch the border keep the scalar path (TMA can only
// zero-fill outside the tensor, OpenCV reflects).  Levels whose width / offset / pyramid stride is not a multiple of 16
// bytes (e.g. 94 x 60 of a 752 x 480 image) have no map and always take the scalar path.
#ifndef CUEMU
struct KltMaps {
    CUtensorMap tmpl[kPyrLevels];     // box 48 x 24
    CUtensorMap srch[kPyrLevels];     // box 48 x 32
    unsigned ok;                      // bit l: level l has maps
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// lane 0: arm the barrier with the tile's byte count and issue the tile load (coordinates x, y, pyramid index)
__device__ __forceinline__ void tma_tile_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int x, int y, int z, unsigned bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the warp's earlier generic-proxy accesses to dst come first
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#else
struct KltMaps { unsigned ok; };
#endif
// what one lk_track_point call needs to use the maps: which pyramid (z) each image is, and the warp's barrier phase
struct KltTma {
    const KltMaps *mapsI, *mapsJ;     // maps of the buffers holding the template image / the searched image (null: scalar path)
    int zI, zJ;
    unsigned phase;
};

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One LK run for one point through all levels.  Returns status; out = tracked point (level 0).
// Lane l owns window pixels i = l + 32 k (k < 14); their (y, x) are computed once (no div/mod in the loops).
__device__ bool lk_track_point(const uint8_t *__restrict__ pyrI, const uint8_t *__restrict__ pyrJ, const PyrLayout &L,
                               float px, float py, float gx, float gy, float &ox, float &oy, KltSmem &sm, KltTma &tma) {
    const int lane = threadIdx.x & 31;
    constexpr int NPX = (WIN * WIN + 31) / 32;      // 14
    int pxy[NPX];                                   // y << 8 | x, or -1 beyond the 441 pixels
#pragma unroll
    for (int k = 0; k < NPX; ++k) {
        const int i = lane + 32 * k;
        pxy[k] = i < WIN * WIN ? (((i / WIN) << 8) | (i % WIN)) : -1;
    }
    bool status = true;
    float outx = 0.f, outy = 0.f;
    for (int level = kPyrLevels - 1; level >= 0; --level) {
        const int cols = L.w[level], rows = L.h[level];
        const uint8_t *I = pyrI + L.off[level];
        const uint8_t *J = pyrJ + L.off[level];
        const float sc = 1.f / (float)(1 << level);
        float ppx = px * sc, ppy = py * sc;
        float nx, ny;
        if (level == kPyrLevels - 1) { nx = gx * sc; ny = gy * sc; }
        else { nx = outx * 2.f; ny = outy * 2.f; }
        outx = nx; outy = ny;
        ppx -= 10.f; ppy -= 10.f;
        const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
        if (ipx < -WIN || ipx >= cols || ipy < -WIN || ipy >= rows) {
            if (level == 0) status = false;
            continue;
        }
        float a = ppx - (float)ipx, b = ppy - (float)ipy;
        int iw00 = cv_round((1.f - a) * (1.f - b) * 16384.f);
        int iw01 = cv_round(a * (1.f - b) * 16384.f);
        int iw10 = cv_round((1.f - a) * b * 16384.f);
        int iw11 = 16384 - iw00 - iw01 - iw10;
        // stage the raw template neighbourhood (24 x 24): loads first, stores after, so the loads overlap
        __syncwarp();
        bool tile_done = false;
        int ro = 0;                                    // column of the tile's first wanted pixel inside the staged rows
#ifndef CUEMU
        if (tma.mapsI != nullptr && ((tma.mapsI->ok >> level) & 1u) && ipx >= 1 && ipx + RAWP - 2 < cols && ipy >= 1 && ipy + RAWP - 2 < rows) {
            ro = (ipx - 1) & 15;
            if (lane == 0) tma_tile_3d(&sm.raw[0][0], &tma.mapsI->tmpl[level], &sm.bar, (ipx - 1) - ro, ipy - 1, tma.zI, RAWP * RAWPITCH);
            mbar_wait(&sm.bar, tma.phase);
            tma.phase ^= 1u;
            tile_done = true;
        }
#endif
        if (!tile_done) {
            constexpr int NR = (RAWP * RAWP + 31) / 32;     // 18
            uint8_t v[NR];
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                const int i = lane + 32 * k;
                const int ly = i / RAWP, lx = i - ly * RAWP;
                v[k] = i < RAWP * RAWP ? I[(size_t)reflect1(ipy - 1 + ly, rows) * cols + reflect1(ipx - 1 + lx, cols)] : 0;
            }
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                const int i = lane + 32 * k;
                const int ly = i / RAWP, lx = i - ly * RAWP;
                if (i < RAWP * RAWP) sm.raw[ly][lx] = v[k];
            }
        }
        __syncwarp();
        for (int i = lane; i < WINP * WINP; i += 32) {
            int ly = i / WINP, lx = i - ly * WINP;
            int X = ipx + lx, Y = ipy + ly;
            short ddx = 0, ddy = 0;
            if (X >= 0 && X < cols && Y >= 0 && Y < rows) {
                int cx = ro + lx + 1, cy = ly + 1;
                int t0m = ((int)sm.raw[cy - 1][cx - 1] + (int)sm.raw[cy + 1][cx - 1]) * 3 + (int)sm.raw[cy][cx - 1] * 10;
                int t0p = ((int)sm.raw[cy - 1][cx + 1] + (int)sm.raw[cy + 1][cx + 1]) * 3 + (int)sm.raw[cy][cx + 1] * 10;
                int t1m = (int)sm.raw[cy + 1][cx - 1] - (int)sm.raw[cy - 1][cx - 1];
                int t1c = (int)sm.raw[cy + 1][cx] - (int)sm.raw[cy - 1][cx];
                int t1p = (int)sm.raw[cy + 1][cx + 1] - (int)sm.raw[cy - 1][cx + 1];
                ddx = (short)(t0p - t0m);
                ddy = (short)((t1p + t1m) * 3 + t1c * 10);
            }
            sm.dx[ly][lx] = ddx; sm.dy[ly][lx] = ddy;
        }
        __syncwarp();
        float A11 = 0.f, A12 = 0.f, A22 = 0.f;
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
            if (pxy[k] < 0) continue;
            const int y = pxy[k] >> 8, x = pxy[k] & 255;
            const int i = lane + 32 * k;
            const int cy = y + 1, cx = ro + x + 1;
            int ival = ((int)sm.raw[cy][cx] * iw00 + (int)sm.raw[cy][cx + 1] * iw01 + (int)sm.raw[cy + 1][cx] * iw10 +
                        (int)sm.raw[cy + 1][cx + 1] * iw11 + (1 << 8)) >> 9;
            int ixval = ((int)sm.dx[y][x] * iw00 + (int)sm.dx[y][x + 1] * iw01 + (int)sm.dx[y + 1][x] * iw10 +
                         (int)sm.dx[y + 1][x + 1] * iw11 + (1 << 13)) >> 14;
            int iyval = ((int)sm.dy[y][x] * iw00 + (int)sm.dy[y][x + 1] * iw01 + (int)sm.dy[y + 1][x] * iw10 +
                         (int)sm.dy[y + 1][x + 1] * iw11 + (1 << 13)) >> 14;
            sm.I[i] = (short)ival; sm.Ix[i] = (short)ixval; sm.Iy[i] = (short)iyval;
            A11 += (float)(ixval * ixval); A12 += (float)(ixval * iyval); A22 += (float)(iyval * iyval);
        }
        const float FLT_SCALE = 1.f / (float)(1 << 20);
        A11 = warp_sum(A11) * FLT_SCALE; A12 = warp_sum(A12) * FLT_SCALE; A22 = warp_sum(A22) * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * WIN * WIN);
        if ((double)minEig < 1e-4 || D < FLT_EPSILON) {
            if (level == 0) status = false;
            continue;
        }
        D = 1.f / D;
        // the template (and its gradients) stay in registers for the iterations
        short tI[NPX], tIx[NPX], tIy[NPX];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < NPX; ++k) {
            const int i = lane + 32 * k;
            tI[k] = pxy[k] >= 0 ? sm.I[i] : (short)0; tIx[k] = pxy[k] >= 0 ? sm.Ix[i] : (short)0; tIy[k] = pxy[k] >= 0 ? sm.Iy[i] : (short)0;
        }
        nx -= 10.f; ny -= 10.f;
        float pdx = 0.f, pdy = 0.f;
        int rox = 0, roy = 0, jo = 0;                  // jo: column of the region's first pixel inside the staged rows
        bool staged = false;
        for (int j = 0; j < 30; ++j) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -WIN || inx >= cols || iny < -WIN || iny >= rows) {
                if (level == 0) status = false;
                break;
            }
            a = nx - (float)inx; b = ny - (float)iny;
            iw00 = cv_round((1.f - a) * (1.f - b) * 16384.f);
            iw01 = cv_round(a * (1.f - b) * 16384.f);
            iw10 = cv_round((1.f - a) * b * 16384.f);
            iw11 = 16384 - iw00 - iw01 - iw10;
            if (!staged || inx < rox || inx > rox + JREG - WINP || iny < roy || iny > roy + JREG - WINP) {
                rox = inx - (JREG - WINP) / 2; roy = iny - (JREG - WINP) / 2;
                __syncwarp();
                bool region_done = false;
                jo = 0;
#ifndef CUEMU
                if (tma.mapsJ != nullptr && ((tma.mapsJ->ok >> level) & 1u) && rox >= 0 && rox + JREG <= cols && roy >= 0 && roy + JREG <= rows) {
                    jo = rox & 15;
                    if (lane == 0) tma_tile_3d(&sm.J[0][0], &tma.mapsJ->srch[level], &sm.bar, rox - jo, roy, tma.zJ, JREG * RAWPITCH);
                    mbar_wait(&sm.bar, tma.phase);
                    tma.phase ^= 1u;
                    region_done = true;
                }
#endif
                if (!region_done) {
                    // lane = column, 32 rows: 32 independent loads per lane
                    uint8_t v[JREG];
                    const int cxg = reflect1(rox + lane, cols);
#pragma unroll
                    for (int r = 0; r < JREG; ++r) v[r] = J[(size_t)reflect1(roy + r, rows) * cols + cxg];
#pragma unroll
                    for (int r = 0; r < JREG; ++r) sm.J[r][lane] = v[r];
                }
                __syncwarp();
                staged = true;
            }
            const uint8_t *Jb = &sm.J[iny - roy][jo + inx - rox];
            float b1 = 0.f, b2 = 0.f;
#pragma unroll
            for (int k = 0; k < NPX; ++k) {
                if (pxy[k] < 0) continue;
                const uint8_t *q = Jb + (pxy[k] >> 8) * RAWPITCH + (pxy[k] & 255);
                const int diff = (((int)q[0] * iw00 + (int)q[1] * iw01 + (int)q[RAWPITCH] * iw10 + (int)q[RAWPITCH + 1] * iw11 + (1 << 8)) >> 9) - (int)tI[k];
                b1 += (float)(diff * (int)tIx[k]);
                b2 += (float)(diff * (int)tIy[k]);
            }
            b1 = warp_sum(b1) * FLT_SCALE; b2 = warp_sum(b2) * FLT_SCALE;
            float dxv = (A12 * b2 - A22 * b1) * D;
            float dyv = (A12 * b1 - A11 * b2) * D;
            nx += dxv; ny += dyv;
            outx = nx + 10.f; outy = ny + 10.f;
            if ((double)dxv * (double)dxv + (double)dyv * (double)dyv <= 0.01 * 0.01) break;
            if (j > 0 && fabsf(dxv + pdx) < 0.01f && fabsf(dyv + pdy) < 0.01f) {
                outx -= dxv * 0.5f; outy -= dyv * 0.5f;
                break;
            }
            pdx = dxv; pdy = dyv;
        }
    }
    ox = outx; oy = outy;
    return status;
}

#ifndef CUEMU
#define XV_GRID_CONSTANT __grid_constant__
#else
#define XV_GRID_CONSTANT
#endif
__global__ void __launch_bounds__(32 * KLT_WARPS) klt_kernel(const uint8_t *__restrict__ prev_pyr,
                                                              const uint8_t *__restrict__ next_pyr,
                                                              const int *__restrict__ pidx, const int *__restrict__ nidx,
                                                              PyrLayout L,
                                                              const float *__restrict__ prev_pts,
                                                              float *__restrict__ next_pts,
                                                              const int *__restrict__ npts,
                                                              uint8_t *__restrict__ status,
                                                              const XV_GRID_CONSTANT KltMaps maps_prev,
                                                              const XV_GRID_CONSTANT KltMaps maps_next) {
    __shared__ KltSmem sm[KLT_WARPS];
    const int img = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const int pt = blockIdx.x * KLT_WARPS + warp;
    if (pt >= npts[img]) return;
    const int zp = pidx ? pidx[img] : img, zn = nidx ? nidx[img] : img;
    const uint8_t *P = prev_pyr + (size_t)zp * L.total;
    const uint8_t *N = next_pyr + (size_t)zn * L.total;
    const size_t o = ((size_t)img * kMaxPts + pt) * 2;
    const float px = prev_pts[o], py = prev_pts[o + 1];
    const float gx = next_pts[o], gy = next_pts[o + 1];
    KltTma tf{maps_prev.ok ? &maps_prev : nullptr, maps_next.ok ? &maps_next : nullptr, zp, zn, 0u};
#ifndef CUEMU
    if (maps_prev.ok | maps_next.ok) {       // the barrier belongs to this warp alone: no CTA-wide synchronisation needed
        if ((threadIdx.x & 31) == 0) mbar_init(&sm[warp].bar, 1);
        __syncwarp();
    }
#endif
    float qx, qy;
    bool st = lk_track_point(P, N, L, px, py, gx, gy, qx, qy, sm[warp], tf);
    const int cols = L.w[0], rows = L.h[0];
    // reference gates, opencv_image.cpp:99-113
    if (qx < 20.f || qx >= (float)(cols - 20) || qy < 20.f || qy >= (float)(rows - 20)) st = false;
    if (st) {
        double fx = (double)(qx - px), fy = (double)(qy - py);
        if (sqrt(fx * fx + fy * fy) > (double)(rows / 4)) st = false;
    }
    if (st) {
        float bx, by;
        KltTma tb{tf.mapsJ, tf.mapsI, zn, zp, tf.phase};
        bool stb = lk_track_point(N, P, L, qx, qy, px, py, bx, by, sm[warp], tb);
        double ex = (double)(px - bx), ey = (double)(py - by);
        if (!stb || sqrt(ex * ex + ey * ey) > 0.5) st = false;
    }
    if ((threadIdx.x & 31) == 0) {
        next_pts[o] = qx; next_pts[o + 1] = qy;
        status[(size_t)img * kMaxPts + pt] = st ? 1 : 0;
    }
}

// Tensor maps over a pyramid buffer of n_pyramids pyramids (see KltMaps).  ok == 0 when the driver entry point is missing or
// the layout is not 16-byte regular: klt_track then runs the scalar staging path for every tile.
KltMapSet klt_make_maps(const uint8_t *pyr, int n_pyramids, const PyrLayout &L) {
    KltMapSet out;
    std::memset(&out, 0, sizeof(out));
#ifndef CUEMU
    static_assert(sizeof(KltMaps) <= sizeof(out.bytes), "KltMapSet too small");
    KltMaps m;
    std::memset(&m, 0, sizeof(m));
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) fn = nullptr;
        return (EncodeFn)fn;
    }();
    if (encode != nullptr && n_pyramids > 0 && L.total % 16 == 0 && ((uintptr_t)pyr % 16) == 0) {
        for (int l = 0; l < kPyrLevels; ++l) {
            if (L.w[l] % 16 != 0 || L.off[l] % 16 != 0 || L.w[l] < RAWPITCH || L.h[l] < JREG) continue;
            const cuuint64_t dims[3] = {(cuuint64_t)L.w[l], (cuuint64_t)L.h[l], (cuuint64_t)n_pyramids};
            const cuuint64_t strides[2] = {(cuuint64_t)L.w[l], (cuuint64_t)L.total};       // bytes, dims 1 and 2
            const cuuint32_t estr[3] = {1, 1, 1};
            const cuuint32_t box_t[3] = {RAWPITCH, RAWP, 1}, box_s[3] = {RAWPITCH, JREG, 1};
            void *base = (void *)(pyr + L.off[l]);
            if (encode(&m.tmpl[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box_t, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) continue;
            if (encode(&m.srch[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box_s, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) continue;
            m.ok |= 1u << l;
        }
    }
    if (const char *e = std::getenv("XRVIO_KLT_TMA")) if (e[0] == '0') m.ok = 0;       // A/B switch for profiling
    std::memcpy(out.bytes, &m, sizeof(m));
    out.levels = m.ok;
#else
    (void)pyr; (void)n_pyramids; (void)L;
#endif
    return out;
}

void klt_track(const uint8_t *prev_pyr, const uint8_t *next_pyr, const int *prev_idx, const int *next_idx, int n,
               const PyrLayout &L, const float *prev_pts,
               float *next_pts, const int *npts, uint8_t *status, cudaStream_t st, const KltMapSet *maps_prev, const KltMapSet *maps_next) {
    dim3 grid(kMaxPts / KLT_WARPS, n);
    KltMaps mp, mn;
    std::memset(&mp, 0, sizeof(mp)); std::memset(&mn, 0, sizeof(mn));
#ifndef CUEMU
    if (maps_prev && maps_prev->levels) std::memcpy(&mp, maps_prev->bytes, sizeof(mp));
    if (maps_next && maps_next->levels) std::memcpy(&mn, maps_next->bytes, sizeof(mn));
#else
    (void)maps_prev; (void)maps_next;
#endif
    klt_kernel<<<grid, 32 * KLT_WARPS, 0, st>>>(prev_pyr, next_pyr, prev_idx, next_idx, L, prev_pts, next_pts, npts, status, mp, mn);
}

// ------------------------------------------------------------------------------------------------
// K5 Poisson-disk filter (poisson_disk_filter.h): one warp per list; the 25-cell neighbourhood scan
// (first cell skipped, one extra cell visited) is spread over the lanes.
// ------------------------------------------------------------------------------------------------
constexpr int PG_W = 96, PG_H = 64;      // grid cells available (cell = radius / sqrt 2): images up to ~1350 x 900 at r = 20
__global__ void __launch_bounds__(32) poisson_kernel(const double *__restrict__ pts, const int *__restrict__ n_preset,
                                                      const int *__restrict__ n_total,
                                                      const uint8_t *__restrict__ veto, uint8_t *__restrict__ keep,
                                                      double radius) {
    extern __shared__ short grid[];      // PG_W * PG_H, -1 = empty
    __shared__ double P[kMaxPts * 2];    // the list, staged once (the scan is strictly sequential)
    __shared__ short cell[kMaxPts * 2];  // grid cell of every point, computed up front by all lanes (two FP64 divisions per point)
    __shared__ uint8_t flag[kMaxPts];    // in: veto, out: keep -- the sequential loop touches shared memory only
    const int list = blockIdx.x, lane = threadIdx.x;
    for (int i = lane; i < PG_W * PG_H; i += 32) grid[i] = -1;
    const double r2 = radius * radius;
    const double gs = radius / sqrt(2.0);
    const int npre = n_preset[list], ntot = n_total[list];
    {
        const double *Pg = pts + (size_t)list * kMaxPts * 2;
        for (int i = lane; i < 2 * ntot; i += 32) P[i] = Pg[i];
        for (int i = lane; i < ntot; i += 32) {
            const double x = Pg[2 * i], y = Pg[2 * i + 1];
            const int ix = (int)floor(x / gs), iy = (int)floor(y / gs);
            cell[2 * i] = (short)min(max(ix, -4), PG_W + 4);          // clamped values stay outside the grid, like the originals
            cell[2 * i + 1] = (short)min(max(iy, -4), PG_H + 4);
            flag[i] = veto != nullptr ? veto[(size_t)list * kMaxPts + i] : 0;
        }
    }
    __syncwarp();
    for (int i = 0; i < ntot; ++i) {
        const double x = P[2 * i], y = P[2 * i + 1];
        const int ix = cell[2 * i], iy = cell[2 * i + 1];
        bool ok = true;
        if (i >= npre) {
            bool hit = false;
            if (lane < 25) {
                // visit order of test_point: all cells of the 5x5 block except (bx,by), plus (bx, ey + 1)
                int cx, cy;
                if (lane == 0) { cx = ix - 2; cy = iy + 3; }
                else { cx = ix - 2 + lane % 5; cy = iy - 2 + lane / 5; }
                if (cx >= 0 && cx < PG_W && cy >= 0 && cy < PG_H) {
                    int k = grid[cy * PG_W + cx];
                    if (k >= 0) {
                        double dx = x - P[2 * k], dy = y - P[2 * k + 1];
                        if (dx * dx + dy * dy < r2) hit = true;
                    }
                }
            }
            ok = !__any_sync(0xffffffffu, hit);
            if (flag[i]) ok = false;
        }
        __syncwarp();
        if (lane == 0) {
            flag[i] = ok ? 1 : 0;
            if (ok && ix >= 0 && ix < PG_W && iy >= 0 && iy < PG_H) grid[iy * PG_W + ix] = (short)i;
        }
        __syncwarp();
    }
    for (int i = lane; i < ntot; i += 32) keep[(size_t)list * kMaxPts + i] = flag[i];
}

void poisson_filter(const double *pts, const int *n_preset, const int *n_total, const uint8_t *veto, uint8_t *keep,
                    int n, double radius, cudaStream_t st) {
    static xv::SmemOptIn attr_done;
    size_t smem = (size_t)PG_W * PG_H * sizeof(short);
    xv::ensure_dynamic_smem(poisson_kernel, (int)smem, attr_done);
    poisson_kernel<<<n, 32, smem, st>>>(pts, n_preset, n_total, veto, keep, radius);
}

// ================================================================================================
// Undistortion: fixed-point bilinear remap with OpenCV's arithmetic (undistort_math.h)
// ================================================================================================
__global__ void __launch_bounds__(256) undistort_map_kernel(xund::Model m, int w, int h, int2 *__restrict__ map) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
    if (u >= w || v >= h) return;
    int qx, qy;
    xund::source_q5(m, u, v, &qx, &qy);
    map[(size_t)v * w + u] = make_int2(qx, qy);
}

__global__ void __launch_bounds__(256) undistort_kernel(const uint8_t *__restrict__ src, const int *__restrict__ sidx, size_t sstride,
                                                        uint8_t *__restrict__ dst, const int *__restrict__ didx, size_t dstride,
                                                        int w, int h, const int2 *__restrict__ map) {
    const int item = blockIdx.z;
    const uint8_t *s = src + (size_t)(sidx ? sidx[item] : item) * sstride;
    uint8_t *d = dst + (size_t)(didx ? didx[item] : item) * dstride;
    const int v = blockIdx.y;
    // four consecutive output pixels per thread: one 32-bit store, map reads as two 128-bit loads
    const int u0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (u0 >= w) return;
    const int2 *mrow = map + (size_t)v * w;
    if (u0 + 3 < w && (w & 3) == 0) {
        const int4 a = *reinterpret_cast<const int4 *>(mrow + u0), b = *reinterpret_cast<const int4 *>(mrow + u0 + 2);
        const unsigned p0 = xund::remap_q5(s, w, w, h, a.x, a.y), p1 = xund::remap_q5(s, w, w, h, a.z, a.w);
        const unsigned p2 = xund::remap_q5(s, w, w, h, b.x, b.y), p3 = xund::remap_q5(s, w, w, h, b.z, b.w);
        *reinterpret_cast<unsigned *>(d + (size_t)v * w + u0) = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
    } else {
        for (int u = u0; u < min(u0 + 4, w); ++u) d[(size_t)v * w + u] = xund::remap_q5(s, w, w, h, mrow[u].x, mrow[u].y);
    }
}

void undistort_build_map(const xund::Model &m, int w, int h, int2 *map, cudaStream_t st) {
    undistort_map_kernel<<<dim3((w + 255) / 256, h), 256, 0, st>>>(m, w, h, map);
}

void undistort(const uint8_t *src, const int *src_idx, size_t src_stride, uint8_t *dst, const int *dst_idx, size_t dst_stride,
               int n, int w, int h, const int2 *map, cudaStream_t st) {
    if (n <= 0) return;
    undistort_kernel<<<dim3(((w + 3) / 4 + 255) / 256, h, n), 256, 0, st>>>(src, src_idx, src_stride, dst, dst_idx, dst_stride, w, h, map);
}

}  // namespace xvfe
