// Large-window variant of the solve kernel (BASELINE.json config 5: 30 keyframes x 2000 landmarks, 58 000 reprojection
// residuals, 29 preintegration edges, dense 435 x 435 prior).  The trust-region loop, the factor evaluation and the
// J^T J assembly are the templated code of solve_core.cuh / backend_solve.cu; what differs from the sliding-window
// instance is where the reduced camera matrix lives and how it is reduced and factorised:
//   * H (packed lower triangle, 480 * 481 / 2 doubles = 0.9 MB) does not fit in shared memory: it lives in the problem's
//     global scratch and stays L2-resident (64 problems x 2 copies = 118 MB against 126 MB of L2);
//   * landmark Schur complement S = H_pp - W diag(1/ht) W^T (solver.cpp:179, marginalization_factor.h:383-398) is a true
//     contraction over the landmarks (180 x 2000 SYRK): FP64 tensor pipe, mma.sync.m8n8k4.f64 (tcgen05 has no FP64 kind);
//   * Cholesky of the 450 x 450 system: right-looking, panels of six columns, the trailing rank-6 update as 8 x 8 DMMA tiles
//     streamed over the L2-resident triangle (the register-tile factorisation of the small instance needs n <= 180).
#pragma once
#include "solve_core.cuh"

namespace xvbe {

constexpr int kMaxActiveLarge = XRVIO_MAX_ACTIVE_LARGE;
using SmemL = SmemT<kMaxActiveLarge, false>;

// ---- W tiles through the bulk-copy engine (north_star: "Hessian blocks staged through TMA into shared memory") ----------
// The Schur complement reads every row of W (6 * n_active rows, one double per landmark) once per 16 x 16 output block that
// touches it -- 12 times for a 30-keyframe window when the fragments come straight from L2.  Instead the CTA streams W through
// shared memory in tiles of 32 landmarks: one cp.async.bulk of 256 bytes per row (plus one for the 32 values 1 / ht_l),
// all signalling one mbarrier per buffer, double buffered so the copy of tile k + 1 runs under the tensor-pipe work on tile k.
// Rows are padded to 36 doubles in shared memory: the four row groups a half-warp reads in one fragment load then fall
// into distinct banks.  A warp carries at most three output blocks (24 accumulator registers), so windows with more than
// 48 blocks take two passes over the tiles: every W element leaves L2 at most twice.
constexpr int WT_BLOCKS = 3;                       // output blocks per warp and pass
// ROWS: pose rows (6 per free frame); TL: landmarks per tile; PITCH: doubles per staged row, chosen so that PITCH * 2 mod 32 = 8:
// the four row groups of a half-warp's fragment load (rows g, columns 4u + t) then hit 16 distinct banks.
template <int ROWS, int TL, int PITCH> struct alignas(128) WStageT {
    static constexpr int kRows = ROWS, kL = TL, kPitch = PITCH;
    double w[2][ROWS][PITCH];
    double ih[2][TL];
    unsigned long long bar[2];
};
using WStage = WStageT<6 * kMaxActiveLarge, 32, 36>;     // large windows: 2 x 55 KB
using WStageS = WStageT<6 * kMaxActive, 16, 20>;         // sliding windows: 2 x 11.5 KB next to the 130 KB packed H
template <class SM> struct WStageOf { using type = WStage; };
template <> struct WStageOf<Smem> { using type = WStageS; };
template <class SM> struct SmemCore { static constexpr size_t bytes = (sizeof(SM) + 127) / 128 * 128; };      // the stage follows SM in the dynamic shared memory
constexpr size_t kSolveLargeSmem = SmemCore<SmemL>::bytes + sizeof(WStage);
constexpr size_t kSolveSmem = SmemCore<Smem>::bytes + sizeof(WStageS);

#ifndef CUEMU
__device__ __forceinline__ unsigned xv_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void xv_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xv_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void xv_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xv_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void xv_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(xv_smem_u32(dst)), "l"(src), "r"(bytes), "r"(xv_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void xv_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}"
                 ::"r"(xv_smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// warp 0: start the copy of the tile of landmarks [l0, l0 + 32) into buffer `buf`
// (W is stored tile-major with the stage's row pitch, solve_core.cuh: the tile is one contiguous block of R * kPitch doubles)
template <class WS> __device__ __forceinline__ void wstage_issue(WS &ws, int buf, const double *W, const double *ihv, int R, int tile, int lane) {
    constexpr int TL = WS::kL, PITCH = WS::kPitch;
    const double *src = W + (size_t)tile * R * PITCH;
#ifndef CUEMU
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const unsigned wb = (unsigned)(R * PITCH * sizeof(double)), ib = (unsigned)(TL * sizeof(double));
        xv_mbar_expect_tx(&ws.bar[buf], wb + ib);
        xv_bulk_g2s(&ws.w[buf][0][0], src, wb, &ws.bar[buf]);
        xv_bulk_g2s(&ws.ih[buf][0], ihv + (size_t)tile * TL, ib, &ws.bar[buf]);
    }
#else
    for (int e = lane; e < R * PITCH; e += 32) (&ws.w[buf][0][0])[e] = src[e];
    if (lane < TL) ws.ih[buf][lane] = ihv[(size_t)tile * TL + lane];
#endif
}

// Same contract as schur_landmarks() (solve_core.cuh).  The 6 * n_active pose rows of W are cut into 16 x 16 blocks of the
// lower triangle, four m8n8k4 accumulators per block; the contraction runs over the landmarks four at a time on the
// FP64 tensor pipe, operands from the staged tiles.  Every output element is owned by exactly one lane.
template <class SM> static __device__ void schur_landmarks_mma(const Prob &P, SM &s, double *rhs) {
    XV_ASSUME_SHARED(&s); XV_ASSUME_SHARED(rhs);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = P.M, na = P.d.n_active, Wp = P.Wp;
    if (M <= 0) return;
    using WS = typename WStageOf<SM>::type;
    constexpr int WT_L = WS::kL, WT_PITCH = WS::kPitch;
    WS &ws = *reinterpret_cast<WS *>(reinterpret_cast<unsigned char *>(&s) + SmemCore<SM>::bytes);
    XV_ASSUME_SHARED(&ws);
    const int R = 6 * na;
    const int nb = (R + 15) / 16;
    const int nblk = nb * (nb + 1) / 2;
    const int g = lane >> 2, t = lane & 3;
    // 1 / ht_l once per landmark (non-finite -> 0: the landmark is skipped, marginalization_factor.h:384-386), 0 in the padding
    // columns; y_l is free until the back-substitution at the end of gauss_newton()
    double *ihv = P.y_l;
    for (int l = tid; l < Wp; l += ST) { double ih = l < M ? 1.0 / P.ht_l[l] : 0.0; if (!isfinite(ih)) ih = 0.0; ihv[l] = ih; }
#ifndef CUEMU
    if (tid == 0) { xv_mbar_init(&ws.bar[0], 1); xv_mbar_init(&ws.bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#endif
    __syncthreads();
    const int T = Wp / WT_L;
    unsigned uses0 = 0, uses1 = 0;                     // completed phases per buffer (parity of the next wait)
    for (int pass = 0; pass * NW * WT_BLOCKS < nblk; ++pass) {
        int bi[WT_BLOCKS], bj[WT_BLOCKS];
        double acc[WT_BLOCKS][2][2][2];
#pragma unroll
        for (int q = 0; q < WT_BLOCKS; ++q) {
            const int blk = warp + NW * (WT_BLOCKS * pass + q);
            int i = -1, j = -1;
            if (blk < nblk) {
                i = (int)((sqrt(8.0 * blk + 1.0) - 1.0) * 0.5);
                while (i * (i + 1) / 2 > blk) --i;
                while ((i + 1) * (i + 2) / 2 <= blk) ++i;
                j = blk - i * (i + 1) / 2;             // i >= j
            }
            bi[q] = i; bj[q] = j;
#pragma unroll
            for (int e = 0; e < 8; ++e) (&acc[q][0][0][0])[e] = 0.0;
        }
        if (warp == 0) wstage_issue(ws, 0, P.W, ihv, R, 0, lane);
        for (int k = 0; k < T; ++k) {
            const int buf = k & 1;
            if (warp == 0 && k + 1 < T) wstage_issue(ws, buf ^ 1, P.W, ihv, R, k + 1, lane);     // its readers passed the barrier below
#ifndef CUEMU
            xv_mbar_wait(&ws.bar[buf], (buf ? uses1 : uses0) & 1u);
#else
            __syncthreads();
#endif
            if (buf) ++uses1; else ++uses0;
            const double(*wt)[WT_PITCH] = ws.w[buf];
            const double *iht = ws.ih[buf];
#pragma unroll
            for (int q = 0; q < WT_BLOCKS; ++q) {
                if (bi[q] < 0) continue;                // warp-uniform
                const int ra0 = bi[q] * 16 + g, ra1 = ra0 + 8;      // rows of the two A fragments
                const int cb0 = bj[q] * 16 + g, cb1 = cb0 + 8;      // columns of the two B fragments
#pragma unroll
                for (int u = 0; u < WT_L / 4; ++u) {
                    const int l = 4 * u + t;
                    const double ih = iht[l];
                    const double a0 = ra0 < R ? wt[ra0][l] : 0.0, a1 = ra1 < R ? wt[ra1][l] : 0.0;
                    const double b0 = cb0 < R ? wt[cb0][l] * ih : 0.0, b1 = cb1 < R ? wt[cb1][l] * ih : 0.0;
                    dmma_m8n8k4(acc[q][0][0][0], acc[q][0][0][1], a0, b0);
                    dmma_m8n8k4(acc[q][0][1][0], acc[q][0][1][1], a0, b1);
                    dmma_m8n8k4(acc[q][1][0][0], acc[q][1][0][1], a1, b0);
                    dmma_m8n8k4(acc[q][1][1][0], acc[q][1][1][1], a1, b1);
                }
            }
            __syncthreads();                            // buffer `buf` may be refilled (tile k + 2) from the next iteration on
        }
#pragma unroll
        for (int q = 0; q < WT_BLOCKS; ++q) {
            if (bi[q] < 0) continue;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = bi[q] * 16 + rr * 8 + g, c = bj[q] * 16 + cc * 8 + 2 * t + i;
                        if (r >= R || c > r) continue;
                        const int fa = r / 6, fb = c / 6;
                        if (!s.pose_free[fa] || !s.pose_free[fb]) continue;
                        s.H[hidx(15 * fa + r % 6, 15 * fb + c % 6)] -= acc[q][rr][cc][i];
                    }
        }
    }
    // b_red = gs_p - sum_l w_l gs_l / ht_l
    for (int c = warp; c < R; c += NW) {
        const int a = c / 6, k = c % 6;
        if (!s.pose_free[a]) continue;
        double v = 0.0;
        for (int l = lane; l < M; l += 32) v += P.W[widx(P, c, l)] * P.g_l[l] * ihv[l];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) rhs[15 * a + k] -= v;
    }
}

// In-place right-looking Cholesky of the packed lower triangle s.H (global memory), panels of six columns.  Per panel:
// (1) thread 0 factorises the 6 x 6 diagonal block, (2) one thread per row below it solves the row against that block and
// publishes the six scaled columns (s.colp, also written back), (3) the trailing triangle takes the rank-6 update on the
// tensor pipe.  s.flag = 0 on a non-positive or non-finite pivot
// (-> LINEAR_SOLVER_FAILURE, the caller raises mu); 1 / L_jj is left in idiag[] for the substitutions.
template <class SM> static __device__ void cholesky_global(SM &s, int n, double *idiag) {
    XV_ASSUME_SHARED(&s); XV_ASSUME_SHARED(idiag);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int TB = 6;
    double *colp = s.colp;                               // colp[c * NPv + row]
    for (int j0 = 0; j0 < n; j0 += TB) {
        const int j1 = min(j0 + TB, n), w = j1 - j0;
        if (tid == 0) {
            double a[TB][TB];
            for (int i = 0; i < w; ++i)
                for (int k = 0; k <= i; ++k) a[i][k] = s.H[hidx(j0 + i, j0 + k)];
            for (int c = 0; c < w && s.flag; ++c) {
                const double d = a[c][c];
                if (!(d > 0.0) || !isfinite(d)) { s.flag = 0; break; }
                const double inv = rsqrt(d);
                idiag[j0 + c] = inv;
                a[c][c] = d * inv;
                for (int i = c + 1; i < w; ++i) a[i][c] *= inv;
                for (int i = c + 1; i < w; ++i)
                    for (int k = c + 1; k <= i; ++k) a[i][k] -= a[i][c] * a[k][c];
            }
            if (s.flag)
                for (int i = 0; i < w; ++i)
                    for (int k = 0; k <= i; ++k) { s.H[hidx(j0 + i, j0 + k)] = a[i][k]; colp[k * SM::NPv + j0 + i] = a[i][k]; }
        }
        __syncthreads();
        if (!s.flag) return;                             // uniform
        // (2) rows below the panel: L(i, j0 + c) = (a(i, c) - sum_{k < c} L(i, j0 + k) L(j0 + c, j0 + k)) / L(j0 + c, j0 + c)
        for (int i = j1 + tid; i < n; i += ST) {
            double v[TB];
            for (int c = 0; c < w; ++c) v[c] = s.H[hidx(i, j0 + c)];
            for (int c = 0; c < w; ++c) {
                v[c] *= idiag[j0 + c];
                for (int k = c + 1; k < w; ++k) v[k] -= v[c] * colp[c * SM::NPv + j0 + k];
            }
            for (int c = 0; c < w; ++c) { s.H[hidx(i, j0 + c)] = v[c]; colp[c * SM::NPv + i] = v[c]; }
        }
        __syncthreads();
        // (3) trailing triangle: A -= Lp Lp^T with Lp the (n - j1) x 6 panel just published -- a rank-6 contraction, on the FP64
        // tensor pipe as 8 x 8 tiles (k padded to 8 with zeros: two m8n8k4 steps).  One warp per row of tiles: its A fragment
        // (-Lp rows) is loaded once, per tile the B fragment comes from shared memory and the C tile from the packed rows of
        // H (two consecutive doubles per lane).  -DXRVIO_LARGE_CHOL_SCALAR keeps the CUDA-core version for A/B timing.
#ifndef XRVIO_LARGE_CHOL_SCALAR
        {
            const int g = lane >> 2, t = lane & 3;
            const int T = (n - j1 + 7) / 8;
            for (int ti = warp; ti < T; ti += NW) {
                const int r = j1 + 8 * ti + g;
                double a0 = 0.0, a1 = 0.0;
                if (r < n) {
                    if (t < w) a0 = -colp[t * SM::NPv + r];
                    if (4 + t < w) a1 = -colp[(4 + t) * SM::NPv + r];
                }
                double *row = &s.H[hidx(r < n ? r : n - 1, 0)];
                // four tiles per pass: loads of the C tiles and B fragments first, then the tensor-pipe instructions, then the
                // stores (spelled out so that the L2 round trips of independent tiles overlap)
                for (int tk0 = 0; tk0 <= ti; tk0 += 4) {
                    double d0[4], d1[4], b0[4], b1[4];
                    bool v0[4], v1[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int tk = tk0 + u;
                        const int cb = j1 + 8 * tk + g, c = j1 + 8 * tk + 2 * t;
                        b0[u] = b1[u] = 0.0;
                        if (tk <= ti && cb < n) {
                            if (t < w) b0[u] = colp[t * SM::NPv + cb];
                            if (4 + t < w) b1[u] = colp[(4 + t) * SM::NPv + cb];
                        }
                        v0[u] = tk <= ti && r < n && c <= r;
                        v1[u] = tk <= ti && r < n && c + 1 <= r;
                        d0[u] = v0[u] ? row[c] : 0.0;
                        d1[u] = v1[u] ? row[c + 1] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        dmma_m8n8k4(d0[u], d1[u], a0, b0[u]);
                        dmma_m8n8k4(d0[u], d1[u], a1, b1[u]);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int c = j1 + 8 * (tk0 + u) + 2 * t;
                        if (v0[u]) row[c] = d0[u];
                        if (v1[u]) row[c + 1] = d1[u];
                    }
                }
            }
        }
#else
        for (int i = j1 + warp; i < n; i += NW) {
            double ci[TB];
            for (int c = 0; c < TB; ++c) ci[c] = c < w ? colp[c * SM::NPv + i] : 0.0;
            double *row = &s.H[hidx(i, 0)];
            for (int k = j1 + lane; k <= i; k += 32) {
                double a = row[k];
#pragma unroll
                for (int c = 0; c < TB; ++c)
                    if (c < w) a -= ci[c] * colp[c * SM::NPv + k];
                row[k] = a;
            }
        }
#endif
        __syncthreads();
    }
}

}  // namespace xvbe
