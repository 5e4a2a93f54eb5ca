// K7-K11 driver: the Ceres-1.14-style trust-region loop around solve_core.cuh (see there).
#include "solve_core.cuh"
#include "solve_large.cuh"
#include "launch_attr.cuh"

namespace xvbe {
using namespace xv;

size_t solve_scratch_doubles(int n_active, int n_lm_free, int n_pre, int marg_k, int n_rpe) {
    size_t M = (size_t)((n_lm_free + 31) / 32) * 32 + 32;
    size_t n = (size_t)15 * marg_k;
    return 8 * M + (size_t)n_active * 6 * (M + M / 4) + HPACK + (size_t)n_pre * kPreScratch + 4 * (n + 32) + 2 * n * n + 64 + (size_t)6 * n_rpe + 2 * M + jb_doubles(n_rpe)
           + kSpec * M + (size_t)16 * n_pre * kSpec;      // inverse depths / preintegration residuals of the speculative candidates
}


// Candidates b = 0 .. nb-1 of a batch (evaluate_cost_multi): q <- normalize(q * exp(d)) for free pose blocks, additive elsewhere
// (quaternion_parameterization.h:11-17), with delta_j = c1 du_j + c2 dw_j, du = -gs / D^2 * scale, dw = -y * scale (prepared per
// linearisation) and (c1, c2) of candidate b in s.sp.  s.red4[(kSpec + b) * NW + warp] receives the warp's part of
// |x - candidate_b|^2 over the free blocks (ParameterToleranceReached), summed with the costs in evaluate_cost_multi.
template <class SM> __device__ void make_candidates(const Prob &P, SM &s, int nb) {
    XV_ASSUME_SHARED(&s);
    const int tid = threadIdx.x;
    double ss[kSpec];
#pragma unroll
    for (int k = 0; k < kSpec; ++k) ss[k] = 0.0;
    const int nf0 = P.d.n_frames, na = P.d.n_active;
    for (int t = tid; t < nf0 + (nb - 1) * na; t += ST) {
        const int b = t < nf0 ? 0 : 1 + (t - nf0) / na;
        const int f = t < nf0 ? t : (t - nf0) % na;
        const double c1 = s.sp[b * 8 + 1], c2 = s.sp[b * 8 + 2];
        double *c = cand_frame_w(s, b, f);
        const double *x = s.xs + f * 16;
        for (int k = 0; k < 16; ++k) c[k] = x[k];
        if (f < na) {
            double dl[15];
            for (int k = 0; k < 15; ++k) { const int j = 15 * f + k; dl[k] = c1 * s.du[j] + c2 * s.dw[j]; }
            double a = 0.0;
            if (s.pose_free[f]) {
                st4(c, qnormalized(qmul(ld4(x), expmap(v3(dl[0], dl[1], dl[2])))));
                for (int k = 0; k < 3; ++k) c[4 + k] = x[4 + k] + dl[3 + k];
                for (int k = 0; k < 7; ++k) { const double d = x[k] - c[k]; a += d * d; }
            }
            if (s.motion_free[f]) {
                for (int k = 0; k < 9; ++k) c[7 + k] = x[7 + k] + dl[6 + k];
                for (int k = 7; k < 16; ++k) { const double d = x[k] - c[k]; a += d * d; }
            }
#pragma unroll
            for (int k = 0; k < kSpec; ++k) if (k == b) ss[k] += a;
        }
    }
    for (int l = tid; l < P.M; l += ST) {
        const double x = P.inv_x[l], du = P.du_l[l], dw = P.dw_l[l];
#pragma unroll
        for (int b = 0; b < kSpec; ++b) {
            if (b >= nb) continue;
            const double c = x + (s.sp[b * 8 + 1] * du + s.sp[b * 8 + 2] * dw);
            P.inv_c[(size_t)b * P.Mp + l] = c;
            const double d = x - c;
            ss[b] += d * d;
        }
    }
#pragma unroll
    for (int b = 0; b < kSpec; ++b) {
        if (b >= nb) continue;
        double v = ss[b];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) s.red4[(kSpec + b) * NW + (tid >> 5)] = v;
    }
    __syncthreads();
}

template <class SM> __device__ double ambient_sq(const Prob &P, SM &s, const double *A, const double *B, const double *ia, const double *ib) {
    // squared norm of (A - B) over the free blocks (B == nullptr: norm of A)
    double a = 0.0;
    for (int f = threadIdx.x; f < P.d.n_active; f += ST) {
        if (s.pose_free[f])
            for (int k = 0; k < 7; ++k) { double d = A[f * 16 + k] - (B ? B[f * 16 + k] : 0.0); a += d * d; }
        if (s.motion_free[f])
            for (int k = 7; k < 16; ++k) { double d = A[f * 16 + k] - (B ? B[f * 16 + k] : 0.0); a += d * d; }
    }
    for (int l = threadIdx.x; l < P.M; l += ST) { double d = ia[l] - (ib ? ib[l] : 0.0); a += d * d; }
    return block_sum(a, s);
}

// Gauss-Newton step: (Hs + mu D^2) y = gs via Schur complement on the landmark block + Cholesky.
// Returns false when the factorisation fails (LINEAR_SOLVER_FAILURE).  H is consumed.
struct alignas(16) Pair64 { double a, b; };      // one 128-bit shared-memory access

template <class SM> __device__ bool gauss_newton(const Prob &P, SM &s, double mu, KProf &kp) {
    XV_ASSUME_SHARED(&s);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = P.NPc, M = P.M, na = P.d.n_active;
    for (int c = tid; c < n; c += ST)
        if (s.col_free[c]) s.H[hidx(c, c)] += mu * s.D[c] * s.D[c];
    for (int l = tid; l < M; l += ST) P.ht_l[l] = P.h_ll[l] + mu * P.D_l[l] * P.D_l[l];
    for (int c = tid; c < n; c += ST) s.bred[c] = s.g[c];
    __syncthreads();
    kp.mark(29); kp.count(27);
    // landmark Schur complement on the FP64 tensor pipe, W streamed through shared memory by the bulk-copy engine (both instances;
    // the scalar schur_landmarks of solve_core.cuh remains the marginalisation kernel's, which has no room for the stage)
    schur_landmarks_mma(P, s, s.bred);
    __syncthreads();
    kp.mark(24);
    // in-place right-looking Cholesky on the packed lower triangle.  The scaled pivot column is mirrored
    // into a contiguous vector (s.du, free at this point) so the trailing update reads it conflict-free;
    // 1/L_jj is kept in s.dw for the substitutions.
    if (tid == 0) s.flag = 1;
    __syncthreads();
    double *col = s.du, *idiag = s.dw;
    // Register-resident, panel-blocked Cholesky: the lower triangle is cut into 6x6 tiles, one tile per thread (n <= 180 ->
    // <= 465 tiles), held in registers for the whole factorisation.  Per PANEL of six columns (three barriers instead of
    // twelve): (1) the owner of the diagonal tile factorises it in registers and publishes its columns, (2) the owners of
    // the tiles below it scale their six columns against that tile and publish them, (3) every tile to the right applies
    // the six rank-1 updates from shared memory.  Every element sees the same operations in the same order as in a
    // column-by-column right-looking factorisation, so the factor is bit-identical to it.
    if constexpr (!SM::kSharedH) cholesky_global(s, n, idiag);
    else {
        constexpr int TB = 6;
        const int nt = (n + TB - 1) / TB;
        int ti = -1, tk = -1;
        if (tid < nt * (nt + 1) / 2) {
            ti = (int)((sqrt(8.0 * tid + 1.0) - 1.0) * 0.5);
            while (ti * (ti + 1) / 2 > tid) --ti;
            while ((ti + 1) * (ti + 2) / 2 <= tid) ++ti;
            tk = tid - ti * (ti + 1) / 2;
        }
        double a[TB][TB];
#pragma unroll
        for (int li = 0; li < TB; ++li)
#pragma unroll
            for (int lk = 0; lk < TB; ++lk) {
                const int gi = ti * TB + li, gk = tk * TB + lk;
                a[li][lk] = (ti >= 0 && gi < n && gk <= gi) ? s.H[hidx(gi, gk)] : 0.0;
            }
        double *colp = s.colp;                       // colp[c * SM::NPv + row]: column (panel * 6 + c), scaled
        __syncthreads();
        for (int p = 0; p < nt; ++p) {
            const int j0 = p * TB;
            // ---- (1) diagonal tile ----
            if (ti == p && tk == p) {
#pragma unroll
                for (int c = 0; c < TB; ++c) {
                    const int j = j0 + c;
                    if (j < n && s.flag) {
                        const double d = a[c][c];
                        if (!(d > 0.0) || !isfinite(d)) s.flag = 0;
                        else {
                            const double inv = rsqrt(d);
                            idiag[j] = inv;
#pragma unroll
                            for (int li = 0; li < TB; ++li) {
                                if (li < c) continue;
                                const int gi = j0 + li;
                                double v = 0.0;
                                if (gi < n) { v = (li == c) ? d * inv : a[li][c] * inv; a[li][c] = v; }
                                colp[c * SM::NPv + gi] = v;
                            }
#pragma unroll
                            for (int li = 0; li < TB; ++li)
#pragma unroll
                                for (int lk = 0; lk < TB; ++lk)
                                    if (li > c && lk > c && lk <= li) a[li][lk] -= a[li][c] * a[lk][c];
                        }
                    }
                }
            }
            __syncthreads();
            if (!s.flag) break;                      // uniform: written before the barrier by one thread
            // ---- (2) tiles below the diagonal tile ----
            if (tk == p && ti > p) {
#pragma unroll
                for (int c = 0; c < TB; ++c) {
                    const int j = j0 + c;
                    if (j >= n) continue;
                    const double inv = idiag[j];
#pragma unroll
                    for (int li = 0; li < TB; ++li) {
                        const int gi = ti * TB + li;
                        double v = 0.0;
                        if (gi < n) { v = a[li][c] * inv; a[li][c] = v; }
                        colp[c * SM::NPv + gi] = v;
                    }
#pragma unroll
                    for (int lk = 0; lk < TB; ++lk) {
                        if (lk <= c) continue;
                        const double ck = (j0 + lk < n) ? colp[c * SM::NPv + j0 + lk] : 0.0;      // L_pp(lk, c), published in (1)
#pragma unroll
                        for (int li = 0; li < TB; ++li) a[li][lk] -= a[li][c] * ck;
                    }
                }
            }
            __syncthreads();
            // ---- (3) tiles to the right of the panel ----
            if (tk > p) {
#pragma unroll
                for (int c = 0; c < TB; ++c) {
                    if (j0 + c >= n) continue;
                    // the six published values of a tile row / tile column are 48 contiguous, 16-byte aligned bytes: three 128-bit
                    // loads each.  Lanes of a warp hold consecutive tile columns (stride 48 bytes): as 128-bit accesses that is
                    // conflict-free, as six 64-bit loads it was an 8-way bank conflict and made this update shared-memory bound
                    // (measured 9 kcycles per panel).  Rows beyond n hold zeros (written by phases 1 and 2), so no bound checks.
                    double ci[TB], ck[TB];
                    {
                        const Pair64 *pi = reinterpret_cast<const Pair64 *>(colp + c * SM::NPv + ti * TB);
                        const Pair64 *pk = reinterpret_cast<const Pair64 *>(colp + c * SM::NPv + tk * TB);
#pragma unroll
                        for (int l = 0; l < TB / 2; ++l) {
                            const Pair64 vi = pi[l], vk = pk[l];
                            ci[2 * l] = vi.a; ci[2 * l + 1] = vi.b; ck[2 * l] = vk.a; ck[2 * l + 1] = vk.b;
                        }
                    }
#pragma unroll
                    for (int li = 0; li < TB; ++li)
#pragma unroll
                        for (int lk = 0; lk < TB; ++lk) a[li][lk] -= ci[li] * ck[lk];
                }
            }
            __syncthreads();                         // colp is rewritten by the next panel
        }
        __syncthreads();
        if (s.flag) {
#pragma unroll
            for (int li = 0; li < TB; ++li)
#pragma unroll
                for (int lk = 0; lk < TB; ++lk) {
                    const int gi = ti * TB + li, gk = tk * TB + lk;
                    if (ti >= 0 && gi < n && gk <= gi) s.H[hidx(gi, gk)] = a[li][lk];
                }
        }
        __syncthreads();
        // inverses of the diagonal blocks for the substitutions, all panels at once (one thread per panel, off the factorisation's
        // critical path): N = L_pp^-1, column by column; rows / columns beyond n act as the identity
        if (s.flag && tid < nt) {
            const int p = tid, j0 = p * TB;
            double Lb[TB][TB], N[TB][TB];
#pragma unroll
            for (int i = 0; i < TB; ++i)
#pragma unroll
                for (int k = 0; k < TB; ++k) Lb[i][k] = (k < i && j0 + i < n) ? s.H[hidx(j0 + i, j0 + k)] : 0.0;
#pragma unroll
            for (int j = 0; j < TB; ++j) {
#pragma unroll
                for (int i = 0; i < TB; ++i) {
                    if (i < j) { N[i][j] = 0.0; continue; }
                    const bool in = j0 + i < n && j0 + j < n;
                    if (i == j) { N[i][j] = in ? idiag[j0 + i] : 1.0; continue; }
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < TB; ++k)
                        if (k >= j && k < i) acc += Lb[i][k] * N[k][j];
                    N[i][j] = in ? -acc * idiag[j0 + i] : 0.0;
                }
            }
#pragma unroll
            for (int i = 0; i < TB; ++i)
#pragma unroll
                for (int j = 0; j < TB; ++j) s.linv[p * 36 + i * 6 + j] = N[i][j];
        }
    }
    __syncthreads();
    kp.mark(25);
    if (!s.flag) { kp.count(28); return false; }
    // (A shared-memory variant -- panels of eight, H left in shared memory, trailing update as 8 x 8 mma.sync.m8n8k4 tiles on the
    // packed rows, tools/chol_shared8.patch -- measured 347 kcycles per call against 210 for the register tiles above: its diagonal
    // block work on one thread and the scattered fragment accesses to the packed triangle cost more than the tensor pipe saves.)
    // forward / backward substitution (L z = b, L^T y = z) in panels of six columns, two barriers per panel.  Sliding-window
    // instance: the panel's six values are a 6 x 6 product with the inverse of the diagonal block (six threads, no dependent
    // chain; the inverses were computed by the factorisation); large instance: thread 0 solves the 6 x 6 triangle.  Then every
    // row outside the panel applies the six values in column order.  (A single-warp version with the work vector in registers
    // and a shuffle broadcast per column measured 1.75x SLOWER on a B200: 241 against 137 kcycles per call.)
    // forward: work vector s.y (starts as b_red), result z in s.bred; backward: work vector s.bred, result in s.y.
    for (int c = tid; c < n; c += ST) s.y[c] = s.bred[c];
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += 6) {
        const int j1 = min(j0 + 6, n);
        if constexpr (SM::kSharedH) {
            if (tid < 6 && j0 + tid < n) {               // z_p = L_pp^-1 y_p
                const double *N = s.linv + (j0 / 6) * 36 + tid * 6;
                double z = 0.0;
                for (int k = 0; k <= tid; ++k) z += N[k] * s.y[j0 + k];
                s.sc[tid] = z;
                s.bred[j0 + tid] = z;
            }
        } else if (tid == 0) {
            for (int j = j0; j < j1; ++j) {
                const double zj = s.y[j] * idiag[j];
                s.sc[j - j0] = zj;
                s.bred[j] = zj;
                for (int i = j + 1; i < j1; ++i) s.y[i] -= s.H[hidx(i, j)] * zj;
            }
        }
        __syncthreads();
        for (int i = j1 + tid; i < n; i += ST) {
            double yi = s.y[i];
            for (int j = j0; j < j1; ++j) yi -= s.H[hidx(i, j)] * s.sc[j - j0];
            s.y[i] = yi;
        }
        __syncthreads();
    }
    for (int j0 = ((n - 1) / 6) * 6; j0 >= 0; j0 -= 6) {       // the panels of the factorisation (multiples of six), last first
        const int j1 = min(j0 + 6, n);
        if constexpr (SM::kSharedH) {
            if (tid < 6 && j0 + tid < j1) {              // y_p = L_pp^-T b_p
                const double *N = s.linv + (j0 / 6) * 36;
                double yv = 0.0;
                for (int k = tid; k < j1 - j0; ++k) yv += N[k * 6 + tid] * s.bred[j0 + k];
                s.sc[tid] = yv;
                s.y[j0 + tid] = yv;
            }
        } else if (tid == 0) {
            for (int j = j1 - 1; j >= j0; --j) {
                const double yj = s.bred[j] * idiag[j];
                s.sc[j - j0] = yj;
                s.y[j] = yj;
                for (int i = j0; i < j; ++i) s.bred[i] -= s.H[hidx(j, i)] * yj;
            }
        }
        __syncthreads();
        for (int i = tid; i < j0; i += ST) {
            double bi = s.bred[i];
            for (int j = j1 - 1; j >= j0; --j) bi -= s.H[hidx(j, i)] * s.sc[j - j0];
            s.bred[i] = bi;
        }
        __syncthreads();
    }
    kp.mark(26);
    bool ok = true;
    for (int c = tid; c < n; c += ST) ok = ok && isfinite(s.y[c]);
    // back-substitute the landmarks: y_l = (gs_l - w_l . y_p) / ht_l
    for (int l = tid; l < M; l += ST) {
        double a = P.g_l[l];
        for (int f = 0; f < na; ++f) {
            if (!s.pose_free[f]) continue;
            for (int k = 0; k < 6; ++k) a -= P.W[widx(P, f * 6 + k, l)] * s.y[15 * f + k];
        }
        a /= P.ht_l[l];
        P.y_l[l] = a;
        ok = ok && isfinite(a);
    }
    const int allok = __syncthreads_and(ok ? 1 : 0);
    return allok != 0;
}

template <class SM> __device__ void solve_body(const SolveBatch &b, const RigParams *__restrict__ rig_g, SM &s, RigParams &rg) {
    XV_ASSUME_SHARED(&s); XV_ASSUME_SHARED(&rg);
    const int tid = threadIdx.x;
    if (tid == 0) rg = *rig_g;
    Prob P;
    P.d = b.desc[blockIdx.x];
    P.states_g = b.states + (size_t)P.d.frame_off * 16;
    P.flags_g = b.frame_flags + P.d.frame_off;
    P.invd_user = b.inv_depth + P.d.lm_off;
    P.rpe = b.rpe; P.rot = b.rot; P.pre = b.pre; P.pre_blocks = b.pre_blocks; P.pre_out = b.pre_out; P.marg_frames = b.marg_frames;
    P.M = P.d.n_lm_free;
    P.Wp = (P.M + 31) / 32 * 32; P.W_sh = WLayout<SM::kSharedH>::sh; P.W_pitch = WLayout<SM::kSharedH>::pitch; P.W_rows = 6 * P.d.n_active;
    P.NPc = 15 * P.d.n_active;
    P.marg_n = 15 * P.d.marg_k;
    {
        const double *m = b.marg + P.d.marg_off;
        const size_t n = P.marg_n;
        P.marg_lin = m; P.marg_S = m + 16 * P.d.marg_k; P.marg_info = P.marg_S + n * n; P.marg_A0 = P.marg_info + n;
        double *sc = b.scratch + P.d.scratch_off;
        const size_t Mp = (size_t)((P.M + 31) / 32) * 32 + 32;
        P.inv_x = sc; sc += Mp; P.inv_c = sc; sc += Mp; P.h_ll = sc; sc += Mp; P.g_l = sc; sc += Mp;
        P.sc_l = sc; sc += Mp; P.D_l = sc; sc += Mp; P.y_l = sc; sc += Mp; P.ht_l = sc; sc += Mp;
        P.W = sc; sc += (size_t)P.d.n_active * 6 * (Mp + Mp / 4);       // >= w_doubles(): room for the row padding of either layout
        if constexpr (!SM::kSharedH) { if (tid == 0) s.H.a = sc; sc += SM::HP; }     // published by the barrier in prepare_tangents()
        P.Hcopy = sc; sc += SM::HP;
        P.pre_s = sc; sc += (size_t)P.d.n_pre * kPreScratch;
        P.m_dx = sc; sc += n + 32; P.m_r = sc; sc += n + 32; P.m_v = sc; sc += n + 32; sc += n + 32;
        P.m_T = sc; sc += n * n;
        prepare_prior(P, m, sc); sc += n * n;
        prepare_tangents(P, sc); sc += (size_t)6 * P.d.n_rpe;
        P.du_l = sc; sc += Mp; P.dw_l = sc; sc += Mp;
        P.Mp = Mp;
        P.inv_c = sc; sc += kSpec * Mp;                                  // the candidates' inverse depths, Mp apart
        P.pre_r = sc; sc += (size_t)16 * P.d.n_pre * kSpec;
        P.JB = nullptr; P.grp_start = nullptr; P.n_grp = 0; P.NJ = 0;
        if constexpr (SM::kSharedH) prepare_groups(P, s, sc);
    }
    SolveOut &out = b.out[blockIdx.x];
    // ---- load user memory ----
    for (int i = tid; i < P.d.n_frames * 16; i += ST) { const double v = P.states_g[i]; s.us[i] = v; s.xs[i] = v; s.cs[i] = v; }
    for (int f = tid; f < SM::kA; f += ST) {
        const unsigned char fl = f < P.d.n_active ? P.flags_g[f] : 0;
        s.pose_free[f] = fl & 1; s.motion_free[f] = (fl >> 1) & 1;
    }
    __syncthreads();
    for (int c = tid; c < SM::NPv; c += ST) {
        const int f = c / 15, k = c % 15;
        s.col_free[c] = (c < P.NPc) && (k < 6 ? s.pose_free[f] : s.motion_free[f]);
        s.scale[c] = 1.0; s.D[c] = 1.0; s.y[c] = 0.0;
    }
    for (int l = tid; l < P.M; l += ST) { P.inv_x[l] = P.invd_user[l]; P.inv_c[l] = P.invd_user[l]; P.y_l[l] = 0.0; P.sc_l[l] = 1.0; P.D_l[l] = 1.0; }
    __syncthreads();
    int nfree = 0;
    for (int c = 0; c < P.NPc; ++c) nfree += s.col_free[c];
    nfree += P.M;

    KProf kp; kp.start(P.d.marg_k > 0);
    kp.count(31);
    // ---- iteration zero ----
    double x_cost = evaluate(P, s, s.xs, P.inv_x, rg, true, true, &kp);
    if (nfree == 0) {
        if (tid == 0) { out.iterations = 1; out.termination = 6; out.initial_cost = out.final_cost = x_cost; out.cost_trace[0] = x_cost; out.n_successful = out.n_unsuccessful = 0; }
        return;
    }
    double gmax;
    {
        double m = 0.0;
        for (int c = tid; c < P.NPc; c += ST) m = fmax(m, fabs(s.g[c]));
        for (int l = tid; l < P.M; l += ST) m = fmax(m, fabs(P.g_l[l]));
        gmax = block_max(m, s);
    }
    // Jacobi scaling, computed once: 1 / (1 + ||column||)
    for (int c = tid; c < P.NPc; c += ST) s.scale[c] = s.col_free[c] ? 1.0 / (1.0 + sqrt(s.H[hidx(c, c)])) : 1.0;
    for (int l = tid; l < P.M; l += ST) P.sc_l[l] = 1.0 / (1.0 + sqrt(P.h_ll[l]));
    __syncthreads();

    auto apply_scaling = [&]() {
        for (int i = tid >> 5; i < P.NPc; i += NW) {               // one warp per row of the packed triangle (no index decode)
            if (!s.col_free[i]) continue;
            const double si = s.scale[i];
            const int r0 = i * (i + 1) / 2;
            for (int j = tid & 31; j <= i; j += 32)
                if (s.col_free[j]) s.H[r0 + j] *= si * s.scale[j];
        }
        for (int c = tid; c < P.NPc; c += ST) s.g[c] *= s.scale[c];
        for (int l = tid; l < P.M; l += ST) { const double sl = P.sc_l[l]; P.h_ll[l] *= sl * sl; P.g_l[l] *= sl; }
        for (int ck = tid >> 5; ck < P.d.n_active * 6; ck += NW) {          // one warp per row of W
            const double sck = s.scale[15 * (ck / 6) + ck % 6];
            for (int l = tid & 31; l < P.M; l += 32) P.W[widx(P, ck, l)] *= P.sc_l[l] * sck;
        }
        __syncthreads();
        for (int e = tid; e < P.NPc * (P.NPc + 1) / 2; e += ST) P.Hcopy[e] = s.H[e];
        __syncthreads();
    };
    apply_scaling();

    double x_norm = sqrt(ambient_sq(P, s, s.xs, nullptr, P.inv_x, nullptr));
    double minimum_cost = x_cost;
    double radius = kInitialRadius, mu = kMinMu, alpha = 0.0, dogleg_step_norm = 0.0;
    bool reuse = false;
    // scalars of the current linearisation
    double gnorm2 = 0.0, uHu = 0.0, gsy = 0.0, D2y2 = 0.0, gn_norm2 = 0.0, grad_dot_gn = 0.0;
    int iteration = 0, termination = -1, n_succ = 0, n_unsucc = 0, n_invalid = 0, n_rec = 0;
    bool it_success = true;
    double it_cost = x_cost;
    int spec_k = 0, spec_n = 0, spec_last = 1;      // candidate batch: next slot to consume, slots filled, size of the last batch
    if (tid == 0) { out.initial_cost = x_cost; }

    while (true) {
        // ---- FinalizeIterationAndCheckIfMinimizerCanContinue ----
        if (it_success) {
            n_succ++;
            if (x_cost < minimum_cost || iteration == 0) {
                minimum_cost = fmin(minimum_cost, x_cost);
                // parameters <- x ; StateUpdatingCallback copies them to user memory
                for (int i = tid; i < P.d.n_active * 16; i += ST) { s.us[i] = s.xs[i]; P.states_g[i] = s.xs[i]; }
                for (int l = tid; l < P.M; l += ST) P.invd_user[l] = P.inv_x[l];
                __syncthreads();
            }
        } else n_unsucc++;
        if (tid == 0 && n_rec < kMaxIter) out.cost_trace[n_rec] = it_cost;
        n_rec++;
        if (iteration >= P.d.max_iterations) { termination = 0; break; }
        if (gmax <= kGradTol) { termination = 1; break; }
        if (radius <= kMinRadius) { termination = 2; break; }
        iteration++;
        it_success = false;
        // ---- DoglegStrategy::ComputeStep ----
        bool solver_failed = false;
        const bool fresh = !reuse;
        if (!reuse) {
            reuse = true;
            spec_n = 0;                                             // candidates of the previous linearisation are void
            // diagonal = sqrt(clamp(diag(Js^T Js)))
            for (int c = tid; c < P.NPc; c += ST)
                s.D[c] = s.col_free[c] ? sqrt(fmin(fmax(s.H[hidx(c, c)], kMinDiag), kMaxDiag)) : 1.0;
            for (int l = tid; l < P.M; l += ST) P.D_l[l] = sqrt(fmin(fmax(P.h_ll[l], kMinDiag), kMaxDiag));
            __syncthreads();
            // gradient / Cauchy point: ||grad||^2 with grad = gs / D ; uHu = v^T Hs v with v = gs / D^2
            double a = 0.0, q = 0.0;
            for (int c = tid; c < P.NPc; c += ST) { const double t = s.g[c] / s.D[c]; a += t * t; s.bred[c] = s.g[c] / (s.D[c] * s.D[c]); }
            for (int l = tid; l < P.M; l += ST) { const double t = P.g_l[l] / P.D_l[l]; a += t * t; }
            gnorm2 = block_sum(a, s);
            for (int i = tid >> 5; i < P.NPc; i += NW) {           // v^T Hs v over the packed triangle, one warp per row
                if (!s.col_free[i]) continue;
                const double vi = s.bred[i];
                const int r0 = i * (i + 1) / 2;
                double qr = 0.0;
                for (int j = tid & 31; j <= i; j += 32)
                    if (s.col_free[j]) qr += (i == j ? 1.0 : 2.0) * s.H[r0 + j] * s.bred[j];
                q += vi * qr;
            }
            for (int l = tid; l < P.M; l += ST) {
                const double vl = P.g_l[l] / (P.D_l[l] * P.D_l[l]);
                double wv = 0.0;
                for (int f = 0; f < P.d.n_active; ++f)
                    if (s.pose_free[f])
                        for (int k = 0; k < 6; ++k) wv += P.W[widx(P, f * 6 + k, l)] * s.bred[15 * f + k];
                q += 2.0 * vl * wv + P.h_ll[l] * vl * vl;
            }
            uHu = block_sum(q, s);
            alpha = gnorm2 / uHu;
            // Gauss-Newton step with the mu loop
            solver_failed = true;
            while (mu < kMaxMu) {
                kp.mark(18);
                const bool gn_ok = gauss_newton(P, s, mu, kp);
                if (gn_ok) { solver_failed = false; break; }
                mu *= kMuInc;
                for (int e = tid; e < P.NPc * (P.NPc + 1) / 2; e += ST) s.H[e] = P.Hcopy[e];
                __syncthreads();
            }
            if (!solver_failed) {
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
                for (int c = tid; c < P.NPc; c += ST)
                    if (s.col_free[c]) { t1 += s.g[c] * s.y[c]; t2 += s.D[c] * s.D[c] * s.y[c] * s.y[c]; }
                for (int l = tid; l < P.M; l += ST) { t1 += P.g_l[l] * P.y_l[l]; t2 += P.D_l[l] * P.D_l[l] * P.y_l[l] * P.y_l[l]; }
                gsy = block_sum(t1, s);
                D2y2 = block_sum(t2, s);
                for (int c = tid; c < P.NPc; c += ST) {
                    const bool fr = s.col_free[c];
                    s.du[c] = fr ? -s.g[c] / (s.D[c] * s.D[c]) * s.scale[c] : 0.0;
                    s.dw[c] = fr ? -s.y[c] * s.scale[c] : 0.0;
                }
                for (int l = tid; l < P.M; l += ST) {
                    P.du_l[l] = -P.g_l[l] / (P.D_l[l] * P.D_l[l]) * P.sc_l[l];
                    P.dw_l[l] = -P.y_l[l] * P.sc_l[l];
                }
                __syncthreads();
                gn_norm2 = D2y2;               // ||gn||^2 with gn = -D y
                grad_dot_gn = -gsy;            // grad . gn = sum (gs/D)(-D y)
                (void)t3;
            }
        }
        double model_cost_change = 0.0, cand_cost = 0.0, step_norm = 0.0;
        bool step_valid = false;
        if (!solver_failed) {
            if (!(spec_k < spec_n && s.sp[spec_k * 8] == radius)) {
                // ---- a new batch of candidates: this radius and, speculatively, the radii of the next rejections ----
                int nb = fresh ? 1 : min(kSpec, 2 * spec_last);    // a step right after a new linearisation is usually accepted
                nb = min(nb, P.d.max_iterations - iteration + 1);
                spec_last = nb;
                __syncthreads();                                    // every thread is done with the previous batch's slots
                if ((tid >> 5) < nb) {
                    // ComputeTraditionalDoglegStep for candidate b on warp b, side by side
                    const int b = tid >> 5;
                    double rad = radius;
                    for (int k = 0; k < b; ++k) rad *= 0.5;
                    double c1, c2, norm;
                    const double gnorm = sqrt(gnorm2), gn_norm = sqrt(gn_norm2);
                    if (gn_norm <= rad) { c1 = 0.0; c2 = 1.0; norm = gn_norm; }
                    else if (gnorm * alpha >= rad) { c1 = rad / gnorm; c2 = 0.0; norm = rad; }
                    else {
                        const double b_dot_a = -alpha * grad_dot_gn;
                        const double a_sq = (alpha * gnorm) * (alpha * gnorm);
                        const double bma_sq = a_sq - 2.0 * b_dot_a + gn_norm2;
                        const double c = b_dot_a - a_sq;
                        const double d = sqrt(c * c + bma_sq * (rad * rad - a_sq));
                        const double beta = (c <= 0) ? (d - c) / bma_sq : (rad * rad - a_sq) / (d + c);
                        c1 = alpha * (1.0 - beta); c2 = beta;
                        norm = sqrt(fmax(c1 * c1 * gnorm2 - 2.0 * c1 * c2 * grad_dot_gn + c2 * c2 * gn_norm2, 0.0));
                    }
                    // step = c1 u + c2 w (Jacobi-scaled space), u = -gs/D^2, w = -y, Hs w = -gs + mu D^2 y
                    const double u_g = -gnorm2, w_g = -gsy;
                    const double uHw = gnorm2 - mu * gsy;
                    const double wHw = gsy - mu * D2y2;
                    const double mcc = -(c1 * u_g + c2 * w_g) - 0.5 * (c1 * c1 * uHu + 2.0 * c1 * c2 * uHw + c2 * c2 * wHw);
                    if ((tid & 31) == 0) {
                        double *sp = s.sp + b * 8;
                        sp[0] = rad; sp[1] = c1; sp[2] = c2; sp[3] = norm; sp[4] = mcc; sp[5] = (mcc > 0.0) ? 1.0 : 0.0;
                    }
                }
                __syncthreads();
                int nbv = 0;
                while (nbv < nb && s.sp[nbv * 8 + 5] != 0.0) ++nbv;  // an invalid step ends the run of rejections (HandleInvalidStep)
                kp.mark(20);
                if (nbv > 0) {
                    make_candidates(P, s, nbv);
                    kp.mark(17);
                    evaluate_cost_multi(P, s, nbv, rg, true);           // -> s.sp[b * 8 + 6 .. 7]
                    kp.mark(9);
                }
                spec_k = 0; spec_n = max(nbv, 1);
            }
            const double *sp = s.sp + spec_k * 8;
            dogleg_step_norm = sp[3]; model_cost_change = sp[4]; step_valid = sp[5] != 0.0;
            cand_cost = sp[6]; step_norm = sqrt(sp[7]);
        }
        if (!step_valid) {
            n_invalid++;
            if (n_invalid >= kMaxInvalid) { termination = 3; if (tid == 0 && n_rec < kMaxIter) out.cost_trace[n_rec] = it_cost; n_rec++; break; }
            mu *= kMuInc; reuse = false;
            // the linearisation is unchanged, but H was consumed by the factorisation
            for (int e = tid; e < P.NPc * (P.NPc + 1) / 2; e += ST) s.H[e] = P.Hcopy[e];
            __syncthreads();
            continue;
        }
        n_invalid = 0;
        // a failed evaluation leaves candidate_cost = DBL_MAX (ComputeCandidatePointAndEvaluateCost): the step is rejected and
        // the radius halved; it must not look like "function tolerance reached"
        if (!isfinite(cand_cost)) cand_cost = DBL_MAX;
        if (step_norm <= kParamTol * (x_norm + kParamTol)) { termination = 4; if (tid == 0 && n_rec < kMaxIter) out.cost_trace[n_rec] = it_cost; n_rec++; break; }
        const double cost_change = x_cost - cand_cost;
        if (fabs(cost_change) <= kFuncTol * x_cost) { termination = 5; if (tid == 0 && n_rec < kMaxIter) out.cost_trace[n_rec] = it_cost; n_rec++; break; }
        const double rel = cost_change / model_cost_change;
        if (rel > kMinRelDecrease) {
            // HandleSuccessfulStep: user memory still holds the previous iterate here (D10)
            // (frames beyond the active ones are constant: equal in every candidate)
            for (int i = tid; i < P.d.n_active * 16; i += ST) s.xs[i] = cand_frame(s, spec_k, i >> 4, P.d.n_active)[i & 15];
            for (int l = tid; l < P.M; l += ST) P.inv_x[l] = P.inv_c[(size_t)spec_k * P.Mp + l];
            __syncthreads();
            x_norm = sqrt(ambient_sq(P, s, s.xs, nullptr, P.inv_x, nullptr));
            x_cost = evaluate(P, s, s.xs, P.inv_x, rg, true, true, &kp);
            {
                double m = 0.0;
                for (int c = tid; c < P.NPc; c += ST) m = fmax(m, fabs(s.g[c]));
                for (int l = tid; l < P.M; l += ST) m = fmax(m, fabs(P.g_l[l]));
                gmax = block_max(m, s);
            }
            kp.mark(22);
            apply_scaling();
            kp.mark(23);
            it_success = true;
            it_cost = x_cost;
            if (rel < 0.25) radius *= 0.5;
            if (rel > 0.75) radius = fmax(radius, 3.0 * dogleg_step_norm);
            mu = fmax(kMinMu, 2.0 * mu / kMuInc);
            reuse = false;
        } else {
            it_success = false;
            it_cost = cand_cost;
            radius *= 0.5;
            reuse = true;
            spec_k++;                                               // the next candidate of the batch, if any, is this radius
        }
    }
    if (tid == 0) {
        out.iterations = n_rec; out.termination = termination; out.n_successful = n_succ; out.n_unsuccessful = n_unsucc;
        out.final_cost = minimum_cost;
    }
}

__global__ void __launch_bounds__(ST, 1) solve_kernel(SolveBatch b, const RigParams *__restrict__ rig_g) {
    extern __shared__ unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    __shared__ RigParams rg;
    solve_body(b, rig_g, s, rg);
}

// Large windows (BASELINE config 5: 30 keyframes x 2000 landmarks): same trust-region loop, H in global scratch, Schur
// complement on the FP64 tensor pipe, panel Cholesky on the L2-resident triangle (solve_large.cuh).
__global__ void __launch_bounds__(ST, 1) solve_large_kernel(SolveBatch b, const RigParams *__restrict__ rig_g) {
    extern __shared__ unsigned char smem_raw[];
    SmemL &s = *reinterpret_cast<SmemL *>(smem_raw);
    __shared__ RigParams rg;
    solve_body(b, rig_g, s, rg);
}

size_t solve_large_scratch_doubles(int n_active, int n_lm_free, int n_pre, int marg_k, int n_rpe) {
    return solve_scratch_doubles(n_active, n_lm_free, n_pre, marg_k, n_rpe) - HPACK + 2 * (size_t)SmemL::HP - jb_doubles(n_rpe);
}

void solve_large(const SolveBatch &b, int n_problems, const RigParams *rig, cudaStream_t st) {
    if (n_problems <= 0) return;
    static xv::SmemOptIn attr_done;
    xv::ensure_dynamic_smem(solve_large_kernel, (int)kSolveLargeSmem, attr_done);
    solve_large_kernel<<<n_problems, ST, kSolveLargeSmem, st>>>(b, rig);
}

void solve(const SolveBatch &b, int n_problems, const RigParams *rig, cudaStream_t st) {
    if (n_problems <= 0) return;
    static xv::SmemOptIn attr_done;
    xv::ensure_dynamic_smem(solve_kernel, (int)kSolveSmem, attr_done);
    solve_kernel<<<n_problems, ST, kSolveSmem, st>>>(b, rig);
}

// A0 = S^T S; blockIdx.x = prior, blockIdx.y = slice of the n^2 entries (one entry per thread and pass, rows of S in order)
__global__ void __launch_bounds__(256) gram_kernel(const double *__restrict__ S, double *__restrict__ A0, int n, size_t stride) {
    const double *Sp = S + (size_t)blockIdx.x * stride;
    double *Ap = A0 + (size_t)blockIdx.x * stride;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < n * n; e += gridDim.y * blockDim.x) {
        const int i = e / n, j = e % n;
        double a = 0.0;
        for (int r = 0; r < n; ++r) a += Sp[(size_t)r * n + i] * Sp[(size_t)r * n + j];
        Ap[e] = a;
    }
}
void gram(const double *S, double *A0, int n, int n_priors, size_t stride, cudaStream_t st) {
    if (n_priors <= 0 || n <= 0) return;
    const int slices = min(max((n * n + 1023) / 1024, 1), 296);
    gram_kernel<<<dim3(n_priors, slices), 256, 0, st>>>(S, A0, n, stride);
}

}  // namespace xvbe

// ---- FP64 peak probe: the denominators of the config-5 roofline (MEASURED_PEAKS.json has no FP64 entry) ----
#ifdef CUEMU
extern "C" int xrvio_debug_fp64_peak(double *, double *) { return -1; }      // a rate measured on the emulator means nothing
#else
namespace xvbe {
__global__ void __launch_bounds__(256) fp64_fma_probe_kernel(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = x[k] * a + b;
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += x[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void __launch_bounds__(256) fp64_mma_probe_kernel(double *out, int iters, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) { c[k][0] = k; c[k][1] = k + 0.5; }
    const double fa = a + threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) dmma_m8n8k4(c[k][0], c[k][1], fa, b);
    }
    double r = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) r += c[k][0] + c[k][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;
}
}  // namespace xvbe

// Measured FP64 rates of this GPU in TFLOP/s: a register-resident DFMA loop and an mma.sync.m8n8k4.f64 loop on every SM
// (8 CTAs of 256 threads per SM, CUDA events).  Diagnostics for bench.py --workload ba_stress.
extern "C" int xrvio_debug_fp64_peak(double *dfma_tflops, double *dmma_tflops) {
    if (!dfma_tflops || !dmma_tflops) return -1;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) return -1;
    const int grid = sms * 8, iters = 8192;
    double *out = nullptr;
    if (cudaMalloc(&out, (size_t)grid * 256 * sizeof(double)) != cudaSuccess) return -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0.f;
    int rc = 0;
    for (int pass = 0; pass < 2; ++pass) {          // pass 0 warms up
        cudaEventRecord(e0, 0);
        xvbe::fp64_fma_probe_kernel<<<grid, 256, 0, 0>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) rc = -1;
        cudaEventElapsedTime(&ms, e0, e1);
    }
    *dfma_tflops = 2.0 * 8.0 * iters * 256.0 * grid / (ms * 1e-3) / 1e12;
    for (int pass = 0; pass < 2; ++pass) {
        cudaEventRecord(e0, 0);
        xvbe::fp64_mma_probe_kernel<<<grid, 256, 0, 0>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, 0);
        if (cudaEventSynchronize(e1) != cudaSuccess) rc = -1;
        cudaEventElapsedTime(&ms, e0, e1);
    }
    *dmma_tflops = 2.0 * 8 * 8 * 4 * 4.0 * iters * 8.0 * grid / (ms * 1e-3) / 1e12;     // 4 mma of 8x8x4 per warp and iteration, 8 warps
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return rc;
}
#endif

extern "C" int xrvio_debug_kprof(unsigned long long *out32, int reset) {
    if (cudaMemcpyFromSymbol(out32, xvbe::g_kprof, 32 * sizeof(unsigned long long)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(xvbe::g_kprof, z, sizeof(z)); }
    return 0;
}
