// K12: marginalisation of the oldest keyframe, one CTA per session.  Restates
// xrslam/src/xrslam/estimation/ceres/marginalization_factor.h:74-475: dense information matrix
// from the old prior, the victim's IMU edge and every reprojection of the victim's tracks (no
// robust loss), scalar-landmark Schur complement, 15x15 frame Schur complement, then the symmetric
// eigen re-factorisation sqrt(L) V^T with eigenvalue cut 1e-8.  The eigen step is done as
// diagonal-pivoted Cholesky (rank revealing) followed by one-sided Jacobi on the factor rows,
// which yields rows sqrt(lambda_i) v_i^T directly (sign / order of rows is immaterial to the
// prior's cost and Jacobian).  Also emits A0 = S^T S for the solver.
#include "solve_core.cuh"
#include "launch_attr.cuh"

namespace xvbe {
using namespace xv;

size_t marg_scratch_doubles(int n_frames, int n_lm, int n_rpe) {
    size_t n = (size_t)15 * n_frames;
    return solve_scratch_doubles(n_frames, n_lm, 2, n_frames, n_rpe) + 2 * n * n + 16 * n + 1024 + jb_doubles(n_rpe);
}

__global__ void __launch_bounds__(ST, 1) marg_kernel(MargBatch b, const RigParams *__restrict__ rig_g) {
    extern __shared__ unsigned char smem_raw[];
    Smem &s = *reinterpret_cast<Smem *>(smem_raw);
    __shared__ RigParams rg;
    __shared__ double Binv[15][30];
    __shared__ int s_p;
    __shared__ double s_dmax;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) rg = *rig_g;
    const MargDesc md = b.desc[blockIdx.x];
    const int N = md.n_frames;
    Prob P;
    P.d.frame_off = md.frame_off; P.d.n_frames = N; P.d.n_active = N;
    P.d.lm_off = md.lm_off; P.d.n_lm = md.n_lm; P.d.n_lm_free = md.n_lm;
    P.d.rpe_off = md.rpe_off; P.d.n_rpe = md.n_rpe; P.d.rot_off = 0; P.d.n_rot = 0;
    P.d.pre_off = md.pre_off; P.d.n_pre = md.n_pre;
    P.d.marg_k = md.marg_k; P.d.marg_frames_off = md.marg_frames_off; P.d.marg_off = md.marg_off;
    P.d.scratch_off = md.scratch_off; P.d.max_iterations = 0;
    P.states_g = const_cast<double *>(b.states) + (size_t)md.frame_off * 16;
    P.flags_g = nullptr;
    P.invd_user = const_cast<double *>(b.inv_depth) + md.lm_off;
    P.rpe = b.rpe; P.rot = nullptr; P.pre = b.pre; P.pre_blocks = b.pre_blocks; P.pre_out = nullptr; P.marg_frames = b.marg_frames;
    P.M = md.n_lm; P.Wp = (P.M + 31) / 32 * 32; P.W_sh = WLayout<true>::sh; P.W_pitch = WLayout<true>::pitch; P.W_rows = 6 * N; P.NPc = 15 * N; P.marg_n = 15 * md.marg_k;
    double *Gm;      // factor rows (global)
    double *Yg;      // n1 x 15
    {
        const double *m = b.marg + md.marg_off;
        const size_t n = P.marg_n;
        P.marg_lin = m; P.marg_S = m + 16 * md.marg_k; P.marg_info = P.marg_S + n * n; P.marg_A0 = P.marg_info + n;
        double *sc = b.scratch + md.scratch_off;
        const size_t Mp = (size_t)((P.M + 31) / 32) * 32 + 32;
        P.inv_x = sc; sc += Mp; P.inv_c = sc; sc += Mp; P.h_ll = sc; sc += Mp; P.g_l = sc; sc += Mp;
        P.sc_l = sc; sc += Mp; P.D_l = sc; sc += Mp; P.y_l = sc; sc += Mp; P.ht_l = sc; sc += Mp;
        P.W = sc; sc += (size_t)N * 6 * (Mp + Mp / 4);
        P.Hcopy = sc; sc += HPACK;
        P.pre_s = sc; sc += (size_t)2 * kPreScratch;
        P.m_dx = sc; sc += n + 32; P.m_r = sc; sc += n + 32; P.m_v = sc; sc += n + 32; sc += n + 32;
        P.m_T = sc; sc += n * n;
        Gm = sc; sc += (size_t)(15 * N) * (15 * N);
        Yg = sc; sc += (size_t)16 * 15 * N;
        prepare_prior(P, m, sc); sc += n * n;
        prepare_tangents(P, sc); sc += (size_t)6 * md.n_rpe;
        prepare_groups(P, s, sc);
        P.du_l = nullptr; P.dw_l = nullptr;
    }
    for (int i = tid; i < N * 16; i += ST) { const double v = P.states_g[i]; s.us[i] = v; s.xs[i] = v; }
    for (int f = tid; f < kMaxActive; f += ST) { s.pose_free[f] = f < N; s.motion_free[f] = f < N; }
    for (int c = tid; c < kNP; c += ST) s.col_free[c] = c < 15 * N;
    for (int l = tid; l < P.M; l += ST) P.inv_x[l] = P.invd_user[l];
    __syncthreads();
    // information matrix / vector at the current states, no robust loss (:108-380)
    evaluate(P, s, s.xs, P.inv_x, rg, true, false);
    // marginalise landmarks (:383-398)
    for (int l = tid; l < P.M; l += ST) P.ht_l[l] = P.h_ll[l];
    __syncthreads();
    schur_landmarks(P, s, s.g);
    __syncthreads();
    // marginalise the victim (last 15 columns) (:400-438)
    const int n1 = 15 * (N - 1);
    for (int e = tid; e < 450; e += ST) {
        const int r = e / 30, c = e % 30;
        if (c < 15) { const int A = n1 + r, B = n1 + c; Binv[r][c] = s.H[A >= B ? hidx(A, B) : hidx(B, A)]; }
        else Binv[r][c] = (c - 15 == r) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (warp == 0) {                       // Gauss-Jordan with partial pivoting on [B | I]
        for (int k = 0; k < 15; ++k) {
            int piv = k;
            double best = fabs(Binv[k][k]);
            for (int r = k + 1; r < 15; ++r) { const double v = fabs(Binv[r][k]); if (v > best) { best = v; piv = r; } }
            if (piv != k && lane < 30) { const double t = Binv[k][lane]; Binv[k][lane] = Binv[piv][lane]; Binv[piv][lane] = t; }
            __syncwarp();
            const double pv = Binv[k][k];
            const double f = (lane < 15 && lane != k) ? Binv[lane][k] / pv : 0.0;
            __syncwarp();
            if (lane < 15 && lane != k)
                for (int c = 0; c < 30; ++c) Binv[lane][c] -= f * Binv[k][c];
            __syncwarp();
            if (lane < 30) Binv[k][lane] /= pv;
            __syncwarp();
        }
    }
    __syncthreads();
    // Y = H_rv * B^-1   (H_rv[i][j] = H[n1 + j][i])
    for (int e = tid; e < n1 * 15; e += ST) {
        const int i = e / 15, c = e % 15;
        double a = 0.0;
        for (int j = 0; j < 15; ++j) a += s.H[hidx(n1 + j, i)] * Binv[j][15 + c];
        Yg[e] = a;
    }
    __syncthreads();
    // Hc = H_rr - Y H_vr ; bc = b_r - Y b_v
    for (int e = tid; e < n1 * (n1 + 1) / 2; e += ST) {
        int i = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        const int k = e - i * (i + 1) / 2;
        double a = 0.0;
        for (int j = 0; j < 15; ++j) a += Yg[i * 15 + j] * s.H[hidx(n1 + j, k)];
        s.H[e] -= a;
    }
    for (int i = tid; i < n1; i += ST) {
        double a = 0.0;
        for (int j = 0; j < 15; ++j) a += Yg[i * 15 + j] * s.g[n1 + j];
        s.bred[i] = s.g[i] - a;                    // bc
    }
    __syncthreads();
    // ---- rank-revealing Cholesky with diagonal pivoting, right-looking and in place on the packed
    //      shared-memory Hc: step k scales the pivot column into row k of G (written to Gm, r x n1) and
    //      subtracts its outer product from the not-yet-pivoted part ----
    for (int j = tid; j < n1; j += ST) s.col_free[j] = 1;      // col_free == "not yet pivoted"
    __syncthreads();
    int rnk = 0;
    for (int k = 0; k < n1; ++k) {
        // argmax of the remaining diagonal
        double bv = -1.0; int bi = -1;
        for (int j = tid; j < n1; j += ST)
            if (s.col_free[j]) { const double dj = s.H[hidx(j, j)]; if (dj > bv) { bv = dj; bi = j; } }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        __syncthreads();
        if (lane == 0) { s.red[warp] = bv; reinterpret_cast<int *>(s.jrinv)[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double m = -1.0; int mi = -1;
            for (int w = 0; w < NW; ++w) {
                const double v = s.red[w]; const int vi = reinterpret_cast<int *>(s.jrinv)[w];
                if (vi >= 0 && (mi < 0 || v > m || (v == m && vi < mi))) { m = v; mi = vi; }
            }
            s_p = mi; s_dmax = m;
        }
        __syncthreads();
        const int p = s_p;
        const double dmax = s_dmax;
        if (p < 0 || !(dmax > 1.0e-8)) break;
        const double inv = 1.0 / sqrt(dmax);
        // g = column p of the current matrix / sqrt(pivot); kept in s.y for the update
        for (int j = tid; j < n1; j += ST) {
            double v = 0.0;
            if (s.col_free[j]) v = s.H[j >= p ? hidx(j, p) : hidx(p, j)] * inv;
            s.y[j] = v;
            Gm[(size_t)k * n1 + j] = v;
        }
        __syncthreads();
        if (tid == 0) { s.col_free[p] = 0; s.g[k] = (double)p; }     // s.g is free now: remember the pivot order
        __syncthreads();
        for (int i = warp; i < n1; i += NW) {
            if (!s.col_free[i]) continue;
            const double gi = s.y[i];
            if (gi == 0.0) continue;
            const int row = i * (i + 1) / 2;
            for (int j = lane; j <= i; j += 32)
                if (s.col_free[j]) s.H[row + j] -= gi * s.y[j];
        }
        __syncthreads();
        rnk++;
    }
    __syncthreads();
    // infovec for the plain Cholesky factor: G^T infovec = bc on the range (forward substitution in pivot order)
    double *info_chol = s.cs;
    if (md.eigen_form == 0 && warp == 0) {
        for (int k = 0; k < rnk; ++k) {
            const int pk = (int)s.g[k];
            double a = 0.0;
            for (int m2 = lane; m2 < k; m2 += 32) a += Gm[(size_t)m2 * n1 + pk] * info_chol[m2];
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) info_chol[k] = (s.bred[pk] - a) / Gm[(size_t)k * n1 + pk];
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- one-sided Jacobi on the rows of G: rows become sqrt(lambda_i) v_i^T ----
    double *Gs = Gm;
    if ((size_t)rnk * n1 <= (size_t)HPACK) {            // factor fits in the (now free) shared H area
        for (int e = tid; e < rnk * n1; e += ST) s.H[e] = Gm[e];
        Gs = s.H.a;
        __syncthreads();
    }
    const int m = rnk + (rnk & 1);
    const bool eigen_form = md.eigen_form != 0;      // refine the Cholesky factor rows to sqrt(lambda_i) v_i^T (one-sided Jacobi)
    for (int sweep = 0; eigen_form && sweep < 40 && rnk > 1; ++sweep) {
        double worst = 0.0;
        for (int t = 0; t < m - 1; ++t) {
            for (int pr = warp; pr < m / 2; pr += NW) {
                int i, j;
                if (pr == 0) { i = m - 1; j = t % (m - 1); }
                else { i = (t + pr) % (m - 1); j = (t + m - 1 - pr) % (m - 1); }
                if (i >= rnk || j >= rnk) continue;
                double *gi = Gs + (size_t)i * n1, *gj = Gs + (size_t)j * n1;
                double a = 0.0, bb = 0.0, c = 0.0;
                for (int q = lane; q < n1; q += 32) { const double x = gi[q], y = gj[q]; a += x * x; bb += y * y; c += x * y; }
                for (int o = 16; o > 0; o >>= 1) {
                    a += __shfl_xor_sync(0xffffffffu, a, o);
                    bb += __shfl_xor_sync(0xffffffffu, bb, o);
                    c += __shfl_xor_sync(0xffffffffu, c, o);
                }
                const double denom = sqrt(a * bb);
                if (denom > 0.0 && fabs(c) > 1e-15 * denom) {
                    worst = fmax(worst, fabs(c) / denom);
                    const double zeta = (bb - a) / (2.0 * c);
                    const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
                    for (int q = lane; q < n1; q += 32) {
                        const double x = gi[q], y = gj[q];
                        gi[q] = cs * x - sn * y;
                        gj[q] = sn * x + cs * y;
                    }
                }
            }
            __syncthreads();
        }
        worst = block_max(worst, s);
        if (worst < 1e-13) break;
    }
    __syncthreads();
    // ---- output: kept rows (lambda > 1e-8) compacted to the top; infovec; A0 = S^T S; lin ----
    double *o = b.out + md.out_off;
    double *o_lin = o, *o_S = o + 16 * (N - 1), *o_info = o_S + (size_t)n1 * n1, *o_A0 = o_info + n1;
    for (int i = tid; i < 16 * (N - 1); i += ST) o_lin[i] = s.xs[i];     // slots 0..N-2 are the remaining frames
    for (int e = tid; e < n1 * n1; e += ST) o_S[e] = 0.0;
    for (int i = tid; i < n1; i += ST) o_info[i] = 0.0;
    __syncthreads();
    // row norms -> s.D ; keep flags -> compact index in s.y
    for (int i = warp; i < rnk; i += NW) {
        double a = 0.0, gb = 0.0;
        for (int q = lane; q < n1; q += 32) { const double x = Gs[(size_t)i * n1 + q]; a += x * x; gb += x * s.bred[q]; }
        for (int o2 = 16; o2 > 0; o2 >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o2); gb += __shfl_xor_sync(0xffffffffu, gb, o2); }
        if (lane == 0) { s.D[i] = a; s.scale[i] = gb; }
    }
    __syncthreads();
    if (tid == 0) {
        int kept = 0;
        for (int i = 0; i < rnk; ++i) {
            if (s.D[i] > 1.0e-8) { s.y[i] = (double)kept; kept++; } else s.y[i] = -1.0;
        }
        b.rank_out[blockIdx.x] = kept;
        s.flag = kept;
        o_A0[(size_t)n1 * n1] = (double)kept;       // leading non-zero rows of S, read back by prepare_prior
    }
    __syncthreads();
    for (int i = warp; i < rnk; i += NW) {
        const int dst = (int)s.y[i];
        if (dst < 0) continue;
        for (int q = lane; q < n1; q += 32) o_S[(size_t)dst * n1 + q] = Gs[(size_t)i * n1 + q];
        if (lane == 0) o_info[dst] = eigen_form ? s.scale[i] / s.D[i] : info_chol[i];
    }
    __syncthreads();
    const int kept = s.flag;
    for (int e = tid; e < n1 * n1; e += ST) {
        const int i = e / n1, j = e % n1;
        double a = 0.0;
        for (int r = 0; r < kept; ++r) a += o_S[(size_t)r * n1 + i] * o_S[(size_t)r * n1 + j];
        o_A0[e] = a;
    }
}

void marginalize(const MargBatch &b, int n_problems, const RigParams *rig, cudaStream_t st) {
    if (n_problems <= 0) return;
    static xv::SmemOptIn attr_done;
    xv::ensure_dynamic_smem(marg_kernel, (int)sizeof(Smem), attr_done);
    marg_kernel<<<n_problems, ST, sizeof(Smem), st>>>(b, rig);
}

// ------------------------------------------------------------------------------------------------
// T: DLT triangulation (track.cpp:46-76, stereo.h:84-94): right singular vector of the smallest
// singular value of the 2k x 4 system, computed as the smallest eigenvector of A^T A by cyclic
// Jacobi; cheirality check in every view.  One thread per track.
// ------------------------------------------------------------------------------------------------
__global__ void tri_kernel(const TriJob *__restrict__ jobs, int n_jobs, const double *__restrict__ views,
                           double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const TriJob tj = jobs[j];
    double A[4][4];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) A[r][c] = 0.0;
    for (int v = 0; v < tj.n_obs; ++v) {
        const double *vw = views + (size_t)(tj.obs_off + v) * 16;
        const M3 R = qmat(qconj(ld4(vw)));
        const V3 T = -(R * ld3(vw + 4));
        const V3 z = ld3(vw + 7);
        double P[3][4];
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) P[r][c] = R(r, c); P[r][3] = T[r]; }
        double r0[4], r1[4];
        for (int c = 0; c < 4; ++c) { r0[c] = z.x * P[2][c] - z.z * P[0][c]; r1[c] = z.y * P[2][c] - z.z * P[1][c]; }
        for (int a = 0; a < 4; ++a) for (int c = 0; c < 4; ++c) A[a][c] += r0[a] * r0[c] + r1[a] * r1[c];
    }
    double V[4][4];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) V[r][c] = (r == c) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 3; ++p) for (int q = p + 1; q < 4; ++q) off += A[p][q] * A[p][q];
        if (off < 1e-40) break;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sN = t * c;
                for (int k = 0; k < 4; ++k) { const double x = A[k][p], y = A[k][q]; A[k][p] = c * x - sN * y; A[k][q] = sN * x + c * y; }
                for (int k = 0; k < 4; ++k) { const double x = A[p][k], y = A[q][k]; A[p][k] = c * x - sN * y; A[q][k] = sN * x + c * y; }
                for (int k = 0; k < 4; ++k) { const double x = V[k][p], y = V[k][q]; V[k][p] = c * x - sN * y; V[k][q] = sN * x + c * y; }
            }
    }
    int mi = 0;
    for (int k = 1; k < 4; ++k) if (A[k][k] < A[mi][mi]) mi = k;
    double h[4] = {V[0][mi], V[1][mi], V[2][mi], V[3][mi]};
    bool valid = true;
    for (int v = 0; v < tj.n_obs && valid; ++v) {
        const double *vw = views + (size_t)(tj.obs_off + v) * 16;
        const M3 R = qmat(qconj(ld4(vw)));
        const V3 T = -(R * ld3(vw + 4));
        const double qz = R(2, 0) * h[0] + R(2, 1) * h[1] + R(2, 2) * h[2] + T.z * h[3];
        if (!(qz * h[3] > 0)) valid = false;
    }
    out[(size_t)j * 4 + 0] = h[0] / h[3]; out[(size_t)j * 4 + 1] = h[1] / h[3]; out[(size_t)j * 4 + 2] = h[2] / h[3];
    out[(size_t)j * 4 + 3] = valid ? 1.0 : 0.0;
}
void triangulate(const TriJob *jobs, int n_jobs, const double *views, double *out_xyz_valid, cudaStream_t st) {
    if (n_jobs <= 0) return;
    tri_kernel<<<(n_jobs + 127) / 128, 128, 0, st>>>(jobs, n_jobs, views, out_xyz_valid);
}

}  // namespace xvbe
