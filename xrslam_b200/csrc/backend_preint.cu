// K6: batched IMU preintegration, one warp per job.  Restates
// xrslam/src/xrslam/estimation/preintegrator.cpp:9-100 (reset / increment / integrate /
// compute_sqrt_inv_cov): left-endpoint Euler steps, 9x9 covariance propagation A S A^T + B Q B^T,
// bias random walk, five 3x3 bias Jacobians, sqrt_inv_cov = chol(cov^-1)^T.
#include "backend.cuh"
#include "vmath.cuh"

namespace xvbe {
using namespace xv;

constexpr int PW = 4;   // warps per block

__device__ __forceinline__ double blkA(int r, int c, const M3 &Rw, const M3 &P1, const M3 &V1, double dt) {
    // A in error-state order q(0-2) p(3-5) v(6-8)   (preintegrator.cpp:33-38)
    int br = r / 3, bc = c / 3, i = r % 3, j = c % 3;
    if (br == 0) return bc == 0 ? Rw(i, j) : 0.0;
    if (br == 1) return bc == 0 ? P1(i, j) : (bc == 1 ? (i == j ? 1.0 : 0.0) : (i == j ? dt : 0.0));
    return bc == 0 ? V1(i, j) : (bc == 2 ? (i == j ? 1.0 : 0.0) : 0.0);
}

__global__ void __launch_bounds__(32 * PW) preint_kernel(const PreintJob *__restrict__ jobs, int n_jobs,
                                                          const double *__restrict__ imu,
                                                          const RigParams *__restrict__ rig,
                                                          double *__restrict__ out) {
    __shared__ double sC[PW][81], sT[PW][81], sA[PW][81], sN[PW][81];
    // Per-sample quantities that do not depend on the running state (dt, a, exp(dt w), its rotation matrix, the right Jacobian and
    // the gyro noise block): 32 samples at a time are computed by the 32 lanes IN PARALLEL and parked here, so the strictly
    // sequential recursion below no longer carries the two sin / cos evaluations per sample on its dependent chain.  Same
    // formulas in the same order as before: results are bit-identical.  The buffer doubles as the 15 x 30 Gauss-Jordan tableau
    // of the epilogue.
    constexpr int SQ = 36;      // dt, inv_dt, a(3), ew(4), Rw(9), Jr(9), NQ(9)
    __shared__ double sS[PW][32 * SQ];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int job = blockIdx.x * PW + warp;
    if (job >= n_jobs) return;
    const PreintJob jb = jobs[job];
    const bool want_jac = jb.flags & 1, want_cov = jb.flags & 2;
    double *C = sC[warp], *T = sT[warp], *A = sA[warp], *Nn = sN[warp];
    for (int e = lane; e < 81; e += 32) C[e] = 0.0;
    M3 cov_w, cov_a, cov_bg, cov_ba;
    for (int i = 0; i < 9; ++i) { cov_w.m[i] = rig->cov_w[i]; cov_a.m[i] = rig->cov_a[i]; cov_bg.m[i] = rig->cov_bg[i]; cov_ba.m[i] = rig->cov_ba[i]; }
    M3 Cbg, Cba;
    for (int i = 0; i < 9; ++i) { Cbg.m[i] = 0; Cba.m[i] = 0; }
    Q4 dq = q4(0, 0, 0, 1);
    V3 dp = v3(0, 0, 0), dv = v3(0, 0, 0);
    double tsum = 0.0;
    M3 dq_dbg, dp_dbg, dp_dba, dv_dbg, dv_dba;
    for (int i = 0; i < 9; ++i) dq_dbg.m[i] = dp_dbg.m[i] = dp_dba.m[i] = dv_dbg.m[i] = dv_dba.m[i] = 0.0;
    const V3 bg = ld3(jb.bg), ba = ld3(jb.ba);
    __syncwarp();
    for (int s0 = 0; s0 < jb.n; s0 += 32) {
        {   // ---- parallel part: lane l prepares sample s0 + l ----
            const int sl = s0 + lane;
            if (sl < jb.n) {
                const double *d = imu + (size_t)(jb.imu_off + sl) * 7;
                const double t0 = d[0];
                const double t1 = (sl + 1 < jb.n) ? d[7] : jb.t_end;
                const double dt = t1 - t0;
                const V3 w = ld3(d + 1) - bg;
                const V3 a = ld3(d + 4) - ba;
                const Q4 ew = expmap(dt * w);
                const M3 Rw = qmat(qconj(ew));
                double *q = sS[warp] + lane * SQ;
                q[0] = dt;
                st3(q + 2, a); st4(q + 5, ew);
                for (int i = 0; i < 9; ++i) q[9 + i] = Rw.m[i];
                if (want_cov || want_jac) {
                    const M3 Jr = right_jacobian(dt * w);
                    for (int i = 0; i < 9; ++i) q[18 + i] = Jr.m[i];
                    if (want_cov) {
                        const double inv_dt = 1.0 / fmax(dt, 1.0e-7);
                        const M3 Bq = dt * Jr;
                        const M3 NQ = Bq * (inv_dt * cov_w) * transpose(Bq);
                        q[1] = inv_dt;
                        for (int i = 0; i < 9; ++i) q[27 + i] = NQ.m[i];
                    }
                }
            }
        }
        __syncwarp();
        // ---- sequential part ----
        const int s1 = min(s0 + 32, jb.n);
        for (int s = s0; s < s1; ++s) {
            const double *q = sS[warp] + (s - s0) * SQ;
            const double dt = q[0];
            const V3 a = ld3(q + 2);
            const Q4 ew = ld4(q + 5);
            M3 Rw, Jr;
            for (int i = 0; i < 9; ++i) { Rw.m[i] = q[9 + i]; Jr.m[i] = q[18 + i]; }
            const M3 R = qmat(dq);
            const M3 Rha = R * hat(a);
            if (want_cov) {
                const M3 P1 = (-0.5 * dt * dt) * Rha, V1 = (-dt) * Rha;
                const double inv_dt = q[1];
                M3 NQ;
                for (int i = 0; i < 9; ++i) NQ.m[i] = q[27 + i];
                const M3 RQ = R * (inv_dt * cov_a) * transpose(R);
                // A (9x9) and the noise term B Q B^T (9x9) are written once to shared memory, then two plain 9x9x9 products
                for (int e = lane; e < 81; e += 32) {
                    const int r = e / 9, c = e % 9;
                    A[e] = blkA(r, c, Rw, P1, V1, dt);
                    const int br = r / 3, bc = c / 3, i = r % 3, j = c % 3;
                    double nz = 0.0;
                    if (br == 0 && bc == 0) nz = NQ(i, j);
                    else if (br == 2 && bc == 2) nz = dt * dt * RQ(i, j);
                    else if (br == 1 && bc == 1) nz = 0.25 * dt * dt * dt * dt * RQ(i, j);
                    else if ((br == 1 && bc == 2) || (br == 2 && bc == 1)) nz = 0.5 * dt * dt * dt * RQ(i, j);
                    Nn[e] = nz;
                }
                __syncwarp();
                for (int e = lane; e < 81; e += 32) {
                    const int r = e / 9, c = e % 9;
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < 9; ++m) acc += A[r * 9 + m] * C[m * 9 + c];
                    T[e] = acc;
                }
                __syncwarp();
                for (int e = lane; e < 81; e += 32) {
                    const int r = e / 9, c = e % 9;
                    double acc = Nn[e];
#pragma unroll
                    for (int m = 0; m < 9; ++m) acc += T[r * 9 + m] * A[c * 9 + m];
                    C[e] = acc;
                }
                __syncwarp();
                Cbg = Cbg + dt * cov_bg;
                Cba = Cba + dt * cov_ba;
            }
            if (want_jac) {
                // preintegrator.cpp:58-69
                dp_dbg = dp_dbg + dt * dv_dbg - (0.5 * dt * dt) * (Rha * dq_dbg);
                dp_dba = dp_dba + dt * dv_dba - (0.5 * dt * dt) * R;
                dv_dbg = dv_dbg - dt * (Rha * dq_dbg);
                dv_dba = dv_dba - dt * R;
                dq_dbg = Rw * dq_dbg - dt * Jr;
            }
            const V3 qa = qrot(dq, a);
            tsum = tsum + dt;
            dp = dp + dt * dv + (0.5 * dt * dt) * qa;
            dv = dv + dt * qa;
            dq = qnormalized(qmul(dq, ew));
        }
        __syncwarp();
    }
    double *o = out + (size_t)job * kPreBlock;
    if (lane == 0) {
        o[PB_T] = tsum;
        st4(o + PB_Q, dq); st3(o + PB_P, dp); st3(o + PB_V, dv);
        for (int i = 0; i < 9; ++i) {
            o[PB_DQ_DBG + i] = dq_dbg.m[i]; o[PB_DP_DBG + i] = dp_dbg.m[i]; o[PB_DP_DBA + i] = dp_dba.m[i];
            o[PB_DV_DBG + i] = dv_dbg.m[i]; o[PB_DV_DBA + i] = dv_dba.m[i];
        }
    }
    for (int e = lane; e < 225; e += 32) {
        int r = e / 15, c = e % 15;
        double v = 0.0;
        if (want_cov) {
            if (r < 9 && c < 9) v = C[r * 9 + c];
            else if (r >= 9 && r < 12 && c >= 9 && c < 12) v = Cbg(r - 9, c - 9);
            else if (r >= 12 && c >= 12) v = Cba(r - 12, c - 12);
        }
        o[PB_COV + e] = v;
        o[PB_SIC + e] = 0.0;
    }
    if (!want_cov) return;
    // sqrt_inv_cov = LLT(cov^-1).matrixL()^T  (preintegrator.cpp:97-100)
    double (*M)[30] = reinterpret_cast<double (*)[30]>(sS[warp]);      // 450 of the 1152 doubles of the sample buffer
    __syncwarp();
    for (int e = lane; e < 450; e += 32) {
        int r = e / 30, c = e % 30;
        M[r][c] = c < 15 ? o[PB_COV + r * 15 + c] : ((c - 15) == r ? 1.0 : 0.0);
    }
    __syncwarp();
    for (int k = 0; k < 15; ++k) {       // Gauss-Jordan, SPD -> no pivoting needed
        double piv = M[k][k];
        double f = (lane < 15 && lane != k) ? M[lane][k] / piv : 0.0;
        __syncwarp();
        if (lane < 15 && lane != k)
            for (int c = 0; c < 30; ++c) M[lane][c] -= f * M[k][c];
        __syncwarp();
        if (lane < 30) M[k][lane] /= piv;
        __syncwarp();
    }
    // Cholesky of the inverse (right half of M), lower L; output U = L^T
    // column by column, lane i owns row i: every element sees the same operations in the same order as a serial
    // Cholesky-Banachiewicz sweep (which one lane used to run with its 15 x 15 factor in local memory)
    double *L = sS[warp] + 512;                     // 15 x 15, behind the tableau
    for (int j = 0; j < 15; ++j) {
        double sv = 0.0;
        if (lane >= j && lane < 15) {
            // symmetrise the two triangles the way a self-adjoint view would (use lower)
            sv = M[lane][15 + j];
            for (int k = 0; k < j; ++k) sv -= L[lane * 15 + k] * L[j * 15 + k];
        }
        const double djj = sqrt(__shfl_sync(0xffffffffu, sv, j));
        if (lane >= j && lane < 15) L[lane * 15 + j] = (lane == j) ? djj : sv / djj;
        __syncwarp();
    }
    for (int e = lane; e < 225; e += 32) {
        const int i = e / 15, j = e % 15;
        if (j <= i) o[PB_SIC + j * 15 + i] = L[i * 15 + j];
    }
}

void preintegrate(const PreintJob *jobs, int n_jobs, const double *imu, const RigParams *rig, double *out,
                  cudaStream_t st) {
    if (n_jobs <= 0) return;
    preint_kernel<<<(n_jobs + PW - 1) / PW, 32 * PW, 0, st>>>(jobs, n_jobs, imu, rig, out);
}

}  // namespace xvbe
