// Device functions for the residuals / analytic Jacobians of the sliding-window factors.
// Restated from the reference's factor headers (cited per function); Jacobians w.r.t. quaternion
// blocks are w.r.t. the right-multiplicative local perturbation (3 columns).
#pragma once
#include "backend.cuh"
#include "vmath.cuh"

namespace xvbe {
using namespace xv;

struct RpeEval {
    double r[2];      // whitened residual (before the robust corrector)
    double Jt[12];    // 2x6 d r / d (q_tgt, p_tgt)      row-major [row*6 + col]
    double Jr[12];    // 2x6 d r / d (q_ref, p_ref)
    double Jl[2];     // d r / d inv_depth
};

// reprojection_factor.h:25-90
// tb: the two tangent vectors b1, b2 of the target bearing (lie_algebra.cpp:47-56), precomputed per observation
XV_HD void rpe_eval(const Q4 &qt, const V3 &pt, const Q4 &qr, const V3 &pr, double inv_depth,
                                         const V3 &zt, const V3 &zr, const double *tb, const RigParams &rg, bool want_jac,
                                         bool want_ref_jac, RpeEval &e) {
    const Q4 cq = ld4(rg.cam_q);
    const V3 cp = ld3(rg.cam_p);
    M3 T;
    T(0, 0) = tb[0]; T(1, 0) = tb[1]; T(2, 0) = tb[2];
    T(0, 1) = tb[3]; T(1, 1) = tb[4]; T(2, 1) = tb[5];
    T(0, 2) = zt.x; T(1, 2) = zt.y; T(2, 2) = zt.z;
    const V3 y_ref = (1.0 / inv_depth) * zr;           // z_ref / inv_depth
    const V3 y_ref_c = qrot(cq, y_ref) + cp;
    const V3 x = qrot(qr, y_ref_c) + pr;
    const V3 y_tgt_c = qrot(qconj(qt), x - pt);
    const V3 y_tgt = qrot(qconj(cq), y_tgt_c - cp);
    const V3 u = transpose(T) * y_tgt;
    e.r[0] = rg.sic[0] * (u.x / u.z);
    e.r[1] = rg.sic[1] * (u.y / u.z);
    if (!want_jac) return;
    // dr_dy_tgt = sqrt_inv_cov * dproj_dp(u) * T^T
    const double iz = 1.0 / u.z, iz2 = 1.0 / (u.z * u.z);
    double D[2][3];   // sqrt_inv_cov * dproj
    D[0][0] = rg.sic[0] * iz; D[0][1] = 0.0; D[0][2] = -rg.sic[0] * u.x * iz2;
    D[1][0] = 0.0; D[1][1] = rg.sic[1] * iz; D[1][2] = -rg.sic[1] * u.y * iz2;
    double A[2][3];   // dr_dy_tgt
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) A[r][c] = D[r][0] * T(c, 0) + D[r][1] * T(c, 1) + D[r][2] * T(c, 2);
    const M3 RcT = qmat(qconj(cq));
    double B[2][3];   // dr_dy_tgt_center
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) B[r][c] = A[r][0] * RcT(0, c) + A[r][1] * RcT(1, c) + A[r][2] * RcT(2, c);
    const M3 RtT = qmat(qconj(qt));
    double X[2][3];   // dr_dx
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) X[r][c] = B[r][0] * RtT(0, c) + B[r][1] * RtT(1, c) + B[r][2] * RtT(2, c);
    const M3 Hy = hat(y_tgt_c);
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) {
            e.Jt[r * 6 + c] = B[r][0] * Hy(0, c) + B[r][1] * Hy(1, c) + B[r][2] * Hy(2, c);
            e.Jt[r * 6 + 3 + c] = -X[r][c];
        }
    if (!want_ref_jac) return;
    const M3 Rr = qmat(qr);
    double Y[2][3];   // dr_dy_ref_center
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) Y[r][c] = X[r][0] * Rr(0, c) + X[r][1] * Rr(1, c) + X[r][2] * Rr(2, c);
    const M3 Hr = hat(y_ref_c);
    const V3 cy = qmat(cq) * y_ref;
    for (int r = 0; r < 2; ++r) {
        for (int c = 0; c < 3; ++c) {
            e.Jr[r * 6 + c] = -(Y[r][0] * Hr(0, c) + Y[r][1] * Hr(1, c) + Y[r][2] * Hr(2, c));
            e.Jr[r * 6 + 3 + c] = X[r][c];
        }
        e.Jl[r] = -(Y[r][0] * cy.x + Y[r][1] * cy.y + Y[r][2] * cy.z) / inv_depth;
    }
}

// rotation_factor.h:23-59 ; Jq: 2x3
XV_HD void rot_eval(const Q4 &qt, const Q4 &qr, const V3 &zt, const V3 &zr, const RigParams &rg,
                                         bool want_jac, double r[2], double Jq[6]) {
    const Q4 cq = ld4(rg.cam_q);
    const V3 cp = ld3(rg.cam_p);
    const M3 T = local_tangent(zt);
    const V3 z_ref_c = qrot(cq, zr) + cp;
    const V3 z_tgt_c = qrot(qmul(qconj(qt), qr), z_ref_c);
    const V3 z_t = qrot(qconj(cq), z_tgt_c - cp);
    const V3 u = transpose(T) * z_t;
    r[0] = rg.sic[0] * (u.x / u.z);
    r[1] = rg.sic[1] * (u.y / u.z);
    if (!want_jac) return;
    const double iz = 1.0 / u.z, iz2 = 1.0 / (u.z * u.z);
    double D[2][3];
    D[0][0] = rg.sic[0] * iz; D[0][1] = 0.0; D[0][2] = -rg.sic[0] * u.x * iz2;
    D[1][0] = 0.0; D[1][1] = rg.sic[1] * iz; D[1][2] = -rg.sic[1] * u.y * iz2;
    double A[2][3];
    for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 3; ++c) A[a][c] = D[a][0] * T(c, 0) + D[a][1] * T(c, 1) + D[a][2] * T(c, 2);
    const M3 RcT = qmat(qconj(cq));
    double B[2][3];
    for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 3; ++c) B[a][c] = A[a][0] * RcT(0, c) + A[a][1] * RcT(1, c) + A[a][2] * RcT(2, c);
    const M3 Hz = hat(z_tgt_c);
    for (int a = 0; a < 2; ++a)
        for (int c = 0; c < 3; ++c) Jq[a * 3 + c] = B[a][0] * Hz(0, c) + B[a][1] * Hz(1, c) + B[a][2] * Hz(2, c);
}

XV_HD void put3x3(double *J, int r0, int c0, const M3 &A) {   // J is 15x15 row-major
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) J[(r0 + r) * 15 + c0 + c] = A(r, c);
}

// preintegration_factor.h:20-159.  si / sj: 16-double states (candidate values); bias0: frame_i's
// bg (3) and ba (3) in user memory (the D10 quirk, :38-39).  r: UNWHITENED residual (15);
// Ji / Jj: UNWHITENED 15x15 Jacobians in error-state order (caller multiplies by sqrt_inv_cov).
// part: -1 everything; 0 / 1 / 2 only the rotation / position / velocity (+ bias) rows of the residual (cost-only evaluations
// spread the three independent chains of a factor over three lanes; each row is computed by exactly the same operations)
__host__ __device__ inline void pre_eval(const double *si, const double *sj, const double *bias0, const double *pb,
                                const RigParams &rg, bool want_jac, double *r, double *Ji, double *Jj, int part = -1) {
    const V3 gravity = v3(0.0, 0.0, -9.80665);
    const Q4 q_ci = ld4(si), q_cj = ld4(sj);
    const V3 p_ci = ld3(si + 4), v_i = ld3(si + 7), bg_i = ld3(si + 10), ba_i = ld3(si + 13);
    const V3 p_cj = ld3(sj + 4), v_j = ld3(sj + 7), bg_j = ld3(sj + 10), ba_j = ld3(sj + 13);
    const Q4 iq = ld4(rg.imu_q);
    const V3 ip = ld3(rg.imu_p);
    const Q4 q_i = qmul(q_ci, iq), q_j = qmul(q_cj, iq);
    const V3 p_i = p_ci + qrot(q_ci, ip), p_j = p_cj + qrot(q_cj, ip);
    const double dt = pb[PB_T];
    const Q4 dq = ld4(pb + PB_Q);
    const V3 dp = ld3(pb + PB_P), dv = ld3(pb + PB_V);
    const V3 dbg = bg_i - ld3(bias0), dba = ba_i - ld3(bias0 + 3);
    M3 dq_dbg, dp_dbg, dp_dba, dv_dbg, dv_dba;
    for (int k = 0; k < 9; ++k) {
        dq_dbg.m[k] = pb[PB_DQ_DBG + k]; dp_dbg.m[k] = pb[PB_DP_DBG + k]; dp_dba.m[k] = pb[PB_DP_DBA + k];
        dv_dbg.m[k] = pb[PB_DV_DBG + k]; dv_dba.m[k] = pb[PB_DV_DBA + k];
    }
    const V3 th = dq_dbg * dbg;
    if (part == 0) { st3(r, logmap(qmul(qmul(qconj(qmul(dq, expmap(th))), qconj(q_i)), q_j))); return; }
    if (part == 1) { st3(r + 3, qrot(qconj(q_i), p_j - p_i - dt * v_i - (0.5 * dt * dt) * gravity) - (dp + dp_dbg * dbg + dp_dba * dba)); return; }
    if (part == 2) {
        st3(r + 6, qrot(qconj(q_i), v_j - v_i - dt * gravity) - (dv + dv_dbg * dbg + dv_dba * dba));
        st3(r + 9, bg_j - bg_i); st3(r + 12, ba_j - ba_i);
        return;
    }
    const V3 rq = logmap(qmul(qmul(qconj(qmul(dq, expmap(th))), qconj(q_i)), q_j));
    const V3 rp = qrot(qconj(q_i), p_j - p_i - dt * v_i - (0.5 * dt * dt) * gravity) - (dp + dp_dbg * dbg + dp_dba * dba);
    const V3 rv = qrot(qconj(q_i), v_j - v_i - dt * gravity) - (dv + dv_dbg * dbg + dv_dba * dba);
    st3(r, rq); st3(r + 3, rp); st3(r + 6, rv); st3(r + 9, bg_j - bg_i); st3(r + 12, ba_j - ba_i);
    if (!want_jac) return;
    for (int k = 0; k < 225; ++k) { Ji[k] = 0.0; Jj[k] = 0.0; }
    const M3 Jr_inv = inverse(right_jacobian(rq));
    const M3 RiT = qmat(qconj(q_i));
    const M3 Rci = qmat(q_ci);
    const M3 RimuT = qmat(qconj(iq));
    const M3 I3 = m3_identity();
    put3x3(Ji, 0, 0, -(Jr_inv * qmat(qconj(q_j)) * Rci));
    put3x3(Ji, 3, 0, RimuT * hat(qrot(qconj(q_ci), p_j - p_ci - dt * v_i - (0.5 * dt * dt) * gravity)));
    put3x3(Ji, 6, 0, RimuT * hat(qrot(qconj(q_ci), v_j - v_i - dt * gravity)));
    put3x3(Ji, 3, 3, -RiT);
    put3x3(Ji, 3, 6, (-dt) * RiT);
    put3x3(Ji, 6, 6, -RiT);
    put3x3(Ji, 0, 9, -(Jr_inv * qmat(qconj(expmap(rq))) * right_jacobian(th) * dq_dbg));
    put3x3(Ji, 3, 9, -dp_dbg);
    put3x3(Ji, 6, 9, -dv_dbg);
    put3x3(Ji, 9, 9, -I3);
    put3x3(Ji, 3, 12, -dp_dba);
    put3x3(Ji, 6, 12, -dv_dba);
    put3x3(Ji, 12, 12, -I3);
    put3x3(Jj, 0, 0, Jr_inv * RimuT);
    put3x3(Jj, 3, 0, -(RiT * qmat(q_cj) * hat(ip)));
    put3x3(Jj, 3, 3, RiT);
    put3x3(Jj, 6, 6, RiT);
    put3x3(Jj, 9, 9, I3);
    put3x3(Jj, 12, 12, I3);
}

// marginalization_factor.h:29-47 : error of one frame against its linearisation point (15), and
// right_jacobian(rq)^-1 (:53-57).
XV_HD void marg_dx(const double *s, const double *lin, double *dx, M3 *Jr_inv) {
    const V3 rq = logmap(qmul(qconj(ld4(lin)), ld4(s)));
    st3(dx, rq);
    for (int k = 0; k < 12; ++k) dx[3 + k] = s[4 + k] - lin[4 + k];
    if (Jr_inv) *Jr_inv = inverse(right_jacobian(rq));
}

}  // namespace xvbe
