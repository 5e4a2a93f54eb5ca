"""Residuals and analytic Jacobians of the sliding-window factors (K7, K8, K8b and the
marginalisation-prior residual of K12).  float64.

Follows
  xrslam/src/xrslam/estimation/ceres/reprojection_factor.h:13-123
  xrslam/src/xrslam/estimation/ceres/preintegration_factor.h:12-199
  xrslam/src/xrslam/estimation/ceres/rotation_factor.h:12-67
  xrslam/src/xrslam/estimation/ceres/marginalization_factor.h:27-72
Jacobians w.r.t. a quaternion block are w.r.t. the local right-multiplicative
perturbation (first 3 of Ceres' 4 columns; the 4th is zero,
quaternion_parameterization.h:18-24).
Oracle code: test infrastructure only.
"""
import numpy as np
from .lie import (hat, expmap, logmap, qconj, qmat, qmul, qrot, right_jacobian,
                  local_tangent)
from .preintegrator import GRAVITY


def dproj_dp(p):
    # geometry/stereo.h:17-21
    return np.array([[1.0 / p[2], 0.0, -p[0] / (p[2] * p[2])],
                     [0.0, 1.0 / p[2], -p[1] / (p[2] * p[2])]])


def reprojection(q_tgt, p_tgt, q_ref, p_ref, inv_depth, z_tgt, z_ref,
                 cam_q_cs, cam_p_cs, sqrt_inv_cov, want_jac=True):
    """reprojection_factor.h:25-90.  Returns r(2) and dict of Jacobians
    'q_tgt'(2x3) 'p_tgt'(2x3) 'q_ref'(2x3) 'p_ref'(2x3) 'inv_depth'(2,)."""
    T = local_tangent(z_tgt)
    y_ref = z_ref / inv_depth
    y_ref_center = qrot(cam_q_cs, y_ref) + cam_p_cs
    x = qrot(q_ref, y_ref_center) + p_ref
    y_tgt_center = qrot(qconj(q_tgt), x - p_tgt)
    y_tgt = qrot(qconj(cam_q_cs), y_tgt_center - cam_p_cs)
    u_tgt = T.T @ y_tgt
    r = u_tgt[:2] / u_tgt[2]
    J = None
    if want_jac:
        dr_dy_tgt = sqrt_inv_cov @ dproj_dp(u_tgt) @ T.T
        dr_dy_tgt_center = dr_dy_tgt @ qmat(qconj(cam_q_cs))
        dr_dx = dr_dy_tgt_center @ qmat(qconj(q_tgt))
        dr_dy_ref_center = dr_dx @ qmat(q_ref)
        J = {
            'q_tgt': dr_dy_tgt_center @ hat(y_tgt_center),
            'p_tgt': -dr_dx,
            'q_ref': -dr_dy_ref_center @ hat(y_ref_center),
            'p_ref': dr_dx,
            'inv_depth': -(dr_dy_ref_center @ qmat(cam_q_cs) @ y_ref) / inv_depth,
        }
    return sqrt_inv_cov @ r, J


def rotation_prior(q_tgt, q_ref, z_tgt, z_ref, cam_q_cs, cam_p_cs, sqrt_inv_cov, want_jac=True):
    """rotation_factor.h:23-59 (note: p_cs is added to a direction vector there; copied)."""
    T = local_tangent(z_tgt)
    z_ref_center = qrot(cam_q_cs, z_ref) + cam_p_cs
    z_tgt_center = qrot(qmul(qconj(q_tgt), q_ref), z_ref_center)
    z_t = qrot(qconj(cam_q_cs), z_tgt_center - cam_p_cs)
    u = T.T @ z_t
    r = u[:2] / u[2]
    J = None
    if want_jac:
        dr_dz = sqrt_inv_cov @ dproj_dp(u) @ T.T
        dr_dzc = dr_dz @ qmat(qconj(cam_q_cs))
        J = {'q_tgt': dr_dzc @ hat(z_tgt_center)}
    return sqrt_inv_cov @ r, J


def preintegration(si, sj, pre, bg_i_0, ba_i_0, imu_q_cs, imu_p_cs, want_jac=True):
    """preintegration_factor.h:20-159.
    si, sj: dicts/tuples (q, p, v, bg, ba) of the two frames (candidate values).
    bg_i_0, ba_i_0: frame_i's bias *in user memory* (D10 quirk, :38-39).
    pre: object with t,q,p,v,sqrt_inv_cov,dq_dbg,dp_dbg,dp_dba,dv_dbg,dv_dba.
    Returns whitened r(15) and J_i(15x15), J_j(15x15) in error-state order."""
    q_ci, p_ci, v_i, bg_i, ba_i = si
    q_cj, p_cj, v_j, bg_j, ba_j = sj
    q_i = qmul(q_ci, imu_q_cs)
    p_i = p_ci + qrot(q_ci, imu_p_cs)
    q_j = qmul(q_cj, imu_q_cs)
    p_j = p_cj + qrot(q_cj, imu_p_cs)
    dt = pre.t
    dbg = bg_i - bg_i_0
    dba = ba_i - ba_i_0
    r = np.empty(15)
    r[0:3] = logmap(qmul(qmul(qconj(qmul(pre.q, expmap(pre.dq_dbg @ dbg))), qconj(q_i)), q_j))
    r[3:6] = qrot(qconj(q_i), p_j - p_i - dt * v_i - 0.5 * dt * dt * GRAVITY) - \
        (pre.p + pre.dp_dbg @ dbg + pre.dp_dba @ dba)
    r[6:9] = qrot(qconj(q_i), v_j - v_i - dt * GRAVITY) - (pre.v + pre.dv_dbg @ dbg + pre.dv_dba @ dba)
    r[9:12] = bg_j - bg_i
    r[12:15] = ba_j - ba_i
    Ji = Jj = None
    if want_jac:
        S = pre.sqrt_inv_cov
        Ji = np.zeros((15, 15))
        Jj = np.zeros((15, 15))
        Jr_inv = np.linalg.inv(right_jacobian(r[0:3]))
        Riq = qmat(qconj(q_i))
        Rci = qmat(q_ci)
        Rimu_T = qmat(qconj(imu_q_cs))
        # d/dq_i
        Ji[0:3, 0:3] = -Jr_inv @ qmat(qconj(q_j)) @ Rci
        Ji[3:6, 0:3] = Rimu_T @ hat(qrot(qconj(q_ci), p_j - p_ci - dt * v_i - 0.5 * dt * dt * GRAVITY))
        Ji[6:9, 0:3] = Rimu_T @ hat(qrot(qconj(q_ci), v_j - v_i - dt * GRAVITY))
        # d/dp_i
        Ji[3:6, 3:6] = -Riq
        # d/dv_i
        Ji[3:6, 6:9] = -dt * Riq
        Ji[6:9, 6:9] = -Riq
        # d/dbg_i
        Ji[0:3, 9:12] = -Jr_inv @ qmat(qconj(expmap(r[0:3]))) @ right_jacobian(pre.dq_dbg @ dbg) @ pre.dq_dbg
        Ji[3:6, 9:12] = -pre.dp_dbg
        Ji[6:9, 9:12] = -pre.dv_dbg
        Ji[9:12, 9:12] = -np.eye(3)
        # d/dba_i
        Ji[3:6, 12:15] = -pre.dp_dba
        Ji[6:9, 12:15] = -pre.dv_dba
        Ji[12:15, 12:15] = -np.eye(3)
        # d/dq_j
        Jj[0:3, 0:3] = Jr_inv @ Rimu_T
        Jj[3:6, 0:3] = -Riq @ qmat(q_cj) @ hat(imu_p_cs)
        Jj[3:6, 3:6] = Riq
        Jj[6:9, 6:9] = Riq
        Jj[9:12, 9:12] = np.eye(3)
        Jj[12:15, 12:15] = np.eye(3)
        Ji = S @ Ji
        Jj = S @ Jj
    return pre.sqrt_inv_cov @ r, Ji, Jj


def marginalization_prior(states, lin_states, sqrt_inv_cov, infovec, want_jac=True):
    """marginalization_factor.h:27-72.  states / lin_states: lists of (q,p,v,bg,ba).
    Returns r(15K) and J(15K x 15K) in error-state order per frame."""
    K = len(states)
    dx = np.empty(15 * K)
    B = np.eye(15 * K)
    for i in range(K):
        q, p, v, bg, ba = states[i]
        q0, p0, v0, bg0, ba0 = lin_states[i]
        rq = logmap(qmul(qconj(q0), q))
        dx[15 * i:15 * i + 3] = rq
        dx[15 * i + 3:15 * i + 6] = p - p0
        dx[15 * i + 6:15 * i + 9] = v - v0
        dx[15 * i + 9:15 * i + 12] = bg - bg0
        dx[15 * i + 12:15 * i + 15] = ba - ba0
        if want_jac:
            B[15 * i:15 * i + 3, 15 * i:15 * i + 3] = np.linalg.inv(right_jacobian(rq))
    r = sqrt_inv_cov @ dx + infovec
    J = sqrt_inv_cov @ B if want_jac else None
    return r, J
