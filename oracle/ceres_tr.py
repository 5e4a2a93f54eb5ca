"""Restatement of the Ceres 1.14 trust-region loop as the reference configures it
(xrslam/src/xrslam/estimation/solver.cpp:176-190: SPARSE_SCHUR, DOGLEG, num_threads 1,
update_state_every_iteration = true, max_num_iterations from config; CauchyLoss(1.0) on all
visual residuals, no loss on IMU / prior residuals, solver.cpp:36-37,112-174).

Ceres Solver 1.14.0 (git e809cf0c, reference cmake/depends/ceres-solver.cmake:8) is a
third-party dependency that is NOT under /root/reference and not installed on this box.
This file restates its published algorithm (trust_region_minimizer.cc, dogleg_strategy.cc,
corrector.cc, loss_function.cc of 1.14) from documentation/memory.  PARITY UNPINNED for
the loop: the reference holds no test that pins a solver iterate or cost.

The linear solve is done on the full normal equations with a dense Cholesky; Ceres'
SPARSE_SCHUR eliminates the inverse-depth blocks first, which is the same linear system
solved exactly (differs in round-off only).

Why a window BA runs to the iteration limit -- the ordering argument, with the Ceres 1.14 function names.
The preintegration factor reads the linearisation bias from USER memory, not from the parameter block:
    dbg = bg_i(parameter) - frame_i->motion.bg,  dba likewise          (preintegration_factor.h:38-39,50-51)
and the reference sets update_state_every_iteration = true (solver.cpp:187), which installs Ceres' StateUpdatingCallback.
Per iteration TrustRegionMinimizer::Minimize (trust_region_minimizer.cc) runs, in this order,
    FinalizeIterationAndCheckIfMinimizerCanContinue()   if the previous step was successful and x_cost_ < minimum_cost_:
                                                        parameters_ = x_;  then RunCallbacks() -> StateUpdatingCallback:
                                                        if (summary.step_is_successful) StateVectorToParameterBlocks(parameters_)
                                                        + CopyParameterBlockStateToUserState()      [user memory := x]
    ComputeTrustRegionStep()                            DoglegStrategy::ComputeStep on the Jacobian of the LAST accepted point
    ComputeCandidatePointAndEvaluateCost()              evaluator_->Evaluate(candidate_x_, &candidate_cost_, no Jacobian)
    ParameterToleranceReached() / FunctionToleranceReached()
    IsStepSuccessful() ? HandleSuccessfulStep() : HandleUnsuccessfulStep()
and HandleSuccessfulStep() does x_ = candidate_x_ followed by EvaluateGradientAndJacobian() IMMEDIATELY, i.e. BEFORE the
callback of the next FinalizeIteration... has copied x_ to user memory.  Consequences for a window BA whose first step is
accepted at iteration 1:
  (1) x_cost_ and the Jacobian of the new point are evaluated with user memory still holding the PREVIOUS iterate, so the
      preintegration residuals see dbg = bg_new - bg_old (the bias-corrected delta, the small cost);
  (2) then the callback rewrites user memory to the new iterate; every later candidate is evaluated with
      dbg = bg_cand - bg_new ~ 0 against deltas that were integrated with the bias at solve start
      (sliding_window_tracker.cpp:315-317), i.e. WITHOUT the first-order correction the accepted cost had -- the same point
      now costs more than x_cost_;
  (3) so relative_decrease = (x_cost_ - candidate_cost_) / model_cost_change is negative for every candidate, however small
      the radius: HandleUnsuccessfulStep() halves the radius (1e4 -> 1e4 / 2^29, far above min_trust_region_radius 1e-32),
      |cost_change| stays above function_tolerance * x_cost_ and the step norm above the parameter tolerance, and the solve
      ends at max_num_iterations = 30 (solver.cpp:181) with the iterate of iteration 1.
`solve` below reproduces (1)-(3) through `finalize` (callback order) and by evaluating the accepted point before it; the CUDA
solver does the same (backend_solve.cu, "HandleSuccessfulStep: user memory still holds the previous iterate").  That the
behaviour really occurs on every window BA of several streams is asserted by tests/test_oracle.py::
test_window_ba_runs_to_the_iteration_limit (oracle, 3 streams) and tests/test_gpu_pipeline.py::
test_window_solves_hit_the_iteration_limit (CUDA, xrvio_stats.window_iterations / window_at_limit).  It remains an argument
about remembered third-party source: Ceres is not available offline, which is why the loop stays "parity unpinned".

Oracle code: test infrastructure only.
"""
import numpy as np
import scipy.sparse as sp
import scipy.linalg as sla
from .lie import qmul, qconj, qrot, qmat, expmap, logmap, qnormalized, right_jacobian, hat, s2_tangential_basis
from . import factors as F
from .preintegrator import GRAVITY

# ---- Ceres 1.14 defaults (Solver::Options) -------------------------------------------------
INITIAL_RADIUS = 1e4
MIN_RADIUS = 1e-32
MIN_RELATIVE_DECREASE = 1e-3
FUNCTION_TOLERANCE = 1e-6
GRADIENT_TOLERANCE = 1e-10
PARAMETER_TOLERANCE = 1e-8
MAX_CONSECUTIVE_INVALID = 5
MIN_LM_DIAGONAL = 1e-6
MAX_LM_DIAGONAL = 1e32
MIN_MU, MAX_MU, MU_INCREASE = 1e-8, 1.0, 10.0


# ---- vectorised quaternion helpers ---------------------------------------------------------
def _qrot_b(q, v):
    u = q[..., :3]
    t = 2.0 * np.cross(u, v)
    return v + q[..., 3:4] * t + np.cross(u, t)


def _qconj_b(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _qmat_b(q):
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - (tyy + tzz); R[..., 0, 1] = txy - twz; R[..., 0, 2] = txz + twy
    R[..., 1, 0] = txy + twz; R[..., 1, 1] = 1 - (txx + tzz); R[..., 1, 2] = tyz - twx
    R[..., 2, 0] = txz - twy; R[..., 2, 1] = tyz + twx; R[..., 2, 2] = 1 - (txx + tyy)
    return R


def _hat_b(w):
    H = np.zeros(w.shape[:-1] + (3, 3))
    H[..., 0, 1] = -w[..., 2]; H[..., 0, 2] = w[..., 1]
    H[..., 1, 0] = w[..., 2]; H[..., 1, 2] = -w[..., 0]
    H[..., 2, 0] = -w[..., 1]; H[..., 2, 1] = w[..., 0]
    return H


def _tangent_b(z):
    """[b1 b2 z] per row (lie_algebra.cpp:47-56)."""
    n = z.shape[0]
    a = np.abs(z)
    d = np.zeros(n, dtype=np.int64)
    for i in (1, 2):
        d = np.where(a[:, i] > a[np.arange(n), d], i, d)
    e = np.zeros((n, 3))
    e[np.arange(n), (d + 1) % 3] = 1.0
    b1 = np.cross(z, e)
    b1 /= np.linalg.norm(b1, axis=1, keepdims=True)
    b2 = np.cross(z, b1)
    b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    T = np.stack([b1, b2, z], axis=2)  # (n,3,3) columns b1,b2,z
    return T


def reprojection_batch(q_tgt, p_tgt, q_ref, p_ref, inv_depth, z_tgt, z_ref, cam_q, cam_p, sic, want_jac=True):
    """Vectorised reprojection_factor.h:25-90.  sic: (2,) diagonal.  Returns r (n,2) and
    Jacobians J_tgt (n,2,6) [q,p], J_ref (n,2,6), J_lm (n,2)."""
    T = _tangent_b(z_tgt)
    y_ref = z_ref / inv_depth[:, None]
    y_ref_c = _qrot_b(cam_q[None], y_ref) + cam_p
    x = _qrot_b(q_ref, y_ref_c) + p_ref
    y_tgt_c = _qrot_b(_qconj_b(q_tgt), x - p_tgt)
    y_tgt = _qrot_b(_qconj_b(cam_q)[None], y_tgt_c - cam_p)
    u = np.einsum('nji,nj->ni', T, y_tgt)
    r = u[:, :2] / u[:, 2:3]
    if not want_jac:
        return r * sic, None, None, None
    n = len(u)
    dproj = np.zeros((n, 2, 3))
    dproj[:, 0, 0] = 1.0 / u[:, 2]; dproj[:, 0, 2] = -u[:, 0] / (u[:, 2] * u[:, 2])
    dproj[:, 1, 1] = 1.0 / u[:, 2]; dproj[:, 1, 2] = -u[:, 1] / (u[:, 2] * u[:, 2])
    dr_dy_tgt = sic[None, :, None] * np.einsum('nij,nkj->nik', dproj, T)   # dproj @ T^T
    Rc_T = qmat(qconj(cam_q))
    dr_dy_tgt_c = dr_dy_tgt @ Rc_T
    dr_dx = dr_dy_tgt_c @ _qmat_b(_qconj_b(q_tgt))
    dr_dy_ref_c = dr_dx @ _qmat_b(q_ref)
    J_tgt = np.concatenate([dr_dy_tgt_c @ _hat_b(y_tgt_c), -dr_dx], axis=2)
    J_ref = np.concatenate([-dr_dy_ref_c @ _hat_b(y_ref_c), dr_dx], axis=2)
    J_lm = -np.einsum('nij,nj->ni', dr_dy_ref_c @ qmat(cam_q), y_ref) / inv_depth[:, None]
    return r * sic, J_tgt, J_ref, J_lm


class _Layout:
    """Reduced program: which blocks are free and where their local columns live."""

    def __init__(self, pb):
        F_ = pb.states.shape[0]
        M = pb.inv_depth.shape[0]
        used_pose = np.zeros(F_, bool)
        used_motion = np.zeros(F_, bool)
        used_lm = np.zeros(M, bool)
        full = ~pb.rpe_prior
        used_pose[pb.rpe_tgt] = True
        used_pose[pb.rpe_ref[full]] = True
        used_lm[pb.rpe_lm[full]] = True
        used_pose[pb.rot_tgt] = True
        for f in pb.pre:
            used_pose[f.j] = used_motion[f.j] = True
            if not f.prior:
                used_pose[f.i] = used_motion[f.i] = True
        if pb.marg is not None:
            used_pose[pb.marg.frames] = True
            used_motion[pb.marg.frames] = True
        self.free_pose = used_pose & ~pb.fix_pose
        self.free_motion = used_motion & ~pb.fix_motion
        self.free_lm = used_lm
        self.pose_col = -np.ones(F_, np.int64)
        self.motion_col = -np.ones(F_, np.int64)
        self.lm_col = -np.ones(M, np.int64)
        c = 0
        for i in range(F_):
            if self.free_pose[i]:
                self.pose_col[i] = c; c += 6
            if self.free_motion[i]:
                self.motion_col[i] = c; c += 9
        self.n_frame_cols = c
        for k in range(M):
            if self.free_lm[k]:
                self.lm_col[k] = c; c += 1
        self.n_cols = c

    def frame_cols(self, i):
        """15 column indices (or -1) for error-state order q,p,v,bg,ba of frame i."""
        out = -np.ones(15, np.int64)
        if self.pose_col[i] >= 0:
            out[0:6] = self.pose_col[i] + np.arange(6)
        if self.motion_col[i] >= 0:
            out[6:15] = self.motion_col[i] + np.arange(9)
        return out


def _plus(pb, lay, states, inv_depth, delta):
    """evaluator->Plus: q <- normalize(q * exp(d)) (quaternion_parameterization.h:11-17), rest additive."""
    s = states.copy()
    d_ = inv_depth.copy()
    for i in range(s.shape[0]):
        if lay.pose_col[i] >= 0:
            c = lay.pose_col[i]
            s[i, 0:4] = qnormalized(qmul(s[i, 0:4], expmap(delta[c:c + 3])))
            s[i, 4:7] += delta[c + 3:c + 6]
        if lay.motion_col[i] >= 0:
            c = lay.motion_col[i]
            s[i, 7:16] += delta[c:c + 9]
    m = lay.lm_col >= 0
    d_[m] += delta[lay.lm_col[m]]
    return s, d_


def _ambient(pb, lay, states, inv_depth):
    parts = []
    for i in range(states.shape[0]):
        if lay.pose_col[i] >= 0:
            parts.append(states[i, 0:7])
        if lay.motion_col[i] >= 0:
            parts.append(states[i, 7:16])
    parts.append(inv_depth[lay.lm_col >= 0])
    return np.concatenate(parts) if parts else np.zeros(0)


def _st(s):
    return (s[0:4], s[4:7], s[7:10], s[10:13], s[13:16])


def evaluate(pb, lay, states, inv_depth, want_jac):
    """ProgramEvaluator::Evaluate at the candidate (states, inv_depth); constant blocks and
    'user memory' reads come from pb.states / pb.inv_depth.
    Returns cost, r (robustified), J (robustified, csr, local columns) or None, gradient or None."""
    user = pb.states
    rows_r = []
    trip_r, trip_c, trip_v = [], [], []
    cost = 0.0
    row0 = 0

    def add_block(r0, cols, Jblk):
        # Jblk (nr, len(cols)); cols may contain -1
        nr = Jblk.shape[0]
        keep = cols >= 0
        if not keep.any():
            return
        rr = np.repeat(np.arange(r0, r0 + nr), keep.sum())
        cc = np.tile(cols[keep], nr)
        trip_r.append(rr); trip_c.append(cc); trip_v.append(Jblk[:, keep].ravel())

    # -- reprojection (Cauchy loss) ---------------------------------------------------------
    O = len(pb.rpe_tgt)
    if O:
        pr = pb.rpe_prior
        qt = states[pb.rpe_tgt, 0:4]; pt = states[pb.rpe_tgt, 4:7]
        ref_src = np.where(pr[:, None], user[pb.rpe_ref], states[pb.rpe_ref])
        qr = ref_src[:, 0:4]; prf = ref_src[:, 4:7]
        invd = np.where(pr, pb.inv_depth[pb.rpe_lm], inv_depth[pb.rpe_lm])
        r, Jt, Jr, Jl = reprojection_batch(qt, pt, qr, prf, invd, pb.rpe_z_tgt, pb.rpe_z_ref,
                                           pb.cam_q_cs, pb.cam_p_cs, pb.sqrt_inv_cov_px, want_jac)
        s = np.sum(r * r, axis=1)
        cost += 0.5 * np.sum(np.log(1.0 + s))          # CauchyLoss(1): rho = log(1+s)
        w = np.sqrt(1.0 / (1.0 + s))                    # corrector: rho'' < 0 -> plain sqrt(rho')
        rows_r.append((r * w[:, None]).ravel())
        if want_jac:
            Jt = Jt * w[:, None, None]; Jr = Jr * w[:, None, None]; Jl = Jl * w[:, None]
            ridx = row0 + 2 * np.arange(O)
            tcol = lay.pose_col[pb.rpe_tgt]
            rcol = np.where(pr, -1, lay.pose_col[pb.rpe_ref])
            lcol = np.where(pr, -1, lay.lm_col[pb.rpe_lm])
            for (col0, Jb) in ((tcol, Jt), (rcol, Jr)):
                m = col0 >= 0
                if m.any():
                    rr = (ridx[m, None, None] + np.arange(2)[None, :, None] + np.zeros((1, 1, 6), np.int64))
                    cc = (col0[m, None, None] + np.zeros((1, 2, 1), np.int64) + np.arange(6)[None, None, :])
                    trip_r.append(rr.ravel()); trip_c.append(cc.ravel()); trip_v.append(Jb[m].ravel())
            m = lcol >= 0
            if m.any():
                rr = ridx[m, None] + np.arange(2)[None, :]
                cc = lcol[m, None] + np.zeros((1, 2), np.int64)
                trip_r.append(rr.ravel()); trip_c.append(cc.ravel()); trip_v.append(Jl[m].ravel())
        row0 += 2 * O

    # -- rotation priors (Cauchy loss; ref quaternion from user memory, rotation_factor.h:29) --
    sic2 = np.diag(pb.sqrt_inv_cov_px)
    for k in range(len(pb.rot_tgt)):
        t, rf = pb.rot_tgt[k], pb.rot_ref[k]
        r, J = F.rotation_prior(states[t, 0:4], user[rf, 0:4], pb.rot_z_tgt[k], pb.rot_z_ref[k],
                                pb.cam_q_cs, pb.cam_p_cs, sic2, want_jac)
        s = float(r @ r)
        cost += 0.5 * np.log(1.0 + s)
        w = np.sqrt(1.0 / (1.0 + s))
        rows_r.append(r * w)
        if want_jac:
            cols = -np.ones(3, np.int64)
            if lay.pose_col[t] >= 0:
                cols = lay.pose_col[t] + np.arange(3)
            add_block(row0, cols, J['q_tgt'] * w)
        row0 += 2

    # -- preintegration (no loss) -------------------------------------------------------------
    for f in pb.pre:
        si = _st(user[f.i]) if f.prior else _st(states[f.i])
        sj = _st(states[f.j])
        r, Ji, Jj = F.preintegration(si, sj, f, user[f.i, 10:13], user[f.i, 13:16],
                                     pb.imu_q_cs, pb.imu_p_cs, want_jac)
        cost += 0.5 * float(r @ r)
        rows_r.append(r)
        if want_jac:
            if not f.prior:
                add_block(row0, lay.frame_cols(f.i), Ji)
            add_block(row0, lay.frame_cols(f.j), Jj)
        row0 += 15

    # -- marginalisation prior (no loss) ------------------------------------------------------
    if pb.marg is not None:
        mg = pb.marg
        sts = [_st(states[i]) for i in mg.frames]
        lin = [_st(l) for l in mg.lin]
        r, J = F.marginalization_prior(sts, lin, mg.sqrt_inv_cov, mg.infovec, want_jac)
        cost += 0.5 * float(r @ r)
        rows_r.append(r)
        if want_jac:
            cols = np.concatenate([lay.frame_cols(i) for i in mg.frames])
            add_block(row0, cols, J)
        row0 += len(r)

    r_all = np.concatenate(rows_r) if rows_r else np.zeros(0)
    if not want_jac:
        return cost, r_all, None, None
    if trip_r:
        J = sp.csr_matrix((np.concatenate(trip_v), (np.concatenate(trip_r), np.concatenate(trip_c))),
                          shape=(row0, lay.n_cols))
    else:
        J = sp.csr_matrix((row0, lay.n_cols))
    g = J.T @ r_all
    return cost, r_all, J, g


class Summary:
    def __init__(self):
        self.iterations = []   # dicts: iteration, cost, cost_change, step_norm, relative_decrease, radius, successful, valid
        self.initial_cost = None
        self.final_cost = None
        self.termination = ''
        self.num_successful = 0
        self.num_unsuccessful = 0


def solve(pb, verbose=False):
    """ceres::Solve equivalent.  Mutates pb.states / pb.inv_depth ("user memory")."""
    lay = _Layout(pb)
    summ = Summary()
    n = lay.n_cols
    x_states = pb.states.copy()
    x_invd = pb.inv_depth.copy()
    if n == 0:
        c, _, _, _ = evaluate(pb, lay, x_states, x_invd, False)
        summ.initial_cost = summ.final_cost = c
        summ.termination = 'no parameters'
        return summ

    # dogleg strategy state
    radius = INITIAL_RADIUS
    mu = MIN_MU
    reuse = False
    dogleg_step_norm = 0.0
    diagonal = grad_s = gn_step = None
    alpha = 0.0

    # ---- iteration zero ----
    x_cost, r, J, g = evaluate(pb, lay, x_states, x_invd, True)
    col_sq = np.asarray(J.multiply(J).sum(axis=0)).ravel()
    jac_scaling = 1.0 / (1.0 + np.sqrt(col_sq))
    Js = J @ sp.diags(jac_scaling)
    x_norm = np.linalg.norm(_ambient(pb, lay, x_states, x_invd))
    minimum_cost = x_cost
    summ.initial_cost = x_cost
    it = dict(iteration=0, cost=x_cost, cost_change=0.0, step_norm=0.0, relative_decrease=0.0,
              radius=radius, successful=True, valid=True, gradient_max_norm=float(np.max(np.abs(g))))
    num_consecutive_invalid = 0

    def finalize(it):
        nonlocal minimum_cost
        if it['successful']:
            summ.num_successful += 1
            if x_cost < minimum_cost or it['iteration'] == 0:
                minimum_cost = min(minimum_cost, x_cost)
                # parameters_ <- x ; StateUpdatingCallback copies it to user memory
                pb.states[:] = x_states
                pb.inv_depth[:] = x_invd
        else:
            summ.num_unsuccessful += 1
        it['radius'] = radius
        summ.iterations.append(it)
        if verbose:
            print(it)
        if it['iteration'] >= pb.max_iterations:
            summ.termination = 'max iterations'; return False
        if it['gradient_max_norm'] <= GRADIENT_TOLERANCE:
            summ.termination = 'gradient tolerance'; return False
        if radius <= MIN_RADIUS:
            summ.termination = 'min radius'; return False
        return True

    while finalize(it):
        prev = it
        it = dict(iteration=prev['iteration'] + 1, cost=prev['cost'], cost_change=0.0, step_norm=0.0,
                  relative_decrease=0.0, radius=radius, successful=False, valid=False,
                  gradient_max_norm=prev['gradient_max_norm'])
        # ---------------- DoglegStrategy::ComputeStep ----------------
        solver_failed = False
        if not reuse:
            reuse = True
            d2 = np.asarray(Js.multiply(Js).sum(axis=0)).ravel()
            diagonal = np.sqrt(np.minimum(np.maximum(d2, MIN_LM_DIAGONAL), MAX_LM_DIAGONAL))
            grad_s = (Js.T @ r) / diagonal
            Jg = Js @ (grad_s / diagonal)
            alpha = float(grad_s @ grad_s) / float(Jg @ Jg)
            H = (Js.T @ Js).toarray()
            rhs = Js.T @ r
            solver_failed = True
            while mu < MAX_MU:
                lm = diagonal * np.sqrt(mu)
                A = H + np.diag(lm * lm)
                ok = True
                try:
                    cf = sla.cho_factor(A, lower=True, check_finite=False)
                    y = sla.cho_solve(cf, rhs, check_finite=False)
                    if not np.all(np.isfinite(y)):
                        ok = False
                except sla.LinAlgError:
                    ok = False
                if not ok:
                    mu *= MU_INCREASE
                    continue
                solver_failed = False
                gn_step = y * (-diagonal)
                break
        if solver_failed:
            # LINEAR_SOLVER_FAILURE: step invalid
            step_valid = False
        else:
            # ComputeTraditionalDoglegStep
            gnorm = np.linalg.norm(grad_s)
            gn_norm = np.linalg.norm(gn_step)
            if gn_norm <= radius:
                step = gn_step.copy(); dogleg_step_norm = gn_norm
            elif gnorm * alpha >= radius:
                step = -(radius / gnorm) * grad_s; dogleg_step_norm = radius
            else:
                b_dot_a = -alpha * float(grad_s @ gn_step)
                a_sq = (alpha * gnorm) ** 2
                bma_sq = a_sq - 2 * b_dot_a + gn_norm ** 2
                c = b_dot_a - a_sq
                d = np.sqrt(c * c + bma_sq * (radius ** 2 - a_sq))
                beta = (d - c) / bma_sq if c <= 0 else (radius * radius - a_sq) / (d + c)
                step = (-alpha * (1.0 - beta)) * grad_s + beta * gn_step
                dogleg_step_norm = np.linalg.norm(step)
            step = step / diagonal
            model_res = Js @ step
            model_cost_change = -float(model_res @ (r + model_res / 2.0))
            step_valid = model_cost_change > 0.0
        it['valid'] = step_valid
        if not step_valid:
            # HandleInvalidStep
            num_consecutive_invalid += 1
            if num_consecutive_invalid >= MAX_CONSECUTIVE_INVALID:
                summ.termination = 'too many invalid steps'
                summ.iterations.append(it)
                break
            mu *= MU_INCREASE; reuse = False      # StepIsInvalid
            continue
        num_consecutive_invalid = 0
        delta = step * jac_scaling
        # ---------------- candidate ----------------
        c_states, c_invd = _plus(pb, lay, x_states, x_invd, delta)
        cand_cost, _, _, _ = evaluate(pb, lay, c_states, c_invd, False)
        if not np.isfinite(cand_cost):
            # ComputeCandidatePointAndEvaluateCost: a failed evaluation (non-finite residuals) leaves candidate_cost =
            # numeric_limits<double>::max(), so the step is rejected and the radius halved instead of ending the solve
            cand_cost = np.finfo(np.float64).max
        # ParameterToleranceReached
        it['step_norm'] = float(np.linalg.norm(_ambient(pb, lay, x_states, x_invd) -
                                               _ambient(pb, lay, c_states, c_invd)))
        if it['step_norm'] <= PARAMETER_TOLERANCE * (x_norm + PARAMETER_TOLERANCE):
            summ.termination = 'parameter tolerance'; summ.iterations.append(it); break
        # FunctionToleranceReached
        it['cost_change'] = x_cost - cand_cost
        if abs(it['cost_change']) <= FUNCTION_TOLERANCE * x_cost:
            summ.termination = 'function tolerance'; summ.iterations.append(it); break
        it['relative_decrease'] = (x_cost - cand_cost) / model_cost_change
        if it['relative_decrease'] > MIN_RELATIVE_DECREASE:
            # HandleSuccessfulStep (user memory is still the previous iterate here: D10)
            x_states, x_invd = c_states, c_invd
            x_norm = np.linalg.norm(_ambient(pb, lay, x_states, x_invd))
            x_cost, r, J, g = evaluate(pb, lay, x_states, x_invd, True)
            Js = J @ sp.diags(jac_scaling)
            it['cost'] = x_cost
            it['gradient_max_norm'] = float(np.max(np.abs(g)))
            it['successful'] = True
            q_ = it['relative_decrease']
            if q_ < 0.25:
                radius *= 0.5
            if q_ > 0.75:
                radius = max(radius, 3.0 * dogleg_step_norm)
            mu = max(MIN_MU, 2.0 * mu / MU_INCREASE)
            reuse = False
        else:
            it['successful'] = False
            it['cost'] = cand_cost
            radius *= 0.5
            reuse = True
    summ.final_cost = minimum_cost
    return summ
