"""IMU preintegration (K6).  Follows xrslam/src/xrslam/estimation/preintegrator.cpp:9-112.
Error-state order q 0-2, p 3-5, v 6-8, bg 9-11, ba 12-14 (estimation/state.h:12-19).
Oracle code: test infrastructure only.
"""
import numpy as np
from .lie import hat, expmap, qconj, qmat, qmul, qrot, qnormalized, right_jacobian

GRAVITY = np.array([0.0, 0.0, -9.80665])  # common.h:41
ES_Q, ES_P, ES_V, ES_BG, ES_BA, ES_SIZE = 0, 3, 6, 9, 12, 15


class PreIntegrator:
    def __init__(self, cov_w=None, cov_a=None, cov_bg=None, cov_ba=None):
        self.cov_w = np.zeros((3, 3)) if cov_w is None else np.array(cov_w, dtype=np.float64).reshape(3, 3)
        self.cov_a = np.zeros((3, 3)) if cov_a is None else np.array(cov_a, dtype=np.float64).reshape(3, 3)
        self.cov_bg = np.zeros((3, 3)) if cov_bg is None else np.array(cov_bg, dtype=np.float64).reshape(3, 3)
        self.cov_ba = np.zeros((3, 3)) if cov_ba is None else np.array(cov_ba, dtype=np.float64).reshape(3, 3)
        self.data = []  # list of (t, w[3], a[3])
        self.reset()

    def copy(self):
        o = PreIntegrator(self.cov_w, self.cov_a, self.cov_bg, self.cov_ba)
        o.data = [(t, w.copy(), a.copy()) for (t, w, a) in self.data]
        for k in ("t", "q", "p", "v", "cov", "sqrt_inv_cov", "dq_dbg", "dp_dbg", "dp_dba", "dv_dbg", "dv_dba"):
            val = getattr(self, k)
            setattr(o, k, val.copy() if isinstance(val, np.ndarray) else val)
        return o

    def reset(self):
        # preintegrator.cpp:9-20
        self.t = 0.0
        self.q = np.array([0.0, 0.0, 0.0, 1.0])
        self.p = np.zeros(3)
        self.v = np.zeros(3)
        self.cov = np.zeros((15, 15))
        self.sqrt_inv_cov = np.zeros((15, 15))
        self.dq_dbg = np.zeros((3, 3))
        self.dp_dbg = np.zeros((3, 3))
        self.dp_dba = np.zeros((3, 3))
        self.dv_dbg = np.zeros((3, 3))
        self.dv_dba = np.zeros((3, 3))

    def increment(self, dt, w_raw, a_raw, bg, ba, compute_jacobian, compute_covariance):
        # preintegrator.cpp:22-76
        assert dt >= 0
        w = w_raw - bg
        a = a_raw - ba
        R = qmat(self.q)
        if compute_covariance:
            A = np.eye(9)
            A[0:3, 0:3] = qmat(qconj(expmap(w * dt)))
            A[6:9, 0:3] = -dt * R @ hat(a)
            A[3:6, 0:3] = -0.5 * dt * dt * R @ hat(a)
            A[3:6, 6:9] = dt * np.eye(3)
            B = np.zeros((9, 6))
            B[0:3, 0:3] = dt * right_jacobian(w * dt)
            B[6:9, 3:6] = dt * R
            B[3:6, 3:6] = 0.5 * dt * dt * R
            inv_dt = 1.0 / max(dt, 1.0e-7)
            Q = np.zeros((6, 6))
            Q[0:3, 0:3] = self.cov_w * inv_dt
            Q[3:6, 3:6] = self.cov_a * inv_dt
            self.cov[0:9, 0:9] = A @ self.cov[0:9, 0:9] @ A.T + B @ Q @ B.T
            self.cov[9:12, 9:12] += self.cov_bg * dt
            self.cov[12:15, 12:15] += self.cov_ba * dt
        if compute_jacobian:
            self.dp_dbg = self.dp_dbg + dt * self.dv_dbg - 0.5 * dt * dt * R @ hat(a) @ self.dq_dbg
            self.dp_dba = self.dp_dba + dt * self.dv_dba - 0.5 * dt * dt * R
            self.dv_dbg = self.dv_dbg - dt * R @ hat(a) @ self.dq_dbg
            self.dv_dba = self.dv_dba - dt * R
            self.dq_dbg = qmat(qconj(expmap(w * dt))) @ self.dq_dbg - dt * right_jacobian(w * dt)
        qa = qrot(self.q, a)
        self.t = self.t + dt
        self.p = self.p + dt * self.v + 0.5 * dt * dt * qa
        self.v = self.v + dt * qa
        self.q = qnormalized(qmul(self.q, expmap(w * dt)))

    def integrate(self, t, bg, ba, compute_jacobian, compute_covariance):
        # preintegrator.cpp:78-95
        if len(self.data) == 0:
            return False
        self.reset()
        for i in range(len(self.data) - 1):
            ti, wi, ai = self.data[i]
            self.increment(self.data[i + 1][0] - ti, wi, ai, bg, ba, compute_jacobian, compute_covariance)
        tl, wl, al = self.data[-1]
        self.increment(t - tl, wl, al, bg, ba, compute_jacobian, compute_covariance)
        if compute_covariance:
            self.compute_sqrt_inv_cov()
        return True

    def compute_sqrt_inv_cov(self):
        # preintegrator.cpp:97-100: LLT(cov^-1).matrixL()^T  (upper triangular U with U^T U = cov^-1)
        L = np.linalg.cholesky(np.linalg.inv(self.cov))
        self.sqrt_inv_cov = L.T.copy()

    def predict(self, q_i, p_i, v_i, bg_i, ba_i):
        """preintegrator.cpp:102-112.  Returns (q_j, p_j, v_j, bg_j, ba_j)."""
        v_j = v_i + GRAVITY * self.t + qrot(q_i, self.v)
        p_j = p_i + 0.5 * GRAVITY * self.t * self.t + v_i * self.t + qrot(q_i, self.p)
        q_j = qmul(q_i, self.q)
        return q_j, p_j, v_j, bg_i.copy(), ba_i.copy()
