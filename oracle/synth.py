"""Synthetic EuRoC-shaped stream generator: a camera+IMU rig moving inside a
textured box room.  Images are ray-cast from procedural wall textures so that KLT tracks are
physically consistent with the trajectory.  Seeded with numpy default_rng(seed).

Oracle-side data generator: test / bench infrastructure only (the product never imports it).
Sensor constants are the reference's configs/euroc_sensor.yaml:3-53.
"""
import numpy as np
from .lie import qmat, qmul, expmap, qnormalized

GRAVITY = np.array([0.0, 0.0, -9.80665])

EUROC = dict(
    cov_g=2.8791302399999997e-08, cov_a=4.0e-6, cov_bg=3.7608844899999997e-10, cov_ba=9.0e-6,
    q_bc=np.array([-7.7071797555374275e-03, 1.0499323370587278e-02, 7.0175280029197162e-01, 7.1230146066895372e-01]),
    p_bc=np.array([-0.0216401454975, -0.064676986768, 0.00981073058949]),
    fx=458.654, fy=457.296, keypoint_cov=0.5,
)


def mat_to_quat(R):
    # robust rotation-matrix -> quaternion xyzw
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


class Stream:
    """One synthetic stream.  width x height pinhole camera, no distortion."""

    ROOM = np.array([6.0, 5.0, 2.5])   # half extents of the box room [m]
    TEX = 1024                         # texels per wall side

    def __init__(self, seed, width=640, height=480, cam_hz=20.0, imu_hz=200.0, static_prefix=0.0, pause=None):
        """pause = (t0, duration, ramp): the rig decelerates smoothly to a full stop at t0 + ramp, stands still for `duration`
        seconds and accelerates again (a time warp of the same trajectory), to exercise the pure-rotation / no-translation
        branches of the tracker (frame.cpp:132-134, sliding_window_tracker.cpp:370-465)."""
        self.seed = seed
        self.pause = pause
        self.w, self.h = width, height
        self.cam_dt, self.imu_dt = 1.0 / cam_hz, 1.0 / imu_hz
        self.imu_per_cam = int(round(imu_hz / cam_hz))
        rng = np.random.default_rng(seed)
        self.rng = rng
        self.K = np.array([EUROC['fx'], EUROC['fy'], width / 2.0, height / 2.0])
        self.q_bc = EUROC['q_bc'] / np.linalg.norm(EUROC['q_bc'])
        self.p_bc = EUROC['p_bc']
        self.R_bc = qmat(self.q_bc)
        # trajectory: sums of sinusoids (|v| <~ 1 m/s, |w| <~ 1 rad/s)
        nk = 4
        self.pa = rng.uniform(0.15, 0.45, (3, nk)) * np.array([[1.0], [1.0], [0.5]])
        self.pf = rng.uniform(0.05, 0.35, (3, nk)) * 2 * np.pi
        self.pp = rng.uniform(0, 2 * np.pi, (3, nk))
        self.aa = rng.uniform(0.05, 0.25, (3, nk)) * np.array([[0.5], [0.5], [1.5]])
        self.af = rng.uniform(0.05, 0.3, (3, nk)) * 2 * np.pi
        self.ap = rng.uniform(0, 2 * np.pi, (3, nk))
        self.yaw0 = rng.uniform(0, 2 * np.pi)
        # base rotation: camera z (optical axis) horizontal, camera y down
        R_wc0 = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
        self.R_wb0 = R_wc0 @ self.R_bc.T
        # IMU biases (random walk) are generated lazily per sample index
        self.bg0 = rng.normal(0, 2e-3, 3)
        self.ba0 = rng.normal(0, 2e-2, 3)
        # wall textures: 6 walls, blocky random gray levels + fine noise
        self.textures = self._make_textures(rng)
        self._imu_cache = None

    # ------------------------------------------------------------------ trajectory
    def _warp(self, t):
        """tau(t), tau'(t), tau''(t) of the pause time warp (identity without a pause)."""
        if self.pause is None:
            return t, 1.0, 0.0
        t0, dur, r = self.pause
        # speed profile s(t) = tau'(t): 1 -> 0 over [t0, t0+r] (quintic smoothstep), 0 during the pause, 0 -> 1 over the next r
        def sm(x):       # smoothstep with zero first and second derivative at both ends, its derivative and its integral
            x = min(max(x, 0.0), 1.0)
            return 6 * x ** 5 - 15 * x ** 4 + 10 * x ** 3, 30 * x ** 4 - 60 * x ** 3 + 30 * x ** 2, x ** 6 - 3 * x ** 5 + 2.5 * x ** 4
        t = float(t)
        if t <= t0:
            return t, 1.0, 0.0
        if t <= t0 + r:
            x = (t - t0) / r
            s, ds, I = sm(x)
            return t0 + r * (x - I), 1.0 - s, -ds / r
        tau_stop = t0 + r * 0.5                     # integral of (1 - smoothstep) over the ramp is r / 2
        if t <= t0 + r + dur:
            return tau_stop, 0.0, 0.0
        t1 = t0 + r + dur
        if t <= t1 + r:
            x = (t - t1) / r
            s, ds, I = sm(x)
            return tau_stop + r * I, s, ds / r
        return tau_stop + r * 0.5 + (t - t1 - r), 1.0, 0.0

    def _p(self, tau):
        tau = np.asarray(tau, dtype=np.float64)[..., None, None]
        return np.sum(self.pa * np.sin(self.pf * tau + self.pp), axis=-1)

    def _dp(self, tau):
        tau = np.asarray(tau, dtype=np.float64)[..., None, None]
        return np.sum(self.pa * self.pf * np.cos(self.pf * tau + self.pp), axis=-1)

    def _ddp(self, tau):
        tau = np.asarray(tau, dtype=np.float64)[..., None, None]
        return np.sum(-self.pa * self.pf ** 2 * np.sin(self.pf * tau + self.pp), axis=-1)

    def position(self, t):
        if self.pause is None:
            return self._p(t)
        return self._p(self._warp(t)[0])

    def velocity(self, t):
        if self.pause is None:
            return self._dp(t)
        tau, s, _ = self._warp(t)
        return self._dp(tau) * s

    def acceleration(self, t):
        if self.pause is None:
            return self._ddp(t)
        tau, s, ds = self._warp(t)
        return self._ddp(tau) * s * s + self._dp(tau) * ds

    def rotation(self, t):
        """R_wb(t) (3x3)."""
        if self.pause is not None:
            t = self._warp(t)[0]
        ang = np.sum(self.aa * np.sin(self.af * t + self.ap), axis=-1)  # roll, pitch, yaw offsets
        yaw = self.yaw0 + ang[2]
        cz, sz = np.cos(yaw), np.sin(yaw)
        Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
        cy, sy = np.cos(ang[1]), np.sin(ang[1])
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        cx, sx = np.cos(ang[0]), np.sin(ang[0])
        Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        return Rz @ Ry @ Rx @ self.R_wb0

    def angular_velocity_body(self, t, h=1e-5):
        R0, R1 = self.rotation(t - h), self.rotation(t + h)
        dR = (R1 - R0) / (2 * h)
        W = self.rotation(t).T @ dR
        return np.array([W[2, 1] - W[1, 2], W[0, 2] - W[2, 0], W[1, 0] - W[0, 1]]) * 0.5

    def gt_state(self, t):
        """(q_wb xyzw, p_wb, v_w) ground truth of the body (IMU) frame."""
        return mat_to_quat(self.rotation(t)), self.position(t), self.velocity(t)

    # ------------------------------------------------------------------ IMU
    def imu(self, n_samples):
        """(n,7) rows t,wx,wy,wz,ax,ay,az at 200 Hz starting at t=0, with noise and bias random walk."""
        if self._imu_cache is not None and len(self._imu_cache) >= n_samples:
            return self._imu_cache[:n_samples]
        rng = np.random.default_rng(self.seed * 7919 + 13)
        dt = self.imu_dt
        sg, sa = np.sqrt(EUROC['cov_g'] / dt), np.sqrt(EUROC['cov_a'] / dt)
        sbg, sba = np.sqrt(EUROC['cov_bg'] * dt), np.sqrt(EUROC['cov_ba'] * dt)
        out = np.empty((n_samples, 7))
        bg, ba = self.bg0.copy(), self.ba0.copy()
        for k in range(n_samples):
            t = k * dt
            R = self.rotation(t)
            w = self.angular_velocity_body(t)
            a = R.T @ (self.acceleration(t) - GRAVITY)
            out[k, 0] = t
            out[k, 1:4] = w + bg + rng.normal(0, sg, 3)
            out[k, 4:7] = a + ba + rng.normal(0, sa, 3)
            bg = bg + rng.normal(0, sbg, 3)
            ba = ba + rng.normal(0, sba, 3)
        self._imu_cache = out
        return out

    # ------------------------------------------------------------------ images
    def _make_textures(self, rng):
        T = self.TEX
        texs = []
        for wall in range(6):
            tex = np.zeros((T, T), np.float32)
            # coarse-to-fine random blocks -> many high-contrast corners
            for cell, amp in ((128, 40.0), (32, 60.0), (8, 25.0)):
                n = T // cell
                blk = rng.uniform(-1, 1, (n, n)).astype(np.float32)
                tex += amp * np.kron(blk, np.ones((cell, cell), np.float32))
            tex += rng.normal(0, 6.0, (T, T)).astype(np.float32)
            tex = 128.0 + tex
            texs.append(np.clip(tex, 5, 250))
        return np.stack(texs)

    def render(self, t, noise_seed=None):
        """u8 (h,w) image seen by the camera at time t."""
        R_wc = self.rotation(t) @ self.R_bc
        c = self.position(t) + self.rotation(t) @ self.p_bc
        fx, fy, cx, cy = self.K
        xs = (np.arange(self.w) - cx) / fx
        ys = (np.arange(self.h) - cy) / fy
        X, Y = np.meshgrid(xs, ys)
        d = np.stack([X, Y, np.ones_like(X)], axis=-1) @ R_wc.T      # (h,w,3) world dirs
        with np.errstate(divide='ignore'):
            tpos = (self.ROOM - c) / d
            tneg = (-self.ROOM - c) / d
        tt = np.where(d > 0, tpos, tneg)                              # exit distance per axis
        axis = np.argmin(tt, axis=-1)
        tmin = np.take_along_axis(tt, axis[..., None], axis=-1)[..., 0]
        hit = c + d * tmin[..., None]
        sign = np.take_along_axis(d, axis[..., None], axis=-1)[..., 0] > 0
        wall = axis * 2 + sign
        a1, a2 = (axis + 1) % 3, (axis + 2) % 3
        u = np.take_along_axis(hit, a1[..., None], axis=-1)[..., 0]
        v = np.take_along_axis(hit, a2[..., None], axis=-1)[..., 0]
        e1, e2 = self.ROOM[a1], self.ROOM[a2]
        T = self.TEX
        tu = (u + e1) / (2 * e1) * (T - 1)
        tv = (v + e2) / (2 * e2) * (T - 1)
        tu = np.clip(tu, 0, T - 1.001); tv = np.clip(tv, 0, T - 1.001)
        iu, iv = np.floor(tu).astype(np.int64), np.floor(tv).astype(np.int64)
        fu, fv = (tu - iu).astype(np.float32), (tv - iv).astype(np.float32)
        tex = self.textures
        val = (tex[wall, iv, iu] * (1 - fu) * (1 - fv) + tex[wall, iv, iu + 1] * fu * (1 - fv) +
               tex[wall, iv + 1, iu] * (1 - fu) * fv + tex[wall, iv + 1, iu + 1] * fu * fv)
        if noise_seed is not None:
            nrng = np.random.default_rng(noise_seed)
            val = val + nrng.normal(0, 1.5, val.shape).astype(np.float32)
        return np.clip(np.rint(val), 0, 255).astype(np.uint8)

    def frame_time(self, k):
        return k * self.cam_dt

    def frame(self, k):
        return self.render(self.frame_time(k), noise_seed=self.seed * 100003 + k)


def _render_chunk(args):
    stream, ks = args
    return [stream.frame(k) for k in ks]


def render_frames(stream, n_frames, workers=None):
    """frames 0..n_frames-1 of `stream`, rendered by a pool of forked workers (every frame has its own noise seed, so the
    result is identical to [stream.frame(k) ...]); the long parity tests use it so that 1200 frames take seconds on the
    GPU box's host cores instead of minutes on one."""
    import multiprocessing as mp
    import os
    workers = workers or max(1, min(os.cpu_count() or 1, 64, n_frames // 8))
    if workers <= 1:
        return [stream.frame(k) for k in range(n_frames)]
    chunks = [list(range(w, n_frames, workers)) for w in range(workers)]
    with mp.get_context('fork').Pool(workers) as pool:
        parts = pool.map(_render_chunk, [(stream, c) for c in chunks])
    out = [None] * n_frames
    for c, p in zip(chunks, parts):
        for k, f in zip(c, p):
            out[k] = f
    return out
