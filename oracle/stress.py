"""TEST INFRASTRUCTURE ONLY.  Generator of the large-window bundle-adjustment problem of BASELINE.json config 5 /
SURVEY.md section 8(d) row 5: N keyframes, M landmarks each observed in every keyframe (reference = keyframe 0,
M (N - 1) reprojection residuals), N - 1 preintegration edges, one dense 15 (N - 1) square marginalisation prior, fixed
iteration count.  The wiring is the one xrslam's Solver gives Ceres for a window (solver.cpp:84-174), only larger than the
10-keyframe window the tracker builds.  Seeded and deterministic, so the GPU box regenerates the same problem."""
import numpy as np
from .synth import Stream
from .pipeline import Config
from .preintegrator import PreIntegrator
from .problem import Problem, PreFactor, MargPrior
from .lie import qmat, qmul, expmap


def make_stress_problem(seed=5, n_kf=30, n_lm=2000, kf_gap=2, max_iterations=10, dense_prior=True):
    rng = np.random.default_rng(1000 + seed)
    st = Stream(seed)
    cfg = Config()
    t_kf = np.array([0.5 + st.cam_dt * kf_gap * k for k in range(n_kf)])
    imu = st.imu(int((t_kf[-1] + 0.1) / st.imu_dt) + 2)
    R_bc, p_bc = qmat(cfg.q_bc), cfg.p_bc

    gt = np.zeros((n_kf, 16))
    for k, t in enumerate(t_kf):
        q, p, v = st.gt_state(t)
        gt[k, 0:4], gt[k, 4:7], gt[k, 7:10] = q, p, v
        gt[k, 10:13], gt[k, 13:16] = st.bg0, st.ba0
    # landmarks: 3 - 8 m in front of the middle keyframe's camera, +-35 degrees
    mid = n_kf // 2
    R_wc = qmat(gt[mid, 0:4]) @ R_bc
    p_wc = gt[mid, 4:7] + qmat(gt[mid, 0:4]) @ p_bc
    depth = rng.uniform(3.0, 8.0, n_lm)
    ang = rng.uniform(-0.6, 0.6, (n_lm, 2))
    dirs = np.column_stack([np.tan(ang[:, 0]), np.tan(ang[:, 1]), np.ones(n_lm)])
    X = p_wc + (dirs * depth[:, None]) @ R_wc.T

    def bearing(k):
        Rwb = qmat(gt[k, 0:4])
        yb = (X - gt[k, 4:7]) @ Rwb            # R_wb^T (X - p)
        yc = (yb - p_bc) @ R_bc
        n = np.linalg.norm(yc, axis=1)
        return yc / n[:, None], n

    sig = 0.5 / cfg.K[0]                        # half a pixel of bearing noise
    z = []
    for k in range(n_kf):
        b, n = bearing(k)
        if k == 0:
            inv_gt = 1.0 / n
        b = b + rng.normal(0, sig, b.shape)
        z.append(b / np.linalg.norm(b, axis=1)[:, None])

    # initial iterate: ground truth + perturbation
    states = gt.copy()
    for k in range(n_kf):
        states[k, 0:4] = qmul(gt[k, 0:4], expmap(rng.normal(0, 5e-3, 3)))
        states[k, 0:4] /= np.linalg.norm(states[k, 0:4])
        states[k, 4:7] += rng.normal(0, 1e-2, 3)
        states[k, 7:10] += rng.normal(0, 1e-2, 3)
        states[k, 10:13] += rng.normal(0, 2e-4, 3)
        states[k, 13:16] += rng.normal(0, 2e-3, 3)
    inv_depth = inv_gt * (1.0 + rng.normal(0, 0.03, n_lm))

    pb = Problem(states=states, fix_pose=np.zeros(n_kf, bool), fix_motion=np.zeros(n_kf, bool), inv_depth=inv_depth,
                 cam_q_cs=cfg.q_bc.copy(), cam_p_cs=cfg.p_bc.copy(), imu_q_cs=cfg.q_bi.copy(), imu_p_cs=cfg.p_bi.copy(),
                 sqrt_inv_cov_px=cfg.sqrt_inv_cov(), max_iterations=max_iterations)
    tg = np.repeat(np.arange(1, n_kf), n_lm)
    lm = np.tile(np.arange(n_lm), n_kf - 1)
    pb.rpe_tgt = tg.astype(np.int32)
    pb.rpe_ref = np.zeros_like(pb.rpe_tgt)
    pb.rpe_lm = lm.astype(np.int32)
    pb.rpe_z_tgt = np.concatenate([z[k] for k in range(1, n_kf)])
    pb.rpe_z_ref = np.tile(z[0], (n_kf - 1, 1))
    pb.rpe_prior = np.zeros(len(tg), bool)

    # preintegration edges between consecutive keyframes, integrated at the initial biases of frame i
    pb.pre = []
    for k in range(n_kf - 1):
        pre = PreIntegrator(cfg.cov_g, cfg.cov_a, cfg.cov_bg, cfg.cov_ba)
        sel = (imu[:, 0] >= t_kf[k] - 1e-9) & (imu[:, 0] < t_kf[k + 1] - 1e-9)
        pre.data = [(float(r[0]), r[1:4].copy(), r[4:7].copy()) for r in imu[sel]]
        pre.integrate(float(t_kf[k + 1]), states[k, 10:13], states[k, 13:16], True, True)
        pb.pre.append(PreFactor.from_preintegrator(k, k + 1, pre))

    # dense marginalisation prior over the first N - 1 keyframes
    K = n_kf - 1
    n = 15 * K
    sd = np.tile(np.concatenate([np.full(3, 1 / 0.02), np.full(3, 1 / 0.05), np.full(3, 1 / 0.1), np.full(3, 1 / 1e-3), np.full(3, 1 / 2e-2)]), K)
    S = np.diag(sd)
    if dense_prior:
        S = S + np.triu(rng.normal(0, 0.02, (n, n)) * np.sqrt(sd[:, None] * sd[None, :]), 1)
    lin = states[:K].copy()
    for k in range(K):                          # linearisation point a little off the current iterate: dx != 0
        lin[k, 0:4] = qmul(lin[k, 0:4], expmap(rng.normal(0, 1e-3, 3)))
        lin[k, 0:4] /= np.linalg.norm(lin[k, 0:4])
        lin[k, 4:10] += rng.normal(0, 2e-3, 6)
    pb.marg = MargPrior(frames=np.arange(K), lin=lin, sqrt_inv_cov=S, infovec=rng.normal(0, 0.1, n))
    return pb
