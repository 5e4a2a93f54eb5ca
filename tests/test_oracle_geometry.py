"""Oracle two-view geometry and initializer (rows N1 / N3 of SURVEY.md 8f) pinned against OpenCV and the
synthetic generator's ground truth.  CPU only."""
import numpy as np
import pytest
import cv2

from oracle import geometry as g
from oracle.lie import expmap, qmat, hat, qmul, qconj, qrot, logmap
from oracle.initializer import mat_to_quat, from_two_vectors


def _two_view(rng, n, planar=False, noise=0.0):
    R = qmat(expmap(rng.normal(size=3) * 0.1))
    t = rng.normal(size=3) * 0.3
    if planar:
        nrm = np.array([0.1, -0.2, 1.0]); nrm /= np.linalg.norm(nrm)
        xy = rng.uniform(-1, 1, size=(n, 2))
        rays = np.concatenate([xy, np.ones((n, 1))], 1)
        X = rays * (4.0 / (rays @ nrm))[:, None]
    else:
        X = np.concatenate([rng.uniform(-2, 2, size=(n, 2)), rng.uniform(3, 8, size=(n, 1))], 1)
    X2 = X @ R.T + t
    p1 = X[:, :2] / X[:, 2:3] + rng.normal(size=(n, 2)) * noise
    p2 = X2[:, :2] / X2[:, 2:3] + rng.normal(size=(n, 2)) * noise
    return R, t, p1, p2


def test_five_point_contains_true_essential():
    rng = np.random.default_rng(1)
    for _ in range(20):
        R, t, p1, p2 = _two_view(rng, 5)
        sols = g.solve_essential_5pt(p1, p2)
        assert 1 <= len(sols) <= 10
        Et = hat(t) @ R
        Et /= np.linalg.norm(Et)
        err = min(min(np.linalg.norm(E / np.linalg.norm(E) - s * Et) for s in (1, -1)) for E in sols)
        assert err < 1e-8
        for E in sols:                                # every solution satisfies the five epipolar constraints
            assert g.essential_geometric_error(E, p1, p2).max() < 1e-20
            # and the essential-matrix constraints  2 E E^T E - tr(E E^T) E = 0,  det E = 0
            En = E / np.linalg.norm(E)
            assert np.abs(2 * En @ En.T @ En - np.trace(En @ En.T) * En).max() < 1e-8
            assert abs(np.linalg.det(En)) < 1e-9


def test_essential_ransac_against_opencv():
    rng = np.random.default_rng(2)
    R, t, p1, p2 = _two_view(rng, 150, noise=0.3 / 458)
    p2[:15] += rng.normal(size=(15, 2)) * 0.1           # gross outliers
    thr = 0.7 / 458
    E, mask = g.find_essential_matrix(p1, p2, thr, 0.999, 1000, 648)
    assert mask[15:].mean() > 0.9 and mask[:15].sum() <= 2
    R1, R2, T = g.decompose_essential(E)
    assert min(np.linalg.norm(R1 - R), np.linalg.norm(R2 - R)) < 0.08      # minimal-sample model, not refined
    assert abs(abs(T @ t) / np.linalg.norm(t) - 1) < 0.05
    Ecv, mcv = cv2.findEssentialMat(p1, p2, np.eye(3), cv2.RANSAC, 0.999, thr * 3)
    Ecv = Ecv[:3] / np.linalg.norm(Ecv[:3])
    En = E / np.linalg.norm(E)
    assert min(np.linalg.norm(En - Ecv), np.linalg.norm(En + Ecv)) < 0.25


def test_homography_against_opencv():
    rng = np.random.default_rng(3)
    R, t, p1, p2 = _two_view(rng, 4, planar=True)
    H = g.solve_homography_4pt(p1, p2)
    Hcv = cv2.getPerspectiveTransform(p1.astype(np.float32), p2.astype(np.float32))
    assert np.abs(H / H[2, 2] - Hcv).max() < 1e-4
    R, t, p1, p2 = _two_view(rng, 100, planar=True, noise=0.2 / 458)
    p2[:10] += 0.2
    H, mask = g.find_homography_matrix(p1, p2, 0.7 / 458, 0.999, 1000, 648)
    assert mask[:10].sum() == 0 and mask[10:].mean() > 0.85
    ok, R1, R2, T1, T2, n1, n2 = g.decompose_homography(H)
    assert ok
    k = int(np.argmin([np.linalg.norm(R1 - R), np.linalg.norm(R2 - R)]))
    Rk, Tk = (R1, T1) if k == 0 else (R2, T2)
    assert np.linalg.norm(Rk - R) < 0.02 and abs(np.linalg.det(Rk) - 1) < 1e-9
    assert abs(abs(Tk @ t) / (np.linalg.norm(Tk) * np.linalg.norm(t)) - 1) < 0.01
    # OpenCV's decomposition holds the same rotation among its four solutions
    _, Rs, Ts, Ns = cv2.decomposeHomographyMat(H, np.eye(3))
    assert min(np.linalg.norm(Rc - Rk) for Rc in Rs) < 1e-6
    # pure rotation is reported as such (homography.cpp:12-32)
    Rr = qmat(expmap(np.array([0.02, -0.05, 0.01])))
    ok, R1, *_ = g.decompose_homography(Rr * 3.0)
    assert not ok and np.linalg.norm(R1 - Rr) < 1e-9


def test_ransac_small_inputs():
    p = np.zeros((3, 2))
    E, m = g.find_essential_matrix(p, p)
    assert E is None and m.shape == (3,) and m.sum() == 0          # ransac.h:46-50
    H, m = g.find_homography_matrix(p, p)
    assert H is None and m.sum() == 0


def test_eigen_helpers():
    rng = np.random.default_rng(4)
    for _ in range(50):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        q2 = mat_to_quat(qmat(q))
        assert min(np.linalg.norm(q - q2), np.linalg.norm(q + q2)) < 1e-12
        a, b = rng.normal(size=3), rng.normal(size=3)
        r = from_two_vectors(a, b)
        assert np.linalg.norm(qrot(r, a / np.linalg.norm(a)) - b / np.linalg.norm(b)) < 1e-12
    r = from_two_vectors(np.array([0, 0, 1.0]), np.array([0, 0, -1.0]))
    assert np.linalg.norm(qrot(r, np.array([0, 0, 1.0])) - np.array([0, 0, -1.0])) < 1e-9


@pytest.mark.timeout(300)
def test_initializer_recovers_metric_state():
    """init_sfm + init_imu on a synthetic stream: metric baseline, relative rotation, gravity and speed
    against the generator's ground truth; then the session tracks."""
    from oracle.synth import Stream
    from oracle.pipeline import Session, Config
    st = Stream(7)
    cfg = Config(st.w, st.h, st.K)
    sess = Session(cfg, gt_fn=None)
    n = 40
    imu = st.imu(n * st.imu_per_cam + 2)
    k_imu = 0
    for k in range(n):
        t_img = st.frame_time(k)
        while k_imu < len(imu) and imu[k_imu, 0] <= t_img + 1e-9:
            sess.track_imu(imu[k_imu, 0], imu[k_imu, 1:4], imu[k_imu, 4:7])
            k_imu += 1
        sess.track_camera(t_img, st.frame(k))
        if sess.swt is not None:
            break
    assert sess.swt is not None and k <= 37
    tr = sess.init_trace
    assert tr['matches'] >= cfg.init_min_matches and 0.001 < tr['scale'] < 1.0 and tr['landmarks'] >= 30
    m = sess.swt.map
    f0, f1 = m.frames[0], m.frames[-1]
    q0, p0, v0 = st.gt_state(f0.t)
    q1, p1, v1 = st.gt_state(f1.t)
    assert abs(np.linalg.norm(f1.p - f0.p) / np.linalg.norm(p1 - p0) - 1) < 0.02
    assert np.linalg.norm(logmap(qmul(qconj(qmul(qconj(f0.q), f1.q)), qmul(qconj(q0), q1)))) < 0.01
    z = np.array([0, 0, 1.0])
    assert np.linalg.norm(qrot(qconj(f1.q), z) - qrot(qconj(q1), z)) < 0.01
    assert abs(np.linalg.norm(f1.v) / np.linalg.norm(v1) - 1) < 0.03


@pytest.mark.timeout(300)
def test_per_frame_essential_mask_is_all_ones():
    """frame.cpp:108-114 runs the 5-point RANSAC with threshold 1.0 (2 * 3.84 in squared normalised units) on every
    frame pair; oracle/ransac.py and the CUDA session replace it by its effect.  Run the real thing on a stream."""
    from oracle.synth import Stream
    from oracle.pipeline import Session, Config
    st = Stream(3)
    cfg = Config(st.w, st.h, st.K)
    cfg.essential_ransac = True
    cfg.essential_masks = []
    sess = Session(cfg, gt_fn=st.gt_state)
    n = 16
    imu = st.imu(n * st.imu_per_cam + 2)
    k_imu = 0
    for k in range(n):
        t_img = st.frame_time(k)
        while k_imu < len(imu) and imu[k_imu, 0] <= t_img + 1e-9:
            sess.track_imu(imu[k_imu, 0], imu[k_imu, 1:4], imu[k_imu, 4:7])
            k_imu += 1
        sess.track_camera(t_img, st.frame(k))
    assert len(cfg.essential_masks) >= n - 3
    assert all(m.all() and len(m) >= 50 for m in cfg.essential_masks)
