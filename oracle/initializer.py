"""Restatement of the reference initializer:

  xrslam/src/xrslam/core/initializer.cpp:78-155   initialize(): init_sfm, init_imu, final visual-inertial BA
  xrslam/src/xrslam/core/initializer.cpp:157-383  init_sfm(): H/E RANSAC, 8 (R,T) candidates, two-view
                                                  triangulation, PnP of the middle keyframes, visual BA
  xrslam/src/xrslam/core/initializer.cpp:385-571  init_imu(): gyro bias, gravity/scale/velocity linear solves,
                                                  gravity refinement on the tangent plane, apply_init

It works on the Map produced by Session.mirror_keyframe_map (pipeline.py) and leaves the frames/tracks in the
state the reference hands to its final BA.  Eigen's JacobiSVD::solve / fullPivHouseholderQr().solve are
numpy least-squares solves (same minimiser for the full-rank systems that occur).
Oracle code: test infrastructure only.
"""
import math
import numpy as np

from .lie import qmul, qconj, qrot, qmat, qnormalized, logmap, s2_tangential_basis
from . import geometry as geo

GRAVITY_NOMINAL = 9.80665


def mat_to_quat(R):
    """Eigen Quaterniond(Matrix3d) (Shepperd's branches as in Eigen/src/Geometry/Quaternion.h)."""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)
    if t > 0:
        t = math.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (R[2, 1] - R[1, 2]) * t
        q[1] = (R[0, 2] - R[2, 0]) * t
        q[2] = (R[1, 0] - R[0, 1]) * t
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (R[k, j] - R[j, k]) * t
        q[j] = (R[j, i] + R[i, j]) * t
        q[k] = (R[k, i] + R[i, k]) * t
    return q


def from_two_vectors(a, b):
    """Eigen Quaterniond::FromTwoVectors."""
    v0 = a / np.linalg.norm(a)
    v1 = b / np.linalg.norm(b)
    c = float(v1 @ v0)
    if c < -1.0 + 1e-12:
        c = max(c, -1.0)
        m = np.stack([v0, v1])
        _, _, Vt = np.linalg.svd(m, full_matrices=True)
        axis = Vt[2]
        w2 = (1.0 + c) * 0.5
        return np.concatenate([axis * math.sqrt(1.0 - w2), [math.sqrt(w2)]])
    axis = np.cross(v0, v1)
    s = math.sqrt((1.0 + c) * 2.0)
    return np.concatenate([axis / s, [s * 0.5]])


def set_camera_pose(frame, q_c, p_c):
    # frame.cpp:183-186 with sensor = camera
    cfg = frame.cfg
    frame.q = qmul(q_c, qconj(cfg.q_bc))
    frame.p = p_c - qrot(frame.q, cfg.p_bc)


def imu_pose(frame):
    cfg = frame.cfg
    return qmul(frame.q, cfg.q_bi), frame.p + qrot(frame.q, cfg.p_bi)


def set_imu_pose(frame, q_i, p_i):
    cfg = frame.cfg
    frame.q = qmul(q_i, qconj(cfg.q_bi))
    frame.p = p_i - qrot(frame.q, cfg.p_bi)


def two_view_seed(pts_i, pts_j, fx, min_triangulation, seed):
    """initializer.cpp:197-285: homography and essential RANSAC, the eight (R, T) candidates, two-view
    triangulation of each and the reference's selection rule.  Pure function of the matches."""
    n = len(pts_i)
    out = dict(fail=None, candidate=-1, counts=[0] * 8)
    Rs, Ts = [], []
    H, _ = geo.find_homography_matrix(pts_i, pts_j, 0.7 / fx, 0.999, 1000, seed)
    ok, RH1, RH2, TH1, TH2, _, _ = geo.decompose_homography(H)
    if not ok:
        out['fail'] = 'pure rotation'
        return out
    TH1 = TH1 / np.linalg.norm(TH1)
    TH2 = TH2 / np.linalg.norm(TH2)
    Rs += [RH1, RH1, RH2, RH2]
    Ts += [TH1, -TH1, TH2, -TH2]
    E, _ = geo.find_essential_matrix(pts_i, pts_j, 0.7 / fx, 0.999, 1000, seed)
    RE1, RE2, TE = geo.decompose_essential(E)
    TE = TE / np.linalg.norm(TE)
    Rs += [RE1, RE1, RE2, RE2]
    Ts += [TE, -TE, TE, -TE]
    P1 = np.concatenate([np.eye(3), np.zeros((3, 1))], axis=1)
    cand_pts, cand_status, counts, scores = [], [], [], []
    best = 0
    for c in range(len(Rs)):
        P2 = np.concatenate([Rs[c], Ts[c][:, None]], axis=1)
        pts = np.zeros((n, 3))
        status = np.zeros(n, np.uint8)
        count, score = 0, 0.0
        for j in range(n):
            x1 = np.append(pts_i[j], 1.0)
            x2 = np.append(pts_j[j], 1.0)
            q = geo.triangulate_point_2v(P1, P2, x1, x2)
            q1, q2 = P1 @ q, P2 @ q
            if q1[2] * q[3] > 0 and q2[2] * q[3] > 0:
                if q1[2] / q[3] < 100 and q2[2] / q[3] < 100:
                    pts[j] = q[:3] / q[3]
                    status[j] = 1
                    count += 1
                    score += 0.5 * (np.sum((q1[:2] / q1[2] - pts_i[j]) ** 2) + np.sum((q2[:2] / q2[2] - pts_j[j]) ** 2))
        cand_pts.append(pts)
        cand_status.append(status)
        counts.append(count)
        scores.append(score)
        if counts[c] > min_triangulation and scores[c] < scores[best]:      # the selection rule of :262-268
            best = c
        elif counts[c] > counts[best]:
            best = c
    out.update(candidate=best, counts=counts, scores=scores, R=Rs[best], T=Ts[best], points=cand_pts[best],
               status=cand_status[best], Rs=Rs, Ts=Ts)
    if counts[best] < min_triangulation:
        out['fail'] = 'triangulation'
    return out


class Initializer:
    def __init__(self, cfg, map_, solve, builder_cls, recorder=None):
        """map_: the mirrored keyframe map; solve(builder, kind) -> summary (writes the states back)."""
        self.cfg, self.map = cfg, map_
        self.solve, self.builder_cls = solve, builder_cls
        self.trace = {}

    # ------------------------------------------------------------------ init_sfm (initializer.cpp:157-383)
    def init_sfm(self):
        cfg, m = self.cfg, self.map
        fi, fj = m.frames[0], m.frames[-1]
        K = cfg.K
        pts_i, pts_j, matches, par_list = [], [], [], []
        total_parallax = 0.0
        for ki in range(fi.keypoint_num()):
            tr = fi.tracks[ki]
            if tr is None:
                continue
            kj = tr.get_keypoint_index(fj)
            if kj is None:
                continue
            bi, bj = fi.bearings[ki], fj.bearings[kj]
            pts_i.append(bi[:2] / bi[2])
            pts_j.append(bj[:2] / bj[2])
            matches.append((ki, kj))
            ui = np.array([bi[0] / bi[2] * K[0] + K[2], bi[1] / bi[2] * K[1] + K[3]])
            uj = np.array([bj[0] / bj[2] * K[0] + K[2], bj[1] / bj[2] * K[1] + K[3]])
            par_list.append(np.linalg.norm(ui - uj))
            total_parallax += par_list[-1]
        n = len(matches)
        self.trace['matches'] = n
        if n < cfg.init_min_matches:
            return False
        total_parallax /= max(n, 1)
        self.trace['parallax'] = total_parallax
        if total_parallax < cfg.init_min_parallax:
            return False
        pts_i = np.array(pts_i)
        pts_j = np.array(pts_j)
        seed = two_view_seed(pts_i, pts_j, K[0], cfg.init_min_triangulation, cfg.random)
        self.trace['sfm_in'] = (pts_i, pts_j, np.array(par_list))
        self.trace['sfm_out'] = seed
        if seed['fail']:
            self.trace['fail'] = seed['fail']
            return False
        best = seed['candidate']
        self.trace['candidate'] = best
        self.trace['counts'] = seed['counts']
        init_R, init_T = seed['R'], seed['T']
        cand_pts = {best: seed['points']}
        cand_status = {best: seed['status']}

        # [2.1] first / last keyframe poses (camera frame of keyframe 0 is the origin)
        set_camera_pose(fi, np.array([0.0, 0, 0, 1]), np.zeros(3))
        set_camera_pose(fj, mat_to_quat(init_R.T), -(init_R.T @ init_T))
        for k in range(n):
            if cand_status[best][k] == 0:
                continue
            tr = fi.tracks[matches[k][0]]
            tr.set_landmark_point(cand_pts[best][k])
            tr.valid = True
            tr.triangulated = True

        # [2.2] the keyframes in between: PnP on the landmarks anchored in keyframe 0, chained
        for j in range(1, m.frame_num() - 1):
            fa, fb = m.frames[j - 1], m.frames[j]
            qa, pa = fa.camera_pose()
            set_camera_pose(fb, qa, pa)
            b = self.builder_cls(cfg)
            b.frame(fb)
            for k in range(fb.keypoint_num()):
                tr = fb.tracks[k]
                if tr is None or not tr.has_keypoint(m.frames[0]):
                    continue
                if tr.valid and tr.triangulated:
                    b.add_rpe(fb, tr, True)
            self.solve(b, 'init_pnp')

        # [2.3] triangulate what is left
        for tr in list(m.tracks):
            if tr.valid:
                continue
            x = tr.triangulate()
            if x is not None:
                tr.set_landmark_point(x)
                tr.valid = True
                tr.triangulated = True

        # [3.1] visual bundle adjustment, poses only (add_frame_states(frame, false)), keyframe 0 fixed
        m.frames[0].fix_pose = True
        b = self.builder_cls(cfg)
        for f in m.frames:
            b.frame(f)
        for f in m.frames:
            for j in range(f.keypoint_num()):
                tr = f.tracks[j]
                if tr is None or not (tr.valid and tr.triangulated):
                    continue
                if f is tr.first_frame():
                    continue
                b.add_rpe(f, tr, False)
        summ = self.solve(b, 'init_sfm_ba')
        if summ.termination == 'too many invalid steps':
            return False
        # [3.2] landmark.reprojection_error is never written by the reference (state.h:41-47), so only
        # tracks that are not VALID are pruned
        for tr in [t for t in m.tracks if not t.valid]:
            m.erase_track(tr)
        return True

    # ------------------------------------------------------------------ init_imu (initializer.cpp:385-571)
    def preintegrate(self):
        m = self.map
        for j in range(1, m.frame_num()):
            fj = m.frames[j]
            fj.preintegration.integrate(fj.t, self.bg, self.ba, True, False)
        # (t, q, p, v, dq_dbg) of every edge, kept for the host-logic tests
        self.trace.setdefault('deltas', []).append([(f.preintegration.t, f.preintegration.q.copy(), f.preintegration.p.copy(),
                                                      f.preintegration.v.copy(), f.preintegration.dq_dbg.copy()) for f in m.frames[1:]])

    def solve_gyro_bias(self):
        self.preintegrate()
        m = self.map
        A = np.zeros((3, 3))
        b = np.zeros(3)
        for j in range(1, m.frame_num()):
            fi, fj = m.frames[j - 1], m.frames[j]
            qi, _ = imu_pose(fi)
            qj, _ = imu_pose(fj)
            pre = fj.preintegration
            A += pre.dq_dbg.T @ pre.dq_dbg
            b += pre.dq_dbg.T @ logmap(qmul(qconj(qmul(qi, pre.q)), qj))
        self.bg = np.linalg.lstsq(A, b, rcond=None)[0]
        self.trace['bg'] = self.bg.copy()

    def solve_gravity_scale_velocity(self):
        self.preintegrate()
        m, cfg = self.map, self.cfg
        N = m.frame_num()
        A = np.zeros(((N - 1) * 6, 3 + 1 + 3 * N))
        b = np.zeros((N - 1) * 6)
        I3 = np.eye(3)
        for j in range(1, N):
            i = j - 1
            fi, fj = m.frames[i], m.frames[j]
            d = fj.preintegration
            _, pci = fi.camera_pose()
            _, pcj = fj.camera_pose()
            A[i * 6:i * 6 + 3, 0:3] = -0.5 * d.t * d.t * I3
            A[i * 6:i * 6 + 3, 3] = pcj - pci
            A[i * 6:i * 6 + 3, 4 + i * 3:7 + i * 3] = -d.t * I3
            b[i * 6:i * 6 + 3] = qrot(fi.q, d.p) + (qrot(fj.q, cfg.p_bc) - qrot(fi.q, cfg.p_bc))
            A[i * 6 + 3:i * 6 + 6, 0:3] = -d.t * I3
            A[i * 6 + 3:i * 6 + 6, 4 + i * 3:7 + i * 3] = -I3
            A[i * 6 + 3:i * 6 + 6, 4 + j * 3:7 + j * 3] = I3
            b[i * 6 + 3:i * 6 + 6] = qrot(fi.q, d.v)
        x = np.linalg.lstsq(A, b, rcond=None)[0]
        g = x[0:3]
        self.gravity = g / np.linalg.norm(g) * GRAVITY_NOMINAL
        self.scale = float(x[3])
        self.velocities = [x[4 + 3 * i:7 + 3 * i].copy() for i in range(N)]
        self.trace['gsv'] = (self.gravity.copy(), self.scale, np.array(self.velocities))

    def refine_scale_velocity_via_gravity(self):
        damp = 0.1
        self.preintegrate()
        m, cfg = self.map, self.cfg
        N = m.frame_num()
        A = np.zeros(((N - 1) * 6, 2 + 1 + 3 * N))
        b = np.zeros((N - 1) * 6)
        I3 = np.eye(3)
        Tg = s2_tangential_basis(self.gravity)
        for j in range(1, N):
            i = j - 1
            fi, fj = m.frames[i], m.frames[j]
            d = fj.preintegration
            _, pci = fi.camera_pose()
            _, pcj = fj.camera_pose()
            A[i * 6:i * 6 + 3, 0:2] = -0.5 * d.t * d.t * Tg
            A[i * 6:i * 6 + 3, 2] = pcj - pci
            A[i * 6:i * 6 + 3, 3 + i * 3:6 + i * 3] = -d.t * I3
            b[i * 6:i * 6 + 3] = 0.5 * d.t * d.t * self.gravity + qrot(fi.q, d.p) + \
                (qrot(fj.q, cfg.p_bc) - qrot(fi.q, cfg.p_bc))
            A[i * 6 + 3:i * 6 + 6, 0:2] = -d.t * Tg
            A[i * 6 + 3:i * 6 + 6, 3 + i * 3:6 + i * 3] = -I3
            A[i * 6 + 3:i * 6 + 6, 3 + j * 3:6 + j * 3] = I3
            b[i * 6 + 3:i * 6 + 6] = d.t * self.gravity + qrot(fi.q, d.v)
        x = np.linalg.lstsq(A, b, rcond=None)[0]
        g = self.gravity + damp * (Tg @ x[0:2])
        self.gravity = g / np.linalg.norm(g) * GRAVITY_NOMINAL
        self.scale = float(x[2])
        self.velocities = [x[3 + 3 * i:6 + 3 * i].copy() for i in range(N)]
        self.trace['refined'] = (self.gravity.copy(), self.scale, np.array(self.velocities))

    def apply_init(self, apply_ba=False, apply_velocity=True):
        m, cfg = self.map, self.cfg
        q = from_two_vectors(self.gravity, np.array([0.0, 0.0, -GRAVITY_NOMINAL]))
        for i, f in enumerate(m.frames):
            qi, pi = imu_pose(f)
            set_imu_pose(f, qmul(q, qi), self.scale * qrot(q, pi))
            f.v = qrot(q, self.velocities[i]) if apply_velocity else np.zeros(3)
            f.bg = self.bg.copy()
            f.ba = self.ba.copy() if apply_ba else np.zeros(3)
        final = 0
        for tr in list(m.tracks):
            x = tr.triangulate()
            if x is not None:
                tr.set_landmark_point(x)
                tr.valid = True
                tr.triangulated = True
                final += 1
            else:
                tr.valid = False
        self.trace['landmarks'] = final
        return final >= cfg.init_min_landmarks

    def init_imu(self):
        N = self.map.frame_num()
        self.bg, self.ba = np.zeros(3), np.zeros(3)
        self.gravity = np.zeros(3)
        self.scale = 1.0
        self.velocities = [np.zeros(3) for _ in range(N)]
        self.trace['align_frames'] = [(f.q.copy(), f.p.copy()) for f in self.map.frames]
        self.solve_gyro_bias()
        self.solve_gravity_scale_velocity()
        self.trace['scale0'] = self.scale
        if self.scale < 0.001 or self.scale > 1.0:
            return False
        if not self.cfg.init_refine_imu:
            return self.apply_init()
        self.refine_scale_velocity_via_gravity()
        self.trace['scale'] = self.scale
        if self.scale < 0.001 or self.scale > 1.0:
            return False
        return self.apply_init()

    def run(self):
        """initialize() up to (not including) the final visual-inertial BA (initializer.cpp:78-84)."""
        if not self.init_sfm():
            return False
        if not self.init_imu():
            return False
        return True
