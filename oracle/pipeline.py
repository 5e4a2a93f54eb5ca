"""CPU restatement of the reference's per-session orchestration around the hot path (row G of
SURVEY.md 8(a)): sensor sync, per-frame front-end driver, keyframe/subframe state machine,
sliding-window solves and marginalisation.

Follows (non-iOS, threading OFF, parsac_flag false):
  xrslam/src/xrslam/core/detail.cpp:15-28,103-177          (track_imu / track_camera / predict_pose)
  xrslam/src/xrslam/core/feature_tracker.cpp:24-153        (FeatureTracker::work)
  xrslam/src/xrslam/map/frame.cpp:55-174                   (detect_keypoints / track_keypoints)
  xrslam/src/xrslam/map/track.cpp:14-101, map/map.cpp:20-63
  xrslam/src/xrslam/core/frontend_worker.cpp:28-86
  xrslam/src/xrslam/core/sliding_window_tracker.cpp:18-474
  xrslam/src/xrslam/core/initializer.cpp:22-76,88-155      (mirror_keyframe_map + final VI-BA)

The SfM/IMU-alignment part of the Initializer (initializer.cpp:157-571, SURVEY.md row N3,
out of the hot path) is replaced by a ground-truth bootstrap: the 8 initial keyframes get the
generator's poses/velocities, tracks are triangulated with Track::triangulate, then the
reference's own final visual-inertial BA runs.  DESIGN.md states this.
Oracle code: test infrastructure only.
"""
import math
import numpy as np
from .lie import qmul, qconj, qrot, qmat, expmap, qnormalized
from .preintegrator import PreIntegrator, GRAVITY
from .frontend import OracleImage
from .poisson import PoissonDiskFilter2
from .ransac import find_rotation_matrix
from .problem import Problem, PreFactor, MargPrior
from . import ceres_tr
from . import marginalization as marg_mod


class Config:
    """configs/euroc_slam.yaml + configs/euroc_sensor.yaml values used on the path."""

    def __init__(self, width=640, height=480, K=(458.654, 457.296, 320.0, 240.0)):
        self.width, self.height = width, height
        self.K = np.array(K, dtype=np.float64)          # fx fy cx cy
        self.keypoint_noise_cov = (0.5, 0.5)
        self.q_bc = np.array([-7.7071797555374275e-03, 1.0499323370587278e-02, 7.0175280029197162e-01, 7.1230146066895372e-01])
        self.p_bc = np.array([-0.0216401454975, -0.064676986768, 0.00981073058949])
        self.q_bi = np.array([0.0, 0.0, 0.0, 1.0])
        self.p_bi = np.zeros(3)
        self.q_bo = np.array([0.0, 0.0, 0.0, 1.0])
        self.p_bo = np.zeros(3)
        self.cov_g = np.eye(3) * 2.8791302399999997e-08
        self.cov_a = np.eye(3) * 4.0e-6
        self.cov_bg = np.eye(3) * 3.7608844899999997e-10
        self.cov_ba = np.eye(3) * 9.0e-6
        self.sliding_window_size = 10
        self.subframe_size = 3
        self.force_keyframe_landmarks = 35
        self.min_keypoint_distance = 20.0
        self.max_keypoint_detection = 200
        self.max_init_frames = 60
        self.max_frames = 20
        self.predict_keypoints = True
        self.clahe_clip_limit, self.clahe_width, self.clahe_height = 6.0, 8, 8
        self.init_keyframe_num, self.init_keyframe_gap = 8, 5
        # configs/euroc_slam.yaml:21-28, config.cpp:66 (random)
        self.init_min_matches, self.init_min_parallax = 50, 10.0
        self.init_min_triangulation, self.init_min_landmarks = 20, 30
        self.init_refine_imu = True
        self.random = 648
        self.solver_iteration_limit = 30
        self.rotation_misalignment_threshold = 0.02
        self.rotation_ransac_threshold = 10.0

    def sqrt_inv_cov(self):
        # core/detail.cpp:107-109
        return np.array([self.K[0] / math.sqrt(self.keypoint_noise_cov[0]), self.K[1] / math.sqrt(self.keypoint_noise_cov[1])])


def apply_k(b, K):
    return np.array([b[0] / b[2] * K[0] + K[2], b[1] / b[2] * K[1] + K[3]])


def remove_k(p, K):
    v = np.array([(p[0] - K[2]) / K[0], (p[1] - K[3]) / K[1], 1.0])
    return v / np.linalg.norm(v)


class Track:
    _next_id = 0

    def __init__(self, map_):
        Track._next_id += 1
        self.id = Track._next_id
        self.map = map_
        self.refs = {}          # Frame -> keypoint index
        self.valid = False
        self.triangulated = False
        self.trash = False
        self.static = True      # track.cpp:7
        self.inv_depth = 0.0
        self.life = 0

    def keypoint_num(self):
        return len(self.refs)

    def items(self):
        return sorted(self.refs.items(), key=lambda kv: kv[0].id)

    def first_frame(self):
        return min(self.refs.keys(), key=lambda f: f.id)

    def first_keypoint(self):
        f = self.first_frame()
        return f, self.refs[f]

    def has_keypoint(self, frame):
        return frame in self.refs

    def get_keypoint_index(self, frame):
        return self.refs.get(frame, None)

    def add_keypoint(self, frame, idx):
        # track.cpp:14-23
        self.refs[frame] = idx
        frame.tracks[idx] = self
        if self.triangulated:
            self.life += 1
        else:
            self.life = 1

    def remove_keypoint(self, frame, suicide_if_empty=True):
        # track.cpp:25-44
        idx = self.refs[frame]
        landmark = None
        if frame is self.first_frame():
            landmark = self.get_landmark_point()
        frame.tracks[idx] = None
        del self.refs[frame]
        if len(self.refs) > 0:
            if landmark is not None:
                self.set_landmark_point(landmark)
        else:
            self.valid = False
            if suicide_if_empty:
                self.map.recycle_track(self)

    def triangulate(self):
        # track.cpp:46-76, geometry/stereo.h:84-94
        rows = []
        Ps = []
        for frame, idx in self.items():
            q, p = frame.camera_pose()
            R = qmat(qconj(q))
            P = np.concatenate([R, (-(R @ p))[:, None]], axis=1)
            z = frame.bearings[idx]
            rows.append(z[0] * P[2] - z[2] * P[0])
            rows.append(z[1] * P[2] - z[2] * P[1])
            Ps.append(P)
        A = np.array(rows)
        _, _, Vt = np.linalg.svd(A, full_matrices=True)
        h = Vt[3]
        for P in Ps:
            qi = P @ h
            if not (qi[2] * h[3] > 0):
                return None
        self.life = 1
        return h[:3] / h[3]

    def get_landmark_point(self):
        # track.cpp:91-96
        frame, idx = self.first_keypoint()
        q, p = frame.camera_pose()
        with np.errstate(divide='ignore', invalid='ignore'):
            return qrot(q, frame.bearings[idx]) / self.inv_depth + p

    def set_landmark_point(self, x):
        # track.cpp:98-102
        frame, _ = self.first_keypoint()
        q, p = frame.camera_pose()
        with np.errstate(divide='ignore', invalid='ignore'):
            self.inv_depth = 1.0 / np.linalg.norm(qrot(qconj(q), x - p))


class Frame:
    _next_id = 0

    def __init__(self, cfg, image=None, clone_of=None):
        if clone_of is None:
            Frame._next_id += 1
            self.id = Frame._next_id
            self.keyframe = False
            self.no_translation = False
        else:
            self.id = clone_of.id
            self.keyframe = clone_of.keyframe
            self.no_translation = clone_of.no_translation
        self.fix_pose = False if clone_of is None else clone_of.fix_pose
        self.fix_motion = False if clone_of is None else clone_of.fix_motion
        self.cfg = cfg
        self.map = None
        self.image = image
        self.K = cfg.K
        self.q = np.array([0.0, 0.0, 0.0, 1.0])
        self.p = np.zeros(3)
        self.v = np.zeros(3)
        self.bg = np.zeros(3)
        self.ba = np.zeros(3)
        self.preintegration = PreIntegrator(cfg.cov_g, cfg.cov_a, cfg.cov_bg, cfg.cov_ba)
        self.keyframe_preintegration = None
        self.subframes = []
        self.bearings = []
        self.tracks = []

    @property
    def t(self):
        return self.image.t

    def clone(self):
        # frame.cpp:20-37
        f = Frame(self.cfg, self.image, clone_of=self)
        f.q, f.p, f.v, f.bg, f.ba = self.q.copy(), self.p.copy(), self.v.copy(), self.bg.copy(), self.ba.copy()
        f.preintegration = self.preintegration.copy()
        f.bearings = [b.copy() for b in self.bearings]
        f.tracks = [None] * len(self.bearings)
        return f

    def keypoint_num(self):
        return len(self.bearings)

    def append_keypoint(self, b):
        self.bearings.append(b)
        self.tracks.append(None)

    def get_track(self, idx, allocation_map=None, create=False):
        if create:
            if allocation_map is None:
                allocation_map = self.map
            if self.tracks[idx] is None:
                tr = allocation_map.create_track()
                tr.add_keypoint(self, idx)
        return self.tracks[idx]

    def camera_pose(self):
        # frame.cpp:176-181 with sensor = camera
        return qmul(self.q, self.cfg.q_bc), self.p + qrot(self.q, self.cfg.p_bc)

    def state16(self):
        return np.concatenate([self.q, self.p, self.v, self.bg, self.ba])

    def set_state16(self, s):
        self.q, self.p, self.v, self.bg, self.ba = s[0:4].copy(), s[4:7].copy(), s[7:10].copy(), s[10:13].copy(), s[13:16].copy()

    # frame.cpp:55-72
    def detect_keypoints(self):
        pk = [tuple(apply_k(b, self.K)) for b in self.bearings]
        old = len(pk)
        pk = self.image.detect_keypoints(pk, self.cfg.max_keypoint_detection, self.cfg.min_keypoint_distance)
        for i in range(old, len(pk)):
            self.append_keypoint(remove_k(pk[i], self.K))

    # frame.cpp:74-174
    def track_keypoints(self, nxt):
        cfg = self.cfg
        n = len(self.bearings)
        curr = np.array([apply_k(b, self.K) for b in self.bearings]).reshape(n, 2)
        pred = None
        if cfg.predict_keypoints:
            dq = qconj(qmul(qmul(qmul(qmul(qconj(cfg.q_bc), cfg.q_bi), nxt.preintegration.q), qconj(cfg.q_bi)), cfg.q_bc))
            pred = np.array([apply_k(qrot(dq, b), nxt.K) for b in self.bearings]).reshape(n, 2)
        next_kps, status = self.image.track_keypoints(nxt.image, curr, pred)
        next_bearings = [remove_k(next_kps[i], nxt.K) for i in range(n)]
        # essential-matrix RANSAC mask (frame.cpp:108-114): see oracle/ransac.py header
        mask_e = np.ones(n, np.uint8) if n >= 5 else np.zeros(n, np.uint8)
        if getattr(cfg, 'essential_ransac', False) and n >= 5:
            # the real thing (stereo.cpp:40-69 with threshold 1.0): used by tests to show that the mask is all ones
            from .geometry import find_essential_matrix
            ch = np.array([b[:2] / b[2] for b in self.bearings]).reshape(n, 2)
            nh = np.array([b[:2] / b[2] for b in next_bearings]).reshape(n, 2)
            _, mask_e = find_essential_matrix(ch, nh, 1.0)
            cfg.essential_masks.append(mask_e.copy())
        status = status.copy()
        status[mask_e == 0] = 0
        b1 = np.array(self.bearings).reshape(n, 3)
        b2 = np.array(next_bearings).reshape(n, 3)
        R, mask = find_rotation_matrix(b1, b2, (math.pi / 180.0) * cfg.rotation_ransac_threshold)
        angles = []
        for i in range(n):
            if mask[i]:
                with np.errstate(invalid='ignore'):
                    angles.append(math.degrees(math.acos(min(max(float((R @ b1[i]) @ b2[i]), -1.0), 1.0))) if abs(float((R @ b1[i]) @ b2[i])) <= 1.0 else float('nan'))
        angles.sort()
        misalignment = angles[len(angles) * 7 // 10] if len(angles) > 0 else 0.0
        self.last_misalignment = misalignment
        if misalignment < cfg.rotation_misalignment_threshold:
            nxt.no_translation = True
        # filter keypoints by track length (frame.cpp:136-163)
        order = []
        for i in range(n):
            if status[i] == 0:
                continue
            tr = self.get_track(i)
            if tr is None:
                continue
            order.append((i, tr.keypoint_num()))
        order.sort(key=lambda kv: -kv[1])      # NOTE std::sort is not stable; Python's is (ties keep index order)
        flt = PoissonDiskFilter2(cfg.min_keypoint_distance)
        for i, _ in order:
            pt = next_kps[i]
            tr = self.get_track(i)
            if flt.permit_point(pt) and (tr is None or not tr.trash):
                flt.preset_point(pt)
            else:
                status[i] = 0
        for i in range(n):
            if status[i]:
                j = nxt.keypoint_num()
                nxt.append_keypoint(next_bearings[i])
                self.get_track(i, None, create=True).add_keypoint(nxt, j)
        return status


class Map:
    def __init__(self):
        self.frames = []
        self.tracks = []
        self.marg = None          # MargPrior-like: dict(frames=[Frame], lin=(K,16), S, infovec)

    def frame_num(self):
        return len(self.frames)

    def attach_frame(self, frame, position=None):
        frame.map = self
        if position is None:
            self.frames.append(frame)
        else:
            self.frames.insert(position, frame)

    def detach_frame(self, index):
        f = self.frames.pop(index)
        f.map = None
        return f

    def untrack_frame(self, frame):
        for i in range(frame.keypoint_num()):
            tr = frame.tracks[i]
            if tr is not None:
                tr.remove_keypoint(frame)

    def erase_frame(self, index):
        self.untrack_frame(self.frames[index])
        self.detach_frame(index)

    def frame_index_by_id(self, id_):
        for i, f in enumerate(self.frames):
            if f.id == id_:
                return i
        return None

    def create_track(self):
        tr = Track(self)
        self.tracks.append(tr)
        return tr

    def recycle_track(self, track):
        self.tracks.remove(track)

    def erase_track(self, track):
        while track.keypoint_num() > 0:
            track.remove_keypoint(track.first_frame(), False)
        self.recycle_track(track)


# ------------------------------------------------------------------------------------------------
class ProblemBuilder:
    """Collects Solver::add_frame_states / add_factor calls into an oracle Problem."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.frames = []
        self.fidx = {}
        self.tracks = []
        self.tidx = {}
        self.rpe = []     # (tgt, ref, lm, z_tgt, z_ref, prior)
        self.rot = []
        self.pre = []
        self.marg = None

    def frame(self, f):
        if f not in self.fidx:
            self.fidx[f] = len(self.frames)
            self.frames.append(f)
        return self.fidx[f]

    def track(self, t):
        if t not in self.tidx:
            self.tidx[t] = len(self.tracks)
            self.tracks.append(t)
        return self.tidx[t]

    def add_rpe(self, frame, track, prior):
        ref, kref = track.first_keypoint()
        self.rpe.append((self.frame(frame), self.frame(ref), self.track(track),
                         frame.bearings[track.refs[frame]], ref.bearings[kref], prior))

    def add_rot(self, frame, track):
        ref, kref = track.first_keypoint()
        self.rot.append((self.frame(frame), self.frame(ref), frame.bearings[track.refs[frame]], ref.bearings[kref]))

    def add_pre(self, fi, fj, pre, prior):
        self.pre.append(PreFactor.from_preintegrator(self.frame(fi), self.frame(fj), pre, prior))

    def add_marg(self, m):
        self.marg = MargPrior(frames=np.array([self.frame(f) for f in m['frames']]), lin=m['lin'].copy(),
                              sqrt_inv_cov=m['S'], infovec=m['infovec'])

    def build(self):
        cfg = self.cfg
        states = np.array([f.state16() for f in self.frames]).reshape(len(self.frames), 16)
        pb = Problem(states=states,
                     fix_pose=np.array([f.fix_pose for f in self.frames], bool),
                     fix_motion=np.array([f.fix_motion for f in self.frames], bool),
                     inv_depth=np.array([t.inv_depth for t in self.tracks], dtype=np.float64),
                     cam_q_cs=cfg.q_bc, cam_p_cs=cfg.p_bc, imu_q_cs=cfg.q_bi, imu_p_cs=cfg.p_bi,
                     sqrt_inv_cov_px=cfg.sqrt_inv_cov(), max_iterations=cfg.solver_iteration_limit)
        if self.rpe:
            pb.rpe_tgt = np.array([r[0] for r in self.rpe], np.int32)
            pb.rpe_ref = np.array([r[1] for r in self.rpe], np.int32)
            pb.rpe_lm = np.array([r[2] for r in self.rpe], np.int32)
            pb.rpe_z_tgt = np.array([r[3] for r in self.rpe])
            pb.rpe_z_ref = np.array([r[4] for r in self.rpe])
            pb.rpe_prior = np.array([r[5] for r in self.rpe], bool)
        if self.rot:
            pb.rot_tgt = np.array([r[0] for r in self.rot], np.int32)
            pb.rot_ref = np.array([r[1] for r in self.rot], np.int32)
            pb.rot_z_tgt = np.array([r[2] for r in self.rot])
            pb.rot_z_ref = np.array([r[3] for r in self.rot])
        pb.pre = self.pre
        pb.marg = self.marg
        return pb

    def writeback(self, pb):
        for i, f in enumerate(self.frames):
            f.set_state16(pb.states[i])
        for k, t in enumerate(self.tracks):
            t.inv_depth = float(pb.inv_depth[k])


class SlidingWindowTracker:
    def __init__(self, map_, cfg, solve_fn=None, marg_fn=None, recorder=None):
        # sliding_window_tracker.cpp:18-27
        self.map = map_
        self.cfg = cfg
        self.solve_fn = solve_fn or ceres_tr.solve
        self.marg_fn = marg_fn or marg_mod.marginalize
        self.recorder = recorder
        for j in range(1, map_.frame_num()):
            fi, fj = map_.frames[j - 1], map_.frames[j]
            fj.preintegration.integrate(fj.t, fi.bg, fi.ba, True, True)

    def _solve(self, builder, kind):
        pb = builder.build()
        if self.recorder is not None:
            self.recorder(kind, 'pre', pb)
        summ = self.solve_fn(pb)
        if self.recorder is not None:
            self.recorder(kind, 'post', pb, summ)
        builder.writeback(pb)
        return summ

    # sliding_window_tracker.cpp:31-80
    def mirror_frame(self, ft_map, frame_id):
        keyframe = self.map.frames[-1]
        new_i = keyframe.subframes[-1] if keyframe.subframes else keyframe
        ii = ft_map.frame_index_by_id(new_i.id)
        jj = ft_map.frame_index_by_id(frame_id)
        if ii is None or jj is None:
            return False
        old_i, old_j = ft_map.frames[ii], ft_map.frames[jj]
        curr = old_j.clone()
        new_data = curr.preintegration.data
        for index in range(jj - 1, ii, -1):
            new_data[0:0] = [(t, w.copy(), a.copy()) for (t, w, a) in ft_map.frames[index].preintegration.data]
        self.map.attach_frame(curr.clone())
        new_j = self.map.frames[-1]
        for ki in range(old_i.keypoint_num()):
            tr = old_i.tracks[ki]
            if tr is not None:
                kj = tr.get_keypoint_index(old_j)
                if kj is not None:
                    nt = new_i.get_track(ki, self.map, create=True)
                    nt.add_keypoint(new_j, kj)
                    tr.trash = nt.trash and not nt.static
        for tr in [t for t in self.map.tracks if t.trash and not t.static]:
            self.map.erase_track(tr)
        new_j.preintegration.integrate(new_j.t, new_i.bg, new_i.ba, True, True)
        new_j.q, new_j.p, new_j.v, new_j.bg, new_j.ba = new_j.preintegration.predict(new_i.q, new_i.p, new_i.v, new_i.bg, new_i.ba)
        return True

    # sliding_window_tracker.cpp:82-117
    def track(self):
        self.localize_newframe()
        if self.manage_keyframe():
            self.track_landmark()
            self.refine_window()
            self.slide_window()
            self.last_kind = 'keyframe'
        else:
            self.refine_subwindow()
            self.last_kind = 'subframe'
        return True

    # :119-143
    def localize_newframe(self):
        fi = self.map.frames[-2]
        if fi.subframes:
            fi = fi.subframes[-1]
        fj = self.map.frames[-1]
        b = ProblemBuilder(self.cfg)
        b.frame(fj)
        b.add_pre(fi, fj, fj.preintegration, True)
        for k in range(fj.keypoint_num()):
            tr = fj.tracks[k]
            if tr is not None and tr.valid and tr.triangulated and tr.static:
                b.add_rpe(fj, tr, True)
        self._solve(b, 'localize')

    # :145-223
    def manage_keyframe(self):
        cfg = self.cfg
        ki = self.map.frames[-2]
        nj = self.map.frames[-1]
        if ki.subframes:
            if ki.subframes[-1].no_translation:
                if nj.no_translation:
                    pass
                else:
                    ki.subframes[-1].keyframe = True
                    f = ki.subframes.pop()
                    self.map.attach_frame(f, self.map.frame_num() - 1)
                    nj.keyframe = True
                    return True
            else:
                if nj.no_translation:
                    lifted = ki.subframes.pop()
                    lifted.keyframe = True
                    lifted.subframes.append(self.map.detach_frame(self.map.frame_num() - 1))
                    self.map.attach_frame(lifted)
                    return True
                else:
                    if len(ki.subframes) >= cfg.subframe_size:
                        nj.keyframe = True
                        return True
        cnt = 0
        for k in range(nj.keypoint_num()):
            tr = nj.tracks[k]
            if tr is not None and tr.valid and tr.triangulated and tr.static:
                cnt += 1
        if cnt < cfg.force_keyframe_landmarks:
            nj.keyframe = True
            return True
        ki.subframes.append(self.map.detach_frame(self.map.frame_num() - 1))
        return False

    # :225-245
    def track_landmark(self):
        nj = self.map.frames[-1]
        for k in range(nj.keypoint_num()):
            tr = nj.tracks[k]
            if tr is not None and not tr.triangulated:
                p = tr.triangulate()
                if p is not None:
                    tr.set_landmark_point(p)
                    tr.triangulated = True
                    tr.valid = True
                    tr.static = True
                else:
                    tr.inv_depth = -1.0
                    tr.triangulated = False
                    tr.valid = False

    # :247-358
    def refine_window(self):
        m = self.map
        cfg = self.cfg
        b = ProblemBuilder(cfg)
        if m.marg is None:
            # estimation/marginalization_factor.h:18-33
            fr = m.frames[:-1]
            K = len(fr)
            S = np.zeros((15 * K, 15 * K))
            S[3:6, 3:6] = 1.0e15 * np.eye(3)
            S[0:3, 0:3] = 1.0e15 * np.eye(3)
            m.marg = dict(frames=list(fr), lin=np.array([f.state16() for f in fr]), S=S, infovec=np.zeros(15 * K))
        for f in m.frames:
            b.frame(f)
        b.add_marg(m.marg)
        for f in m.frames:
            for j in range(f.keypoint_num()):
                tr = f.tracks[j]
                if tr is None:
                    continue
                if not (tr.valid and tr.triangulated and tr.static):
                    continue
                ff = tr.first_frame()
                if not ff.keyframe:
                    continue
                if f is ff:
                    continue
                b.add_rpe(f, tr, False)
        for j in range(1, m.frame_num()):
            fi, fj = m.frames[j - 1], m.frames[j]
            fj.keyframe_preintegration = fj.preintegration.copy()
            if fi.subframes:
                data = []
                for sf in fi.subframes:
                    data.extend([(t, w.copy(), a.copy()) for (t, w, a) in sf.preintegration.data])
                fj.keyframe_preintegration.data[0:0] = data
            if fj.keyframe_preintegration.integrate(fj.t, fi.bg, fi.ba, True, True):
                b.add_pre(fi, fj, fj.keyframe_preintegration, False)
        self.last_ba_summary = self._solve(b, 'window')
        for tr in m.tracks:
            if tr.triangulated:
                is_valid = True
                x = tr.get_landmark_point()
                rpe, cnt = 0.0, 0.0
                for f, kidx in tr.items():
                    if not f.keyframe:
                        continue
                    q, p = f.camera_pose()
                    y = qrot(qconj(q), x - p)
                    if not (y[2] > 1.0e-3 and y[2] <= 50):
                        is_valid = False
                        break
                    rpe += float(np.linalg.norm(apply_k(y, f.K) - apply_k(f.bearings[kidx], f.K)))
                    cnt += 1.0
                is_valid = is_valid and (rpe / max(cnt, 1.0) < 3.0)
                tr.valid = bool(is_valid)
            else:
                tr.inv_depth = -1.0
        for tr in m.tracks:
            if not tr.valid:
                tr.trash = True

    # :360-368 + map.cpp:51-63
    def slide_window(self):
        m = self.map
        while m.frame_num() > self.cfg.sliding_window_size:
            f0 = m.frames[0]
            for sf in f0.subframes:
                m.untrack_frame(sf)
            self.marginalize_frame0()
            for i in range(f0.keypoint_num()):
                tr = f0.tracks[i]
                if tr is not None:
                    tr.remove_keypoint(f0)
            m.frames.pop(0)

    def marginalize_frame0(self):
        m = self.map
        cfg = self.cfg
        N = m.frame_num()
        fidx = {f: i for i, f in enumerate(m.frames)}
        states = np.array([f.state16() for f in m.frames])
        prior = MargPrior(frames=np.array([fidx[f] for f in m.marg['frames']]), lin=m.marg['lin'],
                          sqrt_inv_cov=m.marg['S'], infovec=m.marg['infovec'])
        pre = []
        if N > 1:
            pre.append(PreFactor.from_preintegrator(0, 1, m.frames[1].keyframe_preintegration, False))
        victim = m.frames[0]
        tracks, tix = [], {}
        obs = []
        for j in range(victim.keypoint_num()):
            tr = victim.tracks[j]
            if tr is None or not tr.valid:
                continue
            ref = tr.first_frame()
            if not ref.keyframe:
                continue
            o = dict(lm=len(tracks), ref=fidx[ref], z_ref=ref.bearings[tr.refs[ref]], tgt=[], z_tgt=[])
            tracks.append(tr)
            for f, kidx in tr.items():
                if f is ref or f not in fidx:
                    continue
                o['tgt'].append(fidx[f])
                o['z_tgt'].append(f.bearings[kidx])
            obs.append(o)
        invd = np.array([t.inv_depth for t in tracks], dtype=np.float64)
        if self.recorder is not None:
            self.recorder('marg', 'pre', dict(states=states, prior=prior, pre=pre, obs=obs, inv_depth=invd))
        newp, Hc, bc = self.marg_fn(states, prior, pre, obs, invd, cfg.q_bc, cfg.p_bc, cfg.q_bi, cfg.p_bi, cfg.sqrt_inv_cov(), 0)
        if self.recorder is not None:
            self.recorder('marg', 'post', newp)
        m.marg = dict(frames=list(m.frames[1:]), lin=newp.lin, S=newp.sqrt_inv_cov, infovec=newp.infovec)

    # :370-465
    def refine_subwindow(self):
        cfg = self.cfg
        frame = self.map.frames[-1]
        if not frame.subframes:
            return
        if frame.subframes[0].no_translation:
            if len(frame.subframes) >= 9:
                i = len(frame.subframes) // 3
                while i > 0:
                    tgt = frame.subframes[i * 3 - 1]
                    data = []
                    j = i * 3 - 1
                    while j > (i - 1) * 3:
                        src = frame.subframes[j - 1]
                        data[0:0] = src.preintegration.data
                        self.map.untrack_frame(src)
                        frame.subframes.pop(j - 1)
                        j -= 1
                    tgt.preintegration.data[0:0] = data
                    i -= 1
            b = ProblemBuilder(cfg)
            frame.fix_pose = frame.fix_motion = True
            b.frame(frame)
            for i, sf in enumerate(frame.subframes):
                b.frame(sf)
                prev = frame if i == 0 else frame.subframes[i - 1]
                sf.preintegration.integrate(sf.t, prev.bg, prev.ba, True, True)
                b.add_pre(prev, sf, sf.preintegration, False)
            last = frame.subframes[-1]
            for k in range(last.keypoint_num()):
                tr = last.tracks[k]
                if tr is not None and tr.valid:
                    if tr.triangulated:
                        if tr.static:
                            b.add_rpe(last, tr, True)
                    else:
                        b.add_rot(last, tr)
            self._solve(b, 'subwindow_r')
            frame.fix_pose = frame.fix_motion = False
        else:
            b = ProblemBuilder(cfg)
            frame.fix_pose = frame.fix_motion = True
            b.frame(frame)
            for i, sf in enumerate(frame.subframes):
                b.frame(sf)
                prev = frame if i == 0 else frame.subframes[i - 1]
                sf.preintegration.integrate(sf.t, prev.bg, prev.ba, True, True)
                b.add_pre(prev, sf, sf.preintegration, False)
                for k in range(sf.keypoint_num()):
                    tr = sf.tracks[k]
                    if tr is not None and tr.valid and tr.triangulated and tr.static:
                        if tr.first_frame().keyframe:
                            b.add_rpe(sf, tr, True)
                        elif tr.first_frame().id > frame.id:
                            # sliding_window_tracker.cpp:451-454: unreachable with the default
                            # configuration (a triangulated track never starts after the newest keyframe).
                            raise RuntimeError('refine_subwindow: unreachable reference branch reached')
            self._solve(b, 'subwindow_t')
            frame.fix_pose = frame.fix_motion = False

    def get_latest_state(self):
        f = self.map.frames[-1]
        if f.subframes:
            f = f.subframes[-1]
        return f.t, f.id, f.q.copy(), f.p.copy(), f.v.copy(), f.bg.copy(), f.ba.copy()


# ------------------------------------------------------------------------------------------------
class Session:
    """One VIO session: Detail + FeatureTracker + FrontendWorker in the synchronous PC build."""

    def __init__(self, cfg, gt_fn=None, solve_fn=None, marg_fn=None, recorder=None, image_cls=OracleImage):
        self.cfg = cfg
        self.gt_fn = gt_fn                    # t -> (q,p,v) ground truth for the bootstrap
        self.solve_fn, self.marg_fn, self.recorder = solve_fn, marg_fn, recorder
        self.image_cls = image_cls
        self.ft_map = Map()
        self.swt = None
        self.frames = []                      # Detail::frames (queued, waiting for IMU)
        self.imus = []
        self.frontal_imus = []
        self.latest_state = None              # FeatureTracker::latest_state (t, q, p, v, bg, ba)
        self.swt_latest = None                # FrontendWorker::latest_state
        self.initialized = False
        self.latest_pose = (0.0, None)
        self.outputs = []                     # (t, q, p) poses as the player would write them
        self.processed = 0

    # detail.cpp:103-128
    def track_camera(self, t, gray):
        f = Frame(self.cfg, self.image_cls(t, gray))
        self.frames.append(f)
        out = self.predict_pose(t)
        if t > self.latest_pose[0]:
            self.latest_pose = (t, out)
        if self.swt is not None and out is not None:
            self.outputs.append((t, out[0], out[1]))
        return out

    # detail.cpp:130-142
    def track_imu(self, t, w, a):
        w = np.asarray(w, dtype=np.float64)
        a = np.asarray(a, dtype=np.float64)
        self.frontal_imus.append((t, w, a))
        self.imus.append((t, w, a))
        while self.imus and self.frames:
            if self.imus[0][0] <= self.frames[0].t:
                self.frames[0].preintegration.data.append(self.imus.pop(0))
            else:
                self.feature_tracker_work(self.frames.pop(0))
        return self.predict_pose(t)

    # detail.cpp:15-28,144-177
    def predict_pose(self, t):
        if self.latest_state is None:
            return None
        st, q, p, v, bg, ba = self.latest_state
        q, p, v = q.copy(), p.copy(), v.copy()
        while self.frontal_imus and self.frontal_imus[0][0] <= st:
            self.frontal_imus.pop(0)
        for (ti, w, a) in self.frontal_imus:
            if ti <= t:
                dt = ti - st
                acc = GRAVITY + qrot(q, a - ba)
                p = p + dt * v + 0.5 * dt * dt * acc
                v = v + dt * acc
                q = qnormalized(qmul(q, expmap((w - bg) * dt)))
                st = ti
        return qmul(q, self.cfg.q_bo), p + qrot(q, self.cfg.p_bo)

    # feature_tracker.cpp:24-153
    def feature_tracker_work(self, frame):
        cfg = self.cfg
        m = self.ft_map
        frame.image.preprocess(cfg.clahe_clip_limit, cfg.clahe_width, cfg.clahe_height)
        is_init = self.swt_latest is not None
        if m.frame_num() > 0:
            if is_init:
                lt, lid, lq, lp, lv, lbg, lba = self.swt_latest
                idx = m.frame_index_by_id(lid)
                if idx is not None:
                    lf = m.frames[idx]
                    lf.q, lf.p, lf.v, lf.bg, lf.ba = lq.copy(), lp.copy(), lv.copy(), lbg.copy(), lba.copy()
                    for j in range(idx + 1, m.frame_num()):
                        fi, fj = m.frames[j - 1], m.frames[j]
                        fj.preintegration.integrate(fj.t, fi.bg, fi.ba, False, False)
                        fj.q, fj.p, fj.v, fj.bg, fj.ba = fj.preintegration.predict(fi.q, fi.p, fi.v, fi.bg, fi.ba)
                else:
                    self.latest_state = None
            last = m.frames[-1]
            if last.preintegration.data:
                if (not frame.preintegration.data) or (frame.preintegration.data[0][0] - last.t > 1.0e-5):
                    t_, w_, a_ = last.preintegration.data[-1]
                    frame.preintegration.data.insert(0, (last.t, w_.copy(), a_.copy()))
            frame.preintegration.integrate(frame.t, last.bg, last.ba, False, False)
            last.track_keypoints(frame)
            if is_init:
                frame.q, frame.p, frame.v, frame.bg, frame.ba = frame.preintegration.predict(last.q, last.p, last.v, last.bg, last.ba)
                self.latest_state = (frame.t, frame.q.copy(), frame.p.copy(), frame.v.copy(), frame.bg.copy(), frame.ba.copy())
            last.image.pyramid = None
        frame.detect_keypoints()
        m.attach_frame(frame)
        max_frames = cfg.max_frames if is_init else cfg.max_init_frames
        lid = self.swt_latest[1] if is_init else None
        while m.frame_num() > max_frames and (lid is None or m.frames[0].id < lid):      # nil() == SIZE_MAX before init
            m.erase_frame(0)
        self.processed += 1
        self.frontend_work(m.frames[-1].id)

    # frontend_worker.cpp:28-86
    def frontend_work(self, frame_id):
        if self.swt is None:
            self.swt = self.bootstrap(frame_id)
            if self.swt is not None:
                t, fid, q, p, v, bg, ba = self.swt.get_latest_state()
                self.swt_latest = (t, frame_id, q, p, v, bg, ba)
        else:
            if self.swt.mirror_frame(self.ft_map, frame_id):
                pass
            self.swt.track()
            t, fid, q, p, v, bg, ba = self.swt.get_latest_state()
            self.swt_latest = (t, frame_id, q, p, v, bg, ba)

    # initializer.cpp:22-76
    def mirror_keyframe_map(self, init_frame_id):
        cfg = self.cfg
        ft = self.ft_map
        last = ft.frame_index_by_id(init_frame_id)
        gap = cfg.init_keyframe_gap
        dist = gap * (cfg.init_keyframe_num - 1)
        if last < dist:
            return None
        first = last - dist
        idxs = [first + i * gap for i in range(cfg.init_keyframe_num)]
        m = Map()
        for ix in idxs:
            m.attach_frame(ft.frames[ix].clone())
        for j in range(1, m.frame_num()):
            oi, oj = ft.frames[idxs[j - 1]], ft.frames[idxs[j]]
            ni, nj = m.frames[j - 1], m.frames[j]
            for ki in range(oi.keypoint_num()):
                tr = oi.tracks[ki]
                if tr is not None:
                    kj = tr.get_keypoint_index(oj)
                    if kj is not None:
                        ni.get_track(ki, None, create=True).add_keypoint(nj, kj)
            nj.preintegration.data = []
            for f in range(idxs[j - 1], idxs[j]):
                nj.preintegration.data.extend([(t, w.copy(), a.copy()) for (t, w, a) in ft.frames[f + 1].preintegration.data])
        return m

    def _new_tracker_shell(self, m):
        swt = SlidingWindowTracker.__new__(SlidingWindowTracker)
        swt.map, swt.cfg = m, self.cfg
        swt.solve_fn = self.solve_fn or ceres_tr.solve
        swt.marg_fn = self.marg_fn or marg_mod.marginalize
        swt.recorder = self.recorder
        return swt

    # initializer.cpp:86-135: final visual-inertial BA, then the map becomes the sliding window
    def final_ba(self, swt):
        cfg, m = self.cfg, swt.map
        m.frames[0].fix_pose = True
        b = ProblemBuilder(cfg)
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
        for j in range(1, m.frame_num()):
            fi, fj = m.frames[j - 1], m.frames[j]
            if fj.preintegration.integrate(fj.t, fi.bg, fi.ba, True, True):
                b.add_pre(fi, fj, fj.preintegration, False)
        swt._solve(b, 'init_ba')
        for f in m.frames:
            f.keyframe = True
        # NOTE: FT_FIX_POSE stays set on frame 0 in the reference (initializer.cpp:86 never clears it).
        SlidingWindowTracker.__init__(swt, m, cfg, swt.solve_fn, swt.marg_fn, swt.recorder)
        return swt

    # mirror_keyframe_map + ground-truth replacement of init_sfm / init_imu + final VI-BA
    def bootstrap(self, init_frame_id):
        if self.gt_fn is None:
            return self.initialize(init_frame_id)
        m = self.mirror_keyframe_map(init_frame_id)
        if m is None:
            return None
        for f in m.frames:
            q, p, v = self.gt_fn(f.t)
            f.q, f.p, f.v = q.copy(), p.copy(), v.copy()
            f.bg, f.ba = np.zeros(3), np.zeros(3)
        for tr in list(m.tracks):
            if tr.keypoint_num() >= 2:
                x = tr.triangulate()
                if x is not None:
                    tr.set_landmark_point(x)
                    tr.triangulated = True
                    tr.valid = True
        return self.final_ba(self._new_tracker_shell(m))

    # initializer.cpp:78-155 with the reference's own SfM + IMU alignment (oracle/initializer.py)
    def initialize(self, init_frame_id):
        from .initializer import Initializer
        m = self.mirror_keyframe_map(init_frame_id)
        if m is None:
            return None
        swt = self._new_tracker_shell(m)
        ini = Initializer(self.cfg, m, swt._solve, ProblemBuilder)
        ok = ini.run()
        self.init_trace = ini.trace
        if not ok:
            return None
        return self.final_ba(swt)


def run_stream(stream, n_frames, cfg=None, bootstrap=True, **kw):
    """Drive a synthetic Stream through a Session the way xrslam-pc/player would
    (IMU samples in time order, image, ...).  Returns the Session."""
    cfg = cfg or Config(stream.w, stream.h, stream.K)
    sess = Session(cfg, gt_fn=stream.gt_state if bootstrap else None, **kw)
    imu = stream.imu(n_frames * stream.imu_per_cam + 2)
    k_imu = 0
    for k in range(n_frames):
        t_img = stream.frame_time(k)
        while k_imu < len(imu) and imu[k_imu, 0] <= t_img + 1e-9:
            sess.track_imu(imu[k_imu, 0], imu[k_imu, 1:4], imu[k_imu, 4:7])
            k_imu += 1
        sess.track_camera(t_img, stream.frame(k))
    return sess
