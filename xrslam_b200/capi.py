"""ctypes binding of include/xrvio.h (operator level + session level)."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libxrvio.so")

MAX_CORNERS, MAX_PTS, PRE_BLOCK, MAX_ACTIVE, MAX_FRAMES, MAX_ITER = 256, 512, 512, 12, 96, 64


class XrvioError(RuntimeError):
    pass


# see bench.py / INTEGRATION.md: one hardware work queue per stream (the library uses up to 24); only effective if no CUDA
# context exists yet in this process
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')


class _Lib:
    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise XrvioError("libxrvio.so not built (%s); run `python -m xrslam_b200.build` -- there is no "
                                 "CPU fallback" % LIB_PATH)
            self._dll = C.CDLL(LIB_PATH)
            self._dll.xrvio_last_error.restype = C.c_char_p
        return self._dll

    def __getattr__(self, name):
        return getattr(self.load(), name)

    def check(self, rc):
        if rc != 0:
            raise XrvioError("xrvio error %d: %s" % (rc, self.load().xrvio_last_error().decode()))


lib = _Lib()

rig_dtype = np.dtype([("cam_q", "f8", 4), ("cam_p", "f8", 3), ("imu_q", "f8", 4), ("imu_p", "f8", 3), ("sic", "f8", 2),
                      ("cov_w", "f8", 9), ("cov_a", "f8", 9), ("cov_bg", "f8", 9), ("cov_ba", "f8", 9)], align=True)
preint_job_dtype = np.dtype([("imu_off", "i4"), ("n", "i4"), ("flags", "i4"), ("pad", "i4"), ("t_end", "f8"),
                             ("bg", "f8", 3), ("ba", "f8", 3)], align=True)
rpe_dtype = np.dtype([("tgt", "i2"), ("ref", "i2"), ("lm", "i4"), ("prior", "i4"), ("pad", "i4"), ("zt", "f8", 3),
                      ("zr", "f8", 3)], align=True)
rot_dtype = np.dtype([("tgt", "i2"), ("ref", "i2"), ("pad", "i4"), ("zt", "f8", 3), ("zr", "f8", 3)], align=True)
pre_ref_dtype = np.dtype([("i", "i2"), ("j", "i2"), ("prior", "i4"), ("block", "i4"), ("pad", "i4")], align=True)
solve_desc_dtype = np.dtype([("frame_off", "i4"), ("n_frames", "i4"), ("n_active", "i4"), ("lm_off", "i4"), ("n_lm", "i4"),
                             ("n_lm_free", "i4"), ("rpe_off", "i4"), ("n_rpe", "i4"), ("rot_off", "i4"), ("n_rot", "i4"),
                             ("pre_off", "i4"), ("n_pre", "i4"), ("marg_k", "i4"), ("marg_frames_off", "i4"),
                             ("marg_off", "i8"), ("scratch_off", "i8"), ("max_iterations", "i4"), ("pad", "i4")], align=True)
solve_out_dtype = np.dtype([("iterations", "i4"), ("termination", "i4"), ("n_successful", "i4"), ("n_unsuccessful", "i4"),
                            ("initial_cost", "f8"), ("final_cost", "f8"), ("cost_trace", "f8", MAX_ITER)], align=True)
marg_desc_dtype = np.dtype([("frame_off", "i4"), ("n_frames", "i4"), ("lm_off", "i4"), ("n_lm", "i4"), ("rpe_off", "i4"),
                            ("n_rpe", "i4"), ("pre_off", "i4"), ("n_pre", "i4"), ("marg_k", "i4"), ("marg_frames_off", "i4"),
                            ("marg_off", "i8"), ("out_off", "i8"), ("scratch_off", "i8"), ("eigen_form", "i4"), ("pad", "i4")], align=True)
tri_job_dtype = np.dtype([("obs_off", "i4"), ("n_obs", "i4")], align=True)

assert rig_dtype.itemsize == 8 * (4 + 3 + 4 + 3 + 2 + 36)
assert rpe_dtype.itemsize == 64 and rot_dtype.itemsize == 56 and pre_ref_dtype.itemsize == 16
assert solve_desc_dtype.itemsize == 80 and marg_desc_dtype.itemsize == 72
assert solve_out_dtype.itemsize == 16 + 16 + 8 * MAX_ITER


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_rig(cam_q, cam_p, imu_q, imu_p, sic, cov_w, cov_a, cov_bg, cov_ba):
    r = np.zeros(1, rig_dtype)
    r["cam_q"], r["cam_p"], r["imu_q"], r["imu_p"], r["sic"] = cam_q, cam_p, imu_q, imu_p, sic
    r["cov_w"], r["cov_a"] = np.asarray(cov_w).reshape(9), np.asarray(cov_a).reshape(9)
    r["cov_bg"], r["cov_ba"] = np.asarray(cov_bg).reshape(9), np.asarray(cov_ba).reshape(9)
    return r


def pyramid_layout(w, h):
    ws, hs = (C.c_int * 4)(), (C.c_int * 4)()
    offs = (C.c_longlong * 4)()
    tot = C.c_longlong()
    lib.check(lib.xrvio_pyramid_layout(w, h, ws, hs, offs, C.byref(tot)))
    return list(ws), list(hs), list(offs), tot.value


def op_undistort(images, model, K, D):
    """images: (n, h, w) u8.  model: 1 radtan (cv::undistort), 2 equidistant (ImageUndistorter + cv::remap)."""
    images = np.ascontiguousarray(images, np.uint8)
    n, h, w = images.shape
    out = np.empty_like(images)
    K = np.ascontiguousarray(K, np.float64); D = np.ascontiguousarray(D, np.float64)
    lib.check(lib.xrvio_op_undistort(_p(images), _p(out), n, w, h, int(model), _p(K), _p(D)))
    return out


def op_clahe(imgs, clip=6.0, tiles=(8, 8)):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    out = np.empty_like(imgs)
    lib.check(lib.xrvio_op_clahe(_p(imgs), _p(out), n, w, h, C.c_double(clip), tiles[0], tiles[1]))
    return out


def op_pyramid(imgs, with_deriv=False):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    ws, hs, offs, tot = pyramid_layout(w, h)
    pyr = np.empty((n, tot), np.uint8)
    der = np.empty((n, tot, 2), np.int16) if with_deriv else None
    lib.check(lib.xrvio_op_pyramid(_p(imgs), n, w, h, _p(pyr), _p(der)))
    levels = [[pyr[i, offs[l]:offs[l] + ws[l] * hs[l]].reshape(hs[l], ws[l]) for l in range(4)] for i in range(n)]
    derivs = None
    if with_deriv:
        derivs = [[der[i, offs[l]:offs[l] + ws[l] * hs[l]].reshape(hs[l], ws[l], 2) for l in range(4)] for i in range(n)]
    return levels, derivs


def op_gftt(imgs, max_corners=200, quality=1e-3, min_distance=20.0):
    imgs = np.ascontiguousarray(imgs, np.uint8)
    n, h, w = imgs.shape
    corners = np.zeros((n, MAX_CORNERS, 3), np.float32)
    count = np.zeros(n, np.int32)
    lib.check(lib.xrvio_op_gftt(_p(imgs), n, w, h, max_corners, C.c_double(quality), C.c_double(min_distance), _p(corners), _p(count)))
    return [corners[i, :count[i]].copy() for i in range(n)]


def op_klt(prev_imgs, next_imgs, prev_pts, init_pts):
    """prev_pts / init_pts: lists of (k,2) float arrays.  Returns lists next_pts (k,2) f32, status (k,) u8."""
    prev_imgs = np.ascontiguousarray(prev_imgs, np.uint8)
    next_imgs = np.ascontiguousarray(next_imgs, np.uint8)
    n, h, w = prev_imgs.shape
    pp = np.zeros((n, MAX_PTS, 2), np.float32)
    nn = np.zeros((n, MAX_PTS, 2), np.float32)
    cnt = np.zeros(n, np.int32)
    for i in range(n):
        k = len(prev_pts[i])
        cnt[i] = k
        pp[i, :k] = np.asarray(prev_pts[i], np.float32).reshape(k, 2)
        nn[i, :k] = np.asarray(init_pts[i], np.float32).reshape(k, 2)
    st = np.zeros((n, MAX_PTS), np.uint8)
    lib.check(lib.xrvio_op_klt(_p(prev_imgs), _p(next_imgs), n, w, h, _p(pp), _p(nn), _p(cnt), _p(st)))
    return [nn[i, :cnt[i]].copy() for i in range(n)], [st[i, :cnt[i]].copy() for i in range(n)]


def op_poisson(lists, n_preset, radius, veto=None):
    n = len(lists)
    pts = np.zeros((n, MAX_PTS, 2), np.float64)
    tot = np.zeros(n, np.int32)
    pre = np.asarray(n_preset, np.int32).copy()
    vt = None if veto is None else np.zeros((n, MAX_PTS), np.uint8)
    for i in range(n):
        k = len(lists[i])
        tot[i] = k
        pts[i, :k] = np.asarray(lists[i], np.float64).reshape(k, 2)
        if veto is not None:
            vt[i, :k] = veto[i]
    keep = np.zeros((n, MAX_PTS), np.uint8)
    lib.check(lib.xrvio_op_poisson(_p(pts), _p(pre), _p(tot), _p(vt), _p(keep), n, C.c_double(radius)))
    return [keep[i, :tot[i]].copy() for i in range(n)]


def op_preintegrate(rig, jobs, imu):
    imu = np.ascontiguousarray(imu, np.float64).reshape(-1, 7)
    out = np.zeros((len(jobs), PRE_BLOCK), np.float64)
    lib.check(lib.xrvio_op_preintegrate(_p(rig), _p(jobs), len(jobs), _p(imu), len(imu), _p(out)))
    return out


def op_solve(rig, desc, states, flags, inv_depth, rpe, rot, pre, pre_blocks, marg_frames, marg):
    out = np.zeros(len(desc), solve_out_dtype)
    pre_blocks = np.ascontiguousarray(pre_blocks, np.float64).reshape(-1, PRE_BLOCK)
    lib.check(lib.xrvio_op_solve(_p(rig), _p(desc), len(desc), _p(states), len(states), _p(flags), _p(inv_depth), len(inv_depth),
                                 _p(rpe), len(rpe), _p(rot), len(rot), _p(pre), len(pre), _p(pre_blocks), len(pre_blocks),
                                 _p(marg_frames), len(marg_frames), _p(marg), C.c_longlong(len(marg)), _p(out)))
    return out


def last_solve_ms():
    """Device time of the solve kernel launched by the last op_solve call (CUDA events), ms."""
    f = lib.xrvio_debug_last_solve_ms
    f.restype = C.c_double
    return float(f())


def fp64_peak():
    """(DFMA TFLOP/s, DMMA TFLOP/s) measured on the current device."""
    a, b = C.c_double(0.0), C.c_double(0.0)
    lib.check(lib.xrvio_debug_fp64_peak(C.byref(a), C.byref(b)))
    return a.value, b.value


def op_marginalize(rig, desc, states, inv_depth, rpe, pre, pre_blocks, marg_frames, marg, n_out):
    out = np.zeros(n_out, np.float64)
    rank = np.zeros(len(desc), np.int32)
    pre_blocks = np.ascontiguousarray(pre_blocks, np.float64).reshape(-1, PRE_BLOCK)
    lib.check(lib.xrvio_op_marginalize(_p(rig), _p(desc), len(desc), _p(states), len(states), _p(inv_depth), len(inv_depth),
                                       _p(rpe), len(rpe), _p(pre), len(pre), _p(pre_blocks), len(pre_blocks),
                                       _p(marg_frames), len(marg_frames), _p(marg), C.c_longlong(len(marg)), _p(out),
                                       C.c_longlong(n_out), _p(rank)))
    return out, rank


def op_triangulate(jobs, views):
    views = np.ascontiguousarray(views, np.float64).reshape(-1, 16)
    out = np.zeros((len(jobs), 4), np.float64)
    lib.check(lib.xrvio_op_triangulate(_p(jobs), len(jobs), _p(views), len(views), _p(out)))
    return out


# ------------------------------------------------------------------------------------------------
# session level
# ------------------------------------------------------------------------------------------------
params_dtype = np.dtype([("width", "i4"), ("height", "i4"), ("fx", "f8"), ("fy", "f8"), ("cx", "f8"), ("cy", "f8"),
                         ("q_bc", "f8", 4), ("p_bc", "f8", 3), ("q_bi", "f8", 4), ("p_bi", "f8", 3), ("q_bo", "f8", 4), ("p_bo", "f8", 3),
                         ("cov_g", "f8", 9), ("cov_a", "f8", 9), ("cov_bg", "f8", 9), ("cov_ba", "f8", 9), ("keypoint_cov", "f8", 2),
                         ("sliding_window_size", "i4"), ("subframe_size", "i4"), ("force_keyframe_landmarks", "i4"),
                         ("min_keypoint_distance", "f8"), ("max_keypoint_detection", "i4"), ("max_init_frames", "i4"),
                         ("max_frames", "i4"), ("predict_keypoints", "i4"), ("clahe_clip_limit", "f8"), ("clahe_width", "i4"),
                         ("clahe_height", "i4"), ("init_keyframe_num", "i4"), ("init_keyframe_gap", "i4"),
                         ("solver_iteration_limit", "i4"), ("rotation_misalignment_threshold", "f8"),
                         ("rotation_ransac_threshold", "f8"), ("prior_eigen_form", "i4"), ("initializer", "i4"),
                         ("init_min_matches", "i4"), ("init_min_triangulation", "i4"), ("init_min_landmarks", "i4"),
                         ("init_refine_imu", "i4"), ("init_min_parallax", "f8"), ("random_seed", "i4"), ("distortion_model", "i4"),
                         ("distortion", "f8", 4), ("essential_ransac", "i4"), ("pad", "i4")], align=True)
imu_dtype = np.dtype([("t", "f8"), ("w", "f8", 3), ("a", "f8", 3)], align=True)
pose_dtype = np.dtype([("q", "f8", 4), ("p", "f8", 3), ("t", "f8"), ("valid", "i4"), ("pad", "i4")], align=True)
init_info_dtype = np.dtype([("attempts", "i4"), ("last_fail", "i4"), ("matches", "i4"), ("candidate", "i4"), ("landmarks", "i4"),
                            ("essential_rejected", "i4"), ("parallax", "f8"), ("scale", "f8")], align=True)
stats_dtype = np.dtype([("frames_processed", "i8"), ("kernels_launched", "i8"), ("h2d_bytes", "i8"), ("d2h_bytes", "i8"),
                        ("window_solves", "i8"), ("localize_solves", "i8"), ("subwindow_solves", "i8"), ("marginalizations", "i8"),
                        ("last_window_cost_initial", "f8"), ("last_window_cost_final", "f8"),
                        ("window_iterations", "i8"), ("window_at_limit", "i8"), ("graph_launches", "i8")], align=True)


def default_params(width=640, height=480):
    p = np.zeros(1, params_dtype)
    lib.check(lib.xrvio_default_params(_p(p), width, height))
    return p


class Context:
    """n_sessions VIO sessions on one GPU (xrvio_create ... xrvio_destroy)."""

    def __init__(self, params, n_sessions, device=0):
        self.params = params
        self.n = n_sessions
        self._h = C.c_void_p()
        lib.check(lib.xrvio_create(_p(params), n_sessions, device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib.xrvio_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push_imu(self, sessions, samples):
        sessions = np.ascontiguousarray(sessions, np.int32)
        samples = np.ascontiguousarray(samples)
        assert samples.dtype == imu_dtype and len(samples) == len(sessions)
        lib.check(lib.xrvio_push_imu(self._h, _p(sessions), _p(samples), len(sessions)))

    def push_bootstrap_state(self, session, t, q, p, v):
        q, p, v = (np.ascontiguousarray(x, np.float64) for x in (q, p, v))
        lib.check(lib.xrvio_push_bootstrap_state(self._h, int(session), C.c_double(t), _p(q), _p(p), _p(v)))

    def push_images(self, sessions, pointers, times, stride=None, on_device=False):
        """pointers: integer addresses (host numpy .ctypes.data or device data_ptr) of 8-bit images."""
        n = len(sessions)
        sessions = np.ascontiguousarray(sessions, np.int32)
        ptrs = (C.c_void_p * n)(*[int(x) for x in pointers])
        times = np.ascontiguousarray(times, np.float64)
        st = None if stride is None else np.ascontiguousarray(stride, np.int32)
        lib.check(lib.xrvio_push_images(self._h, _p(sessions), ptrs, _p(st), _p(times), n, 1 if on_device else 0))

    def step(self):
        lib.check(lib.xrvio_step(self._h))

    def get_pose(self, session):
        p = np.zeros(1, pose_dtype)
        lib.check(lib.xrvio_get_pose(self._h, int(session), _p(p)))
        return p[0]

    def get_latest_state(self, session):
        t = C.c_double()
        s = np.zeros(16)
        rc = lib.xrvio_get_latest_state(self._h, int(session), C.byref(t), _p(s))
        return (t.value, s) if rc == 0 else None

    def get_state(self, session):
        return lib.xrvio_get_state(self._h, int(session))

    def get_landmarks(self, session, capacity=4096):
        xyz = np.zeros((capacity, 3))
        n = lib.xrvio_get_landmarks(self._h, int(session), _p(xyz), capacity)
        return xyz[:min(n, capacity)]

    KERNEL_FAMILIES = ('clahe', 'pyramid', 'gftt', 'klt', 'poisson', 'preintegrate', 'solve', 'marginalize', 'triangulate')

    def set_profiling(self, on=True):
        lib.check(lib.xrvio_set_profiling(self._h, 1 if on else 0))

    def kernel_times(self):
        ms = np.zeros(16)
        ln = np.zeros(16, np.int64)
        n = lib.xrvio_get_kernel_times(self._h, _p(ms), _p(ln), 16)
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(self.KERNEL_FAMILIES[:n])}

    def phase_times(self):
        ms = np.zeros(32)
        lib.xrvio_get_phase_times(self._h, _p(ms), 32)
        return ms

    def set_async(self, enable=True):
        lib.check(lib.xrvio_set_async(self._h, 1 if enable else 0))

    def sync(self):
        lib.check(lib.xrvio_sync(self._h))

    def get_features(self, session):
        """(xy [n,2], track ids [n], track lengths [n]) of the newest front-end frame."""
        n = lib.xrvio_get_features(self._h, int(session), None, None, None, 0)
        if n < 0:
            lib.check(n)
        xy = np.zeros((max(n, 1), 2)); tid = np.zeros(max(n, 1), np.int64); tl = np.zeros(max(n, 1), np.int32)
        n = lib.xrvio_get_features(self._h, int(session), _p(xy), _p(tid), _p(tl), int(n))
        return xy[:n], tid[:n], tl[:n]

    def get_init_info(self, session):
        s = np.zeros(1, init_info_dtype)
        lib.check(lib.xrvio_get_init_info(self._h, int(session), _p(s)))
        return {k: s[0][k].item() for k in init_info_dtype.names}

    def stats(self):
        s = np.zeros(1, stats_dtype)
        lib.check(lib.xrvio_get_stats(self._h, _p(s)))
        return s[0]
