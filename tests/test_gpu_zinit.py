"""Row N3 of SURVEY.md 8(f): a session that is given NO bootstrap state runs the reference's initializer
(two-view SfM seed, PnP, visual BA, IMU alignment; xrslam_b200/csrc/session_init.inl + init_host.hpp) and is
compared with the oracle pipeline doing the same (tests/golden/stream7_init_traj.npz, made by
tests/golden/make_golden.py init)."""
import os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle.synth import Stream
from oracle import drive


def _ate(a, b):
    """RMSE of position differences after rigid SE3 alignment (evo_ape -a semantics)."""
    a, b = np.asarray(a), np.asarray(b)
    ca, cb = a.mean(0), b.mean(0)
    H = (b - cb).T @ (a - ca)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    R = (U @ D @ Vt).T
    al = (R @ (b - cb).T).T + ca
    return float(np.sqrt(np.mean(np.sum((al - a) ** 2, axis=1))))


def test_initializer_matches_oracle(xrvio, golden_dir):
    g = np.load(os.path.join(golden_dir, 'stream7_init_traj.npz'))
    ref, info_ref = g['traj'], g['info']
    s = Stream(7)
    ctx, outs = drive.run_streams_cuda([s], 70, bootstrap=False)
    info = ctx.get_init_info(0)
    assert info['last_fail'] == 0 and info['attempts'] == 1, info
    assert abs(info['matches'] - info_ref[0]) <= 2 and info['candidate'] == int(info_ref[1])
    assert abs(info['landmarks'] - info_ref[2]) <= 10
    assert abs(info['parallax'] - info_ref[3]) < 0.05 and abs(info['scale'] / info_ref[4] - 1) < 1e-2
    got = np.array([[t] + list(q) + list(p) for (t, q, p) in outs[0]])
    assert got.shape == ref.shape and np.allclose(got[:, 0], ref[:, 0])        # initialised on the same frame
    ate = _ate(ref[:, 5:8], got[:, 5:8])
    raw = np.linalg.norm(got[:, 5:8] - ref[:, 5:8], axis=1).max()
    assert ate <= 1e-3 and raw <= 2e-3, (ate, raw)
    gt = np.array([s.gt_state(t)[1] for t in got[:, 0]])
    assert _ate(gt, got[:, 5:8]) < 0.02                                         # metric scale and gravity are right
    st = ctx.stats()
    assert st['window_solves'] >= 5 and st['localize_solves'] >= 30
    ctx.close()


def test_initializer_can_be_disabled(xrvio):
    s = Stream(7)
    p = drive.params_for(s)
    p['initializer'] = 0
    from xrslam_b200 import capi
    ctx = capi.Context(p, 1)
    ctx, outs = drive.run_streams_cuda([s], 40, ctx=ctx, bootstrap=False)
    assert ctx.get_state(0) == 0 and len(outs[0]) == 0 and ctx.get_init_info(0)['attempts'] == 0
    ctx.close()


def test_per_frame_essential_ransac_changes_nothing(xrvio, golden_dir):
    """xrvio_params.essential_ransac = 1 runs the reference's per-frame 5-point RANSAC (frame.cpp:108-114) on the host; with its
    threshold every match is an inlier, so the trajectory is the one of the default path (the committed golden)."""
    ref = np.load(os.path.join(golden_dir, 'stream3_traj.npz'))['traj']
    s = Stream(3)
    p = drive.params_for(s)
    p['essential_ransac'] = 1
    from xrslam_b200 import capi
    ctx = capi.Context(p, 1)
    ctx, outs = drive.run_streams_cuda([s], 50, ctx=ctx)
    got = np.array([[t] + list(q) + list(pp) for (t, q, pp) in outs[0]])
    m = len(got)
    assert m >= 10 and np.allclose(got[:, 0], ref[:m, 0])
    assert np.linalg.norm(got[:, 5:8] - ref[:m, 5:8], axis=1).max() <= 2e-3
    assert ctx.get_init_info(0)['essential_rejected'] == 0
    ctx.close()


def test_session_api_error_paths(xrvio):
    """Errors are return codes with a message, never exceptions through the C boundary or silent acceptance."""
    from xrslam_b200 import capi
    p = capi.default_params(640, 480)
    ctx = capi.Context(p, 2)
    img = np.zeros((480, 640), np.uint8)
    with pytest.raises(capi.XrvioError):
        ctx.push_images([2], [img.ctypes.data], [0.0])                      # session index out of range
    with pytest.raises(capi.XrvioError):
        ctx.get_init_info(-1)
    for k in range(8):                                                       # the image queue holds 8 frames per session
        ctx.push_images([0], [img.ctypes.data], [0.05 * k])
    with pytest.raises(capi.XrvioError) as e:
        ctx.push_images([0], [img.ctypes.data], [0.45])
    with pytest.raises(capi.XrvioError):
        ctx.push_images([0, 0], [img.ctypes.data, img.ctypes.data], [0.5, 0.55])   # a session twice in one call
    assert 'queue full' in str(e.value)
    ctx.step()                                                               # no IMU yet: nothing is released, nothing breaks
    assert ctx.get_state(0) == 0 and not ctx.get_pose(0)['valid']
    ctx.close()
    bad = capi.default_params(640, 480)
    bad['sliding_window_size'] = 40                                          # beyond XRVIO_MAX_ACTIVE
    with pytest.raises(capi.XrvioError):
        capi.Context(bad, 1)
    bad = capi.default_params(640, 480)
    bad['distortion_model'] = 7
    with pytest.raises(capi.XrvioError):
        capi.Context(bad, 1)


def test_uninitialised_front_end_map_is_capped(xrvio):
    """Before initialisation latest_optimized_frame_id is nil() == SIZE_MAX, so FeatureTracker::work trims the front-end map to
    feature_tracker.max_init_frames = 60 (feature_tracker.cpp:46,132-135; ADVICE r1: the trim used to be gated on "initialised").
    A session that is never initialised (no bootstrap states, initializer off) must stop growing: no track can be longer than
    60 frames, and the oracle's map holds exactly 60 frames."""
    from oracle.synth import Stream
    from oracle.pipeline import Session, Config
    from oracle import drive
    s = Stream(3)
    n = 72
    p = drive.params_for(s)
    p['initializer'] = 0
    ctx = xrvio.Context(p, 1)
    ctx, outs = drive.run_streams_cuda([s], n, ctx=ctx, bootstrap=False)
    assert ctx.get_state(0) == 0 and len(outs[0]) == 0
    xy, tid, tl = ctx.get_features(0)
    assert len(tl) > 30 and 50 <= int(tl.max()) <= 60, int(tl.max())
    ctx.close()
    sess = Session(Config(s.w, s.h, s.K), gt_fn=None)
    sess.initialize = lambda init_frame_id: None            # same switch as params['initializer'] = 0
    imu = s.imu(n * s.imu_per_cam + 2)
    k_imu = 0
    for k in range(n):
        t_img = s.frame_time(k)
        while k_imu < len(imu) and imu[k_imu, 0] <= t_img + 1e-9:
            sess.track_imu(imu[k_imu, 0], imu[k_imu, 1:4], imu[k_imu, 4:7]); k_imu += 1
        sess.track_camera(t_img, s.frame(k))
    assert sess.swt is None and sess.ft_map.frame_num() == 60
    lens = [len(tr.refs) for tr in sess.ft_map.tracks]
    assert max(lens) == int(tl.max())
