"""Regenerates the committed golden fixtures under tests/golden/ from the CPU oracle.

Run in the build container (needs /root/reference for the two EuRoC frames of the reference's
test_feature_track known-answer test):   python tests/golden/make_golden.py
Outputs:
  euroc_v101_a.png / euroc_v101_b.png   the reference's fixture frames (xrslam-test/data/*.png, test DATA)
  frontend_kat.npz      oracle front-end results on that pair (counts the reference asserts: 164 / 161)
  backend_problems.npz  one recorded problem of every solve kind + a marginalisation, before / after the oracle
  preint.npz            preintegration jobs and oracle results
  stream3_traj.npz      oracle trajectory of synthetic stream seed 3 (first 66 frames)
  stream7_traj1200.npz  oracle trajectory of stream seed 7 over 60 s (BASELINE config 2; `make_golden.py minute`)
  stream7_pause_traj.npz oracle trajectory of stream seed 7 with a 1.5 s full stop (pure-rotation branches)
  stream7_init_traj.npz oracle trajectory of stream seed 7 initialised by the reference initializer (no bootstrap states)
  stream7_traj140.npz   oracle trajectory of synthetic stream seed 7, 140 frames (steady-state marginalisation)
  stress_config5.npz    oracle cost trace / final states of BASELINE config 5 (30 keyframes x 2000 landmarks; `make_golden.py stress`)
"""
import os
import shutil
import sys
import copy
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..', '..'))
import cv2  # noqa: E402
from oracle.synth import Stream  # noqa: E402
from oracle.pipeline import run_stream, Config, Frame, Map  # noqa: E402
from oracle.frontend import OracleImage  # noqa: E402
from oracle.preintegrator import PreIntegrator  # noqa: E402
from oracle import problem_io  # noqa: E402

REF = '/root/reference/xrslam-test/data'


def frontend_kat():
    a = os.path.join(REF, '1403715282262142976.png')
    b = os.path.join(REF, '1403715282312143104.png')
    if os.path.exists(a):
        shutil.copyfile(a, os.path.join(HERE, 'euroc_v101_a.png'))
        shutil.copyfile(b, os.path.join(HERE, 'euroc_v101_b.png'))
        os.chmod(os.path.join(HERE, 'euroc_v101_a.png'), 0o644)
        os.chmod(os.path.join(HERE, 'euroc_v101_b.png'), 0o644)
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    ims = [cv2.undistort(cv2.imread(os.path.join(HERE, n), 0), K, D) for n in ('euroc_v101_a.png', 'euroc_v101_b.png')]
    cfg = Config(752, 480, (458.654, 457.296, 367.215, 248.375))
    f0 = Frame(cfg, OracleImage(0.0, ims[0]))
    f0.image.preprocess()
    f0.detect_keypoints()
    m = Map()
    m.attach_frame(f0)
    f1 = Frame(cfg, OracleImage(0.05, ims[1]))
    f1.image.preprocess()
    f0.track_keypoints(f1)
    tracked = sum(1 for t in f1.tracks if t is not None)
    np.savez_compressed(os.path.join(HERE, 'frontend_kat.npz'), detected=f0.keypoint_num(), tracked=tracked,
                        no_translation=f1.no_translation, misalignment=f0.last_misalignment,
                        keypoints=np.array([[b[0] / b[2] * cfg.K[0] + cfg.K[2], b[1] / b[2] * cfg.K[1] + cfg.K[3]] for b in f0.bearings]),
                        clahe_sum=np.array([int(f0.image.image.astype(np.int64).sum()), int(f1.image.image.astype(np.int64).sum())]))
    print('frontend KAT: detected %d (reference asserts 164), tracked %d (reference asserts 161)' % (f0.keypoint_num(), tracked))


def backend_problems():
    s = Stream(3)
    saved, post = {}, {}
    count = {}

    def rec(kind, phase, pb, summ=None):
        count[(kind, phase)] = count.get((kind, phase), 0) + 1
        want = 4 if kind in ('window', 'marg') else 1     # a later window / marginalisation (full 10-frame prior)
        if count[(kind, phase)] != want:
            return
        if phase == 'pre':
            saved[kind] = copy.deepcopy(pb)
        else:
            post[kind] = (copy.deepcopy(pb), summ)
    sess = run_stream(s, 66, recorder=rec)
    out = {}
    for kind in ('init_ba', 'localize', 'subwindow_t', 'window'):
        out.update(problem_io.problem_to_dict(saved[kind], kind + '/pre/'))
        pb, summ = post[kind]
        out[kind + '/post/states'] = pb.states
        out[kind + '/post/inv_depth'] = pb.inv_depth
        out[kind + '/post/cost'] = np.array([summ.initial_cost, summ.final_cost])
        out[kind + '/post/iterations'] = np.array(len(summ.iterations))
        out[kind + '/post/trace'] = np.array([it['cost'] for it in summ.iterations])
    for k, v in problem_io.marg_inputs_to_dict(saved['marg']).items():
        out['marg/pre/' + k] = v
    newp = post['marg'][0]
    out['marg/post/S'] = newp.sqrt_inv_cov
    out['marg/post/infovec'] = newp.infovec
    out['marg/post/lin'] = newp.lin
    np.savez_compressed(os.path.join(HERE, 'backend_problems.npz'), **out)
    traj = np.array([[t] + list(q) + list(p) for (t, q, p) in sess.outputs])
    np.savez_compressed(os.path.join(HERE, 'stream3_traj.npz'), traj=traj)
    print('backend problems saved; trajectory rows', len(traj))


def preint():
    rng = np.random.default_rng(11)
    cfg = Config()
    jobs = []
    imu_all = []
    res = []
    for n in (1, 3, 10, 37):
        t = 0.3
        data = []
        for k in range(n):
            data.append((t, rng.normal(size=3) * 0.4, np.array([0.1, -0.2, 9.7]) + rng.normal(size=3)))
            t += 0.005 * (1 + 0.1 * rng.random())
        t_end = t + 0.002
        bg, ba = rng.normal(size=3) * 0.01, rng.normal(size=3) * 0.05
        pre = PreIntegrator(cfg.cov_g, cfg.cov_a, cfg.cov_bg, cfg.cov_ba)
        pre.data = data
        flags = 3 if n > 1 else 1      # a single Euler step leaves the 9x9 covariance rank-deficient
        pre.integrate(t_end, bg, ba, True, n > 1)
        jobs.append([len(imu_all), n, t_end] + list(bg) + list(ba) + [flags])
        imu_all.extend([[d[0]] + list(d[1]) + list(d[2]) for d in data])
        res.append(np.concatenate([[pre.t], pre.q, pre.p, pre.v, pre.cov.ravel(), pre.sqrt_inv_cov.ravel(), pre.dq_dbg.ravel(),
                                   pre.dp_dbg.ravel(), pre.dp_dba.ravel(), pre.dv_dbg.ravel(), pre.dv_dba.ravel()]))
    np.savez_compressed(os.path.join(HERE, 'preint.npz'), jobs=np.array(jobs), imu=np.array(imu_all), result=np.array(res))
    print('preintegration goldens saved')


def long_trajectory():
    s = Stream(7)
    sess = run_stream(s, 140)
    traj = np.array([[t] + list(q) + list(p) for (t, q, p) in sess.outputs])
    np.savez_compressed(os.path.join(HERE, 'stream7_traj140.npz'), traj=traj)
    print('long trajectory rows', len(traj))


def minute_trajectory():
    """BASELINE config 2: one stream, 60 s = 1200 frames (about 11 min of oracle time)."""
    s = Stream(7)
    sess = run_stream(s, 1200)
    traj = np.array([[t] + list(q) + list(p) for (t, q, p) in sess.outputs])
    np.savez_compressed(os.path.join(HERE, 'stream7_traj1200.npz'), traj=traj)
    print('60 s trajectory rows', len(traj))


def pause_trajectory():
    """Stream 7 with a 1.5 s full stop: NO_TRANSLATION frames, pure-rotation sub-windows, subframe compaction."""
    s = Stream(7, pause=(3.0, 1.5, 0.5))
    sess = run_stream(s, 135)
    traj = np.array([[t] + list(q) + list(p) for (t, q, p) in sess.outputs])
    np.savez_compressed(os.path.join(HERE, 'stream7_pause_traj.npz'), traj=traj)
    print('pause trajectory rows', len(traj))


def initializer_trajectory():
    """Stream 7 WITHOUT ground-truth bootstrap: the oracle runs the reference initializer (oracle/initializer.py)."""
    s = Stream(7)
    sess = run_stream(s, 70, bootstrap=False)
    traj = np.array([[t] + list(q) + list(p) for (t, q, p) in sess.outputs])
    tr = sess.init_trace
    np.savez_compressed(os.path.join(HERE, 'stream7_init_traj.npz'), traj=traj,
                        info=np.array([tr['matches'], tr['candidate'], tr['landmarks'], tr['parallax'], tr['scale']], np.float64))
    print('initializer trajectory rows', len(traj), tr['matches'], tr['candidate'], tr['landmarks'], tr['scale'])


def stress_config5():
    """BASELINE config 5: oracle cost trace and final states of the 30-keyframe x 2000-landmark solve (oracle/stress.py)."""
    from oracle import stress, ceres_tr
    seed, n_kf, n_lm = 5, 30, 2000
    pb = stress.make_stress_problem(seed, n_kf, n_lm)
    summ = ceres_tr.solve(pb)
    np.savez_compressed(os.path.join(HERE, 'stress_config5.npz'), seed=seed, n_kf=n_kf, n_lm=n_lm,
                        trace=np.array([it['cost'] for it in summ.iterations]), successful=np.array([it['successful'] for it in summ.iterations]),
                        states=pb.states, inv_depth_sum=pb.inv_depth.sum(), final_cost=summ.final_cost)
    print('stress config 5: cost %.6f -> %.6f in %d iterations' % (summ.initial_cost, summ.final_cost, len(summ.iterations) - 1))


if __name__ == '__main__':
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == 'stress':
        stress_config5()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'init':
        initializer_trajectory()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'minute':
        minute_trajectory()
        sys.exit(0)
    long_trajectory()
    pause_trajectory()
    initializer_trajectory()
    frontend_kat()
    preint()
    backend_problems()
    stress_config5()
