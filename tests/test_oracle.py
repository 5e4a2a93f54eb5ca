"""CPU tests (no GPU): the oracle against the reference's own known answers and the committed
golden vectors, the finite-difference Jacobian recipe of the reference's validator, host logic, and
the C ABI surface."""
import os
import re
import ctypes
import numpy as np
import pytest

from oracle.lie import expmap, logmap, qmul, qconj, qrot, qnormalized, right_jacobian
from oracle import factors as F
from oracle.preintegrator import PreIntegrator
from oracle.pipeline import Config
from oracle import problem_io, ceres_tr
from oracle import marginalization as marg_mod
from oracle.poisson import PoissonDiskFilter2
from oracle.ransac import LotBox, find_rotation_matrix


def _plus(q, d):
    return qnormalized(qmul(q, expmap(d)))


def _fd(f, n, eps=1e-6):
    cols = []
    for k in range(n):
        d = np.zeros(n); d[k] = eps
        cols.append((f(d) - f(-d)) / (2 * eps))
    return np.stack(cols, axis=1)


# -- reference KAT: xrslam-test/test/src/test_feature_track.cpp:41,55,64 -------------------------
def test_frontend_known_answer(golden_dir):
    """164 detected / 161 tracked / not NO_TRANSLATION on the two EuRoC V1_01 frames.  The reference pins no
    OpenCV version; cv2 4.13 gives 165 / 162 -> +-2."""
    import cv2
    from oracle.pipeline import Frame, Map
    from oracle.frontend import OracleImage
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    ims = [cv2.undistort(cv2.imread(os.path.join(golden_dir, n), 0), K, D) for n in ('euroc_v101_a.png', 'euroc_v101_b.png')]
    cfg = Config(752, 480, (458.654, 457.296, 367.215, 248.375))
    f0 = Frame(cfg, OracleImage(0.0, ims[0])); f0.image.preprocess(); f0.detect_keypoints()
    assert abs(f0.keypoint_num() - 164) <= 2
    m = Map(); m.attach_frame(f0)
    f1 = Frame(cfg, OracleImage(0.05, ims[1])); f1.image.preprocess()
    f0.track_keypoints(f1)
    assert abs(sum(t is not None for t in f1.tracks) - 161) <= 2
    assert not f1.no_translation
    g = np.load(os.path.join(golden_dir, 'frontend_kat.npz'))
    assert int(g['detected']) == f0.keypoint_num() and int(g['tracked']) == sum(t is not None for t in f1.tracks)


# -- reference validator recipe: estimation/ceres/cost_function_validator.h:22-26 -------------------
def test_factor_jacobians_finite_difference():
    rng = np.random.default_rng(0)
    cfg = Config()
    S = np.diag(cfg.sqrt_inv_cov())
    qt = qnormalized(rng.normal(size=4)); pt = rng.normal(size=3)
    qr = _plus(qt, rng.normal(size=3) * 0.1); pr = pt + rng.normal(size=3) * 0.3
    zr = np.array([0.1, -0.2, 1.0]); zr /= np.linalg.norm(zr); invd = 0.25
    y = qrot(qconj(cfg.q_bc), qrot(qconj(qt), qrot(qr, qrot(cfg.q_bc, zr / invd) + cfg.p_bc) + pr - pt) - cfg.p_bc)
    zt = y / np.linalg.norm(y)
    args = (zt, zr, cfg.q_bc, cfg.p_bc, S)
    r, J = F.reprojection(qt, pt, qr, pr, invd, *args)
    tol = 2e-6
    assert np.abs(_fd(lambda d: F.reprojection(_plus(qt, d), pt, qr, pr, invd, *args, False)[0], 3) - J['q_tgt']).max() < tol * np.abs(J['q_tgt']).max()
    assert np.abs(_fd(lambda d: F.reprojection(qt, pt + d, qr, pr, invd, *args, False)[0], 3) - J['p_tgt']).max() < tol * np.abs(J['p_tgt']).max()
    assert np.abs(_fd(lambda d: F.reprojection(qt, pt, _plus(qr, d), pr, invd, *args, False)[0], 3) - J['q_ref']).max() < tol * np.abs(J['q_ref']).max()
    assert np.abs(_fd(lambda d: F.reprojection(qt, pt, qr, pr + d, invd, *args, False)[0], 3) - J['p_ref']).max() < tol * np.abs(J['p_ref']).max()
    assert np.abs(_fd(lambda d: F.reprojection(qt, pt, qr, pr, invd + d[0], *args, False)[0], 1)[:, 0] - J['inv_depth']).max() < tol * np.abs(J['inv_depth']).max()
    r2, J2 = F.rotation_prior(qt, qr, zt, zr, cfg.q_bc, cfg.p_bc, S)
    assert np.abs(_fd(lambda d: F.rotation_prior(_plus(qt, d), qr, zt, zr, cfg.q_bc, cfg.p_bc, S, False)[0], 3) - J2['q_tgt']).max() < tol * np.abs(J2['q_tgt']).max()
    # preintegration factor
    pre = PreIntegrator(cfg.cov_g, cfg.cov_a, cfg.cov_bg, cfg.cov_ba)
    t = 0.0
    for k in range(10):
        pre.data.append((t, rng.normal(size=3) * 0.3, np.array([0, 0, 9.8]) + rng.normal(size=3))); t += 0.005
    bg0, ba0 = rng.normal(size=3) * 0.01, rng.normal(size=3) * 0.05
    pre.integrate(0.05, bg0, ba0, True, True)
    qi, pi, vi = qnormalized(rng.normal(size=4)), rng.normal(size=3), rng.normal(size=3)
    qj, pj, vj, _, _ = pre.predict(qi, pi, vi, bg0, ba0)
    bgi, bai = bg0 + rng.normal(size=3) * 1e-3, ba0 + rng.normal(size=3) * 1e-2
    bgj, baj = bgi + 1e-4, bai - 1e-3
    iq, ip = _plus(np.array([0, 0, 0, 1.0]), np.array([0.01, 0.02, -0.03])), np.array([0.01, -0.02, 0.03])

    def ev(di, dj, jac=False):
        si = (_plus(qi, di[0:3]), pi + di[3:6], vi + di[6:9], bgi + di[9:12], bai + di[12:15])
        sj = (_plus(qj, dj[0:3]), pj + dj[3:6], vj + dj[6:9], bgj + dj[9:12], baj + dj[12:15])
        return F.preintegration(si, sj, pre, bg0, ba0, iq, ip, jac)
    r, Ji, Jj = ev(np.zeros(15), np.zeros(15), True)
    z = np.zeros(15)
    assert np.abs(_fd(lambda d: ev(d, z)[0], 15) - Ji).max() < tol * np.abs(Ji).max()
    assert np.abs(_fd(lambda d: ev(z, d)[0], 15) - Jj).max() < tol * np.abs(Jj).max()


def test_preintegration_golden(golden_dir):
    d = np.load(os.path.join(golden_dir, 'preint.npz'))
    cfg = Config()
    for k, j in enumerate(d['jobs']):
        pre = PreIntegrator(cfg.cov_g, cfg.cov_a, cfg.cov_bg, cfg.cov_ba)
        off, n = int(j[0]), int(j[1])
        pre.data = [(r[0], r[1:4].copy(), r[4:7].copy()) for r in d['imu'][off:off + n]]
        pre.integrate(j[2], j[3:6], j[6:9], True, bool(int(j[9]) & 2))
        got = np.concatenate([[pre.t], pre.q, pre.p, pre.v])
        assert np.allclose(got, d['result'][k][:11], rtol=0, atol=1e-14)
        # preintegrated rotation is a unit quaternion; zero-motion consistency of predict()
        assert abs(np.linalg.norm(pre.q) - 1) < 1e-12


def test_solver_golden_and_properties(golden_dir):
    g = np.load(os.path.join(golden_dir, 'backend_problems.npz'))
    for kind in ('localize', 'subwindow_t', 'window'):
        pb = problem_io.problem_from_dict(g, kind + '/pre/')
        summ = ceres_tr.solve(pb)
        ref = g[kind + '/post/cost']
        assert abs(summ.final_cost - ref[1]) <= 1e-9 * ref[1]
        assert summ.final_cost <= summ.initial_cost
        assert len(summ.iterations) == int(g[kind + '/post/iterations'])
        assert np.allclose(pb.states, g[kind + '/post/states'], atol=1e-10)
        # accepted iterations never increase the cost (monotonic trust region)
        costs = [it['cost'] for it in summ.iterations if it['successful']]
        assert all(b <= a + 1e-12 for a, b in zip(costs, costs[1:]))


def test_stress_problem_oracle_golden_and_schur_identity(golden_dir):
    """BASELINE config 5 (30 keyframes x 2000 landmarks x 29 targets, 29 IMU edges, dense 435 x 435 prior): the generator is
    deterministic and the oracle reproduces the committed cost trace; at reduced size, the landmark Schur complement the CUDA
    kernel forms (marginalization_factor.h:383-398) solves the same system as the oracle's full Cholesky."""
    from oracle import stress
    g = np.load(os.path.join(golden_dir, 'stress_config5.npz'))
    pb = stress.make_stress_problem(int(g['seed']), int(g['n_kf']), int(g['n_lm']))
    assert len(pb.rpe_tgt) == 58000 and len(pb.pre) == 29 and pb.marg.sqrt_inv_cov.shape == (435, 435) and pb.max_iterations == 10
    summ = ceres_tr.solve(pb)
    trace = np.array([it['cost'] for it in summ.iterations])
    assert np.allclose(trace, g['trace'], rtol=1e-9, atol=0)
    assert np.array_equal(np.array([it['successful'] for it in summ.iterations]), g['successful'])
    assert np.abs(pb.states - g['states']).max() < 1e-8
    # Schur identity on a small instance
    small = stress.make_stress_problem(3, 14, 60)
    lay = ceres_tr._Layout(small)
    _, r, J, gvec = ceres_tr.evaluate(small, lay, small.states, small.inv_depth, True)
    H = (J.T @ J).toarray() + 1e-6 * np.eye(lay.n_cols)
    n = lay.n_frame_cols
    Hpp, Hpl, hll = H[:n, :n], H[:n, n:], np.diag(H[n:, n:])
    assert np.abs(H[n:, n:] - np.diag(hll)).max() == 0.0          # inverse depths are independent scalar blocks
    S = Hpp - (Hpl / hll) @ Hpl.T
    b = gvec[:n] - Hpl @ (gvec[n:] / hll)
    yp = np.linalg.solve(S, b)
    yl = (gvec[n:] - Hpl.T @ yp) / hll
    full = np.linalg.solve(H, gvec)
    assert np.allclose(np.concatenate([yp, yl]), full, rtol=1e-6, atol=1e-9)


def test_marginalization_properties(golden_dir):
    g = np.load(os.path.join(golden_dir, 'backend_problems.npz'))
    d = {k[len('marg/pre/'):]: g[k] for k in g.files if k.startswith('marg/pre/')}
    m = problem_io.marg_inputs_from_dict(d)
    cfg = Config()
    newp, Hc, bc = marg_mod.marginalize(m['states'], m['prior'], m['pre'], m['obs'], m['inv_depth'], cfg.q_bc, cfg.p_bc, cfg.q_bi, cfg.p_bi,
                                        cfg.sqrt_inv_cov())
    assert np.allclose(newp.sqrt_inv_cov, g['marg/post/S'], atol=1e-6 * np.abs(g['marg/post/S']).max())
    # S^T S equals the Schur-complemented information on its range; it is symmetric positive semi-definite
    H = newp.sqrt_inv_cov.T @ newp.sqrt_inv_cov
    assert np.abs(H - Hc).max() < 1e-6 * np.abs(Hc).max()
    assert np.linalg.eigvalsh(Hc).min() > -1e-6
    assert newp.lin.shape == (m['states'].shape[0] - 1, 16)


def test_poisson_filter_scan_order_quirk():
    """test_point skips the first cell of the 5x5 block and visits (bx, ey + 1) instead
    (poisson_disk_filter.h:81-93)."""
    r = 20.0
    gs = r / np.sqrt(2)
    f = PoissonDiskFilter2(r)
    f.preset_point((2.5 * gs, 2.5 * gs))
    # candidate whose cell is 2 cells right and 2 down of the preset: preset sits in the skipped (bx,by) cell
    p = (4.01 * gs, 4.01 * gs)
    d2 = (p[0] - 2.5 * gs) ** 2 + (p[1] - 2.5 * gs) ** 2
    assert f.permit_point(p) == (d2 >= r * r or True)      # skipped cell -> never rejected by that preset
    assert f.permit_point((2.6 * gs, 2.6 * gs)) is False
    # a cell remembers only the last point stored in it
    f2 = PoissonDiskFilter2(r)
    f2.preset_point((0.1, 0.1)); f2.preset_point((13.9, 13.9))
    assert f2.permit_point((0.5, 0.5)) == ((0.5 - 13.9) ** 2 * 2 >= r * r)


def test_lotbox_matches_libstdcxx():
    """Sample sequence of std::default_random_engine seeded 0 through LotBox (values from g++ 13 libstdc++)."""
    lot = LotBox(150, 0)
    seq = []
    for _ in range(3):
        lot.refill_all(); seq += [lot.draw_without_replacement(), lot.draw_without_replacement()]
    assert seq == [0, 20, 113, 69, 79, 33]


def test_rotation_ransac_recovers_rotation():
    rng = np.random.default_rng(4)
    b1 = rng.normal(size=(100, 3)) + np.array([0, 0, 4.0]); b1 /= np.linalg.norm(b1, axis=1, keepdims=True)
    from oracle.lie import qmat
    R = qmat(expmap(np.array([0.01, -0.02, 0.015])))
    b2 = b1 @ R.T
    Rest, mask = find_rotation_matrix(b1, b2, np.deg2rad(10.0))
    assert mask.all() and np.abs(Rest - R).max() < 1e-9
    assert find_rotation_matrix(b1[:1], b2[:1], 0.1)[1].sum() == 0      # fewer than 2 points -> empty mask


def test_lie_roundtrips():
    rng = np.random.default_rng(5)
    for _ in range(20):
        w = rng.normal(size=3)
        w = w / np.linalg.norm(w) * rng.uniform(1e-9, 3.0)      # logmap returns the angle in [0, pi]
        assert np.allclose(logmap(expmap(w)), w, atol=1e-12)
        J = right_jacobian(w)
        e = 1e-7
        num = np.stack([logmap(qmul(qconj(expmap(w)), expmap(w + e * np.eye(3)[k]))) / e for k in range(3)], axis=1)
        assert np.allclose(num, J, atol=1e-5)
    assert np.allclose(right_jacobian(np.zeros(3)), np.eye(3))


# -- C ABI surface ----------------------------------------------------------------------------------
def test_c_abi_exports_every_declared_symbol():
    from xrslam_b200 import capi
    hdr = open(os.path.join(os.path.dirname(capi._HERE), 'include', 'xrvio.h')).read()
    names = set(re.findall(r'\b(xrvio_[a-z_0-9]+)\s*\(', hdr))
    hdr2 = open(os.path.join(os.path.dirname(capi._HERE), 'include', 'xrslam_compat.h')).read()
    names |= set(re.findall(r'\b(XRSLAM[A-Za-z]+)\s*\(', hdr2))
    assert len(names) >= 31 and 'XRSLAMCreate' in names
    dll = ctypes.CDLL(capi.LIB_PATH)
    for n in sorted(names):
        assert hasattr(dll, n), 'libxrvio.so does not export ' + n


def test_struct_layouts_match_header():
    """numpy dtypes used by the binding mirror the C structs (sizes checked against a compiled probe)."""
    import subprocess, tempfile
    from xrslam_b200 import capi
    inc = os.path.join(os.path.dirname(capi._HERE), 'include')
    src = '#include <stdio.h>\n#include "xrvio.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",sizeof(xrvio_rig),sizeof(xrvio_preint_job),' \
          'sizeof(xrvio_rpe),sizeof(xrvio_rot),sizeof(xrvio_pre_ref),sizeof(xrvio_solve_desc),sizeof(xrvio_solve_out),sizeof(xrvio_marg_desc),' \
          'sizeof(xrvio_params),sizeof(xrvio_pose),sizeof(xrvio_stats),sizeof(xrvio_init_info));}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, 'p.c'); open(c, 'w').write(src)
        subprocess.check_call(['gcc', '-I', inc, c, '-o', os.path.join(td, 'p')])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(td, 'p')]).split()]
    want = [capi.rig_dtype, capi.preint_job_dtype, capi.rpe_dtype, capi.rot_dtype, capi.pre_ref_dtype, capi.solve_desc_dtype,
            capi.solve_out_dtype, capi.marg_desc_dtype, capi.params_dtype, capi.pose_dtype, capi.stats_dtype, capi.init_info_dtype]
    assert sizes == [d.itemsize for d in want]


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product refuses to run instead of falling back to a CPU path."""
    from xrslam_b200 import capi
    if capi.lib.xrvio_device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(capi.XrvioError):
        capi.op_clahe(np.zeros((1, 480, 640), np.uint8))
    with pytest.raises(capi.XrvioError):
        capi.Context(capi.default_params(), 2)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, 'xrslam_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.hpp', '.inl', '.h')):
                txt = open(os.path.join(dp, f), errors='ignore').read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'oracle/' not in txt, f


def test_window_ba_runs_to_the_iteration_limit():
    """D10 (oracle/ceres_tr.py header): after the first accepted step of a window BA the StateUpdatingCallback rewrites the
    linearisation bias in user memory, every later candidate is rejected and the solve ends at max_num_iterations with the
    iterate of the first accepted step.  Asserted on every window BA of three different streams."""
    from oracle.synth import Stream
    from oracle.pipeline import run_stream
    seen = 0
    for seed in (3, 7, 11):
        rec = []
        run_stream(Stream(seed), 48, recorder=lambda kind, when, pb, summ=None: rec.append((kind, summ)) if when == 'post' else None)
        wins = [s for k, s in rec if k == 'window']
        assert len(wins) >= 2, (seed, [k for k, _ in rec])
        for s in wins:
            # (over longer runs a few solves end at the same 30th iteration as "parameter tolerance": the radius has been
            # halved 29 times by then -- tests/test_gpu_pipeline.py counts both kinds on 60 frames)
            assert s.termination in ('max iterations', 'parameter tolerance') and len(s.iterations) == 31, (seed, s.termination, len(s.iterations))
            ok = [it['iteration'] for it in s.iterations if it['successful']]
            assert ok[0] == 0 and len(ok) >= 2 and max(ok) <= 3, (seed, ok)       # accepted steps only at the very start
            tail = s.iterations[max(ok) + 1:]
            assert all((not it['successful']) and it['cost'] > s.final_cost for it in tail)
            assert tail[-1]['radius'] < tail[0]['radius'] * 2.0 ** -(len(tail) - 2)   # halved at every rejection
        seen += len(wins)
    assert seen >= 6


def test_failed_candidate_evaluation_is_a_rejected_step(monkeypatch):
    """Ceres 1.14 (ComputeCandidatePointAndEvaluateCost) leaves candidate_cost = numeric_limits<double>::max() when the
    evaluation of a candidate fails, so the step is rejected and the radius halved; it must not end the solve as "function
    tolerance reached".  Natural inputs do not produce a non-finite candidate (it takes a landmark exactly at depth 0), so the
    evaluation is stubbed here: the second cost-only evaluation of a recorded sub-window problem returns inf.  The kernel holds the
    same statement (backend_solve.cu: `if (!isfinite(cand_cost)) cand_cost = DBL_MAX`)."""
    from oracle import ceres_tr, problem_io
    g = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'backend_problems.npz'))
    pb = problem_io.problem_from_dict(g, 'localize/pre/')
    clean = ceres_tr.solve(problem_io.problem_from_dict(g, 'localize/pre/'))
    real = ceres_tr.evaluate
    calls = {'n': 0}

    def stub(pb_, lay, states, inv_depth, want_jac):
        out = real(pb_, lay, states, inv_depth, want_jac)
        if not want_jac:
            calls['n'] += 1
            if calls['n'] == 2:
                return (float('inf'),) + tuple(out[1:])
        return out
    monkeypatch.setattr(ceres_tr, 'evaluate', stub)
    summ = ceres_tr.solve(pb)
    its = summ.iterations
    assert calls['n'] >= 3                                        # the solve went on after the failed evaluation
    bad = its[2]
    assert bad['successful'] is False and bad['cost'] == np.finfo(np.float64).max
    assert bad['radius'] == 0.5 * its[1]['radius']                 # rejected: the radius (recorded after the update) is halved ...
    assert len(its) > 3 and summ.termination == clean.termination  # ... and the loop goes on to the same end
    assert abs(summ.final_cost - clean.final_cost) <= 1e-9 * clean.final_cost
