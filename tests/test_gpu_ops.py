"""Parity of every CUDA operator (through the C ABI) against the CPU oracle.  Needs a GPU."""
import os
import numpy as np
import pytest
import cv2

pytestmark = pytest.mark.gpu

from oracle.synth import Stream
from oracle.frontend import OracleImage, LK_CRITERIA
from oracle.poisson import PoissonDiskFilter2
from oracle import problem_io, ceres_tr, bridge
from oracle import marginalization as marg_mod
from oracle.pipeline import Config


def _fixture_images(golden_dir):
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1]], np.float32)
    D = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05], np.float32)
    return [cv2.undistort(cv2.imread(os.path.join(golden_dir, n), 0), K, D) for n in ('euroc_v101_a.png', 'euroc_v101_b.png')]


@pytest.fixture(scope='module')
def synth_frames():
    s = Stream(5)
    return np.stack([s.frame(k) for k in range(3)])


# ---------------------------------------------------------------- K1
def test_clahe_bit_exact(xrvio, golden_dir, synth_frames):
    cl = cv2.createCLAHE(6.0, (8, 8))
    for imgs in (synth_frames, np.stack(_fixture_images(golden_dir))):     # 640x480 and 752x480
        out = xrvio.op_clahe(imgs)
        ref = np.stack([cl.apply(im) for im in imgs])
        assert np.array_equal(out, ref), 'CLAHE differs in %d pixels (max %d)' % (
            (out != ref).sum(), np.abs(out.astype(int) - ref.astype(int)).max())


def test_clahe_edge_cases(xrvio):
    cl = cv2.createCLAHE(6.0, (8, 8))
    flat = np.full((1, 480, 640), 77, np.uint8)
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (1, 480, 640)).astype(np.uint8)
    odd = rng.integers(0, 256, (1, 123, 250)).astype(np.uint8)          # not a multiple of the tile grid
    for imgs in (flat, noise, odd):
        out = xrvio.op_clahe(imgs)
        assert np.array_equal(out[0], cl.apply(imgs[0]))


# ---------------------------------------------------------------- K2
def test_pyramid_and_scharr_bit_exact(xrvio, synth_frames, golden_dir):
    for imgs in (synth_frames[:2], np.stack(_fixture_images(golden_dir))[:1]):
        levels, derivs = xrvio.op_pyramid(imgs, with_deriv=True)
        for i, im in enumerate(imgs):
            _, pyr = cv2.buildOpticalFlowPyramid(im, (21, 21), 3, None, True)
            for l in range(4):
                assert np.array_equal(levels[i][l], pyr[2 * l]), 'level %d' % l
                assert np.array_equal(derivs[i][l], pyr[2 * l + 1]), 'deriv %d' % l


# ---------------------------------------------------------------- K3
def test_gftt_matches_opencv(xrvio, synth_frames, golden_dir):
    cl = cv2.createCLAHE(6.0, (8, 8))
    g = cv2.GFTTDetector_create(200, 1e-3, 20, 3, True)
    for imgs in (np.stack([cl.apply(f) for f in synth_frames]), np.stack([cl.apply(f) for f in _fixture_images(golden_dir)])):
        res = xrvio.op_gftt(imgs, 200, 1e-3, 20.0)
        for i, im in enumerate(imgs):
            ref = g.detect(im)
            ref_xy = np.array([k.pt for k in ref])
            ref_resp = np.array([k.response for k in ref])
            got = res[i]
            assert abs(len(got) - len(ref_xy)) <= 2
            # float32 Harris responses can differ in the last bits between SIMD paths: compare as sets
            sref = {(int(x), int(y)) for x, y in ref_xy}
            sgot = {(int(x), int(y)) for x, y, _ in got}
            overlap = len(sref & sgot) / max(len(sref), 1)
            assert overlap >= 0.97, 'corner overlap %.3f' % overlap
            # responses of common corners agree to float precision
            dref = {(int(x), int(y)): r for (x, y), r in zip(ref_xy, ref_resp)}
            rel = [abs(r - dref[(int(x), int(y))]) / dref[(int(x), int(y))] for x, y, r in got if (int(x), int(y)) in dref]
            assert np.max(rel) < 1e-4
            # strongest-first ordering
            assert np.all(np.diff(got[:, 2]) <= 0)


# ---------------------------------------------------------------- K4
def _klt_oracle(prev, nxt, pts, init):
    a, b = OracleImage(0, prev), OracleImage(0.05, nxt)
    return a.track_keypoints(b, pts.astype(np.float64), init.astype(np.float64))


def test_klt_matches_opencv(xrvio, synth_frames, golden_dir):
    cl = cv2.createCLAHE(6.0, (8, 8))
    g = cv2.GFTTDetector_create(200, 1e-3, 20, 3, True)
    cases = []
    fr = [cl.apply(f) for f in synth_frames]
    cases.append((fr[0], fr[1]))
    cases.append((fr[1], fr[2]))
    fx = [cl.apply(f) for f in _fixture_images(golden_dir)]
    cases.append((fx[0], fx[1]))
    for prev, nxt in cases:
        pts = np.array([k.pt for k in g.detect(prev)], np.float32)
        pts = pts[(pts[:, 0] >= 20) & (pts[:, 1] >= 20) & (pts[:, 0] < prev.shape[1] - 20) & (pts[:, 1] < prev.shape[0] - 20)]
        rng = np.random.default_rng(1)
        init = pts + rng.normal(0, 1.5, pts.shape).astype(np.float32)
        ref_next, ref_st = _klt_oracle(prev, nxt, pts, init)
        got_next, got_st = xrvio.op_klt(prev[None], nxt[None], [pts], [init])
        got_next, got_st = got_next[0], got_st[0]
        agree = (got_st != 0) == (ref_st != 0)
        assert agree.mean() >= 0.97, 'status agreement %.3f' % agree.mean()
        both = (got_st != 0) & (ref_st != 0)
        assert both.sum() >= 0.8 * len(pts)
        err = np.linalg.norm(got_next[both].astype(np.float64) - ref_next[both], axis=1)
        assert np.percentile(err, 99) < 0.05 and np.median(err) < 5e-3, (np.median(err), np.percentile(err, 99), err.max())


def test_klt_edge_cases(xrvio, synth_frames):
    cl = cv2.createCLAHE(6.0, (8, 8))
    prev, nxt = cl.apply(synth_frames[0]), cl.apply(synth_frames[1])
    # no points at all, and points at the border / outside the gates
    got_next, got_st = xrvio.op_klt(prev[None], nxt[None], [np.zeros((0, 2), np.float32)], [np.zeros((0, 2), np.float32)])
    assert len(got_st[0]) == 0
    pts = np.array([[2.0, 3.0], [638.0, 478.0], [320.0, 5.0], [25.0, 25.0]], np.float32)
    ref_next, ref_st = _klt_oracle(prev, nxt, pts, pts.copy())
    got_next, got_st = xrvio.op_klt(prev[None], nxt[None], [pts], [pts.copy()])
    assert np.array_equal(got_st[0] != 0, ref_st != 0)


# ---------------------------------------------------------------- K5
def test_poisson_exact(xrvio):
    rng = np.random.default_rng(3)
    lists, npre, ref = [], [], []
    for case in range(6):
        n_pre = [0, 40, 150, 0, 5, 200][case]
        n_new = [200, 200, 200, 1, 0, 300][case]
        pts = np.column_stack([rng.uniform(20, 732, n_pre + n_new), rng.uniform(20, 460, n_pre + n_new)])
        if case == 2:
            pts[n_pre:] = pts[:n_new] + rng.normal(0, 8, (n_new, 2))        # many near-collisions
        f = PoissonDiskFilter2(20.0)
        f.preset_points([tuple(p) for p in pts[:n_pre]])
        keep = [1] * n_pre + [1 if f.insert_point(tuple(p)) else 0 for p in pts[n_pre:]]
        lists.append(pts); npre.append(n_pre); ref.append(np.array(keep, np.uint8))
    got = xrvio.op_poisson(lists, npre, 20.0)
    for g, r in zip(got, ref):
        assert np.array_equal(g, r)


# ---------------------------------------------------------------- K6
def test_preintegration(xrvio, golden_dir):
    d = np.load(os.path.join(golden_dir, 'preint.npz'))
    cfg = Config()
    rig = bridge.rig_from_cfg(cfg)
    jobs = np.zeros(len(d['jobs']), xrvio.preint_job_dtype)
    for k, j in enumerate(d['jobs']):
        jobs[k]['imu_off'], jobs[k]['n'], jobs[k]['t_end'] = int(j[0]), int(j[1]), j[2]
        jobs[k]['bg'], jobs[k]['ba'], jobs[k]['flags'] = j[3:6], j[6:9], int(j[9])
    out = xrvio.op_preintegrate(rig, jobs, d['imu'])
    for k in range(len(jobs)):
        ref, got = d['result'][k], out[k][:506]
        sl = dict(state=slice(0, 11), cov=slice(11, 236), sic=slice(236, 461), jac=slice(461, 506))
        for name, s in sl.items():
            if name in ('cov', 'sic') and not (jobs[k]['flags'] & 2):
                continue
            scale = np.abs(ref[s]).max() + 1e-300
            tol = 1e-9 if name != 'sic' else 1e-7        # sqrt_inv_cov goes through a 15x15 inverse (cond ~1e8)
            assert np.abs(got[s] - ref[s]).max() / scale < tol, (k, name, np.abs(got[s] - ref[s]).max() / scale)


# ---------------------------------------------------------------- K7-K11
@pytest.fixture(scope='module')
def golden_problems(golden_dir):
    return np.load(os.path.join(golden_dir, 'backend_problems.npz'))


@pytest.mark.parametrize('kind', ['localize', 'subwindow_t', 'init_ba', 'window'])
def test_solver_matches_oracle(xrvio, golden_problems, kind):
    pb = problem_io.problem_from_dict(golden_problems, kind + '/pre/')
    out = bridge.solve_problem_cuda(pb)
    ref_cost = golden_problems[kind + '/post/cost']
    ref_trace = golden_problems[kind + '/post/trace']
    assert int(out['iterations']) == int(golden_problems[kind + '/post/iterations'])
    # BA cost within 1e-6 relative after the same iteration count (north_star)
    assert abs(out['initial_cost'] - ref_cost[0]) <= 1e-9 * ref_cost[0]
    assert abs(out['final_cost'] - ref_cost[1]) <= 1e-6 * ref_cost[1]
    n = len(ref_trace)
    assert np.allclose(out['cost_trace'][:n], ref_trace, rtol=1e-6, atol=0)
    ref_states = golden_problems[kind + '/post/states']
    assert np.abs(pb.states[:, 0:4] - ref_states[:, 0:4]).max() < 1e-7
    assert np.abs(pb.states[:, 4:7] - ref_states[:, 4:7]).max() < 1e-6
    assert np.abs(pb.states[:, 7:16] - ref_states[:, 7:16]).max() < 1e-5
    ref_inv = golden_problems[kind + '/post/inv_depth']
    assert np.abs(pb.inv_depth - ref_inv).max() < 1e-6


def test_solver_rotation_prior_residuals(xrvio, golden_problems):
    """K8b: rotation-prior residuals (rotation_factor.h:23-59) in the solve kernel.  The tracker only creates them for valid
    but un-triangulated tracks of a pure-rotation sub-window (sliding_window_tracker.cpp:455-461), which the synthetic streams
    never produce, so the problem is made here: a third of a sub-window problem's reprojection priors become rotation priors."""
    import copy
    from oracle import ceres_tr
    pb = problem_io.problem_from_dict(golden_problems, 'subwindow_t/pre/')
    idx = np.arange(len(pb.rpe_tgt))
    to_rot = (idx % 3 == 0) & pb.rpe_prior
    assert to_rot.sum() >= 10
    pb.rot_tgt, pb.rot_ref = pb.rpe_tgt[to_rot].copy(), pb.rpe_ref[to_rot].copy()
    pb.rot_z_tgt, pb.rot_z_ref = pb.rpe_z_tgt[to_rot].copy(), pb.rpe_z_ref[to_rot].copy()
    keep = ~to_rot
    for name in ('rpe_tgt', 'rpe_ref', 'rpe_lm', 'rpe_z_tgt', 'rpe_z_ref', 'rpe_prior'):
        setattr(pb, name, getattr(pb, name)[keep].copy())
    ref = copy.deepcopy(pb)
    summ = ceres_tr.solve(ref)
    out = bridge.solve_problem_cuda(pb)
    assert int(out['iterations']) == len(summ.iterations)
    assert abs(out['initial_cost'] - summ.initial_cost) <= 1e-9 * summ.initial_cost
    assert abs(out['final_cost'] - summ.final_cost) <= 1e-6 * summ.final_cost
    assert np.abs(pb.states[:, 0:4] - ref.states[:, 0:4]).max() < 1e-7 and np.abs(pb.states[:, 4:7] - ref.states[:, 4:7]).max() < 1e-6


def test_solver_batch_of_mixed_problems(xrvio, golden_problems):
    pbs = [problem_io.problem_from_dict(golden_problems, k + '/pre/') for k in ('localize', 'window', 'subwindow_t', 'localize', 'init_ba')]
    pk = bridge.PackedProblems(pbs)
    out = pk.solve(bridge.rig_from_problem(pbs[0]))
    for i, k in enumerate(('localize', 'window', 'subwindow_t', 'localize', 'init_ba')):
        ref = golden_problems[k + '/post/cost']
        assert abs(out[i]['final_cost'] - ref[1]) <= 1e-6 * ref[1]


def test_solver_empty_problem(xrvio, golden_problems):
    """A problem whose only frame is fixed has no free parameters (termination 6)."""
    pb = problem_io.problem_from_dict(golden_problems, 'localize/pre/')
    pb.fix_pose[:] = True
    pb.fix_motion[:] = True
    out = bridge.solve_problem_cuda(pb)
    assert int(out['termination']) == 6


# ---------------------------------------------------------------- K12
def test_marginalization_matches_oracle(xrvio, golden_problems):
    d = {k[len('marg/pre/'):]: golden_problems[k] for k in golden_problems.files if k.startswith('marg/pre/')}
    m = problem_io.marg_inputs_from_dict(d)
    cfg = Config()
    args = (m['states'], m['prior'], m['pre'], m['obs'], m['inv_depth'], cfg.q_bc, cfg.p_bc, cfg.q_bi, cfg.p_bi, cfg.sqrt_inv_cov())
    ref, Hc, bc = marg_mod.marginalize(*args)
    got, A0, rank = bridge.marginalize_cuda(*args)
    lam = np.linalg.eigvalsh(Hc)
    assert rank == int((lam > 1e-8).sum()) or abs(rank - int((lam > 1e-8).sum())) <= 3     # noise-level eigenvalues may differ
    scale = np.abs(Hc).max()
    # S^T S reproduces the Schur-complemented information matrix (range part), as does the reference's sqrt(L) V^T
    Href = ref.sqrt_inv_cov.T @ ref.sqrt_inv_cov
    Hgot = got.sqrt_inv_cov.T @ got.sqrt_inv_cov
    assert np.abs(Hgot - Href).max() / scale < 1e-9
    assert np.abs(A0 - Hgot).max() / scale < 1e-12
    # S^T infovec = b restricted to the range, and the constant term of the prior cost agree
    bref = ref.sqrt_inv_cov.T @ ref.infovec
    bgot = got.sqrt_inv_cov.T @ got.infovec
    assert np.abs(bgot - bref).max() / np.abs(bref).max() < 1e-7
    assert abs(got.infovec @ got.infovec - ref.infovec @ ref.infovec) / (ref.infovec @ ref.infovec) < 1e-6
    assert np.array_equal(got.lin, ref.lin)
    # rows of S are mutually orthogonal (eigen form sqrt(lambda) v^T)
    G = got.sqrt_inv_cov[:rank]
    GG = G @ G.T
    off = GG - np.diag(np.diag(GG))
    assert np.abs(off).max() / np.abs(np.diag(GG)).max() < 1e-9
    assert np.allclose(np.sort(np.diag(GG)), np.sort(lam[lam > 1e-8])[-rank:], rtol=1e-6) or rank != int((lam > 1e-8).sum())
    # the plain rank-revealing Cholesky form (session default) defines the same prior: same S^T S, S^T infovec, |infovec|^2
    got2, A02, rank2 = bridge.marginalize_cuda(*args, eigen_form=False)
    H2 = got2.sqrt_inv_cov.T @ got2.sqrt_inv_cov
    assert np.abs(H2 - Href).max() / scale < 1e-9 and np.abs(A02 - H2).max() / scale < 1e-12
    b2 = got2.sqrt_inv_cov.T @ got2.infovec
    assert np.abs(b2 - bref).max() / np.abs(bref).max() < 1e-7
    assert abs(got2.infovec @ got2.infovec - ref.infovec @ ref.infovec) / (ref.infovec @ ref.infovec) < 1e-6


# ---------------------------------------------------------------- T
def test_triangulation(xrvio):
    from oracle.lie import expmap, qmul, qconj, qrot, qmat
    rng = np.random.default_rng(2)
    jobs = np.zeros(20, xrvio.tri_job_dtype)
    views = []
    ref = []
    for j in range(20):
        X = np.array([rng.uniform(-2, 2), rng.uniform(-2, 2), rng.uniform(3, 8)])
        k = int(rng.integers(2, 9))
        jobs[j]['obs_off'], jobs[j]['n_obs'] = len(views), k
        rows, Ps = [], []
        for v in range(k):
            q = expmap(rng.normal(size=3) * 0.1)
            p = rng.normal(size=3) * 0.5
            if j == 19 and v == 0:
                p = p + np.array([0, 0, 20.0])          # camera beyond the point: cheirality must fail
            z = qrot(qconj(q), X - p)
            z = z / np.linalg.norm(z) + rng.normal(size=3) * 1e-3
            z /= np.linalg.norm(z)
            views.append(np.concatenate([q, p, z, np.zeros(6)]))
            R = qmat(qconj(q))
            P = np.concatenate([R, (-(R @ p))[:, None]], axis=1)
            rows += [z[0] * P[2] - z[2] * P[0], z[1] * P[2] - z[2] * P[1]]
            Ps.append(P)
        h = np.linalg.svd(np.array(rows))[2][3]
        valid = all((P @ h)[2] * h[3] > 0 for P in Ps)
        ref.append(np.concatenate([h[:3] / h[3], [1.0 if valid else 0.0]]))
    out = xrvio.op_triangulate(jobs, np.array(views))
    ref = np.array(ref)
    assert np.array_equal(out[:, 3], ref[:, 3])
    ok = ref[:, 3] > 0
    assert np.abs(out[ok, :3] - ref[ok, :3]).max() < 1e-7


@pytest.mark.parametrize('limit', [1, 2, 3, 4, 5, 7, 12])
def test_solver_iteration_limits_with_batched_candidates(xrvio, golden_problems, limit):
    """The trust-region loop evaluates the candidates of consecutive rejections in batches (evaluate_cost_multi) and consumes them
    one iteration at a time; a batch may not reach past max_num_iterations.  Every limit must give the oracle's iteration count,
    termination and cost trace -- in particular the limits that end in the middle of a batch (the window problem accepts two
    steps and then rejects: batches 1, 1, 1, 2, 4, 4, ...)."""
    import copy
    from oracle import ceres_tr
    pb = problem_io.problem_from_dict(golden_problems, 'window/pre/')
    pb.max_iterations = limit
    ref = copy.deepcopy(pb)
    summ = ceres_tr.solve(ref)
    out = bridge.solve_problem_cuda(pb)
    assert int(out['iterations']) == len(summ.iterations)
    trace = np.array([it['cost'] for it in summ.iterations])
    assert np.allclose(out['cost_trace'][:len(trace)], trace, rtol=1e-6, atol=0)
    assert abs(out['final_cost'] - summ.final_cost) <= 1e-6 * summ.final_cost
    flags = [bool(it.get('successful', False)) for it in summ.iterations]
    assert int(out['n_successful']) == sum(flags) and int(out['n_unsuccessful']) == len(flags) - sum(flags)
    assert np.abs(pb.states[:, 0:4] - ref.states[:, 0:4]).max() < 1e-7 and np.abs(pb.states[:, 4:7] - ref.states[:, 4:7]).max() < 1e-6


def _perturb_problem(pb, seed, amp):
    rng = np.random.default_rng(seed)
    na = pb.states.shape[0]
    pb.states[:, 4:7] += amp * rng.standard_normal((na, 3))
    dq = amp * 0.5 * rng.standard_normal((na, 3))
    q = pb.states[:, 0:4].copy()
    for i in range(na):
        d = np.array([dq[i, 0] / 2, dq[i, 1] / 2, dq[i, 2] / 2, 1.0])
        d /= np.linalg.norm(d)
        x1, y1, z1, w1 = q[i]
        x2, y2, z2, w2 = d
        pb.states[i, 0:4] = [w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2, w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                             w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2, w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2]
    pb.inv_depth = pb.inv_depth * np.exp(amp * rng.standard_normal(pb.inv_depth.shape))


@pytest.mark.parametrize('seed,amp', [(2, 0.05), (3, 0.05), (4, 0.05), (5, 0.5)])
def test_solver_accepts_inside_candidate_batches(xrvio, golden_problems, seed, amp):
    """Window problems started far from the optimum: the oracle's runs mix accepted and rejected steps (e.g. AAAAAArrrAAAArAAAArrr...,
    AAAArrrrrAArAArrAAArrr...), so candidates are accepted out of the middle of a speculative batch and the rest of the batch
    is discarded.  Iteration count, accept / reject totals and the whole cost trace must equal the oracle's."""
    import copy
    from oracle import ceres_tr
    pb = problem_io.problem_from_dict(golden_problems, 'window/pre/')
    _perturb_problem(pb, seed, amp)
    ref = copy.deepcopy(pb)
    summ = ceres_tr.solve(ref)
    flags = [bool(it.get('successful', False)) for it in summ.iterations]
    pattern = ''.join('A' if f else 'r' for f in flags)
    assert 'rrA' in pattern and sum(flags) >= 10, pattern           # the case is what the docstring says
    out = bridge.solve_problem_cuda(pb)
    assert int(out['iterations']) == len(summ.iterations)
    assert int(out['n_successful']) == sum(flags) and int(out['n_unsuccessful']) == len(flags) - sum(flags)
    trace = np.array([it['cost'] for it in summ.iterations])
    assert np.allclose(out['cost_trace'][:len(trace)], trace, rtol=1e-6, atol=0)
    assert abs(out['final_cost'] - summ.final_cost) <= 1e-6 * summ.final_cost


def test_preintegration_chunk_boundaries(xrvio):
    """preint_kernel precomputes the per-sample quantities 32 samples at a time (one per lane) before the sequential recursion:
    job lengths around the chunk boundaries, many chunks, and a batch of jobs of mixed lengths against the oracle's PreIntegrator."""
    from oracle.preintegrator import PreIntegrator
    rng = np.random.default_rng(23)
    cfg = Config()
    rig = bridge.rig_from_cfg(cfg)
    lengths = [2, 31, 32, 33, 63, 64, 65, 100, 200, 5]
    jobs = np.zeros(len(lengths), xrvio.preint_job_dtype)
    imu_all, ref = [], []
    for k, n in enumerate(lengths):
        t = 0.3 + k
        data = []
        for _ in range(n):
            data.append((t, rng.normal(size=3) * 0.4, np.array([0.1, -0.2, 9.7]) + rng.normal(size=3)))
            t += 0.005 * (1 + 0.1 * rng.random())
        t_end = t + 0.002
        bg, ba = rng.normal(size=3) * 0.01, rng.normal(size=3) * 0.05
        pre = PreIntegrator(cfg.cov_g, cfg.cov_a, cfg.cov_bg, cfg.cov_ba)
        pre.data = data
        pre.integrate(t_end, bg, ba, True, True)
        jobs[k]['imu_off'], jobs[k]['n'], jobs[k]['t_end'] = len(imu_all), n, t_end
        jobs[k]['bg'], jobs[k]['ba'], jobs[k]['flags'] = bg, ba, 3
        imu_all.extend([[d[0]] + list(d[1]) + list(d[2]) for d in data])
        ref.append(np.concatenate([[pre.t], pre.q, pre.p, pre.v, pre.cov.ravel(), pre.sqrt_inv_cov.ravel(), pre.dq_dbg.ravel(),
                                   pre.dp_dbg.ravel(), pre.dp_dba.ravel(), pre.dv_dbg.ravel(), pre.dv_dba.ravel()]))
    out = xrvio.op_preintegrate(rig, jobs, np.array(imu_all))
    sl = dict(state=slice(0, 11), cov=slice(11, 236), sic=slice(236, 461), jac=slice(461, 506))
    for k in range(len(lengths)):
        for name, s in sl.items():
            scale = np.abs(ref[k][s]).max() + 1e-300
            tol = 1e-9 if name != 'sic' else 1e-7
            err = np.abs(out[k][:506][s] - ref[k][s]).max() / scale
            assert err < tol, (lengths[k], name, err)


def test_solver_rejects_preintegration_block_out_of_range(xrvio, golden_problems):
    """xrvio_pre_ref::block must index pre_blocks (0 <= block < n_pre_blocks); the negative values the session layer uses internally
    for blocks that never left the device are not part of the operator API."""
    pb = problem_io.problem_from_dict(golden_problems, 'window/pre/')
    for bad in (-1, 10 ** 6):
        pk = bridge.PackedProblems([pb])
        assert len(pk.pre) > 0
        pk.pre[0]['block'] = bad
        with pytest.raises(Exception) as e:
            pk.solve(bridge.rig_from_problem(pb))
        assert 'block index out of range' in str(e.value)
