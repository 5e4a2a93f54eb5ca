"""Large-window solve (BASELINE.json config 5 / SURVEY.md section 8(d) row 5: 30 keyframes x 2000 landmarks, 58 000 reprojection
residuals, 29 preintegration edges, dense 435 x 435 prior, fixed 10 iterations) through xrvio_op_solve, which routes a batch
holding a problem with more than XRVIO_MAX_ACTIVE free frames to the large-window kernel (H in global memory, Schur complement
on the FP64 tensor pipe, panel Cholesky).  Parity bar as for the sliding-window solver: same iteration count, cost trace within
1e-6 relative of the oracle (north_star), states within 1e-6.
Written after the round-1 GPU budget was spent: validated on the SIMT emulator (mma.sync spelled with shuffles there), hence
last in the run order."""
import copy
import os
import numpy as np
import pytest

from oracle import bridge, ceres_tr, problem_io, stress

pytestmark = pytest.mark.gpu
EMU = os.environ.get('XRVIO_EMU') == '1'


def _check(out, pb, ref, summ):
    trace = np.array([it['cost'] for it in summ.iterations])
    assert int(out['iterations']) == len(trace)
    assert int(out['n_successful']) == summ.num_successful and int(out['n_unsuccessful']) == summ.num_unsuccessful
    assert abs(out['initial_cost'] - summ.initial_cost) <= 1e-9 * summ.initial_cost
    assert abs(out['final_cost'] - summ.final_cost) <= 1e-6 * summ.final_cost
    assert np.allclose(out['cost_trace'][:len(trace)], trace, rtol=1e-6, atol=0)
    assert np.abs(pb.states[:, 0:4] - ref.states[:, 0:4]).max() < 1e-7
    assert np.abs(pb.states[:, 4:7] - ref.states[:, 4:7]).max() < 1e-6
    assert np.abs(pb.states[:, 7:16] - ref.states[:, 7:16]).max() < 1e-5
    assert np.abs(pb.inv_depth - ref.inv_depth).max() < 1e-6


@pytest.mark.parametrize('n_kf,n_lm', [(13, 40), (20, 150), (32, 90)])
def test_large_window_matches_oracle(xrvio, n_kf, n_lm):
    pb = stress.make_stress_problem(5, n_kf, n_lm)
    ref = copy.deepcopy(pb)
    summ = ceres_tr.solve(ref)
    out = bridge.solve_problem_cuda(pb)
    _check(out, pb, ref, summ)


def test_large_window_with_fixed_gauge_and_rotation_priors(xrvio):
    """Constant blocks inside a large window: keyframe 0 fixed (pose and motion: its columns are unit-diagonal in the reduced
    system and masked out of the tensor-pipe Schur update), keyframe 1 with a fixed pose only, and some observations turned
    into rotation priors (rotation_factor.h:23-59)."""
    pb = stress.make_stress_problem(8, 16, 80)
    pb.fix_pose[0] = pb.fix_motion[0] = True
    pb.fix_pose[1] = True
    idx = np.arange(len(pb.rpe_tgt))
    to_rot = idx % 11 == 0
    pb.rot_tgt, pb.rot_ref = pb.rpe_tgt[to_rot].copy(), pb.rpe_ref[to_rot].copy()
    pb.rot_z_tgt, pb.rot_z_ref = pb.rpe_z_tgt[to_rot].copy(), pb.rpe_z_ref[to_rot].copy()
    for name in ('rpe_tgt', 'rpe_ref', 'rpe_lm', 'rpe_z_tgt', 'rpe_z_ref', 'rpe_prior'):
        setattr(pb, name, getattr(pb, name)[~to_rot].copy())
    ref = copy.deepcopy(pb)
    summ = ceres_tr.solve(ref)
    out = bridge.solve_problem_cuda(pb)
    _check(out, pb, ref, summ)
    assert np.array_equal(pb.states[0], ref.states[0]) and np.array_equal(pb.states[1, 0:7], ref.states[1, 0:7])


def test_large_window_limits(xrvio):
    pb = stress.make_stress_problem(5, 33, 20)
    with pytest.raises(xrvio.XrvioError):
        bridge.solve_problem_cuda(pb)


def test_large_window_batch_with_a_sliding_window_problem(xrvio, golden_dir):
    """One launch: two large windows of different shapes and a recorded 10-keyframe window problem; the batch goes to the
    large-window kernel as a whole and every problem keeps its own scratch."""
    gp = np.load(os.path.join(golden_dir, 'backend_problems.npz'))
    pbs = [stress.make_stress_problem(6, 18, 120), problem_io.problem_from_dict(gp, 'window/pre/'), stress.make_stress_problem(7, 25, 60, dense_prior=False)]
    refs = [copy.deepcopy(p) for p in pbs]
    summs = [ceres_tr.solve(r) for r in refs]
    pk = bridge.PackedProblems(pbs)
    out = pk.solve(bridge.rig_from_problem(pbs[0]))
    for i in range(3):
        st, ld = pk.result(i)
        pbs[i].states[:] = st
        pbs[i].inv_depth[:] = ld
        _check(out[i], pbs[i], refs[i], summs[i])
    ref_cost = gp['window/post/cost']
    assert abs(out[1]['final_cost'] - ref_cost[1]) <= 1e-6 * ref_cost[1]


@pytest.mark.skipif(EMU and os.environ.get('XRVIO_LONG') != '1', reason='full size takes minutes on the emulator')
def test_large_window_config5_full_size(xrvio, golden_dir):
    """BASELINE config 5 at full size against the live oracle and the committed oracle cost trace."""
    g = np.load(os.path.join(golden_dir, 'stress_config5.npz'))
    pb = stress.make_stress_problem(int(g['seed']), int(g['n_kf']), int(g['n_lm']))
    assert len(pb.rpe_tgt) == 58000 and len(pb.pre) == 29 and pb.marg.sqrt_inv_cov.shape == (435, 435)
    ref = copy.deepcopy(pb)
    summ = ceres_tr.solve(ref)
    out = bridge.solve_problem_cuda(pb)
    _check(out, pb, ref, summ)
    assert np.allclose(out['cost_trace'][:len(g['trace'])], g['trace'], rtol=1e-6, atol=0)
    assert np.abs(pb.states - g['states']).max() < 1e-5
