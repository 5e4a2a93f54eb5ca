"""The device residual / Jacobian code of the solve and marginalisation kernels (xrslam_b200/csrc/factors.cuh: K7
reprojection, K8 preintegration, K8b rotation prior, marginalisation error) compiled for the host with g++ and
compared with the oracle's factors at random states.  CPU only; the same functions run inside solve_kernel."""
import os
import subprocess
import types
import numpy as np
import pytest

from oracle import factors as F
from oracle.lie import expmap, qmul, qnormalized, logmap, right_jacobian, qconj

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp('host') / 'factor_probe'
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-I', '/usr/local/cuda/include', os.path.join(ROOT, 'tests', 'host', 'factor_probe.cpp'), '-o', str(out)])

    def run(text):
        r = subprocess.run([str(out)], input=text, capture_output=True, text=True, check=True)
        return [np.array([float(x) for x in ln.split()]) for ln in r.stdout.strip().split('\n')]
    return run


def _f(*a):
    return ' '.join('%.17g' % v for x in a for v in np.atleast_1d(np.asarray(x, np.float64)).ravel())


def _rq(rng, s=1.0):
    return qnormalized(expmap(rng.normal(size=3) * s))


CAM_Q = np.array([-7.7071797555374275e-03, 1.0499323370587278e-02, 7.0175280029197162e-01, 7.1230146066895372e-01])
CAM_P = np.array([-0.0216401454975, -0.064676986768, 0.00981073058949])
SIC = np.array([458.654, 457.296]) / np.sqrt(0.5)


def _rig(imu_q, imu_p):
    return _f(CAM_Q, CAM_P, imu_q, imu_p, SIC)


def test_reprojection_and_rotation_factor_code(probe):
    rng = np.random.default_rng(0)
    imu_q, imu_p = np.array([0, 0, 0, 1.0]), np.zeros(3)
    lines, refs = [], []
    for _ in range(40):
        qt, qr = _rq(rng, 0.3), _rq(rng, 0.3)
        pt, prf = rng.normal(size=3) * 0.3, rng.normal(size=3) * 0.3
        zt = rng.normal(size=3) * 0.2 + np.array([0, 0, 1.0]); zt /= np.linalg.norm(zt)
        zr = rng.normal(size=3) * 0.2 + np.array([0, 0, 1.0]); zr /= np.linalg.norm(zr)
        invd = rng.uniform(0.1, 0.5)
        lines.append('rpe %s %s' % (_rig(imu_q, imu_p), _f(qt, pt, qr, prf, invd, zt, zr)))
        refs.append(('rpe', F.reprojection(qt, pt, qr, prf, invd, zt, zr, CAM_Q, CAM_P, np.diag(SIC))))
        lines.append('rot %s %s' % (_rig(imu_q, imu_p), _f(qt, qr, zt, zr)))
        refs.append(('rot', F.rotation_prior(qt, qr, zt, zr, CAM_Q, CAM_P, np.diag(SIC))))
    outs = probe('\n'.join(lines) + '\n')
    for o, (kind, (r, J)) in zip(outs, refs):
        scale = max(1.0, np.abs(r).max())
        assert np.abs(o[:2] - r).max() < 1e-9 * scale
        if kind == 'rpe':
            Jt = o[2:14].reshape(2, 6); Jr = o[14:26].reshape(2, 6); Jl = o[26:28]
            for got, ref in ((Jt[:, :3], J['q_tgt']), (Jt[:, 3:], J['p_tgt']), (Jr[:, :3], J['q_ref']), (Jr[:, 3:], J['p_ref']), (Jl, J['inv_depth'])):
                assert np.abs(got - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())
        else:
            Jq = J['q_tgt'] if isinstance(J, dict) else J
            assert np.abs(o[2:8].reshape(2, 3) - Jq).max() < 1e-9 * max(1.0, np.abs(Jq).max())


def test_preintegration_factor_code(probe):
    rng = np.random.default_rng(1)
    lines, refs = [], []
    for k in range(25):
        imu_q = _rq(rng, 0.05) if k % 2 else np.array([0, 0, 0, 1.0])
        imu_p = rng.normal(size=3) * 0.02 if k % 2 else np.zeros(3)
        si = (_rq(rng, 0.5), rng.normal(size=3), rng.normal(size=3), rng.normal(size=3) * 0.01, rng.normal(size=3) * 0.05)
        sj = (qnormalized(qmul(si[0], expmap(rng.normal(size=3) * 0.1))), si[1] + rng.normal(size=3) * 0.1, si[2] + rng.normal(size=3) * 0.1,
              si[3] + rng.normal(size=3) * 1e-3, si[4] + rng.normal(size=3) * 1e-3)
        bias0 = (si[3] + rng.normal(size=3) * 1e-3, si[4] + rng.normal(size=3) * 1e-2)
        pre = types.SimpleNamespace(t=0.05 * (1 + k % 5), q=_rq(rng, 0.1), p=rng.normal(size=3) * 0.05, v=rng.normal(size=3) * 0.1,
                                    sqrt_inv_cov=np.eye(15), dq_dbg=rng.normal(size=(3, 3)) * 0.05, dp_dbg=rng.normal(size=(3, 3)) * 0.01,
                                    dp_dba=rng.normal(size=(3, 3)) * 0.01, dv_dbg=rng.normal(size=(3, 3)) * 0.05, dv_dba=rng.normal(size=(3, 3)) * 0.05)
        lines.append('pre %s %s' % (_rig(imu_q, imu_p), _f(*si, *sj, *bias0, pre.t, pre.q, pre.p, pre.v, pre.dq_dbg, pre.dp_dbg, pre.dp_dba, pre.dv_dbg, pre.dv_dba)))
        refs.append(F.preintegration(si, sj, pre, bias0[0], bias0[1], imu_q, imu_p))       # sqrt_inv_cov = I: unwhitened, like pre_eval
    outs = probe('\n'.join(lines) + '\n')
    for o, (r, Ji, Jj) in zip(outs, refs):
        assert np.abs(o[:15] - r).max() < 1e-11
        assert np.abs(o[15:240].reshape(15, 15) - Ji).max() < 1e-9
        assert np.abs(o[240:465].reshape(15, 15) - Jj).max() < 1e-9


def test_marginalization_error_code(probe):
    rng = np.random.default_rng(2)
    imu_q, imu_p = np.array([0, 0, 0, 1.0]), np.zeros(3)
    lines, refs = [], []
    for _ in range(20):
        lin = np.concatenate([_rq(rng, 0.5), rng.normal(size=12)])
        s = np.concatenate([qnormalized(qmul(lin[:4], expmap(rng.normal(size=3) * 0.05))), lin[4:] + rng.normal(size=12) * 0.01])
        lines.append('marg %s %s' % (_rig(imu_q, imu_p), _f(s, lin)))
        rq = logmap(qmul(qconj(lin[:4]), s[:4]))
        refs.append((np.concatenate([rq, s[4:] - lin[4:]]), np.linalg.inv(right_jacobian(rq))))
    outs = probe('\n'.join(lines) + '\n')
    for o, (dx, Jri) in zip(outs, refs):
        assert np.abs(o[:15] - dx).max() < 1e-12 and np.abs(o[15:24].reshape(3, 3) - Jri).max() < 1e-10
