"""(De)serialisation of oracle Problems / marginalisation inputs to flat dicts of numpy arrays
(for the committed golden fixtures under tests/golden/).  Test infrastructure only."""
import numpy as np
from .problem import Problem, PreFactor, MargPrior

_PF = ('t', 'q', 'p', 'v', 'sqrt_inv_cov', 'dq_dbg', 'dp_dbg', 'dp_dba', 'dv_dbg', 'dv_dba')


def problem_to_dict(pb, prefix=''):
    d = {}
    for k in ('states', 'fix_pose', 'fix_motion', 'inv_depth', 'cam_q_cs', 'cam_p_cs', 'imu_q_cs', 'imu_p_cs', 'sqrt_inv_cov_px',
              'rpe_tgt', 'rpe_ref', 'rpe_lm', 'rpe_z_tgt', 'rpe_z_ref', 'rpe_prior', 'rot_tgt', 'rot_ref', 'rot_z_tgt', 'rot_z_ref'):
        d[prefix + k] = np.asarray(getattr(pb, k))
    d[prefix + 'max_iterations'] = np.array(pb.max_iterations)
    d[prefix + 'pre_ij'] = np.array([[f.i, f.j, int(f.prior)] for f in pb.pre], np.int64).reshape(-1, 3)
    for k in _PF:
        d[prefix + 'pre_' + k] = np.array([np.asarray(getattr(f, k), dtype=np.float64) for f in pb.pre])
    if pb.marg is not None:
        d[prefix + 'marg_frames'] = np.asarray(pb.marg.frames)
        d[prefix + 'marg_lin'] = pb.marg.lin
        d[prefix + 'marg_S'] = pb.marg.sqrt_inv_cov
        d[prefix + 'marg_infovec'] = pb.marg.infovec
    return d


def problem_from_dict(d, prefix=''):
    g = lambda k: d[prefix + k]
    pb = Problem(states=np.array(g('states'), dtype=np.float64), fix_pose=np.array(g('fix_pose'), bool),
                 fix_motion=np.array(g('fix_motion'), bool), inv_depth=np.array(g('inv_depth'), dtype=np.float64),
                 cam_q_cs=np.array(g('cam_q_cs')), cam_p_cs=np.array(g('cam_p_cs')), imu_q_cs=np.array(g('imu_q_cs')),
                 imu_p_cs=np.array(g('imu_p_cs')), sqrt_inv_cov_px=np.array(g('sqrt_inv_cov_px')),
                 max_iterations=int(g('max_iterations')))
    pb.rpe_tgt = np.array(g('rpe_tgt'), np.int32); pb.rpe_ref = np.array(g('rpe_ref'), np.int32); pb.rpe_lm = np.array(g('rpe_lm'), np.int32)
    pb.rpe_z_tgt = np.array(g('rpe_z_tgt'), dtype=np.float64).reshape(-1, 3); pb.rpe_z_ref = np.array(g('rpe_z_ref'), dtype=np.float64).reshape(-1, 3)
    pb.rpe_prior = np.array(g('rpe_prior'), bool)
    pb.rot_tgt = np.array(g('rot_tgt'), np.int32); pb.rot_ref = np.array(g('rot_ref'), np.int32)
    pb.rot_z_tgt = np.array(g('rot_z_tgt'), dtype=np.float64).reshape(-1, 3); pb.rot_z_ref = np.array(g('rot_z_ref'), dtype=np.float64).reshape(-1, 3)
    ij = g('pre_ij')
    pb.pre = []
    for k in range(len(ij)):
        vals = {n: (float(g('pre_' + n)[k]) if n == 't' else np.array(g('pre_' + n)[k], dtype=np.float64)) for n in _PF}
        pb.pre.append(PreFactor(int(ij[k, 0]), int(ij[k, 1]), bool(ij[k, 2]), **vals))
    if prefix + 'marg_frames' in d:
        pb.marg = MargPrior(frames=np.array(g('marg_frames')), lin=np.array(g('marg_lin')), sqrt_inv_cov=np.array(g('marg_S')),
                            infovec=np.array(g('marg_infovec')))
    return pb


def marg_inputs_to_dict(m):
    d = dict(states=m['states'], inv_depth=m['inv_depth'], prior_frames=np.asarray(m['prior'].frames), prior_lin=m['prior'].lin,
             prior_S=m['prior'].sqrt_inv_cov, prior_infovec=m['prior'].infovec)
    d['pre_ij'] = np.array([[f.i, f.j, int(f.prior)] for f in m['pre']], np.int64).reshape(-1, 3)
    for k in _PF:
        d['pre_' + k] = np.array([np.asarray(getattr(f, k), dtype=np.float64) for f in m['pre']])
    rows = []
    for o in m['obs']:
        for tgt, zt in zip(o['tgt'], o['z_tgt']):
            rows.append([o['lm'], o['ref'], tgt] + list(o['z_ref']) + list(zt))
    d['obs'] = np.array(rows, dtype=np.float64).reshape(-1, 9)
    return d


def marg_inputs_from_dict(d):
    prior = MargPrior(frames=np.array(d['prior_frames']), lin=np.array(d['prior_lin']), sqrt_inv_cov=np.array(d['prior_S']),
                      infovec=np.array(d['prior_infovec']))
    ij = d['pre_ij']
    pre = []
    for k in range(len(ij)):
        vals = {n: (float(d['pre_' + n][k]) if n == 't' else np.array(d['pre_' + n][k], dtype=np.float64)) for n in _PF}
        pre.append(PreFactor(int(ij[k, 0]), int(ij[k, 1]), bool(ij[k, 2]), **vals))
    obs = {}
    for row in d['obs']:
        lm, ref, tgt = int(row[0]), int(row[1]), int(row[2])
        o = obs.setdefault(lm, dict(lm=lm, ref=ref, z_ref=row[3:6].copy(), tgt=[], z_tgt=[]))
        o['tgt'].append(tgt)
        o['z_tgt'].append(row[6:9].copy())
    return dict(states=np.array(d['states'], dtype=np.float64), inv_depth=np.array(d['inv_depth'], dtype=np.float64), prior=prior, pre=pre,
                obs=[obs[k] for k in sorted(obs)])
