"""Test-side bridge: converts oracle Problem / marginalisation inputs into the flat arrays of
include/xrvio.h and runs the CUDA operators through the C ABI (xrslam_b200.capi), so that the
oracle and the CUDA path can be compared on identical inputs.  Test infrastructure only.
"""
import numpy as np
from xrslam_b200 import capi
from .ceres_tr import _Layout
from .problem import MargPrior

PB = dict(T=0, Q=1, P=5, V=8, COV=11, SIC=236, DQ_DBG=461, DP_DBG=470, DP_DBA=479, DV_DBG=488, DV_DBA=497)


def rig_from_cfg(cfg):
    return capi.make_rig(cfg.q_bc, cfg.p_bc, cfg.q_bi, cfg.p_bi, cfg.sqrt_inv_cov(), cfg.cov_g, cfg.cov_a, cfg.cov_bg, cfg.cov_ba)


def rig_from_problem(pb):
    z = np.zeros((3, 3))
    return capi.make_rig(pb.cam_q_cs, pb.cam_p_cs, pb.imu_q_cs, pb.imu_p_cs, pb.sqrt_inv_cov_px, z, z, z, z)


def preblock(f):
    b = np.zeros(capi.PRE_BLOCK)
    b[PB['T']] = f.t
    b[PB['Q']:PB['Q'] + 4] = f.q
    b[PB['P']:PB['P'] + 3] = f.p
    b[PB['V']:PB['V'] + 3] = f.v
    b[PB['SIC']:PB['SIC'] + 225] = np.asarray(f.sqrt_inv_cov).reshape(225)
    for k in ('DQ_DBG', 'DP_DBG', 'DP_DBA', 'DV_DBG', 'DV_DBA'):
        b[PB[k]:PB[k] + 9] = np.asarray(getattr(f, k.lower())).reshape(9)
    return b


def unpack_preblock(b):
    return dict(t=b[0], q=b[1:5], p=b[5:8], v=b[8:11], cov=b[11:236].reshape(15, 15), sqrt_inv_cov=b[236:461].reshape(15, 15),
                dq_dbg=b[461:470].reshape(3, 3), dp_dbg=b[470:479].reshape(3, 3), dp_dba=b[479:488].reshape(3, 3),
                dv_dbg=b[488:497].reshape(3, 3), dv_dba=b[497:506].reshape(3, 3))


def marg_arena(prior_lin, S, infovec):
    K = len(prior_lin)
    n = 15 * K
    return np.concatenate([np.asarray(prior_lin).reshape(-1), np.asarray(S).reshape(-1), np.asarray(infovec).reshape(-1), np.zeros(n * n + 1)])


class PackedProblems:
    """A batch of oracle Problems packed for xrvio_op_solve."""

    def __init__(self, problems):
        self.problems = problems
        descs = np.zeros(len(problems), capi.solve_desc_dtype)
        states, flags, invd, rpes, rots, pres, blocks, mframes, margs = [], [], [], [], [], [], [], [], []
        self.frame_perm, self.lm_perm = [], []
        f_off = l_off = r_off = ro_off = p_off = mf_off = 0
        m_off = 0
        for i, pb in enumerate(problems):
            lay = _Layout(pb)
            active = lay.free_pose | lay.free_motion
            order = np.concatenate([np.nonzero(active)[0], np.nonzero(~active)[0]]).astype(np.int64)
            inv_order = np.empty_like(order)
            inv_order[order] = np.arange(len(order))
            lm_order = np.concatenate([np.nonzero(lay.free_lm)[0], np.nonzero(~lay.free_lm)[0]]).astype(np.int64)
            lm_inv = np.empty_like(lm_order)
            lm_inv[lm_order] = np.arange(len(lm_order))
            self.frame_perm.append(order)
            self.lm_perm.append(lm_order)
            d = descs[i]
            d['frame_off'], d['n_frames'], d['n_active'] = f_off, len(order), int(active.sum())
            d['lm_off'], d['n_lm'], d['n_lm_free'] = l_off, len(lm_order), int(lay.free_lm.sum())
            states.append(pb.states[order])
            fl = (lay.free_pose.astype(np.uint8) | (lay.free_motion.astype(np.uint8) << 1))[order]
            flags.append(fl)
            invd.append(pb.inv_depth[lm_order])
            O = len(pb.rpe_tgt)
            r = np.zeros(O, capi.rpe_dtype)
            if O:
                r['tgt'] = inv_order[pb.rpe_tgt]
                r['ref'] = inv_order[pb.rpe_ref]
                r['lm'] = lm_inv[pb.rpe_lm]
                r['prior'] = pb.rpe_prior
                r['zt'] = pb.rpe_z_tgt
                r['zr'] = pb.rpe_z_ref
                r = r[np.lexsort((r['ref'], r['tgt']))]
            rpes.append(r)
            d['rpe_off'], d['n_rpe'] = r_off, O
            R = len(pb.rot_tgt)
            ro = np.zeros(R, capi.rot_dtype)
            if R:
                ro['tgt'] = inv_order[pb.rot_tgt]
                ro['ref'] = inv_order[pb.rot_ref]
                ro['zt'] = pb.rot_z_tgt
                ro['zr'] = pb.rot_z_ref
            rots.append(ro)
            d['rot_off'], d['n_rot'] = ro_off, R
            pr = np.zeros(len(pb.pre), capi.pre_ref_dtype)
            for k, f in enumerate(pb.pre):
                pr[k]['i'], pr[k]['j'], pr[k]['prior'] = inv_order[f.i], inv_order[f.j], int(f.prior)
                pr[k]['block'] = sum(len(b) for b in blocks) + k
            blocks.append([preblock(f) for f in pb.pre])
            pres.append(pr)
            d['pre_off'], d['n_pre'] = p_off, len(pb.pre)
            if pb.marg is not None:
                K = len(pb.marg.frames)
                d['marg_k'], d['marg_frames_off'], d['marg_off'] = K, mf_off, m_off
                mframes.append(inv_order[pb.marg.frames].astype(np.int32))
                a = marg_arena(pb.marg.lin, pb.marg.sqrt_inv_cov, pb.marg.infovec)
                margs.append(a)
                mf_off += K
                m_off += len(a)
            d['max_iterations'] = pb.max_iterations
            f_off += len(order); l_off += len(lm_order); r_off += O; ro_off += R; p_off += len(pb.pre)
        cat = lambda xs, dt, shape=None: (np.concatenate(xs) if len(xs) and sum(len(x) for x in xs) else np.zeros((0,) + (shape or ()), dt))
        self.desc = descs
        self.states = np.ascontiguousarray(np.concatenate(states), np.float64)
        self.flags = np.ascontiguousarray(np.concatenate(flags), np.uint8)
        self.inv_depth = np.ascontiguousarray(cat(invd, np.float64), np.float64)
        self.rpe = cat(rpes, capi.rpe_dtype)
        self.rot = cat(rots, capi.rot_dtype)
        self.pre = cat(pres, capi.pre_ref_dtype)
        bl = [b for bs in blocks for b in bs]
        self.pre_blocks = np.array(bl).reshape(len(bl), capi.PRE_BLOCK) if bl else np.zeros((0, capi.PRE_BLOCK))
        self.marg_frames = cat(mframes, np.int32).astype(np.int32)
        self.marg = cat(margs, np.float64)

    def solve(self, rig):
        out = capi.op_solve(rig, self.desc, self.states, self.flags, self.inv_depth, self.rpe, self.rot, self.pre,
                            self.pre_blocks, self.marg_frames, self.marg)
        return out

    def result(self, i):
        """States / inverse depths of problem i in the ORACLE's ordering."""
        d = self.desc[i]
        s = self.states[d['frame_off']:d['frame_off'] + d['n_frames']]
        st = np.empty_like(s)
        st[self.frame_perm[i]] = s
        l = self.inv_depth[d['lm_off']:d['lm_off'] + d['n_lm']]
        ld = np.empty_like(l)
        ld[self.lm_perm[i]] = l
        return st, ld


def solve_problem_cuda(pb):
    """Drop-in for oracle.ceres_tr.solve: mutates pb.states / pb.inv_depth using the CUDA solver."""
    pk = PackedProblems([pb])
    out = pk.solve(rig_from_problem(pb))
    st, ld = pk.result(0)
    pb.states[:] = st
    pb.inv_depth[:] = ld
    return out[0]


def marginalize_cuda(states, prior, pre_factors, obs, inv_depth, cam_q, cam_p, imu_q, imu_p, sic, index=0, eigen_form=True):
    """Drop-in for oracle.marginalization.marginalize (index 0 only) using xrvio_op_marginalize."""
    assert index == 0
    N = states.shape[0]
    order = np.concatenate([np.arange(1, N), [0]])          # victim last
    inv_order = np.empty(N, np.int64)
    inv_order[order] = np.arange(N)
    z = np.zeros((3, 3))
    rig = capi.make_rig(cam_q, cam_p, imu_q, imu_p, sic, z, z, z, z)
    desc = np.zeros(1, capi.marg_desc_dtype)
    rp = []
    for tr in obs:
        for tgt, zt in zip(tr['tgt'], tr['z_tgt']):
            rp.append((inv_order[tgt], inv_order[tr['ref']], tr['lm'], zt, tr['z_ref']))
    rpe = np.zeros(len(rp), capi.rpe_dtype)
    for k, (t, r, l, zt, zr) in enumerate(rp):
        rpe[k]['tgt'], rpe[k]['ref'], rpe[k]['lm'], rpe[k]['zt'], rpe[k]['zr'] = t, r, l, zt, zr
    if len(rp):
        rpe = rpe[np.lexsort((rpe['ref'], rpe['tgt']))]
    pre = np.zeros(len(pre_factors), capi.pre_ref_dtype)
    blocks = np.zeros((max(len(pre_factors), 1), capi.PRE_BLOCK))
    for k, f in enumerate(pre_factors):
        pre[k]['i'], pre[k]['j'], pre[k]['block'] = inv_order[f.i], inv_order[f.j], k
        blocks[k] = preblock(f)
    K = len(prior.frames)
    marg = marg_arena(prior.lin, prior.sqrt_inv_cov, prior.infovec)
    mframes = inv_order[prior.frames].astype(np.int32)
    n1 = 15 * (N - 1)
    n_out = 16 * (N - 1) + 2 * n1 * n1 + n1 + 1
    d = desc[0]
    d['n_frames'], d['n_lm'], d['n_rpe'], d['n_pre'], d['marg_k'] = N, len(inv_depth), len(rpe), len(pre), K
    d['eigen_form'] = 1 if eigen_form else 0
    out, rank = capi.op_marginalize(rig, desc, np.ascontiguousarray(states[order]), np.ascontiguousarray(inv_depth, np.float64),
                                    rpe, pre, blocks, mframes, marg, n_out)
    lin = out[:16 * (N - 1)].reshape(N - 1, 16)
    S = out[16 * (N - 1):16 * (N - 1) + n1 * n1].reshape(n1, n1)
    info = out[16 * (N - 1) + n1 * n1:16 * (N - 1) + n1 * n1 + n1]
    A0 = out[16 * (N - 1) + n1 * n1 + n1:16 * (N - 1) + 2 * n1 * n1 + n1].reshape(n1, n1)
    newp = MargPrior(frames=np.arange(N - 1), lin=lin.copy(), sqrt_inv_cov=S.copy(), infovec=info.copy())
    return newp, A0, int(rank[0])
