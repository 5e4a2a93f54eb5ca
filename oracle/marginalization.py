"""Marginalisation of the oldest keyframe (K12).
Follows xrslam/src/xrslam/estimation/ceres/marginalization_factor.h:74-475 and the initial
prior of xrslam/src/xrslam/estimation/marginalization_factor.h:18-33.
Oracle code: test infrastructure only.
"""
import numpy as np
from . import factors as F
from .problem import MargPrior


def initial_prior(states_first_n_minus_1):
    """estimation/marginalization_factor.h:18-33: zero information except 1e15*I on frame-0 q and p.
    `states_first_n_minus_1`: (K,16) states of all map frames but the newest."""
    K = states_first_n_minus_1.shape[0]
    S = np.zeros((15 * K, 15 * K))
    S[3:6, 3:6] = 1.0e15 * np.eye(3)
    S[0:3, 0:3] = 1.0e15 * np.eye(3)
    return MargPrior(frames=np.arange(K), lin=states_first_n_minus_1.copy(), sqrt_inv_cov=S,
                     infovec=np.zeros(15 * K))


def _st(s):
    return (s[0:4], s[4:7], s[7:10], s[10:13], s[13:16])


def marginalize(states, prior, pre_factors, obs, inv_depth, cam_q, cam_p, imu_q, imu_p, sic, index=0):
    """Marginalise map frame `index` (only 0 is used by the reference, map.cpp:51-53).

    states       (N,16) current states of all N map keyframes (user memory)
    prior        MargPrior over prior.frames (indices into the N frames)
    pre_factors  list of PreFactor for the edges (index-1,index) and (index,index+1) that exist
                 (keyframe_preintegration of frame j, marginalization_factor.h:164-231)
    obs          list of tracks seen by the victim, each a dict
                   {'lm': landmark id, 'ref': map frame index of the track's first frame,
                    'z_ref': (3,), 'tgt': [frame indices], 'z_tgt': [(3,)...]}  -- valid tracks whose
                 first frame is a keyframe; tgt lists every other map keyframe observing the track
                 (marginalization_factor.h:234-256).
    Returns a new MargPrior over the N-1 remaining frames (indices re-based to the new map order),
    linearised at the current states (:456-473).
    """
    N = states.shape[0]
    last = N - 1
    fidx = np.empty(N, np.int64)
    for i in range(N):
        fidx[i] = i if i < index else (i - 1 if i > index else last)
    Hm = np.zeros((15 * N, 15 * N))
    bv = np.zeros(15 * N)

    # ---- old prior (:108-161) ----
    sts = [_st(states[i]) for i in prior.frames]
    lin = [_st(l) for l in prior.lin]
    r, J = F.marginalization_prior(sts, lin, prior.sqrt_inv_cov, prior.infovec, True)
    K = len(prior.frames)
    for a in range(K):
        ia = fidx[prior.frames[a]]
        Ja = J[:, 15 * a:15 * a + 15]
        for b in range(K):
            ib = fidx[prior.frames[b]]
            Hm[15 * ia:15 * ia + 15, 15 * ib:15 * ib + 15] += Ja.T @ J[:, 15 * b:15 * b + 15]
        bv[15 * ia:15 * ia + 15] += Ja.T @ r

    # ---- preintegration factors touching the victim (:164-231); dbg = dba = 0 here ----
    for f in pre_factors:
        r, Ji, Jj = F.preintegration(_st(states[f.i]), _st(states[f.j]), f, states[f.i, 10:13],
                                     states[f.i, 13:16], imu_q, imu_p, True)
        ii, jj = fidx[f.i], fidx[f.j]
        si, sj = slice(15 * ii, 15 * ii + 15), slice(15 * jj, 15 * jj + 15)
        Hm[si, si] += Ji.T @ Ji
        Hm[si, sj] += Ji.T @ Jj
        Hm[sj, si] += Jj.T @ Ji
        Hm[sj, sj] += Jj.T @ Jj
        bv[si] += Ji.T @ r
        bv[sj] += Jj.T @ r

    # ---- reprojection factors of the victim's tracks, no robust loss (:234-380) ----
    sic2 = np.diag(sic)
    linfo = {}
    for tr in obs:
        ref = tr['ref']
        iref = fidx[ref]
        for tgt, z_tgt in zip(tr['tgt'], tr['z_tgt']):
            itgt = fidx[tgt]
            r, J = F.reprojection(states[tgt, 0:4], states[tgt, 4:7], states[ref, 0:4], states[ref, 4:7],
                                  inv_depth[tr['lm']], z_tgt, tr['z_ref'], cam_q, cam_p, sic2, True)
            Jt = np.concatenate([J['q_tgt'], J['p_tgt']], axis=1)
            Jr = np.concatenate([J['q_ref'], J['p_ref']], axis=1)
            Jl = J['inv_depth']
            st, sr = slice(15 * itgt, 15 * itgt + 6), slice(15 * iref, 15 * iref + 6)
            Hm[st, st] += Jt.T @ Jt
            Hm[sr, st] += Jr.T @ Jt
            Hm[st, sr] += Jt.T @ Jr
            Hm[sr, sr] += Jr.T @ Jr
            bv[st] += Jt.T @ r
            bv[sr] += Jr.T @ r
            li = linfo.setdefault(tr['lm'], {'mat': 0.0, 'vec': 0.0, 'h': {}})
            li['mat'] += float(Jl @ Jl)
            li['vec'] += float(Jl @ r)
            li['h'].setdefault(itgt, np.zeros(6))
            li['h'].setdefault(iref, np.zeros(6))
            li['h'][itgt] = li['h'][itgt] + Jl @ Jt
            li['h'][iref] = li['h'][iref] + Jl @ Jr

    # ---- marginalise landmarks (:383-398) ----
    for lm in sorted(linfo.keys()):
        li = linfo[lm]
        with np.errstate(divide='ignore'):
            inv = np.float64(1.0) / np.float64(li['mat'])
        if not np.isfinite(inv):
            continue
        for fi, hi in li['h'].items():
            for fj, hj in li['h'].items():
                Hm[15 * fi:15 * fi + 6, 15 * fj:15 * fj + 6] -= np.outer(hi, hj) * inv
            bv[15 * fi:15 * fi + 6] -= hi * inv * li['vec']

    # ---- marginalise the victim frame (:400-438) ----
    n1 = 15 * last
    inv = np.linalg.inv(Hm[n1:, n1:])
    Hc = Hm[:n1, :n1] - Hm[:n1, n1:] @ inv @ Hm[n1:, :n1]
    bc = bv[:n1] - Hm[:n1, n1:] @ inv @ bv[n1:]

    # ---- eigen re-factorisation (:440-454) ----
    lam, V = np.linalg.eigh(Hc)
    keep = lam > 1.0e-8
    lam_k = np.where(keep, lam, 0.0)
    lam_inv = np.where(keep, 1.0 / np.where(keep, lam, 1.0), 0.0)
    S = np.sqrt(lam_k)[:, None] * V.T
    infovec = np.sqrt(lam_inv) * (V.T @ bc)

    remain = [i for i in range(N) if i != index]
    new_lin = states[remain].copy()
    return MargPrior(frames=np.arange(N - 1), lin=new_lin, sqrt_inv_cov=S, infovec=infovec), Hc, bc
