"""Flat description of one sliding-window solve, shared vocabulary between the oracle
and the tests of the CUDA solver.  It mirrors what xrslam's Solver wires into Ceres
(xrslam/src/xrslam/estimation/solver.cpp:84-174): frame blocks q(4)/p(3)/v(3)/bg(3)/ba(3),
1-D inverse-depth landmark blocks, reprojection / reprojection-prior / rotation-prior /
preintegration / preintegration-prior / marginalisation residual blocks.
Oracle code: test infrastructure only.
"""
from dataclasses import dataclass, field
import numpy as np


@dataclass
class PreFactor:
    i: int                 # frame index of frame_i
    j: int                 # frame index of frame_j
    prior: bool            # True: CeresPreIntegrationPriorFactor (frame_i read from user memory, constant)
    t: float
    q: np.ndarray          # delta q xyzw
    p: np.ndarray
    v: np.ndarray
    sqrt_inv_cov: np.ndarray   # 15x15 upper
    dq_dbg: np.ndarray
    dp_dbg: np.ndarray
    dp_dba: np.ndarray
    dv_dbg: np.ndarray
    dv_dba: np.ndarray

    @staticmethod
    def from_preintegrator(i, j, pre, prior=False):
        return PreFactor(i, j, prior, float(pre.t), pre.q.copy(), pre.p.copy(), pre.v.copy(),
                         pre.sqrt_inv_cov.copy(), pre.dq_dbg.copy(), pre.dp_dbg.copy(),
                         pre.dp_dba.copy(), pre.dv_dbg.copy(), pre.dv_dba.copy())


@dataclass
class MargPrior:
    frames: np.ndarray         # (K,) frame indices
    lin: np.ndarray            # (K,16) linearisation states
    sqrt_inv_cov: np.ndarray   # (15K,15K)
    infovec: np.ndarray        # (15K,)


@dataclass
class Problem:
    states: np.ndarray                     # (F,16) q xyzw, p, v, bg, ba  == "user memory"
    fix_pose: np.ndarray                   # (F,) bool  FT_FIX_POSE
    fix_motion: np.ndarray                 # (F,) bool  FT_FIX_MOTION
    inv_depth: np.ndarray                  # (M,)      == "user memory"
    cam_q_cs: np.ndarray
    cam_p_cs: np.ndarray
    imu_q_cs: np.ndarray
    imu_p_cs: np.ndarray
    sqrt_inv_cov_px: np.ndarray            # (2,) diagonal of Frame::sqrt_inv_cov (core/detail.cpp:107-109)
    # reprojection residuals (CeresReprojectionErrorFactor / ...PriorFactor)
    rpe_tgt: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rpe_ref: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rpe_lm: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rpe_z_tgt: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    rpe_z_ref: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    rpe_prior: np.ndarray = field(default_factory=lambda: np.zeros(0, bool))
    # rotation prior residuals
    rot_tgt: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rot_ref: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    rot_z_tgt: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    rot_z_ref: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    pre: list = field(default_factory=list)          # list[PreFactor]
    marg: MargPrior = None
    max_iterations: int = 30                          # configs/euroc_slam.yaml:32
