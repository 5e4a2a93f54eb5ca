// Back-end kernels (K6-K12, T): IMU preintegration, sliding-window trust-region solve
// (residuals + Jacobians, block-sparse J^T J, scalar-landmark Schur complement, reduced-system
// Cholesky, dogleg step, accept/reject), marginalisation of the oldest keyframe, DLT
// triangulation.  One CTA per problem / one warp per small job; float64 throughout.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/xrvio.h"

namespace xvbe {

constexpr int kMaxActive = 12;              // free frames per solve (window 10 + the new keyframe)
constexpr int kNP = 15 * kMaxActive;        // 180 pose/motion columns
constexpr int kMaxFrames = 96;              // frames referenced by one problem (active + constant refs)
constexpr int kMaxIter = 64;

using RigParams = xrvio_rig;

// ---- K6 ---------------------------------------------------------------------------------------
using PreintJob = xrvio_preint_job;
constexpr int kPreBlock = 512;              // doubles per preintegration result
// layout of one result block:
//   [0] t  [1..4] q xyzw  [5..7] p  [8..10] v  [11..235] cov 15x15  [236..460] sqrt_inv_cov 15x15
//   [461..469] dq_dbg  [470..478] dp_dbg  [479..487] dp_dba  [488..496] dv_dbg  [497..505] dv_dba
constexpr int PB_T = 0, PB_Q = 1, PB_P = 5, PB_V = 8, PB_COV = 11, PB_SIC = 236, PB_DQ_DBG = 461, PB_DP_DBG = 470,
              PB_DP_DBA = 479, PB_DV_DBG = 488, PB_DV_DBA = 497;

void preintegrate(const PreintJob *jobs, int n_jobs, const double *imu, const RigParams *rig, double *out,
                  cudaStream_t st);

// ---- K7-K11 -----------------------------------------------------------------------------------
using RpeObs = xrvio_rpe;
using RotObs = xrvio_rot;
using PreRef = xrvio_pre_ref;
using SolveDesc = xrvio_solve_desc;
using SolveOut = xrvio_solve_out;
struct SolveBatch {
    const SolveDesc *desc;
    double *states;                  // [frames][16] q xyzw, p, v, bg, ba   (user memory, updated)
    const uint8_t *frame_flags;      // bit0 pose free, bit1 motion free (active frames only)
    double *inv_depth;               // user memory, updated
    const RpeObs *rpe;
    const RotObs *rot;
    const PreRef *pre;
    const double *pre_blocks;
    const int *marg_frames;
    const double *marg;
    double *scratch;
    SolveOut *out;
    const double *pre_out = nullptr; // optional: preintegration results that never left the device; a PreRef with block < 0 reads block ~block here
};
size_t solve_scratch_doubles(int n_active, int n_lm_free, int n_pre, int marg_k, int n_rpe);
void solve(const SolveBatch &b, int n_problems, const RigParams *rig, cudaStream_t st);
// large windows (n_active <= XRVIO_MAX_ACTIVE_LARGE): H in global scratch, FP64 tensor-pipe Schur complement
size_t solve_large_scratch_doubles(int n_active, int n_lm_free, int n_pre, int marg_k, int n_rpe);
void solve_large(const SolveBatch &b, int n_problems, const RigParams *rig, cudaStream_t st);
// A0 = S^T S for a prior (helper used when a prior arrives without it)
void gram(const double *S, double *A0, int n, int n_priors, size_t stride, cudaStream_t st);

// ---- K12 --------------------------------------------------------------------------------------
using MargDesc = xrvio_marg_desc;
struct MargBatch {
    const MargDesc *desc;
    const double *states;
    const double *inv_depth;
    const RpeObs *rpe;
    const PreRef *pre;
    const double *pre_blocks;
    const int *marg_frames;
    const double *marg;
    double *out;
    double *scratch;
    int *rank_out;                          // number of rows kept per problem
};
size_t marg_scratch_doubles(int n_frames, int n_lm, int n_rpe);
void marginalize(const MargBatch &b, int n_problems, const RigParams *rig, cudaStream_t st);

// ---- T ----------------------------------------------------------------------------------------
using TriJob = xrvio_tri_job;       // obs: per view 16 doubles: q_wc xyzw (4), p_wc (3), bearing (3), pad
void triangulate(const TriJob *jobs, int n_jobs, const double *views, double *out_xyz_valid, cudaStream_t st);

}  // namespace xvbe
