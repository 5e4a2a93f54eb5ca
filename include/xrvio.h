/* xrvio.h -- C ABI of the B200-native VIO core (libxrvio.so).
 *
 * Two layers:
 *   (1) operator level (xrvio_op_*): one call = one batched CUDA operator of the hot path on HOST
 *       buffers (the library does the H2D/D2H copies).  These are the entry points the reference's
 *       operator plug-in interface would bind, and what the parity tests call:
 *         xrslam::Image::preprocess          xrslam/include/xrslam/xrslam.h:137-161
 *                                            (impl. xrslam-extra/src/xrslam/extra/opencv_image.cpp:156-161)
 *             -> xrvio_op_clahe, xrvio_op_pyramid
 *         xrslam::Image::detect_keypoints    (opencv_image.cpp:38-73)    -> xrvio_op_gftt, xrvio_op_poisson
 *         xrslam::Image::track_keypoints     (opencv_image.cpp:75-154)   -> xrvio_op_klt
 *         PreIntegrator::integrate           xrslam/src/xrslam/estimation/preintegrator.h:33-38
 *             -> xrvio_op_preintegrate
 *         Solver::solve                      xrslam/src/xrslam/estimation/solver.h:16-71 (solver.cpp:176-190)
 *             -> xrvio_op_solve
 *         MarginalizationFactor::marginalize xrslam/src/xrslam/estimation/marginalization_factor.h:13
 *             -> xrvio_op_marginalize
 *         Track::triangulate                 xrslam/src/xrslam/map/track.h:71      -> xrvio_op_triangulate
 *   (2) session level (xrvio_create / push / step / get): many independent VIO sessions advanced per
 *       call, the batched replacement for XRSLAM::Detail::track_gyroscope / track_accelerometer /
 *       track_camera / get_system_state (xrslam/src/xrslam/core/detail.h:33-45).
 *
 * Plain C structs, int error codes (0 = ok, < 0 = error, text via xrvio_last_error), no exceptions
 * cross the boundary, caller owns every buffer.  Not re-entrant per context.  There is NO CPU
 * fallback: every call fails with XRVIO_ERR_CUDA when no CUDA device is usable.
 */
#ifndef XRVIO_H
#define XRVIO_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XRVIO_OK 0
#define XRVIO_ERR_CUDA (-1)
#define XRVIO_ERR_ARG (-2)
#define XRVIO_ERR_LIMIT (-3)
#define XRVIO_ERR_STATE (-4)

#define XRVIO_PYR_LEVELS 4
#define XRVIO_MAX_CORNERS 256     /* slots per image in xrvio_op_gftt output   */
#define XRVIO_MAX_PTS 512         /* keypoint slots per image (klt / poisson)  */
#define XRVIO_PRE_BLOCK 512       /* doubles per preintegration result         */
#define XRVIO_MAX_ACTIVE 11       /* free frames per solve                     */
#define XRVIO_MAX_FRAMES 96       /* frames referenced by one solve            */
#define XRVIO_MAX_ITER 64

/* Rig / noise constants shared by all back-end operators (core/detail.cpp:103-118). */
typedef struct xrvio_rig {
    double cam_q[4], cam_p[3];    /* camera-to-body  (q xyzw) */
    double imu_q[4], imu_p[3];    /* imu-to-body              */
    double sic[2];                /* diag of Frame::sqrt_inv_cov = (fx, fy) / sqrt(keypoint noise cov) */
    double cov_w[9], cov_a[9], cov_bg[9], cov_ba[9];   /* continuous-time IMU noise covariances */
} xrvio_rig;

/* K6.  Result block layout (XRVIO_PRE_BLOCK doubles):
 *  [0] t [1..4] q xyzw [5..7] p [8..10] v [11..235] cov 15x15 [236..460] sqrt_inv_cov 15x15 (upper)
 *  [461..469] dq_dbg [470..478] dp_dbg [479..487] dp_dba [488..496] dv_dbg [497..505] dv_dba          */
typedef struct xrvio_preint_job {
    int imu_off, n, flags, pad;   /* flags bit0: bias Jacobians, bit1: covariance + sqrt_inv_cov */
    double t_end;
    double bg[3], ba[3];
} xrvio_preint_job;

typedef struct xrvio_rpe {        /* reprojection residual (reprojection_factor.h) */
    short tgt, ref;
    int lm;
    int prior;                    /* 1: ReprojectionPriorFactor (reference pose + depth constant) */
    int pad;
    double zt[3], zr[3];
} xrvio_rpe;
typedef struct xrvio_rot {        /* rotation prior residual (rotation_factor.h) */
    short tgt, ref;
    int pad;
    double zt[3], zr[3];
} xrvio_rot;
typedef struct xrvio_pre_ref {    /* preintegration residual (preintegration_factor.h) */
    short i, j;
    int prior;                    /* 1: PreIntegrationPriorFactor (frame_i constant) */
    int block;                    /* index of its result block */
    int pad;
} xrvio_pre_ref;

/* One sliding-window problem.  Frames [0,n_active) are the free ones (<= XRVIO_MAX_ACTIVE), the
 * rest are constant reference frames.  Landmarks [0,n_lm_free) are free.  rpe must be sorted by
 * (tgt, ref).  Marginalisation prior arena (doubles at marg_off): lin[16K] S[(15K)^2] infovec[15K]
 * A0[(15K)^2] (A0 = S^T S is filled by the library in xrvio_op_solve). */
typedef struct xrvio_solve_desc {
    int frame_off, n_frames, n_active;
    int lm_off, n_lm, n_lm_free;
    int rpe_off, n_rpe;
    int rot_off, n_rot;
    int pre_off, n_pre;
    int marg_k, marg_frames_off;
    long long marg_off;
    long long scratch_off;        /* filled by the library */
    int max_iterations;
    int pad;
} xrvio_solve_desc;
typedef struct xrvio_solve_out {
    int iterations, termination, n_successful, n_unsuccessful;
    double initial_cost, final_cost;
    double cost_trace[XRVIO_MAX_ITER];
} xrvio_solve_out;

/* K12.  Frames: the N map keyframes with the victim LAST (marginalization_factor.h:95-105). */
typedef struct xrvio_marg_desc {
    int frame_off, n_frames;
    int lm_off, n_lm;
    int rpe_off, n_rpe;
    int pre_off, n_pre;
    int marg_k, marg_frames_off;
    long long marg_off;
    long long out_off;            /* new prior: lin[16(N-1)] S[n^2] infovec[n] A0[n^2], n = 15(N-1) */
    long long scratch_off;        /* filled by the library */
    int pad[2];
} xrvio_marg_desc;

typedef struct xrvio_tri_job { int obs_off, n_obs; } xrvio_tri_job;   /* views: 16 doubles each: q_wc, p_wc, bearing */

const char *xrvio_last_error(void);
int xrvio_device_count(void);
int xrvio_set_device(int device);

int xrvio_pyramid_layout(int w, int h, int *widths, int *heights, long long *offsets, long long *total);

int xrvio_op_clahe(const uint8_t *src, uint8_t *dst, int n, int w, int h, double clip_limit, int tiles_x, int tiles_y);
int xrvio_op_pyramid(const uint8_t *level0, int n, int w, int h, uint8_t *pyr, int16_t *deriv /* may be NULL */);
int xrvio_op_gftt(const uint8_t *img, int n, int w, int h, int max_corners, double quality, double min_distance,
                  float *corners /* n*XRVIO_MAX_CORNERS*3 */, int *count);
int xrvio_op_klt(const uint8_t *prev_img, const uint8_t *next_img, int n, int w, int h, const float *prev_pts,
                 float *next_pts /* in: initial guess, out: tracked */, const int *npts, uint8_t *status);
int xrvio_op_poisson(const double *pts, const int *n_preset, const int *n_total, const uint8_t *veto /* may be NULL */,
                     uint8_t *keep, int n, double radius);
int xrvio_op_preintegrate(const xrvio_rig *rig, const xrvio_preint_job *jobs, int n_jobs, const double *imu,
                          int n_imu, double *out);
int xrvio_op_solve(const xrvio_rig *rig, xrvio_solve_desc *desc, int n_problems, double *states, int n_frames,
                   const uint8_t *frame_flags, double *inv_depth, int n_lm, const xrvio_rpe *rpe, int n_rpe,
                   const xrvio_rot *rot, int n_rot, const xrvio_pre_ref *pre, int n_pre, const double *pre_blocks,
                   int n_pre_blocks, const int *marg_frames, int n_marg_frames, const double *marg,
                   long long n_marg_doubles, xrvio_solve_out *out);
int xrvio_op_marginalize(const xrvio_rig *rig, xrvio_marg_desc *desc, int n_problems, const double *states,
                         int n_frames, const double *inv_depth, int n_lm, const xrvio_rpe *rpe, int n_rpe,
                         const xrvio_pre_ref *pre, int n_pre, const double *pre_blocks, int n_pre_blocks,
                         const int *marg_frames, int n_marg_frames, const double *marg, long long n_marg_doubles,
                         double *out, long long n_out_doubles, int *rank_out);
int xrvio_op_triangulate(const xrvio_tri_job *jobs, int n_jobs, const double *views, int n_views, double *out_xyz_valid);

#ifdef __cplusplus
}
#endif
#endif /* XRVIO_H */
