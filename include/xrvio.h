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
#define XRVIO_MAX_ACTIVE 12       /* free frames per solve (sliding window)    */
#define XRVIO_MAX_ACTIVE_LARGE 32 /* free frames per solve, large-window kernel */
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
    int block;                    /* index of its result block in pre_blocks (0 <= block < n_pre_blocks; anything else is XRVIO_ERR_ARG) */
    int pad;
} xrvio_pre_ref;

/* One sliding-window problem.  Frames [0,n_active) are the free ones (<= XRVIO_MAX_ACTIVE), the
 * rest are constant reference frames.  Landmarks [0,n_lm_free) are free.  rpe must be sorted by
 * (tgt, ref).  Marginalisation prior arena (doubles at marg_off): lin[16K] S[(15K)^2] infovec[15K]
 * A0[(15K)^2] rank[1]  (A0 = S^T S is filled by the library in xrvio_op_solve; rank = number of leading
 * non-zero rows of S, 0 = unknown -> all rows are used). */
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
    long long out_off;            /* new prior: lin[16(N-1)] S[n^2] infovec[n] A0[n^2] rank[1], n = 15(N-1) */
    long long scratch_off;        /* filled by the library */
    int eigen_form;               /* != 0: rows of S are refined to sqrt(lambda_i) v_i^T (eigen form of the reference); 0: rows of
                                     the rank-revealing Cholesky factor (same S^T S, same prior cost and Jacobian) */
    int pad;
} xrvio_marg_desc;

typedef struct xrvio_tri_job { int obs_off, n_obs; } xrvio_tri_job;   /* views: 16 doubles each: q_wc, p_wc, bearing */

const char *xrvio_last_error(void);
int xrvio_device_count(void);
int xrvio_set_device(int device);

int xrvio_pyramid_layout(int w, int h, int *widths, int *heights, long long *offsets, long long *total);

/* Undistortion of n w*h images (row N2).  model XRVIO_DISTORTION_RADTAN restates cv::undistort(img, K, D) as the EuRoC
 * reader calls it (euroc_dataset_reader.cpp:62-72), XRVIO_DISTORTION_EQUIDISTANT restates xrslam::extra::ImageUndistorter
 * + cv::remap (image_undistorter.h:13-97).  K = fx fy cx cy. */
#define XRVIO_DISTORTION_NONE 0
#define XRVIO_DISTORTION_RADTAN 1
#define XRVIO_DISTORTION_EQUIDISTANT 2
int xrvio_op_undistort(const uint8_t *src, uint8_t *dst, int n, int w, int h, int model, const double K[4], const double D[4]);
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

/* ------------------------------------------------------------------------------------------------
 * Session level: n_sessions independent VIO sessions advanced together on one GPU.
 * Mirrors, per session, XRSLAM::Detail (xrslam/src/xrslam/core/detail.h:33-45):
 *   xrvio_push_imu     <-> track_imu after gyro/accelerometer interpolation (detail.cpp:130-142)
 *   xrvio_push_images  <-> track_camera (detail.cpp:103-128): queues the frame, returns nothing; the
 *                          IMU-propagated output pose for that image is read with xrvio_get_pose
 *   xrvio_step         <-> the Worker loops FeatureTracker::work + FrontendWorker::work for every frame
 *                          released by a later IMU sample (feature_tracker.cpp:24-153,
 *                          frontend_worker.cpp:28-86), batched over sessions
 *   xrvio_get_state    <-> get_system_state (frontend_worker.cpp:120-127)
 * The reference's Initializer (SfM + IMU alignment) is not part of this hot path:
 * a session leaves INITIALIZING through xrvio_push_bootstrap_state (poses/velocities of the frames,
 * e.g. ground truth) followed by the reference's own mirror_keyframe_map + final visual-inertial BA.
 */
typedef struct xrvio_ctx xrvio_ctx;

typedef struct xrvio_params {
    int width, height;
    double fx, fy, cx, cy;
    double q_bc[4], p_bc[3], q_bi[4], p_bi[3], q_bo[4], p_bo[3];
    double cov_g[9], cov_a[9], cov_bg[9], cov_ba[9];
    double keypoint_cov[2];
    int sliding_window_size, subframe_size, force_keyframe_landmarks;
    double min_keypoint_distance;
    int max_keypoint_detection, max_init_frames, max_frames, predict_keypoints;
    double clahe_clip_limit;
    int clahe_width, clahe_height;
    int init_keyframe_num, init_keyframe_gap;
    int solver_iteration_limit;
    double rotation_misalignment_threshold, rotation_ransac_threshold;
    int prior_eigen_form;         /* 1: marginalisation prior kept in eigen form (see xrvio_marg_desc.eigen_form); default 0 */
    int initializer;              /* 1 (default): a session that was given no bootstrap state runs the reference's own
                                     initializer (core/initializer.cpp:78-571: two-view SfM seed, PnP, visual BA, IMU
                                     alignment) before its first visual-inertial BA; 0: it waits for bootstrap states */
    int init_min_matches, init_min_triangulation, init_min_landmarks, init_refine_imu;   /* configs/euroc_slam.yaml:21-28 */
    double init_min_parallax;
    int random_seed;              /* Config::random() (config.cpp:66), seed of the RANSAC samplers of the initializer */
    int distortion_model;         /* 0 (default): pushed images are already undistorted, as with the reference's readers;
                                     XRVIO_DISTORTION_RADTAN / _EQUIDISTANT: the library undistorts every pushed image on the
                                     device with the readers' arithmetic (xrvio_op_undistort) before CLAHE */
    double distortion[4];         /* k1 k2 p1 p2 (radtan) or k1 k2 k3 k4 (equidistant); intrinsics are fx fy cx cy above */
    int essential_ransac;         /* 1: run the per-frame 5-point essential RANSAC of Frame::track_keypoints (frame.cpp:108-114,
                                     threshold 1.0) on the host; 0 (default): use its observable effect -- with that threshold
                                     (2 * 3.84 in squared normalised image units) every match is an inlier, so the mask is all
                                     ones, or all zeros below 5 points (tests/test_oracle_geometry.py shows it on a stream) */
    int pad;
} xrvio_params;

typedef struct xrvio_imu { double t, w[3], a[3]; } xrvio_imu;
typedef struct xrvio_pose { double q[4], p[3], t; int valid; int pad; } xrvio_pose;   /* q xyzw, like XRSLAMPose */
typedef struct xrvio_stats {
    long long frames_processed;         /* frames through the front end over all sessions */
    long long kernels_launched;         /* CUDA kernel launches issued by the library */
    long long h2d_bytes, d2h_bytes;
    long long window_solves, localize_solves, subwindow_solves, marginalizations;
    double last_window_cost_initial, last_window_cost_final;
    long long window_iterations;        /* trust-region iterations (accepted + rejected) summed over the refine_window solves */
    long long window_at_limit;          /* refine_window solves that ended at solver_iteration_limit (all of them) */
    long long graph_launches;           /* CUDA graph launches (each replays the 12-13 image-stream operations of one push batch) */
} xrvio_stats;

/* Outcome of the initializer attempts of a session (core/initializer.cpp:78-571).  last_fail: 0 initialised (or not yet
 * tried), 1 too few matches, 2 too little parallax, 3 pure rotation, 4 two-view triangulation, 5 visual BA unusable,
 * 6 scale outside [0.001, 1], 7 too few landmarks, 8 keyframe without IMU data, 9 fewer frames than the keyframe span */
typedef struct xrvio_init_info {
    int attempts, last_fail, matches, candidate, landmarks;
    int essential_rejected;       /* matches rejected so far by the optional per-frame essential RANSAC (xrvio_params.essential_ransac) */
    double parallax, scale;
} xrvio_init_info;

#define XRVIO_STATE_INITIALIZING 0   /* SysState::SYS_INITIALIZING */
#define XRVIO_STATE_TRACKING 1       /* SysState::SYS_TRACKING */

int xrvio_default_params(xrvio_params *p, int width, int height);   /* configs/euroc_slam.yaml + euroc_sensor.yaml */
int xrvio_create(const xrvio_params *p, int n_sessions, int device, xrvio_ctx **ctx);
int xrvio_destroy(xrvio_ctx *ctx);
int xrvio_push_imu(xrvio_ctx *ctx, const int *session, const xrvio_imu *samples, int n);
int xrvio_push_bootstrap_state(xrvio_ctx *ctx, int session, double t, const double q[4], const double p[3], const double v[3]);
/* data[i]: pointer to a height*stride[i] 8-bit image; on_device != 0: device pointers */
int xrvio_push_images(xrvio_ctx *ctx, const int *session, const uint8_t *const *data, const int *stride, const double *t,
                      int n, int on_device);
int xrvio_step(xrvio_ctx *ctx);
/* Asynchronous mode (off by default; needs a context of >= 8 sessions, i.e. more than one lane).  xrvio_push_imu,
 * xrvio_push_images and xrvio_step only QUEUE their work per lane and return; every lane of sessions works through its queue at
 * its own pace, at most six steps behind the caller (xrvio_step blocks beyond that).  Image buffers handed to
 * xrvio_push_images must stay valid until the lane has consumed them (xrvio_sync, or six further steps).  xrvio_get_pose does
 * not wait and returns the newest pose available (pose->t says for which frame) -- the behaviour of the reference's threaded
 * build, where Detail::get_latest_pose reads whatever the tracker has reached; every other getter waits for the session's lane.
 * Errors of queued work are reported by xrvio_sync or the next waiting call.  Results per session are identical to synchronous
 * mode: sessions never interact, only the interleaving of lanes changes. */
int xrvio_set_async(xrvio_ctx *ctx, int enable);
int xrvio_sync(xrvio_ctx *ctx);
int xrvio_get_pose(xrvio_ctx *ctx, int session, xrvio_pose *pose);          /* output pose of the latest pushed image */
int xrvio_get_latest_state(xrvio_ctx *ctx, int session, double *t, double state16[16]);   /* last optimised frame */
int xrvio_get_state(xrvio_ctx *ctx, int session);                           /* XRVIO_STATE_* or < 0 */
int xrvio_get_landmarks(xrvio_ctx *ctx, int session, double *xyz, int capacity);   /* returns count */
/* Keypoints of the newest frame of the session's front-end map (what FeatureTracker::work leaves behind, feature_tracker.cpp:
 * 128-130): pixel position (apply_k of the bearing), id of the track through it (-1: none yet) and that track's length in
 * frames.  Returns the number of keypoints (may exceed capacity; only capacity entries are written).  The reference's
 * XRSLAM_RESULT_FEATURES is declared but empty (XRSLAMManager.cpp:217); this is the data it would carry. */
int xrvio_get_features(xrvio_ctx *ctx, int session, double *xy, long long *track_id, int *track_len, int capacity);
int xrvio_get_init_info(xrvio_ctx *ctx, int session, xrvio_init_info *out);
int xrvio_get_stats(xrvio_ctx *ctx, xrvio_stats *stats);
/* Per-kernel-family device time measured with CUDA events on the launch stream (off by default).
 * Families: 0 clahe 1 pyramid 2 gftt 3 klt 4 poisson 5 preintegration 6 solve 7 marginalize 8 triangulate.
 * Returns the number of families. */
int xrvio_set_profiling(xrvio_ctx *ctx, int enable);
int xrvio_get_kernel_times(xrvio_ctx *ctx, double *ms, long long *launches, int capacity);
/* Host wall time accumulated per internal step phase (diagnostics; up to 32 slots). */
int xrvio_get_phase_times(xrvio_ctx *ctx, double *ms, int capacity);
/* Diagnostics: wall time each lane of sessions has spent inside its steps so far (ms); returns the number of lanes. */
int xrvio_debug_lane_busy(xrvio_ctx *ctx, double *ms_per_lane, int capacity);
/* Diagnostics: clock64 cycles accumulated per section of the solve kernel over all window problems (32 slots, slot 31 =
 * number of problems); see tools/phase_probe.py for the slot names. */
int xrvio_debug_kprof(unsigned long long *out32, int reset);
/* Diagnostics: device time (CUDA events on the launch stream) of the solve kernel launched by the last xrvio_op_solve call of this
 * thread's process, in ms; < 0 before the first call.  bench.py --workload ba_stress reads it for the config-5 roofline line. */
double xrvio_debug_last_solve_ms(void);
/* Diagnostics: measured FP64 rates of the current device in TFLOP/s (register-resident DFMA loop; mma.sync.m8n8k4.f64 loop), the
 * roofline denominators of the large-window solve -- MEASURED_PEAKS.json has no FP64 entry. */
int xrvio_debug_fp64_peak(double *dfma_tflops, double *dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* XRVIO_H */
