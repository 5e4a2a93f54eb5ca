// Operator-level C ABI (include/xrvio.h, layer 1): host buffers in, host buffers out.  Every call
// allocates device memory, copies, launches the CUDA operator and copies back; this path exists
// for parity tests and plug-in style use.  The session layer (capi_session.cu) keeps everything
// resident instead.  No CPU fallback: without a usable CUDA device every call fails.
#include "../../include/xrvio.h"
#include "backend.cuh"
#include "frontend.cuh"
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

static thread_local std::string g_err;
void xrvio_set_error(const std::string &e) { g_err = e; }

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            xrvio_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                      \
            return XRVIO_ERR_CUDA;                                                                    \
        }                                                                                             \
    } while (0)

namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16); }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};
int fail_arg(const char *m) { xrvio_set_error(m); return XRVIO_ERR_ARG; }
}  // namespace

extern "C" {

const char *xrvio_last_error(void) { return g_err.c_str(); }

int xrvio_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { xrvio_set_error("cudaGetDeviceCount failed: no CUDA device"); return 0; }
    return n;
}
int xrvio_set_device(int device) { CK(cudaSetDevice(device)); return XRVIO_OK; }

int xrvio_pyramid_layout(int w, int h, int *widths, int *heights, long long *offsets, long long *total) {
    if (w <= 0 || h <= 0) return fail_arg("bad image size");
    xvfe::PyrLayout L = xvfe::make_pyr_layout(w, h);
    for (int l = 0; l < xvfe::kPyrLevels; ++l) {
        if (widths) widths[l] = L.w[l];
        if (heights) heights[l] = L.h[l];
        if (offsets) offsets[l] = (long long)L.off[l];
    }
    if (total) *total = (long long)L.total;
    return XRVIO_OK;
}

int xrvio_op_undistort(const uint8_t *src, uint8_t *dst, int n, int w, int h, int model, const double K[4], const double D[4]) {
    if (!src || !dst || !K || !D || n <= 0 || w <= 0 || h <= 0 || (model != XRVIO_DISTORTION_RADTAN && model != XRVIO_DISTORTION_EQUIDISTANT))
        return fail_arg("xrvio_op_undistort: bad argument");
    const size_t bytes = (size_t)n * w * h;
    DevBuf d_src, d_dst, d_map;
    CK(d_src.alloc(bytes)); CK(d_dst.alloc(bytes)); CK(d_map.alloc((size_t)w * h * sizeof(int2)));
    CK(cudaMemcpy(d_src.p, src, bytes, cudaMemcpyHostToDevice));
    xvfe::undistort_build_map(xund::make_model(model, K, D), w, h, d_map.as<int2>(), 0);
    xvfe::undistort(d_src.as<uint8_t>(), nullptr, (size_t)w * h, d_dst.as<uint8_t>(), nullptr, (size_t)w * h, n, w, h, d_map.as<int2>(), 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(dst, d_dst.p, bytes, cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

int xrvio_op_clahe(const uint8_t *src, uint8_t *dst, int n, int w, int h, double clip_limit, int tiles_x, int tiles_y) {
    if (!src || !dst || n <= 0 || w <= 0 || h <= 0 || tiles_x <= 0 || tiles_y <= 0) return fail_arg("xrvio_op_clahe: bad argument");
    const size_t bytes = (size_t)n * w * h;
    DevBuf d_src, d_dst, d_lut;
    CK(d_src.alloc(bytes)); CK(d_dst.alloc(bytes)); CK(d_lut.alloc((size_t)n * tiles_x * tiles_y * 256));
    CK(cudaMemcpy(d_src.p, src, bytes, cudaMemcpyHostToDevice));
    xvfe::clahe(d_src.as<uint8_t>(), nullptr, (size_t)w * h, d_dst.as<uint8_t>(), nullptr, (size_t)w * h, d_lut.as<uint8_t>(), n, w, h, clip_limit, tiles_x, tiles_y, 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(dst, d_dst.p, bytes, cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

int xrvio_op_pyramid(const uint8_t *level0, int n, int w, int h, uint8_t *pyr, int16_t *deriv) {
    if (!level0 || !pyr || n <= 0 || w <= 0 || h <= 0) return fail_arg("xrvio_op_pyramid: bad argument");
    xvfe::PyrLayout L = xvfe::make_pyr_layout(w, h);
    DevBuf d_pyr, d_der;
    CK(d_pyr.alloc((size_t)n * L.total));
    for (int i = 0; i < n; ++i)
        CK(cudaMemcpy(d_pyr.as<uint8_t>() + (size_t)i * L.total, level0 + (size_t)i * w * h, (size_t)w * h, cudaMemcpyHostToDevice));
    xvfe::pyramid(d_pyr.as<uint8_t>(), nullptr, n, L, 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(pyr, d_pyr.p, (size_t)n * L.total, cudaMemcpyDeviceToHost));
    if (deriv) {
        CK(d_der.alloc((size_t)L.total * 2 * sizeof(int16_t)));
        for (int i = 0; i < n; ++i) {
            for (int l = 0; l < xvfe::kPyrLevels; ++l)
                xvfe::scharr(d_pyr.as<uint8_t>() + (size_t)i * L.total + L.off[l], d_der.as<int16_t>() + L.off[l] * 2, 1, L.w[l], L.h[l], 0);
            CK(cudaGetLastError());
            CK(cudaMemcpy(deriv + (size_t)i * L.total * 2, d_der.p, (size_t)L.total * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost));
        }
    }
    return XRVIO_OK;
}

int xrvio_op_gftt(const uint8_t *img, int n, int w, int h, int max_corners, double quality, double min_distance,
                  float *corners, int *count) {
    if (!img || !corners || !count || n <= 0 || w < 8 || h < 8) return fail_arg("xrvio_op_gftt: bad argument");
    if (max_corners > XRVIO_MAX_CORNERS) { xrvio_set_error("xrvio_op_gftt: max_corners > XRVIO_MAX_CORNERS"); return XRVIO_ERR_LIMIT; }
    const size_t bytes = (size_t)n * w * h;
    DevBuf d_img, d_cor, d_cnt, d_cand, d_cc, d_max;
    CK(d_img.alloc(bytes)); CK(d_cor.alloc((size_t)n * xvfe::kMaxCorners * 3 * sizeof(float))); CK(d_cnt.alloc(n * sizeof(int)));
    CK(d_cand.alloc((size_t)n * xvfe::kMaxCand * 8)); CK(d_cc.alloc(n * sizeof(int))); CK(d_max.alloc(n * sizeof(unsigned)));
    CK(cudaMemcpy(d_img.p, img, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_cor.p, 0, (size_t)n * xvfe::kMaxCorners * 3 * sizeof(float)));
    xvfe::gftt_harris(d_img.as<uint8_t>(), nullptr, (size_t)w * h, n, w, h, max_corners, (float)quality, (float)min_distance,
                      d_cor.as<float>(), d_cnt.as<int>(), d_cand.as<unsigned long long>(), d_cc.as<int>(), d_max.as<unsigned>(), 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(corners, d_cor.p, (size_t)n * xvfe::kMaxCorners * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(count, d_cnt.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

int xrvio_op_klt(const uint8_t *prev_img, const uint8_t *next_img, int n, int w, int h, const float *prev_pts,
                 float *next_pts, const int *npts, uint8_t *status) {
    if (!prev_img || !next_img || !prev_pts || !next_pts || !npts || !status || n <= 0) return fail_arg("xrvio_op_klt: bad argument");
    for (int i = 0; i < n; ++i)
        if (npts[i] < 0 || npts[i] > XRVIO_MAX_PTS) { xrvio_set_error("xrvio_op_klt: npts > XRVIO_MAX_PTS"); return XRVIO_ERR_LIMIT; }
    xvfe::PyrLayout L = xvfe::make_pyr_layout(w, h);
    DevBuf d_p, d_n, d_pp, d_np, d_cnt, d_st;
    CK(d_p.alloc((size_t)n * L.total)); CK(d_n.alloc((size_t)n * L.total));
    const size_t pb = (size_t)n * xvfe::kMaxPts * 2 * sizeof(float);
    CK(d_pp.alloc(pb)); CK(d_np.alloc(pb)); CK(d_cnt.alloc(n * sizeof(int))); CK(d_st.alloc((size_t)n * xvfe::kMaxPts));
    for (int i = 0; i < n; ++i) {
        CK(cudaMemcpy(d_p.as<uint8_t>() + (size_t)i * L.total, prev_img + (size_t)i * w * h, (size_t)w * h, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_n.as<uint8_t>() + (size_t)i * L.total, next_img + (size_t)i * w * h, (size_t)w * h, cudaMemcpyHostToDevice));
    }
    xvfe::pyramid(d_p.as<uint8_t>(), nullptr, n, L, 0);
    xvfe::pyramid(d_n.as<uint8_t>(), nullptr, n, L, 0);
    CK(cudaMemcpy(d_pp.p, prev_pts, pb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_np.p, next_pts, pb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cnt.p, npts, n * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_st.p, 0, (size_t)n * xvfe::kMaxPts));
    const xvfe::KltMapSet mp = xvfe::klt_make_maps(d_p.as<uint8_t>(), n, L), mn = xvfe::klt_make_maps(d_n.as<uint8_t>(), n, L);
    xvfe::klt_track(d_p.as<uint8_t>(), d_n.as<uint8_t>(), nullptr, nullptr, n, L, d_pp.as<float>(), d_np.as<float>(), d_cnt.as<int>(), d_st.as<uint8_t>(), 0, &mp, &mn);
    CK(cudaGetLastError());
    CK(cudaMemcpy(next_pts, d_np.p, pb, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(status, d_st.p, (size_t)n * xvfe::kMaxPts, cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

int xrvio_op_poisson(const double *pts, const int *n_preset, const int *n_total, const uint8_t *veto, uint8_t *keep,
                     int n, double radius) {
    if (!pts || !n_preset || !n_total || !keep || n <= 0 || !(radius > 0)) return fail_arg("xrvio_op_poisson: bad argument");
    for (int i = 0; i < n; ++i)
        if (n_total[i] > XRVIO_MAX_PTS || n_preset[i] > n_total[i]) { xrvio_set_error("xrvio_op_poisson: list too long"); return XRVIO_ERR_LIMIT; }
    DevBuf d_pts, d_a, d_b, d_v, d_k;
    const size_t pb = (size_t)n * xvfe::kMaxPts * 2 * sizeof(double);
    CK(d_pts.alloc(pb)); CK(d_a.alloc(n * sizeof(int))); CK(d_b.alloc(n * sizeof(int))); CK(d_k.alloc((size_t)n * xvfe::kMaxPts));
    CK(cudaMemcpy(d_pts.p, pts, pb, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_a.p, n_preset, n * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b.p, n_total, n * sizeof(int), cudaMemcpyHostToDevice));
    if (veto) { CK(d_v.alloc((size_t)n * xvfe::kMaxPts)); CK(cudaMemcpy(d_v.p, veto, (size_t)n * xvfe::kMaxPts, cudaMemcpyHostToDevice)); }
    CK(cudaMemset(d_k.p, 0, (size_t)n * xvfe::kMaxPts));
    xvfe::poisson_filter(d_pts.as<double>(), d_a.as<int>(), d_b.as<int>(), veto ? d_v.as<uint8_t>() : nullptr, d_k.as<uint8_t>(), n, radius, 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(keep, d_k.p, (size_t)n * xvfe::kMaxPts, cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

int xrvio_op_preintegrate(const xrvio_rig *rig, const xrvio_preint_job *jobs, int n_jobs, const double *imu, int n_imu,
                          double *out) {
    if (!rig || !jobs || !imu || !out || n_jobs <= 0 || n_imu <= 0) return fail_arg("xrvio_op_preintegrate: bad argument");
    for (int i = 0; i < n_jobs; ++i)
        if (jobs[i].n < 1 || jobs[i].imu_off < 0 || jobs[i].imu_off + jobs[i].n > n_imu) return fail_arg("xrvio_op_preintegrate: job out of range");
    DevBuf d_rig, d_jobs, d_imu, d_out;
    CK(d_rig.alloc(sizeof(xrvio_rig))); CK(d_jobs.alloc(n_jobs * sizeof(xrvio_preint_job)));
    CK(d_imu.alloc((size_t)n_imu * 7 * sizeof(double))); CK(d_out.alloc((size_t)n_jobs * XRVIO_PRE_BLOCK * sizeof(double)));
    CK(cudaMemcpy(d_rig.p, rig, sizeof(xrvio_rig), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_jobs.p, jobs, n_jobs * sizeof(xrvio_preint_job), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_imu.p, imu, (size_t)n_imu * 7 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemset(d_out.p, 0, (size_t)n_jobs * XRVIO_PRE_BLOCK * sizeof(double)));
    xvbe::preintegrate(d_jobs.as<xrvio_preint_job>(), n_jobs, d_imu.as<double>(), d_rig.as<xrvio_rig>(), d_out.as<double>(), 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, d_out.p, (size_t)n_jobs * XRVIO_PRE_BLOCK * sizeof(double), cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

static double g_last_solve_ms = -1.0;
double xrvio_debug_last_solve_ms(void) { return g_last_solve_ms; }

int xrvio_op_solve(const xrvio_rig *rig, xrvio_solve_desc *desc, int n_problems, double *states, int n_frames,
                   const uint8_t *frame_flags, double *inv_depth, int n_lm, const xrvio_rpe *rpe, int n_rpe,
                   const xrvio_rot *rot, int n_rot, const xrvio_pre_ref *pre, int n_pre, const double *pre_blocks,
                   int n_pre_blocks, const int *marg_frames, int n_marg_frames, const double *marg,
                   long long n_marg_doubles, xrvio_solve_out *out) {
    if (!rig || !desc || !states || !frame_flags || !out || n_problems <= 0 || n_frames <= 0) return fail_arg("xrvio_op_solve: bad argument");
    // (a negative block means "still on the device" inside the session layer only; here every block comes from pre_blocks)
    for (int i = 0; i < n_pre; ++i)
        if (pre[i].block < 0 || pre[i].block >= n_pre_blocks) return fail_arg("xrvio_op_solve: preintegration block index out of range");
    long long scratch = 0;
    bool large = false;
    for (int p = 0; p < n_problems; ++p) large = large || desc[p].n_active > XRVIO_MAX_ACTIVE;
    for (int p = 0; p < n_problems; ++p) {
        xrvio_solve_desc &d = desc[p];
        const int max_active = large ? XRVIO_MAX_ACTIVE_LARGE : XRVIO_MAX_ACTIVE;
        if (d.n_active > max_active || d.marg_k > max_active || d.n_pre > 34 || d.n_frames > XRVIO_MAX_FRAMES || d.n_active > d.n_frames ||
            d.max_iterations >= XRVIO_MAX_ITER - 2) {
            xrvio_set_error("xrvio_op_solve: problem exceeds XRVIO_MAX_ACTIVE(_LARGE) / XRVIO_MAX_FRAMES / XRVIO_MAX_ITER");
            return XRVIO_ERR_LIMIT;
        }
        if (d.frame_off < 0 || d.frame_off + d.n_frames > n_frames || d.lm_off + d.n_lm > n_lm || d.rpe_off + d.n_rpe > n_rpe ||
            d.rot_off + d.n_rot > n_rot || d.pre_off + d.n_pre > n_pre || d.n_lm_free > d.n_lm || d.marg_frames_off + d.marg_k > n_marg_frames)
            return fail_arg("xrvio_op_solve: descriptor out of range");
        d.scratch_off = scratch;
        scratch += (long long)(large ? xvbe::solve_large_scratch_doubles(d.n_active, d.n_lm_free, d.n_pre, d.marg_k, d.n_rpe)
                                     : xvbe::solve_scratch_doubles(d.n_active, d.n_lm_free, d.n_pre, d.marg_k, d.n_rpe));
    }
    DevBuf d_rig, d_desc, d_states, d_flags, d_inv, d_rpe, d_rot, d_pre, d_pb, d_mf, d_marg, d_scr, d_out;
    CK(d_rig.alloc(sizeof(xrvio_rig))); CK(cudaMemcpy(d_rig.p, rig, sizeof(xrvio_rig), cudaMemcpyHostToDevice));
    CK(d_desc.alloc(n_problems * sizeof(xrvio_solve_desc))); CK(cudaMemcpy(d_desc.p, desc, n_problems * sizeof(xrvio_solve_desc), cudaMemcpyHostToDevice));
    CK(d_states.alloc((size_t)n_frames * 16 * 8)); CK(cudaMemcpy(d_states.p, states, (size_t)n_frames * 16 * 8, cudaMemcpyHostToDevice));
    CK(d_flags.alloc(n_frames)); CK(cudaMemcpy(d_flags.p, frame_flags, n_frames, cudaMemcpyHostToDevice));
    CK(d_inv.alloc((size_t)(n_lm + 1) * 8)); if (n_lm) CK(cudaMemcpy(d_inv.p, inv_depth, (size_t)n_lm * 8, cudaMemcpyHostToDevice));
    CK(d_rpe.alloc((size_t)(n_rpe + 1) * sizeof(xrvio_rpe))); if (n_rpe) CK(cudaMemcpy(d_rpe.p, rpe, (size_t)n_rpe * sizeof(xrvio_rpe), cudaMemcpyHostToDevice));
    CK(d_rot.alloc((size_t)(n_rot + 1) * sizeof(xrvio_rot))); if (n_rot) CK(cudaMemcpy(d_rot.p, rot, (size_t)n_rot * sizeof(xrvio_rot), cudaMemcpyHostToDevice));
    CK(d_pre.alloc((size_t)(n_pre + 1) * sizeof(xrvio_pre_ref))); if (n_pre) CK(cudaMemcpy(d_pre.p, pre, (size_t)n_pre * sizeof(xrvio_pre_ref), cudaMemcpyHostToDevice));
    CK(d_pb.alloc((size_t)(n_pre_blocks + 1) * XRVIO_PRE_BLOCK * 8)); if (n_pre_blocks) CK(cudaMemcpy(d_pb.p, pre_blocks, (size_t)n_pre_blocks * XRVIO_PRE_BLOCK * 8, cudaMemcpyHostToDevice));
    CK(d_mf.alloc((size_t)(n_marg_frames + 1) * sizeof(int))); if (n_marg_frames) CK(cudaMemcpy(d_mf.p, marg_frames, (size_t)n_marg_frames * sizeof(int), cudaMemcpyHostToDevice));
    CK(d_marg.alloc((size_t)(n_marg_doubles + 1) * 8)); if (n_marg_doubles) CK(cudaMemcpy(d_marg.p, marg, (size_t)n_marg_doubles * 8, cudaMemcpyHostToDevice));
    CK(d_scr.alloc((size_t)scratch * 8)); CK(cudaMemset(d_scr.p, 0, (size_t)scratch * 8));
    CK(d_out.alloc(n_problems * sizeof(xrvio_solve_out))); CK(cudaMemset(d_out.p, 0, n_problems * sizeof(xrvio_solve_out)));
    for (int p = 0; p < n_problems; ++p) {          // A0 = S^T S per prior
        const xrvio_solve_desc &d = desc[p];
        if (d.marg_k <= 0) continue;
        const size_t n = (size_t)15 * d.marg_k;
        if (d.marg_off + (long long)(16 * d.marg_k + 2 * n * n + n + 1) > n_marg_doubles) return fail_arg("xrvio_op_solve: prior arena too small");
        double *S = d_marg.as<double>() + d.marg_off + 16 * d.marg_k;
        xvbe::gram(S, S + n * n + n, (int)n, 1, 0, 0);
    }
    xvbe::SolveBatch b;
    b.desc = d_desc.as<xrvio_solve_desc>(); b.states = d_states.as<double>(); b.frame_flags = d_flags.as<uint8_t>();
    b.inv_depth = d_inv.as<double>(); b.rpe = d_rpe.as<xrvio_rpe>(); b.rot = d_rot.as<xrvio_rot>(); b.pre = d_pre.as<xrvio_pre_ref>();
    b.pre_blocks = d_pb.as<double>(); b.marg_frames = d_mf.as<int>(); b.marg = d_marg.as<double>(); b.scratch = d_scr.as<double>();
    b.out = d_out.as<xrvio_solve_out>();
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    CK(cudaEventCreate(&ev0)); CK(cudaEventCreate(&ev1));
    CK(cudaEventRecord(ev0, 0));
    if (large) xvbe::solve_large(b, n_problems, d_rig.as<xrvio_rig>(), 0);
    else xvbe::solve(b, n_problems, d_rig.as<xrvio_rig>(), 0);
    CK(cudaEventRecord(ev1, 0));
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    { float ms = -1.f; if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) g_last_solve_ms = ms; }
    cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    CK(cudaMemcpy(states, d_states.p, (size_t)n_frames * 16 * 8, cudaMemcpyDeviceToHost));
    if (n_lm) CK(cudaMemcpy(inv_depth, d_inv.p, (size_t)n_lm * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out, d_out.p, n_problems * sizeof(xrvio_solve_out), cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

int xrvio_op_marginalize(const xrvio_rig *rig, xrvio_marg_desc *desc, int n_problems, const double *states, int n_frames,
                         const double *inv_depth, int n_lm, const xrvio_rpe *rpe, int n_rpe, const xrvio_pre_ref *pre,
                         int n_pre, const double *pre_blocks, int n_pre_blocks, const int *marg_frames, int n_marg_frames,
                         const double *marg, long long n_marg_doubles, double *out, long long n_out_doubles, int *rank_out) {
    if (!rig || !desc || !states || !marg || !out || !rank_out || n_problems <= 0) return fail_arg("xrvio_op_marginalize: bad argument");
    for (int i = 0; i < n_pre; ++i)
        if (pre[i].block < 0 || pre[i].block >= n_pre_blocks) return fail_arg("xrvio_op_marginalize: preintegration block index out of range");
    long long scratch = 0;
    for (int p = 0; p < n_problems; ++p) {
        xrvio_marg_desc &d = desc[p];
        if (d.n_frames > XRVIO_MAX_ACTIVE || d.n_frames < 2) { xrvio_set_error("xrvio_op_marginalize: 2 <= n_frames <= XRVIO_MAX_ACTIVE"); return XRVIO_ERR_LIMIT; }
        const long long n1 = 15LL * (d.n_frames - 1);
        if (d.out_off + 16 * (d.n_frames - 1) + 2 * n1 * n1 + n1 + 1 > n_out_doubles) return fail_arg("xrvio_op_marginalize: output arena too small");
        if (d.frame_off + d.n_frames > n_frames || d.lm_off + d.n_lm > n_lm || d.rpe_off + d.n_rpe > n_rpe || d.pre_off + d.n_pre > n_pre ||
            d.n_pre > 2 || d.marg_frames_off + d.marg_k > n_marg_frames)
            return fail_arg("xrvio_op_marginalize: descriptor out of range");
        d.scratch_off = scratch;
        scratch += (long long)xvbe::marg_scratch_doubles(d.n_frames, d.n_lm, d.n_rpe);
    }
    DevBuf d_rig, d_desc, d_states, d_inv, d_rpe, d_pre, d_pb, d_mf, d_marg, d_scr, d_out, d_rank;
    CK(d_rig.alloc(sizeof(xrvio_rig))); CK(cudaMemcpy(d_rig.p, rig, sizeof(xrvio_rig), cudaMemcpyHostToDevice));
    CK(d_desc.alloc(n_problems * sizeof(xrvio_marg_desc))); CK(cudaMemcpy(d_desc.p, desc, n_problems * sizeof(xrvio_marg_desc), cudaMemcpyHostToDevice));
    CK(d_states.alloc((size_t)n_frames * 16 * 8)); CK(cudaMemcpy(d_states.p, states, (size_t)n_frames * 16 * 8, cudaMemcpyHostToDevice));
    CK(d_inv.alloc((size_t)(n_lm + 1) * 8)); if (n_lm) CK(cudaMemcpy(d_inv.p, inv_depth, (size_t)n_lm * 8, cudaMemcpyHostToDevice));
    CK(d_rpe.alloc((size_t)(n_rpe + 1) * sizeof(xrvio_rpe))); if (n_rpe) CK(cudaMemcpy(d_rpe.p, rpe, (size_t)n_rpe * sizeof(xrvio_rpe), cudaMemcpyHostToDevice));
    CK(d_pre.alloc((size_t)(n_pre + 1) * sizeof(xrvio_pre_ref))); if (n_pre) CK(cudaMemcpy(d_pre.p, pre, (size_t)n_pre * sizeof(xrvio_pre_ref), cudaMemcpyHostToDevice));
    CK(d_pb.alloc((size_t)(n_pre_blocks + 1) * XRVIO_PRE_BLOCK * 8)); if (n_pre_blocks) CK(cudaMemcpy(d_pb.p, pre_blocks, (size_t)n_pre_blocks * XRVIO_PRE_BLOCK * 8, cudaMemcpyHostToDevice));
    CK(d_mf.alloc((size_t)(n_marg_frames + 1) * sizeof(int))); if (n_marg_frames) CK(cudaMemcpy(d_mf.p, marg_frames, (size_t)n_marg_frames * sizeof(int), cudaMemcpyHostToDevice));
    CK(d_marg.alloc((size_t)(n_marg_doubles + 1) * 8)); CK(cudaMemcpy(d_marg.p, marg, (size_t)n_marg_doubles * 8, cudaMemcpyHostToDevice));
    CK(d_scr.alloc((size_t)scratch * 8)); CK(cudaMemset(d_scr.p, 0, (size_t)scratch * 8));
    CK(d_out.alloc((size_t)n_out_doubles * 8)); CK(cudaMemset(d_out.p, 0, (size_t)n_out_doubles * 8));
    CK(d_rank.alloc(n_problems * sizeof(int)));
    for (int p = 0; p < n_problems; ++p) {
        const xrvio_marg_desc &d = desc[p];
        const size_t n = (size_t)15 * d.marg_k;
        if (d.marg_off + (long long)(16 * d.marg_k + 2 * n * n + n + 1) > n_marg_doubles) return fail_arg("xrvio_op_marginalize: prior arena too small");
        double *S = d_marg.as<double>() + d.marg_off + 16 * d.marg_k;
        xvbe::gram(S, S + n * n + n, (int)n, 1, 0, 0);
    }
    xvbe::MargBatch b;
    b.desc = d_desc.as<xrvio_marg_desc>(); b.states = d_states.as<double>(); b.inv_depth = d_inv.as<double>();
    b.rpe = d_rpe.as<xrvio_rpe>(); b.pre = d_pre.as<xrvio_pre_ref>(); b.pre_blocks = d_pb.as<double>(); b.marg_frames = d_mf.as<int>();
    b.marg = d_marg.as<double>(); b.out = d_out.as<double>(); b.scratch = d_scr.as<double>(); b.rank_out = d_rank.as<int>();
    xvbe::marginalize(b, n_problems, d_rig.as<xrvio_rig>(), 0);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, d_out.p, (size_t)n_out_doubles * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rank_out, d_rank.p, n_problems * sizeof(int), cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

int xrvio_op_triangulate(const xrvio_tri_job *jobs, int n_jobs, const double *views, int n_views, double *out_xyz_valid) {
    if (!jobs || !views || !out_xyz_valid || n_jobs <= 0 || n_views <= 0) return fail_arg("xrvio_op_triangulate: bad argument");
    for (int i = 0; i < n_jobs; ++i)
        if (jobs[i].obs_off < 0 || jobs[i].obs_off + jobs[i].n_obs > n_views) return fail_arg("xrvio_op_triangulate: job out of range");
    DevBuf d_jobs, d_views, d_out;
    CK(d_jobs.alloc(n_jobs * sizeof(xrvio_tri_job))); CK(d_views.alloc((size_t)n_views * 16 * 8)); CK(d_out.alloc((size_t)n_jobs * 4 * 8));
    CK(cudaMemcpy(d_jobs.p, jobs, n_jobs * sizeof(xrvio_tri_job), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_views.p, views, (size_t)n_views * 16 * 8, cudaMemcpyHostToDevice));
    xvbe::triangulate(d_jobs.as<xrvio_tri_job>(), n_jobs, d_views.as<double>(), d_out.as<double>(), 0);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out_xyz_valid, d_out.p, (size_t)n_jobs * 4 * 8, cudaMemcpyDeviceToHost));
    return XRVIO_OK;
}

}  // extern "C"
