// Outer drop-in boundary: the reference's process-wide C ABI (xrslam-interface/include/XRSLAM.h:201-229,
// implementation xrslam-interface/src/XRSLAMInternal.cpp:4-88 and XRSLAMManager.cpp:85-240) on top of a
// one-session xrvio context.  The sensor synchronisation restates XRSLAM::Detail::track_gyroscope /
// track_accelerometer (xrslam/src/xrslam/core/detail.cpp:46-101): gyroscope samples are linearly
// interpolated to accelerometer timestamps before they enter the session.
#include "../../include/xrslam_compat.h"
#include "../../include/xrvio.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Gyr { double t, w[3]; };
struct Acc { double t, a[3]; };

struct Manager {
    xrvio_ctx *ctx = nullptr;
    xrvio_params prm{};
    std::deque<Gyr> gyroscopes;
    std::deque<Acc> accelerometers;
    std::vector<unsigned char> image;       // grey copy of the pushed image (XRSLAMManager::PushImage, :104-136)
    double image_t = 0;
    bool has_image = false;
    std::mutex input_mutex;
    XRSLAMPose last_pose{};
    int last_rc = 0;
    char version[32] = "xrvio-b200 0.1";
};
Manager *g_mgr = nullptr;

// minimal reader for the OpenCV-style yaml files (configs/euroc_*.yaml): finds `key:` and the numbers that follow
std::string slurp(const char *path) {
    std::ifstream f(path);
    std::stringstream ss;
    std::string line;
    while (std::getline(f, line)) {
        size_t h = line.find('#');
        if (h != std::string::npos) line = line.substr(0, h);
        ss << line << '\n';
    }
    return ss.str();
}
bool numbers_after(const std::string &txt, const std::string &key, int n, double *out, size_t from = 0) {
    size_t p = from;
    while (true) {      // whole-key match: "size" must not hit "subframe_size"
        p = txt.find(key + ":", p);
        if (p == std::string::npos) return false;
        if (p == 0 || !(isalnum((unsigned char)txt[p - 1]) || txt[p - 1] == '_')) break;
        p += key.size();
    }
    p += key.size() + 1;
    int got = 0;
    while (got < n && p < txt.size()) {
        while (p < txt.size() && !(isdigit((unsigned char)txt[p]) || txt[p] == '-' || txt[p] == '+' || txt[p] == '.')) {
            if (isalpha((unsigned char)txt[p]) && txt[p] != 'e' && txt[p] != 'E') return false;   // ran into the next key
            ++p;
        }
        if (p >= txt.size()) break;
        char *end = nullptr;
        out[got++] = std::strtod(txt.c_str() + p, &end);
        p = end - txt.c_str();
    }
    return got == n;
}

void track_imu(Manager *m, double t, const double w[3], const double a[3]) {
    int s = 0;
    xrvio_imu smp;
    smp.t = t;
    for (int k = 0; k < 3; ++k) { smp.w[k] = w[k]; smp.a[k] = a[k]; }
    // Worker::resume runs the workers inline in the PC build (utility/worker.h:35-42).  The reference's entry points return
    // void, so an error can only be logged; it is logged once per kind so a broken stream does not flood stderr.
    int rc = xrvio_push_imu(m->ctx, &s, &smp, 1);
    if (rc == XRVIO_OK) rc = xrvio_step(m->ctx);
    if (rc != XRVIO_OK && rc != m->last_rc) std::fprintf(stderr, "XRSLAM: tracking step failed (%d): %s\n", rc, xrvio_last_error());
    m->last_rc = rc;
}

// detail.cpp:46-70
void track_gyroscope(Manager *m, double t, double x, double y, double z) {
    auto &G = m->gyroscopes;
    auto &A = m->accelerometers;
    if (!A.empty()) {
        if (t < A.front().t) G.clear();
        else {
            while (!A.empty() && t >= A.front().t) {
                const Acc acc = A.front();
                const double lambda = (acc.t - G[0].t) / (t - G[0].t);
                const double w[3] = {G[0].w[0] + lambda * (x - G[0].w[0]), G[0].w[1] + lambda * (y - G[0].w[1]), G[0].w[2] + lambda * (z - G[0].w[2])};
                track_imu(m, acc.t, w, acc.a);
                A.pop_front();
            }
            if (!A.empty()) while (!G.empty() && G.front().t < t) G.pop_front();
        }
    }
    G.push_back(Gyr{t, {x, y, z}});
}
// detail.cpp:72-101
void track_accelerometer(Manager *m, double t, double x, double y, double z) {
    auto &G = m->gyroscopes;
    auto &A = m->accelerometers;
    if (!G.empty() && t >= G.front().t) {
        if (t > G.back().t) {
            while (G.size() > 1) G.pop_front();
            A.push_back(Acc{t, {x, y, z}});
        } else if (t == G.back().t) {
            while (G.size() > 1) G.pop_front();
            const double a[3] = {x, y, z};
            track_imu(m, t, G.front().w, a);
        } else {
            while (t >= G[1].t) G.pop_front();
            const double lambda = (t - G[0].t) / (G[1].t - G[0].t);
            const double w[3] = {G[0].w[0] + lambda * (G[1].w[0] - G[0].w[0]), G[0].w[1] + lambda * (G[1].w[1] - G[0].w[1]),
                                 G[0].w[2] + lambda * (G[1].w[2] - G[0].w[2])};
            const double a[3] = {x, y, z};
            track_imu(m, t, w, a);
        }
    }
}

// (q, p) o (qx, px): q * qx, p + q * px  (quaternions xyzw)
void compose_pose(const double q[4], const double p[3], const double qx[4], const double px[3], double qo[4], double po[3]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    qo[0] = w * qx[0] + x * qx[3] + y * qx[2] - z * qx[1];
    qo[1] = w * qx[1] - x * qx[2] + y * qx[3] + z * qx[0];
    qo[2] = w * qx[2] + x * qx[1] - y * qx[0] + z * qx[3];
    qo[3] = w * qx[3] - x * qx[0] - y * qx[1] - z * qx[2];
    // rotate px by q: v + 2 w (u x v) + 2 u x (u x v)
    const double cx = y * px[2] - z * px[1], cy = z * px[0] - x * px[2], cz = x * px[1] - y * px[0];
    po[0] = p[0] + px[0] + 2.0 * (w * cx + y * cz - z * cy);
    po[1] = p[1] + px[1] + 2.0 * (w * cy + z * cx - x * cz);
    po[2] = p[2] + px[2] + 2.0 * (w * cz + x * cy - y * cx);
}

}  // namespace

extern "C" {

int XRSLAMCreate(const char *slam_config_path, const char *device_config_path, const char *, const char *, void **config) {
    if (g_mgr) return 1;
    if (!slam_config_path || !device_config_path) return 0;
    const std::string slam = slurp(slam_config_path), dev = slurp(device_config_path);
    if (slam.empty() || dev.empty()) return 0;
    Manager *m = new Manager();
    double v[16];
    int w = 640, h = 480;
    if (numbers_after(dev, "resolution", 2, v)) { w = (int)v[0]; h = (int)v[1]; }
    xrvio_default_params(&m->prm, w, h);
    xrvio_params &p = m->prm;
    if (numbers_after(dev, "intrinsics", 4, v)) { p.fx = v[0]; p.fy = v[1]; p.cx = v[2]; p.cy = v[3]; }
    if (numbers_after(dev, "q_bc", 4, v)) for (int k = 0; k < 4; ++k) p.q_bc[k] = v[k];
    if (numbers_after(dev, "p_bc", 3, v)) for (int k = 0; k < 3; ++k) p.p_bc[k] = v[k];
    if (numbers_after(dev, "q_bi", 4, v)) for (int k = 0; k < 4; ++k) p.q_bi[k] = v[k];
    if (numbers_after(dev, "p_bi", 3, v)) for (int k = 0; k < 3; ++k) p.p_bi[k] = v[k];
    if (numbers_after(dev, "cov_g", 9, v)) for (int k = 0; k < 9; ++k) p.cov_g[k] = v[k];
    if (numbers_after(dev, "cov_a", 9, v)) for (int k = 0; k < 9; ++k) p.cov_a[k] = v[k];
    if (numbers_after(dev, "cov_bg", 9, v)) for (int k = 0; k < 9; ++k) p.cov_bg[k] = v[k];
    if (numbers_after(dev, "cov_ba", 9, v)) for (int k = 0; k < 9; ++k) p.cov_ba[k] = v[k];
    {   // cam0.noise: [sx2, 0, 0, sy2]
        size_t c0 = dev.find("cam0:");
        if (c0 != std::string::npos && numbers_after(dev, "noise", 4, v, dev.find("camera_readout_time", c0))) { p.keypoint_cov[0] = v[0]; p.keypoint_cov[1] = v[3]; }
    }
    {   // Player option --device-undistort: the caller pushes RAW frames and the library runs the reader's undistortion
        // (euroc_dataset_reader.cpp:62-72 radtan / tum_dataset_reader.cpp:62-76 equidistant) as a CUDA kernel before CLAHE
        const char *du = std::getenv("XRVIO_DEVICE_UNDISTORT");
        size_t c0 = dev.find("cam0:");
        if (du && du[0] == '1' && c0 != std::string::npos && numbers_after(dev, "camera_distortion_flag", 1, v, c0) && v[0] != 0 &&
            numbers_after(dev, "distortion", 4, v, c0)) {
            for (int k = 0; k < 4; ++k) p.distortion[k] = v[k];
            const size_t dm = dev.find("distortion_model:", c0);
            const bool equi = dm != std::string::npos && dev.compare(dm + 17, 12, " equidistant") == 0;
            p.distortion_model = equi ? XRVIO_DISTORTION_EQUIDISTANT : XRVIO_DISTORTION_RADTAN;
        }
    }
    if (numbers_after(slam, "q_bo", 4, v)) for (int k = 0; k < 4; ++k) p.q_bo[k] = v[k];
    if (numbers_after(slam, "p_bo", 3, v)) for (int k = 0; k < 3; ++k) p.p_bo[k] = v[k];
    if (numbers_after(slam, "size", 1, v)) p.sliding_window_size = (int)v[0];
    if (numbers_after(slam, "subframe_size", 1, v)) p.subframe_size = (int)v[0];
    if (numbers_after(slam, "force_keyframe_landmarks", 1, v)) p.force_keyframe_landmarks = (int)v[0];
    if (numbers_after(slam, "min_keypoint_distance", 1, v)) p.min_keypoint_distance = v[0];
    if (numbers_after(slam, "max_keypoint_detection", 1, v)) p.max_keypoint_detection = (int)v[0];
    if (numbers_after(slam, "max_init_frames", 1, v)) p.max_init_frames = (int)v[0];
    if (numbers_after(slam, "max_frames", 1, v)) p.max_frames = (int)v[0];
    if (numbers_after(slam, "clahe_clip_limit", 1, v)) p.clahe_clip_limit = v[0];
    { size_t r = slam.find("predict_keypoints:"); if (r != std::string::npos) p.predict_keypoints = slam.compare(r + 18, 6, " false") == 0 ? 0 : 1; }
    if (numbers_after(slam, "clahe_width", 1, v)) p.clahe_width = (int)v[0];
    if (numbers_after(slam, "clahe_height", 1, v)) p.clahe_height = (int)v[0];
    if (numbers_after(slam, "keyframe_num", 1, v)) p.init_keyframe_num = (int)v[0];
    if (numbers_after(slam, "keyframe_gap", 1, v)) p.init_keyframe_gap = (int)v[0];
    if (numbers_after(slam, "min_matches", 1, v)) p.init_min_matches = (int)v[0];
    if (numbers_after(slam, "min_parallax", 1, v)) p.init_min_parallax = v[0];
    if (numbers_after(slam, "min_triangulation", 1, v)) p.init_min_triangulation = (int)v[0];
    if (numbers_after(slam, "min_landmarks", 1, v)) p.init_min_landmarks = (int)v[0];
    { size_t r = slam.find("refine_imu:"); if (r != std::string::npos) p.init_refine_imu = slam.compare(r + 11, 6, " false") == 0 ? 0 : 1; }
    if (numbers_after(slam, "iteration_limit", 1, v)) p.solver_iteration_limit = (int)v[0];
    if (numbers_after(slam, "misalignment_threshold", 1, v)) p.rotation_misalignment_threshold = v[0];
    if (numbers_after(slam, "ransac_threshold", 1, v)) p.rotation_ransac_threshold = v[0];
    if (xrvio_create(&p, 1, 0, &m->ctx) != XRVIO_OK) { delete m; return 0; }
    g_mgr = m;
    if (config) *config = &m->prm;      // NOTE: the reference hands out its YamlConfig*; see INTEGRATION.md
    return 1;
}

void XRSLAMPushSensorData(XRSLAMSensorType sensor_type, void *sensor_data) {
    Manager *m = g_mgr;
    if (!m || !sensor_data) return;
    switch (sensor_type) {
    case XRSLAM_SENSOR_CAMERA: {
        const XRSLAMImage *img = static_cast<const XRSLAMImage *>(sensor_data);
        const int w = m->prm.width, h = m->prm.height;
        // XRSLAMImage carries no width / height: like XRSLAMManager::PushImage (:104-110) the size is the configured camera
        // resolution, so the only thing that can be validated is that a row holds that many pixels
        if (!img->data || (img->channel != 1 && img->channel != 3 && img->channel != 4 ? false : img->stride < w * img->channel)) {
            std::fprintf(stderr, "XRSLAM: image rejected (null data or stride %d < %d pixels x %d channels)\n", img->stride, w, img->channel);
            break;
        }
        std::lock_guard<std::mutex> lk(m->input_mutex);
        m->image.resize((size_t)w * h);
        if (img->channel == 1) {
            for (int y = 0; y < h; ++y) std::memcpy(&m->image[(size_t)y * w], img->data + (size_t)y * img->stride, w);
        } else if (img->channel == 3 || img->channel == 4) {
            // cv::cvtColor(BGR(A)2GRAY): (B*1868 + G*9617 + R*4899 + 8192) >> 14
            for (int y = 0; y < h; ++y)
                for (int x = 0; x < w; ++x) {
                    const unsigned char *px = img->data + (size_t)y * img->stride + (size_t)x * img->channel;
                    m->image[(size_t)y * w + x] = (unsigned char)((px[0] * 1868 + px[1] * 9617 + px[2] * 4899 + 8192) >> 14);
                }
        } else {
            std::fprintf(stderr, "XRSLAM: unsupported image channel count %d\n", img->channel);
            std::exit(-1);          // XRSLAMManager.cpp:126-129
        }
        m->image_t = img->timeStamp;
        m->has_image = true;
    } break;
    case XRSLAM_SENSOR_ACCELERATION: {
        const XRSLAMAcceleration *a = static_cast<const XRSLAMAcceleration *>(sensor_data);
        track_accelerometer(m, a->timestamp, a->data[0], a->data[1], a->data[2]);
    } break;
    case XRSLAM_SENSOR_GYROSCOPE: {
        const XRSLAMGyroscope *g = static_cast<const XRSLAMGyroscope *>(sensor_data);
        track_gyroscope(m, g->timestamp, g->data[0], g->data[1], g->data[2]);
    } break;
    default: break;
    }
}

void XRSLAMRunOneFrame(void) {
    Manager *m = g_mgr;
    if (!m) return;
    std::lock_guard<std::mutex> lk(m->input_mutex);
    if (!m->has_image) return;
    int s = 0;
    const uint8_t *ptr = m->image.data();
    int stride = m->prm.width;
    const int rc = xrvio_push_images(m->ctx, &s, &ptr, &stride, &m->image_t, 1, 0);
    if (rc != XRVIO_OK) std::fprintf(stderr, "XRSLAM: frame at t = %.6f dropped (%d): %s\n", m->image_t, rc, xrvio_last_error());
    m->has_image = false;
}

void XRSLAMGetResult(XRSLAMResultType result_type, void *result_data) {
    Manager *m = g_mgr;
    if (!m || !result_data) return;
    switch (result_type) {
    case XRSLAM_RESULT_BODY_POSE:       // XRSLAMManager::GetResultBodyPose (:155-167): latest pose o imu_to_body
    case XRSLAM_RESULT_CAMERA_POSE: {   // XRSLAMManager::GetResultCameraPose (:169-181): latest pose o camera_to_body
        xrvio_pose p;
        XRSLAMPose *o = static_cast<XRSLAMPose *>(result_data);
        xrvio_get_pose(m->ctx, 0, &p);
        const double *qx = result_type == XRSLAM_RESULT_BODY_POSE ? m->prm.q_bi : m->prm.q_bc;
        const double *px = result_type == XRSLAM_RESULT_BODY_POSE ? m->prm.p_bi : m->prm.p_bc;
        compose_pose(p.q, p.p, qx, px, o->quaternion, o->translation);
        o->timestamp = p.valid ? p.t : 0.0;
    } break;
    case XRSLAM_RESULT_FEATURES:        // XRSLAMManager::GetResultFeatures (:217) has an empty body: the output is left untouched
        break;
    case XRSLAM_RESULT_STATE:
        *static_cast<XRSLAMState *>(result_data) = xrvio_get_state(m->ctx, 0) == XRVIO_STATE_TRACKING ? XRSLAM_STATE_TRACKING_SUCCESS : XRSLAM_STATE_INITIALIZING;
        break;
    case XRSLAM_RESULT_LANDMARKS: {
        XRSLAMLandmarks *o = static_cast<XRSLAMLandmarks *>(result_data);
        const int n = xrvio_get_landmarks(m->ctx, 0, nullptr, 0);
        o->landmarks = n > 0 ? new XRSLAMLandmark[n] : nullptr;     // caller never frees it in the reference either (:207)
        o->num_landmarks = n > 0 ? xrvio_get_landmarks(m->ctx, 0, reinterpret_cast<double *>(o->landmarks), n) : 0;
    } break;
    case XRSLAM_RESULT_BIAS: {
        double t, s16[16];
        XRSLAMIMUBias *o = static_cast<XRSLAMIMUBias *>(result_data);
        if (xrvio_get_latest_state(m->ctx, 0, &t, s16) == XRVIO_OK)
            for (int k = 0; k < 3; ++k) { o->gyr_bias.data[k] = s16[10 + k]; o->acc_bias.data[k] = s16[13 + k]; }
    } break;
    case XRSLAM_RESULT_VERSION: {
        XRSLAMStringOutput *o = static_cast<XRSLAMStringOutput *>(result_data);
        o->str_length = (int)std::strlen(m->version);
        o->data = m->version;
    } break;
    case XRSLAM_INFO_INTRINSICS: {
        XRSLAMIntrinsics *o = static_cast<XRSLAMIntrinsics *>(result_data);
        o->fx = m->prm.fx; o->fy = m->prm.fy; o->cx = m->prm.cx; o->cy = m->prm.cy;
    } break;
    default: break;
    }
}

void XRSLAMDestroy(void) {
    if (!g_mgr) return;
    xrvio_destroy(g_mgr->ctx);
    delete g_mgr;
    g_mgr = nullptr;
}

void XRSLAMExtPushBootstrapState(double t, const double q[4], const double p[3], const double v[3]) {
    if (g_mgr) xrvio_push_bootstrap_state(g_mgr->ctx, 0, t, q, p, v);
}

}  // extern "C"
