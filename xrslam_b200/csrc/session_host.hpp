// Host-side restatement of the reference's per-session bookkeeping (row G of SURVEY.md 8(a)):
// Map / Frame / Track graph as index-friendly C++ objects, sensor sync, keyframe / subframe state
// machine, Poisson ordering, 2-point rotation RANSAC.  No arithmetic of the hot operators lives
// here: images, preintegration, solves, marginalisation and triangulation run as CUDA kernels.
// Follows xrslam/src/xrslam/map/{frame,track,map}.cpp, core/detail.cpp, core/feature_tracker.cpp,
// core/sliding_window_tracker.cpp (cited per function in capi_session.cu).
#pragma once
#include "../../include/xrvio.h"
#include "vmath.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>
#include <memory>
#include <random>
#include <vector>

namespace xvs {
using namespace xv;

struct Imu { double t; V3 w, a; };

struct State16 {
    double d[16];   // q xyzw, p, v, bg, ba
    State16() { std::memset(d, 0, sizeof(d)); d[3] = 1.0; }
    Q4 q() const { return ld4(d); }
    V3 p() const { return ld3(d + 4); }
    V3 v() const { return ld3(d + 7); }
    V3 bg() const { return ld3(d + 10); }
    V3 ba() const { return ld3(d + 13); }
};

struct Map;
struct Track;

struct Frame {
    long long id = 0;
    double t = 0;
    State16 st;
    std::vector<Imu> imu;                  // preintegration.data
    std::vector<double> pre;               // last result of preintegration.integrate (XRVIO_PRE_BLOCK doubles)
    bool pre_full = false;                 // pre holds Jacobians + covariance (flags 3), computed with the bias pre_bias
    double pre_bias[6] = {0, 0, 0, 0, 0, 0};
    std::vector<double> kf_pre;            // keyframe_preintegration result
    std::vector<Imu> kf_imu;
    std::vector<V3> bearings;
    std::vector<Track *> tracks;
    std::vector<std::unique_ptr<Frame>> subframes;
    bool keyframe = false, no_translation = false, fix_pose = false, fix_motion = false;
    Map *map = nullptr;
    int pyr_slot = -1;                      // device pyramid buffer of this frame's image (front-end map only)
    void *img_batch = nullptr;              // the push batch whose image kernels produced that pyramid and the corners (capi_session.cu)
    int img_index = -1;                     // position of this frame in that batch
    int tmp = -1;                           // scratch index used while building problems

    size_t keypoint_num() const { return bearings.size(); }
    void append_keypoint(const V3 &b) { bearings.push_back(b); tracks.push_back(nullptr); }
    std::unique_ptr<Frame> clone() const {      // frame.cpp:20-37
        auto f = std::make_unique<Frame>();
        f->id = id; f->t = t; f->st = st; f->imu = imu; f->pre = pre; f->bearings = bearings;
        f->tracks.assign(bearings.size(), nullptr);
        f->keyframe = keyframe; f->no_translation = no_translation; f->fix_pose = fix_pose; f->fix_motion = fix_motion;
        f->pre_full = pre_full; std::memcpy(f->pre_bias, pre_bias, sizeof(pre_bias));
        f->pyr_slot = pyr_slot;
        return f;
    }
};

struct Track {
    long long id = 0;
    Map *map = nullptr;
    size_t map_index = 0;
    std::vector<std::pair<Frame *, int>> refs;    // sorted by frame id (std::map<Frame*, size_t, compare<Frame*>>)
    bool valid = false, triangulated = false, trash = false, is_static = true;   // track.cpp:7
    double inv_depth = 0.0;
    int life = 0;
    int tmp = -1;

    size_t keypoint_num() const { return refs.size(); }
    Frame *first_frame() const { return refs.front().first; }
    int find(const Frame *f) const {
        for (size_t i = 0; i < refs.size(); ++i) if (refs[i].first == f) return (int)i;
        return -1;
    }
    int get_keypoint_index(const Frame *f) const { int k = find(f); return k < 0 ? -1 : refs[k].second; }
};

struct Rig {     // extrinsics needed by host-side landmark bookkeeping
    Q4 q_bc; V3 p_bc;
};

inline void camera_pose(const Frame *f, const Rig &rig, Q4 &q, V3 &p) {       // frame.cpp:176-181
    q = qmul(f->st.q(), rig.q_bc);
    p = f->st.p() + qrot(f->st.q(), rig.p_bc);
}

struct MargInfo {        // host view of the device-resident prior
    std::vector<Frame *> frames;
    int slot = 0;        // which of the two device buffers currently holds it
    bool present = false;
};

struct Map {
    std::deque<std::unique_ptr<Frame>> frames;
    std::vector<std::unique_ptr<Track>> tracks;
    MargInfo marg;
    long long *track_id_counter = nullptr;

    size_t frame_num() const { return frames.size(); }
    void attach_frame(std::unique_ptr<Frame> f, long pos = -1) {
        f->map = this;
        if (pos < 0) frames.push_back(std::move(f));
        else frames.insert(frames.begin() + pos, std::move(f));
    }
    std::unique_ptr<Frame> detach_frame(size_t index) {
        std::unique_ptr<Frame> f = std::move(frames[index]);
        frames.erase(frames.begin() + index);
        f->map = nullptr;
        return f;
    }
    long frame_index_by_id(long long id) const {
        for (size_t i = 0; i < frames.size(); ++i) if (frames[i]->id == id) return (long)i;
        return -1;
    }
    Track *create_track() {
        auto t = std::make_unique<Track>();
        t->id = ++(*track_id_counter);
        t->map = this;
        t->map_index = tracks.size();
        tracks.push_back(std::move(t));
        return tracks.back().get();
    }
    void recycle_track(Track *t) {     // map.cpp:123-130
        if (t->map_index != tracks.back()->map_index) {
            size_t i = t->map_index;
            tracks[i].swap(tracks.back());
            tracks[i]->map_index = i;
        }
        tracks.pop_back();
    }
};

// landmark <-> inverse depth (track.cpp:91-101)
inline V3 get_landmark_point(const Track *tr, const Rig &rig) {
    Frame *f = tr->refs.front().first;
    Q4 q; V3 p;
    camera_pose(f, rig, q, p);
    return (1.0 / tr->inv_depth) * qrot(q, f->bearings[tr->refs.front().second]) + p;
}
inline void set_landmark_point(Track *tr, const V3 &x, const Rig &rig) {
    Frame *f = tr->refs.front().first;
    Q4 q; V3 p;
    camera_pose(f, rig, q, p);
    tr->inv_depth = 1.0 / norm(qrot(qconj(q), x - p));
}
inline void track_add_keypoint(Track *tr, Frame *f, int idx) {        // track.cpp:14-23
    auto it = std::lower_bound(tr->refs.begin(), tr->refs.end(), f->id,
                               [](const std::pair<Frame *, int> &a, long long id) { return a.first->id < id; });
    if (it != tr->refs.end() && it->first == f) it->second = idx;
    else tr->refs.insert(it, std::make_pair(f, idx));
    f->tracks[idx] = tr;
    if (tr->triangulated) tr->life++;
    else tr->life = 1;
}
inline void track_remove_keypoint(Track *tr, Frame *f, const Rig &rig, bool suicide_if_empty = true) {   // track.cpp:25-44
    int k = tr->find(f);
    int idx = tr->refs[k].second;
    bool was_first = (tr->refs.front().first == f);
    V3 landmark = v3(0, 0, 0);
    if (was_first) landmark = get_landmark_point(tr, rig);
    f->tracks[idx] = nullptr;
    tr->refs.erase(tr->refs.begin() + k);
    if (!tr->refs.empty()) {
        if (was_first) set_landmark_point(tr, landmark, rig);
    } else {
        tr->valid = false;
        if (suicide_if_empty) tr->map->recycle_track(tr);
    }
}
inline Track *frame_get_track(Frame *f, int idx, Map *allocation_map) {    // frame.cpp:45-53 (creating overload)
    if (!allocation_map) allocation_map = f->map;
    if (f->tracks[idx] == nullptr) {
        Track *t = allocation_map->create_track();
        track_add_keypoint(t, f, idx);
    }
    return f->tracks[idx];
}
inline void map_untrack_frame(Map *m, Frame *f, const Rig &rig) {
    (void)m;
    for (size_t i = 0; i < f->keypoint_num(); ++i)
        if (Track *t = f->tracks[i]) track_remove_keypoint(t, f, rig);
}
inline void map_erase_track(Map *m, Track *t, const Rig &rig) {       // map.cpp:100-105
    while (t->keypoint_num() > 0) track_remove_keypoint(t, t->refs.front().first, rig, false);
    m->recycle_track(t);
}

inline void apply_k(const V3 &b, const double K[4], double px[2]) {    // stereo.h:8-10
    px[0] = b.x / b.z * K[0] + K[2];
    px[1] = b.y / b.z * K[1] + K[3];
}
inline V3 remove_k(const double px[2], const double K[4]) {             // stereo.h:12-15
    return normalized(v3((px[0] - K[2]) / K[0], (px[1] - K[3]) / K[1], 1.0));
}

// ---- 2-point rotation RANSAC (geometry/stereo.cpp:71-97, wahba.h:9-27, utility/ransac.h:26-83) ----
inline void jacobi_eig3(double A[3][3], double V[3][3]) {     // symmetric eigen-decomposition, A -> diag
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                double th = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) { double x = A[k][p], y = A[k][q]; A[k][p] = c * x - s * y; A[k][q] = s * x + c * y; }
                for (int k = 0; k < 3; ++k) { double x = A[p][k], y = A[q][k]; A[p][k] = c * x - s * y; A[q][k] = s * x + c * y; }
                for (int k = 0; k < 3; ++k) { double x = V[k][p], y = V[k][q]; V[k][p] = c * x - s * y; V[k][q] = s * x + c * y; }
            }
    }
}
inline M3 solve_rotation_2pt(const V3 a[2], const V3 b[2]) {
    // cov = 0.5 * sum a_i b_i^T = U S V^T ; R = V diag(1,1,det(V U^T)) U^T.  cov has rank <= 2.
    double C[3][3] = {{0}};
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i][j] += 0.5 * a[k][i] * b[k][j];
    // V from eig(C^T C), U columns = C v / sigma for the two leading singular values, third by cross product
    double G[3][3], V[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { G[i][j] = 0; for (int k = 0; k < 3; ++k) G[i][j] += C[k][i] * C[k][j]; }
    jacobi_eig3(G, V);
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int x, int y) { return G[x][x] > G[y][y]; });
    V3 v[3], u[3];
    for (int k = 0; k < 3; ++k) v[k] = v3(V[0][ord[k]], V[1][ord[k]], V[2][ord[k]]);
    for (int k = 0; k < 2; ++k) {
        V3 cv = v3(C[0][0] * v[k].x + C[0][1] * v[k].y + C[0][2] * v[k].z, C[1][0] * v[k].x + C[1][1] * v[k].y + C[1][2] * v[k].z,
                   C[2][0] * v[k].x + C[2][1] * v[k].y + C[2][2] * v[k].z);
        double n = norm(cv);
        u[k] = n > 0 ? (1.0 / n) * cv : v3(k == 0, k == 1, 0);
    }
    u[1] = normalized(u[1] - dot(u[1], u[0]) * u[0]);
    u[2] = cross(u[0], u[1]);
    v[2] = cross(v[0], v[1]);                  // right-handed completion: det(U) = det(V) = +1
    // with det U = det V = 1 the reflection fix E(2,2) = det(V U^T) = 1
    M3 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R(i, j) = v[0][i] * u[0][j] + v[1][i] * u[1][j] + v[2][i] * u[2][j];
    return R;
}

struct LotBox {       // utility/random.h:80-126 over std::default_random_engine
    size_t cap = 0;
    std::vector<size_t> lots;
    std::default_random_engine engine;
    std::uniform_int_distribution<size_t> dist;
    explicit LotBox(size_t n) : lots(n) { for (size_t i = 0; i < n; ++i) lots[i] = i; }
    void seed(unsigned v) { engine.seed(v); }
    void refill_all() { cap = 0; }
    size_t draw_without_replacement() {
        size_t rem = lots.size() - cap;
        if (rem > 1) {
            size_t k = dist(engine, std::uniform_int_distribution<size_t>::param_type(cap, lots.size() - 1));
            std::swap(lots[cap], lots[k]);
            return lots[cap++];
        } else if (rem == 1) { cap++; return lots.back(); }
        return (size_t)-1;
    }
};

inline M3 find_rotation_matrix(const std::vector<V3> &p1, const std::vector<V3> &p2, std::vector<char> &mask, double threshold_rad) {
    const size_t n = p1.size();
    const double thr = 5.99 * threshold_rad * threshold_rad;
    mask.assign(n, 0);
    M3 model = m3_identity();
    if (n < 2) return model;
    LotBox lot(n);
    lot.seed(0);
    const double K = std::log(std::max(1 - 0.999, 1.0e-5));
    size_t best = 0, iter_max = 1000;
    for (size_t it = 0; it < iter_max; ++it) {
        lot.refill_all();
        size_t i0 = lot.draw_without_replacement(), i1 = lot.draw_without_replacement();
        V3 a[2] = {p1[i0], p1[i1]}, b[2] = {p2[i0], p2[i1]};
        M3 R = solve_rotation_2pt(a, b);
        size_t cnt = 0;
        std::vector<char> cur(n, 0);
        for (size_t i = 0; i < n; ++i) {
            double err = std::acos(dot(R * p1[i], p2[i]));
            if (err <= thr) { cnt++; cur[i] = 1; }
        }
        if (cnt > best) {
            model = R; best = cnt; mask.swap(cur);
            double ratio = best / (double)n;
            double N = K / std::log(1 - std::pow(ratio, 5));
            if (N < (double)iter_max) iter_max = (size_t)std::ceil(N);
        }
    }
    return model;
}

}  // namespace xvs
