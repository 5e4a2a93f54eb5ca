// Session layer of the C ABI (include/xrvio.h, layer 2): n_sessions independent VIO sessions are
// advanced together.  The per-session control flow restates, in batched phases,
//   XRSLAM::Detail::track_imu / track_camera / predict_pose        core/detail.cpp:15-28,103-177
//   FeatureTracker::work                                            core/feature_tracker.cpp:24-153
//   Frame::track_keypoints / detect_keypoints                       map/frame.cpp:55-174
//   FrontendWorker::work                                            core/frontend_worker.cpp:28-86
//   SlidingWindowTracker::{mirror_frame,track,localize_newframe,manage_keyframe,track_landmark,
//                          refine_window,slide_window,refine_subwindow}   core/sliding_window_tracker.cpp:31-465
//   Initializer::mirror_keyframe_map + final BA                    core/initializer.cpp:22-76,88-155
// Every numeric operator is a CUDA kernel launch on the context's stream (frontend.cuh,
// backend.cuh); the host only does bookkeeping between launches.  No CPU fallback exists.
#include "../../include/xrvio.h"
#include "backend.cuh"
#include "frontend.cuh"
#include "session_host.hpp"
#include "init_host.hpp"
#include "thread_pool.hpp"
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <thread>
#include <unordered_set>

void xrvio_set_error(const std::string &e);
struct Lane;

using namespace xvs;

#define CKC(call)                                                                                     \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            xrvio_set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                      \
            return XRVIO_ERR_CUDA;                                                                    \
        }                                                                                             \
    } while (0)

namespace {

// pinned host + device pair that grows on demand
template <class T> struct Arena {
    T *h = nullptr, *d = nullptr;
    size_t cap = 0, n = 0;
    ~Arena() { if (h) cudaFreeHost(h); if (d) cudaFree(d); }
    cudaError_t reserve(size_t want) {
        if (want <= cap) return cudaSuccess;
        size_t nc = std::max(want, cap * 2 + 64);
        T *nh = nullptr, *nd = nullptr;
        cudaError_t e = cudaHostAlloc((void **)&nh, nc * sizeof(T), cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        e = cudaMalloc((void **)&nd, nc * sizeof(T));
        if (e != cudaSuccess) return e;
        if (h) { std::memcpy(nh, h, n * sizeof(T)); cudaFreeHost(h); }
        if (d) cudaFree(d);           // device contents are transient per phase
        h = nh; d = nd; cap = nc;
        return cudaSuccess;
    }
    void clear() { n = 0; }
    size_t push(const T &v) { h[n] = v; return n++; }
};

struct DevOnly {
    void *p = nullptr; size_t bytes = 0;
    ~DevOnly() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        bytes = want + want / 4 + 4096;
        return cudaMalloc(&p, bytes);
    }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

// One pinned host block and its device twin per lane.  A back-end phase packs all of its small inputs back to back
// (put), uploads them with ONE copy, and reads its results back with ONE copy of a contiguous output prefix, instead of a
// copy per array.  Every user ends with a stream synchronisation, so the block is free again when the next phase packs.
struct Stage {
    unsigned char *h = nullptr, *d = nullptr;
    size_t cap = 0, off = 0;
    ~Stage() { if (h) cudaFreeHost(h); if (d) cudaFree(d); }
    cudaError_t reserve(size_t want) {
        if (want <= cap) return cudaSuccess;
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        h = d = nullptr; cap = 0;
        const size_t nc = want + want / 2 + 4096;
        cudaError_t e = cudaHostAlloc((void **)&h, nc, cudaHostAllocDefault);
        if (e != cudaSuccess) return e;
        e = cudaMalloc((void **)&d, nc);
        if (e != cudaSuccess) return e;
        cap = nc;
        return cudaSuccess;
    }
    void reset() { off = 0; }
    static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
    // reserve room for `count` elements; returns the device pointer, *host receives the matching host pointer
    template <class T> T *slot(size_t count, T **host) {
        T *dp = reinterpret_cast<T *>(d + off);
        *host = reinterpret_cast<T *>(h + off);
        off += padded(count * sizeof(T));
        return dp;
    }
    template <class T> T *put(const T *src, size_t count) {
        T *hp;
        T *dp = slot<T>(count, &hp);
        if (count) std::memcpy(hp, src, count * sizeof(T));
        return dp;
    }
};

constexpr int kImgQueue = 8;       // frames per session that may wait on the device between steps (asynchronous lanes run up to 6 steps behind)
constexpr int kPyrSlots = kImgQueue + 2;   // pyramid buffers per session: every waiting frame has its pyramid already (built at push time), + previous + spare

// Image kernels run when the images are PUSHED (xrvio_push_images), on the lane's image stream: CLAHE, pyramid and GFTT need
// nothing but the image, so by the time xrvio_step looks at the frame they are done and the step starts at KLT.  A push call
// of a lane is one batch; its corner lists come back into a pinned block of this kind, which lives until the last frame of the
// batch has been consumed (blocks are pooled: pinned allocations are expensive).
struct ImgBatch {
    float *corners_h = nullptr;        // n * kMaxCorners * 3
    int *count_h = nullptr;            // n
    int *idx_h = nullptr;              // 2 * cap: raw-slot indices, pyramid indices (pinned source of the index upload)
    cudaEvent_t done = nullptr;        // recorded on the image stream after the read-back
#ifndef CUEMU
    cudaGraphExec_t gexec = nullptr;   // the image kernels of a push of graph_n images into this block, captured once
#endif
    int graph_n = 0, uses = 0;
    int cap = 0, pending = 0;          // frames of this batch not consumed yet
    bool waited = false;               // the lane stream already waits for `done` / the host has synchronised with it
    bool host_ready = false;
};

// Per-session staging written by worker threads, merged into the batch arenas by the main thread.
struct Staged {
    std::vector<xvbe::PreintJob> pj; std::vector<double> imu;       // imu_off is session-local until merged
    std::vector<xvbe::TriJob> tj; std::vector<double> tv;
    bool has_prob = false;
    xvbe::SolveDesc d{};
    std::vector<double> states, invd, pblocks;
    std::vector<unsigned char> flags;
    std::vector<xvbe::RpeObs> rpe; std::vector<xvbe::RotObs> rot; std::vector<xvbe::PreRef> pref;
    std::vector<int> mframes;
    // marginalisation job
    bool has_marg = false;
    xvbe::MargDesc md{};
    void clear_jobs() { pj.clear(); imu.clear(); tj.clear(); tv.clear(); }
    void clear_prob() { has_prob = false; has_marg = false; states.clear(); invd.clear(); pblocks.clear(); flags.clear(); rpe.clear(); rot.clear(); pref.clear(); mframes.clear(); }
};

struct Session {
    int idx = 0;
    Staged stg;
    int rc = 0; std::string err;
    int pre_base = 0, tri_base = 0;               // global index of this session's first staged job after a merge
    Map ft_map;                                   // FeatureTracker::map
    std::unique_ptr<Map> sw_map;                  // SlidingWindowTracker::map (null while INITIALIZING)
    std::deque<std::unique_ptr<Frame>> queued;    // Detail::frames (waiting for IMU)
    std::deque<std::unique_ptr<Frame>> released;  // handed to FeatureTracker::track_frame
    std::deque<Imu> imus, frontal_imus;
    long long frame_id_counter = 0, track_id_counter = 0;
    // FeatureTracker::latest_state / FrontendWorker::latest_state
    bool has_latest = false; double latest_t = 0; State16 latest;
    bool has_swt_latest = false; double swt_t = 0; long long swt_frame_id = -1; State16 swt_state;
    xrvio_pose out_pose{};
    double out_t = 0;
    std::map<long long, State16> bootstrap;       // keyed by round(t * 1e6)
    xrvio_init_info init{};                       // outcome of the initializer attempts (session_init.inl)
    long long essential_rejected = 0;             // matches rejected by the optional per-frame essential RANSAC
    bool init_now = false;                        // the initializer succeeded in this round
    int img_slot_next = 0;
    int pyr_next = 0;
    // per-round scratch
    std::unique_ptr<Frame> cur;                   // frame being processed by the front end
    Frame *last = nullptr;
    bool fe_has_last = false;
    int pre_job = -1;
    std::vector<double> klt_next;                 // px of tracked points (f64), status
    std::vector<unsigned char> klt_status;
    std::vector<std::pair<int, int>> order;       // (keypoint index, track length) sorted
    std::vector<V3> next_bearings;
    int n_corners = 0;
    // back end round
    int be_mode = 0;                              // 0 none, 1 bootstrap, 2 track
    bool is_keyframe = false;
    Frame *new_i = nullptr, *new_j = nullptr;
    bool mirror_prepended = false;                // mirror_frame merged the IMU samples of skipped frames into new_j
    int prob1 = -1, prob2 = -1, marg_job = -1;
    std::vector<int> sort_cnt;
    std::vector<Frame *> prob_frames;             // frame order of the last built problem
    std::vector<Track *> prob_tracks;
    std::vector<Track *> tri_tracks; int tri_off = 0;
    std::vector<int> pre_jobs;                    // preintegration job ids launched for this session
    std::vector<Frame *> pre_job_frames; std::vector<int> pre_job_kind;
};

}  // namespace

struct Lane {
    xrvio_params prm;
    std::string err;
    int S = 0, device = 0;
    cudaStream_t st = nullptr;
    xvfe::PyrLayout L;
    xvbe::RigParams rig_h;
    xvbe::RigParams *rig_d = nullptr;
    Rig rig;
    double K[4];
    double sic[2];
    std::vector<Session> ses;
    std::unique_ptr<ThreadPool> pool;
    struct StepBuffers { Arena<int> idx_raw, idx_next, idx_prev; } sbuf;
    xrvio_stats stats{};
    // ---- device image / front-end buffers ----
    uint8_t *d_raw = nullptr;          // S * kImgQueue * w*h
    uint8_t *d_pyr = nullptr;          // S * kPyrSlots * L.total
    uint8_t *d_lut = nullptr;
    xvfe::KltMapSet klt_maps{};        // TMA tensor maps over d_pyr, one pair per pyramid level
    uint8_t *d_und = nullptr;          // S * w*h undistorted images of the round (only with prm.distortion_model != 0)
    int2 *d_und_map = nullptr;         // w*h source positions in 1/32 px
    unsigned long long *d_cand = nullptr; int *d_cand_count = nullptr; unsigned *d_maxval = nullptr;
    Arena<float> corners; Arena<int> corner_count;
    Arena<float> klt_prev, klt_next; Arena<int> klt_n; Arena<unsigned char> klt_status;
    Arena<double> poi_pts; Arena<int> poi_npre, poi_ntot; Arena<unsigned char> poi_veto, poi_keep;
    // ---- back-end arenas ----
    Arena<xvbe::PreintJob> pre_jobs; Arena<double> imu; Arena<double> pre_out;
    Arena<xvbe::SolveDesc> sdesc; Arena<double> states; Arena<unsigned char> fflags; Arena<double> invd;
    Arena<xvbe::RpeObs> rpe; Arena<xvbe::RotObs> rot; Arena<xvbe::PreRef> pref; Arena<double> pblocks;
    Arena<int> mframes; Arena<xvbe::SolveOut> sout;
    Arena<xvbe::MargDesc> mdesc; Arena<int> mrank;
    Arena<xvbe::TriJob> tjobs; Arena<double> tviews; Arena<double> tout;
    DevOnly scratch;
    Stage stage;                       // packed inputs / outputs of one back-end launch
    Stage stage_pre;                   // inputs of a preintegration launch the host does not wait for (its copy may still be in flight
                                       // when the solve's inputs are packed into `stage`)
    // Marginalisation runs on its own stream with its own staging block and scratch: its result (the new prior) stays on the
    // device and is first needed by the NEXT window solve of the session, so the host never waits for it (ev_marg_done orders
    // the solves behind it, ev_marg_h2d protects the pinned staging block until its upload has been consumed).
    cudaStream_t st_marg = nullptr;
    bool own_marg_stream = true;
    cudaEvent_t ev_marg_done = nullptr, ev_marg_h2d = nullptr;
    bool marg_pending = false, marg_h2d_pending = false;
    DevOnly scratch_marg;
    Stage stage_marg;
    // ---- image stream (see ImgBatch) ----
    cudaStream_t st_img = nullptr;
    std::vector<ImgBatch *> img_pool;  // free blocks
    std::vector<ImgBatch *> img_all;   // every block ever made (for destruction)
    int *d_idx_img = nullptr;          // 2 * S ints: raw-slot indices, pyramid indices of the batch being processed
    // ---- optional per-kernel-family timing with CUDA events on the launch stream ----
    bool prof = false;
    struct ProfEv { int fam; cudaEvent_t a, b; };
    std::vector<ProfEv> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[16] = {0};
    double phase_ms[32] = {0};         // host wall time per step phase (diagnostics)
    bool dbg_on = false; std::string dbg_line;      // XRVIO_SLOW_STEP_MS: per-step notes printed with a slow step
    long long prof_launches[16] = {0};
    double *d_prior = nullptr;         // S * 2 * prior_stride doubles, device-resident priors
    size_t prior_stride = 0;
    std::vector<double> prior_init_h;
};

namespace {

inline long long tkey(double t) { return (long long)std::llround(t * 1e6); }

struct PhaseTimer {
    Lane *c; int id; std::chrono::steady_clock::time_point t0;
    PhaseTimer(Lane *c_, int id_) : c(c_), id(id_), t0(std::chrono::steady_clock::now()) {}
    ~PhaseTimer() { c->phase_ms[id] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

// consecutive phases of a step: pt.next(id) closes the running phase and opens the next one
struct PhaseClock {
    Lane *c; int id; std::chrono::steady_clock::time_point t0;
    PhaseClock(Lane *c_, int id_) : c(c_), id(id_), t0(std::chrono::steady_clock::now()) {}
    void next(int nid) { stop(); id = nid; t0 = std::chrono::steady_clock::now(); }
    void stop() { if (id >= 0) c->phase_ms[id] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); id = -1; }
    ~PhaseClock() { stop(); }
};

// kernel families for xrvio_get_kernel_times
enum { FAM_CLAHE = 0, FAM_PYRAMID, FAM_GFTT, FAM_KLT, FAM_POISSON, FAM_PREINT, FAM_SOLVE, FAM_MARG, FAM_TRI, FAM_COUNT };
cudaEvent_t prof_event(Lane *c) {
    if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ProfScope {
    Lane *c; int fam; int launches; cudaEvent_t a{}; cudaStream_t st;
    ProfScope(Lane *c_, int fam_, int launches_, cudaStream_t st_ = nullptr) : c(c_), fam(fam_), launches(launches_), st(st_ ? st_ : c_->st) {
        if (c->prof) { a = prof_event(c); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        c->prof_launches[fam] += launches;
        if (c->prof) { cudaEvent_t b = prof_event(c); cudaEventRecord(b, st); c->prof_pending.push_back({fam, a, b}); }
    }
};
void prof_collect(Lane *c) {     // call after the stream has been synchronised; events of the marginalisation stream may still be pending
    std::vector<Lane::ProfEv> later;
    for (auto &e : c->prof_pending) {
        if (cudaEventQuery(e.b) == cudaErrorNotReady) { later.push_back(e); continue; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) c->prof_ms[e.fam] += ms;
        c->prof_pool.push_back(e.a); c->prof_pool.push_back(e.b);
    }
    c->prof_pending.swap(later);
}

// ---------------------------------------------------------------------------------------------
// Detail::predict_pose + propagate_state (detail.cpp:15-28,144-177)
// ---------------------------------------------------------------------------------------------
void predict_pose(Lane *c, Session &s, double t, xrvio_pose &out) {
    out.valid = 0; out.t = t;
    for (int k = 0; k < 4; ++k) out.q[k] = 0;
    for (int k = 0; k < 3; ++k) out.p[k] = 0;
    if (!s.has_latest) return;
    double st = s.latest_t;
    Q4 q = s.latest.q(); V3 p = s.latest.p(), v = s.latest.v();
    const V3 bg = s.latest.bg(), ba = s.latest.ba();
    const V3 g = v3(0, 0, -9.80665);
    while (!s.frontal_imus.empty() && s.frontal_imus.front().t <= st) s.frontal_imus.pop_front();
    for (const Imu &m : s.frontal_imus) {
        if (m.t <= t) {
            const double dt = m.t - st;
            const V3 acc = g + qrot(q, m.a - ba);
            p = p + dt * v + (0.5 * dt * dt) * acc;
            v = v + dt * acc;
            q = qnormalized(qmul(q, expmap(dt * (m.w - bg))));
            st = m.t;
        }
    }
    const Q4 qo = qmul(q, ld4(c->prm.q_bo));
    const V3 po = p + qrot(q, ld3(c->prm.p_bo));
    st4(out.q, qo); st3(out.p, po);
    out.valid = 1;
}

// PreIntegrator::predict (preintegrator.cpp:102-112)
void pre_predict(const std::vector<double> &pb, const State16 &a, State16 &b) {
    const V3 g = v3(0, 0, -9.80665);
    const double dt = pb[xvbe::PB_T];
    const Q4 dq = ld4(&pb[xvbe::PB_Q]);
    const V3 dp = ld3(&pb[xvbe::PB_P]), dv = ld3(&pb[xvbe::PB_V]);
    const V3 v = a.v() + dt * g + qrot(a.q(), dv);
    const V3 p = a.p() + (0.5 * dt * dt) * g + dt * a.v() + qrot(a.q(), dp);
    const Q4 q = qmul(a.q(), dq);
    st4(b.d, q); st3(b.d + 4, p); st3(b.d + 7, v);
    for (int k = 0; k < 6; ++k) b.d[10 + k] = a.d[10 + k];
}

// stage a preintegration job in the session (thread-safe across sessions); returns the session-local index
int add_preint_job(Session &s, const std::vector<Imu> &data, double t_end, const State16 &bias_from, int flags) {
    xvbe::PreintJob j{};
    j.imu_off = (int)(s.stg.imu.size() / 7); j.n = (int)data.size(); j.flags = flags; j.t_end = t_end;
    for (int k = 0; k < 3; ++k) { j.bg[k] = bias_from.d[10 + k]; j.ba[k] = bias_from.d[13 + k]; }
    for (const Imu &m : data) {
        const double o[7] = {m.t, m.w.x, m.w.y, m.w.z, m.a.x, m.a.y, m.a.z};
        s.stg.imu.insert(s.stg.imu.end(), o, o + 7);
    }
    s.stg.pj.push_back(j);
    return (int)s.stg.pj.size() - 1;
}
// merge the staged preintegration jobs of the active sessions into the batch arenas
int merge_preint_jobs(Lane *c, std::vector<Session *> &act) {
    c->pre_jobs.clear(); c->imu.clear();
    size_t nj = 0, ni = 0;
    for (Session *s : act) { nj += s->stg.pj.size(); ni += s->stg.imu.size(); }
    if (c->pre_jobs.reserve(nj + 1) != cudaSuccess || c->imu.reserve(ni + 7) != cudaSuccess) { xrvio_set_error("arena allocation failed"); return XRVIO_ERR_CUDA; }
    for (Session *s : act) {
        s->pre_base = (int)c->pre_jobs.n;
        const int imu_base = (int)(c->imu.n / 7);
        for (xvbe::PreintJob j : s->stg.pj) { j.imu_off += imu_base; c->pre_jobs.push(j); }
        if (!s->stg.imu.empty()) { std::memcpy(c->imu.h + c->imu.n, s->stg.imu.data(), s->stg.imu.size() * sizeof(double)); c->imu.n += s->stg.imu.size(); }
        s->stg.pj.clear(); s->stg.imu.clear();
    }
    return XRVIO_OK;
}

// wait = false: the launch and the copy of the results to the host are only queued; the caller reads pre_out.h after its next
// synchronisation of the lane's stream (the window solve reads the blocks on the device in the meantime, PreRef::block < 0)
int run_preint_jobs(Lane *c, bool wait = true) {
    const int n = (int)c->pre_jobs.n;
    if (n == 0) return XRVIO_OK;
    CKC(c->pre_out.reserve((size_t)n * XRVIO_PRE_BLOCK));
    Stage &g = wait ? c->stage : c->stage_pre;
    CKC(g.reserve(Stage::padded(n * sizeof(xvbe::PreintJob)) + Stage::padded(c->imu.n * sizeof(double)) + 512));
    g.reset();
    const xvbe::PreintJob *d_jobs = g.put(c->pre_jobs.h, (size_t)n);
    const double *d_imu = g.put(c->imu.h, c->imu.n);
    CKC(cudaMemcpyAsync(g.d, g.h, g.off, cudaMemcpyHostToDevice, c->st));
    { ProfScope ps(c, FAM_PREINT, 1); xvbe::preintegrate(d_jobs, n, d_imu, c->rig_d, c->pre_out.d, c->st); }
    c->stats.kernels_launched += 1;
    CKC(cudaMemcpyAsync(c->pre_out.h, c->pre_out.d, (size_t)n * XRVIO_PRE_BLOCK * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (wait) CKC(cudaStreamSynchronize(c->st));
    c->stats.h2d_bytes += n * sizeof(xvbe::PreintJob) + c->imu.n * sizeof(double);
    c->stats.d2h_bytes += (size_t)n * XRVIO_PRE_BLOCK * sizeof(double);
    return XRVIO_OK;
}
void fetch_pre(Lane *c, int job, std::vector<double> &dst) {
    dst.assign(c->pre_out.h + (size_t)job * XRVIO_PRE_BLOCK, c->pre_out.h + (size_t)(job + 1) * XRVIO_PRE_BLOCK);
}

// ---------------------------------------------------------------------------------------------
// Problem builder (what Solver::add_frame_states / add_factor wire into Ceres, solver.cpp:84-174)
// ---------------------------------------------------------------------------------------------
struct Builder {
    Lane *c;
    Session *s;
    std::vector<Frame *> active, consts;
    std::vector<Track *> free_lm, const_lm;
    struct R { Frame *tgt, *ref; Track *tr; bool prior; V3 zt, zr; };
    struct Ro { Frame *tgt, *ref; V3 zt, zr; };
    struct P { Frame *i, *j; bool prior; const std::vector<double> *blk; int dev_job; };   // blk == nullptr: block dev_job of the lane's pre_out, on the device
    std::vector<R> rpe; std::vector<Ro> rot; std::vector<P> pre;
    bool use_marg = false;
    bool with_motion = true;       // false: Solver::add_frame_states(frame, false) -- pose blocks only (initializer.cpp:342)
    void add_active(Frame *f) { if (f->tmp == -1) { f->tmp = -2; active.push_back(f); } }
    void touch(Frame *f) { if (f->tmp == -1) { f->tmp = -3; consts.push_back(f); } }
    // kp: index of the observation among f's keypoints when the caller already has it (it usually iterates over them)
    void add_rpe(Frame *f, Track *tr, bool prior, int kp = -1) {
        Frame *ref = tr->refs.front().first;
        touch(f); touch(ref);
        if (tr->tmp == -1) { tr->tmp = prior ? -3 : -2; (prior ? const_lm : free_lm).push_back(tr); }
        if (kp < 0) kp = tr->get_keypoint_index(f);
        rpe.push_back({f, ref, tr, prior, f->bearings[kp], ref->bearings[tr->refs.front().second]});
    }
    void add_rot(Frame *f, Track *tr) {
        Frame *ref = tr->refs.front().first;
        touch(f); touch(ref);
        rot.push_back({f, ref, f->bearings[tr->get_keypoint_index(f)], ref->bearings[tr->refs.front().second]});
    }
    void add_pre(Frame *i, Frame *j, const std::vector<double> *blk, bool prior) { touch(i); touch(j); pre.push_back({i, j, prior, blk, -1}); }
    void add_pre_dev(Frame *i, Frame *j, int lane_job, bool prior) { touch(i); touch(j); pre.push_back({i, j, prior, nullptr, lane_job}); }
};

// Packs one problem into the session's staging buffers (thread-safe across sessions).  Returns 0 or < 0.
int pack_problem(Lane *c, Session &s, Builder &b, int max_iterations) {
    Staged &g = s.stg;
    g.clear_prob();
    // a frame is active (gets columns) iff it was added with add_frame_states and has a free block
    std::vector<Frame *> &frames = s.prob_frames;
    frames.clear();
    auto all_fixed = [&](const Frame *f) { return f->fix_pose && (f->fix_motion || !b.with_motion); };
    for (Frame *f : b.active) if (!all_fixed(f)) frames.push_back(f);
    const int n_active = (int)frames.size();
    for (Frame *f : b.active) if (all_fixed(f)) frames.push_back(f);
    for (Frame *f : b.consts) frames.push_back(f);
    if (n_active > XRVIO_MAX_ACTIVE || (int)frames.size() > XRVIO_MAX_FRAMES) { s.err = "solve problem exceeds XRVIO_MAX_ACTIVE / XRVIO_MAX_FRAMES"; return XRVIO_ERR_LIMIT; }
    for (size_t i = 0; i < frames.size(); ++i) frames[i]->tmp = (int)i;
    std::vector<Track *> &lms = s.prob_tracks;
    lms = b.free_lm;
    lms.insert(lms.end(), b.const_lm.begin(), b.const_lm.end());
    for (size_t i = 0; i < lms.size(); ++i) lms[i]->tmp = (int)i;
    xvbe::SolveDesc d{};
    d.n_frames = (int)frames.size(); d.n_active = n_active;
    d.n_lm = (int)lms.size(); d.n_lm_free = (int)b.free_lm.size();
    g.states.resize(frames.size() * 16);
    g.flags.resize(frames.size());
    for (size_t i = 0; i < frames.size(); ++i) {
        std::memcpy(&g.states[i * 16], frames[i]->st.d, 16 * sizeof(double));
        g.flags[i] = (unsigned char)((frames[i]->fix_pose ? 0 : 1) | ((frames[i]->fix_motion || !b.with_motion) ? 0 : 2));
    }
    g.invd.resize(lms.size());
    for (size_t i = 0; i < lms.size(); ++i) g.invd[i] = lms[i]->inv_depth;
    // reprojection residuals ordered by (tgt, ref): counting sort over the (<= 96 x 96) key space
    const int nf = (int)frames.size();
    g.rpe.resize(b.rpe.size());
    {
        std::vector<int> &cnt = s.sort_cnt;
        cnt.assign((size_t)nf * nf + 1, 0);
        for (const Builder::R &r : b.rpe) cnt[(size_t)r.tgt->tmp * nf + r.ref->tmp + 1]++;
        for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
        for (const Builder::R &r : b.rpe) {
            xvbe::RpeObs &o = g.rpe[cnt[(size_t)r.tgt->tmp * nf + r.ref->tmp]++];
            o.tgt = (short)r.tgt->tmp; o.ref = (short)r.ref->tmp; o.lm = r.tr->tmp; o.prior = r.prior ? 1 : 0; o.pad = 0;
            st3(o.zt, r.zt); st3(o.zr, r.zr);
        }
    }
    d.n_rpe = (int)g.rpe.size();
    d.n_rot = (int)b.rot.size();
    for (auto &r : b.rot) { xvbe::RotObs o{}; o.tgt = (short)r.tgt->tmp; o.ref = (short)r.ref->tmp; st3(o.zt, r.zt); st3(o.zr, r.zr); g.rot.push_back(o); }
    d.n_pre = (int)b.pre.size();
    g.pblocks.clear();
    for (size_t k = 0; k < b.pre.size(); ++k) {
        const Builder::P &p = b.pre[k];
        xvbe::PreRef r{};
        r.i = (short)p.i->tmp; r.j = (short)p.j->tmp; r.prior = p.prior ? 1 : 0;
        if (p.blk) {
            r.block = (int)(g.pblocks.size() / XRVIO_PRE_BLOCK);
            g.pblocks.insert(g.pblocks.end(), p.blk->begin(), p.blk->begin() + XRVIO_PRE_BLOCK);
        } else r.block = ~p.dev_job;                                 // stays on the device (SolveBatch::pre_out)
        g.pref.push_back(r);
    }
    if (b.use_marg) {
        Map *m = s.sw_map.get();
        d.marg_k = (int)m->marg.frames.size();
        for (Frame *f : m->marg.frames) g.mframes.push_back(f->tmp);
        d.marg_off = (long long)(((size_t)s.idx * 2 + m->marg.slot) * c->prior_stride);
    }
    d.max_iterations = max_iterations;
    for (Frame *f : frames) f->tmp = -1;
    for (Track *t : lms) t->tmp = -1;
    g.d = d;
    g.has_prob = true;
    return 0;
}

// merge the staged problems of the active sessions into the batch arenas; sets s.probN (global index or -1)
int merge_problems(Lane *c, std::vector<Session *> &act, int which) {
    size_t nst = 0, nl = 0, nr = 0, nro = 0, np = 0, npb = 0, nmf = 0, nprob = 0;
    for (Session *s : act) {
        const Staged &g = s->stg;
        if (!g.has_prob) continue;
        nprob++; nst += g.states.size(); nl += g.invd.size(); nr += g.rpe.size(); nro += g.rot.size(); np += g.pref.size(); npb += g.pblocks.size(); nmf += g.mframes.size();
    }
    c->sdesc.clear(); c->states.clear(); c->fflags.clear(); c->invd.clear(); c->rpe.clear(); c->rot.clear(); c->pref.clear(); c->pblocks.clear(); c->mframes.clear();
    if (c->sdesc.reserve(nprob + 1) != cudaSuccess || c->states.reserve(nst + 16) != cudaSuccess || c->fflags.reserve(nst / 16 + 1) != cudaSuccess ||
        c->invd.reserve(nl + 1) != cudaSuccess || c->rpe.reserve(nr + 1) != cudaSuccess || c->rot.reserve(nro + 1) != cudaSuccess ||
        c->pref.reserve(np + 1) != cudaSuccess || c->pblocks.reserve(npb + XRVIO_PRE_BLOCK) != cudaSuccess || c->mframes.reserve(nmf + 1) != cudaSuccess) {
        xrvio_set_error("arena allocation failed"); return XRVIO_ERR_CUDA;
    }
    for (Session *sp : act) {
        Session &s = *sp;
        Staged &g = s.stg;
        int &slot = which == 1 ? s.prob1 : s.prob2;
        slot = -1;
        if (!g.has_prob) continue;
        xvbe::SolveDesc d = g.d;
        d.frame_off = (int)(c->states.n / 16); d.lm_off = (int)c->invd.n; d.rpe_off = (int)c->rpe.n; d.rot_off = (int)c->rot.n;
        d.pre_off = (int)c->pref.n; d.marg_frames_off = (int)c->mframes.n;
        const int blk_base = (int)(c->pblocks.n / XRVIO_PRE_BLOCK);
        auto app = [](auto &arena, const auto &v) { if (!v.empty()) { std::memcpy(arena.h + arena.n, v.data(), v.size() * sizeof(v[0])); arena.n += v.size(); } };
        app(c->states, g.states); app(c->fflags, g.flags); app(c->invd, g.invd); app(c->rpe, g.rpe); app(c->rot, g.rot);
        for (xvbe::PreRef r : g.pref) { if (r.block >= 0) r.block += blk_base; c->pref.push(r); }
        app(c->pblocks, g.pblocks); app(c->mframes, g.mframes);
        slot = (int)c->sdesc.push(d);
        g.has_prob = false;
    }
    return XRVIO_OK;
}

void clear_solve_arenas(Lane *c) {
    c->sdesc.clear(); c->states.clear(); c->fflags.clear(); c->invd.clear(); c->rpe.clear(); c->rot.clear();
    c->pref.clear(); c->pblocks.clear(); c->mframes.clear();
}

int run_solves(Lane *c) {
    const int n = (int)c->sdesc.n;
    if (n == 0) return XRVIO_OK;
    size_t scratch = 0;
    for (int p = 0; p < n; ++p) {
        xvbe::SolveDesc &d = c->sdesc.h[p];
        d.scratch_off = (long long)scratch;
        scratch += xvbe::solve_scratch_doubles(d.n_active, d.n_lm_free, d.n_pre, d.marg_k, d.n_rpe);
    }
    CKC(c->scratch.reserve(scratch * sizeof(double)));
    CKC(c->sout.reserve(n));
    // one upload: [states | inv_depth | summaries] (also the download, they are updated in place) then the read-only inputs
    Stage &g = c->stage;
    const size_t in_bytes = c->states.n * 8 + c->invd.n * 8 + n * sizeof(xvbe::SolveOut) + n * sizeof(xvbe::SolveDesc) + c->fflags.n + c->rpe.n * sizeof(xvbe::RpeObs) +
                            c->rot.n * sizeof(xvbe::RotObs) + c->pref.n * sizeof(xvbe::PreRef) + c->pblocks.n * 8 + c->mframes.n * sizeof(int);
    CKC(g.reserve(in_bytes + 16 * 256));
    g.reset();
    xvbe::SolveBatch b;
    b.states = g.put(c->states.h, c->states.n);
    b.inv_depth = g.put(c->invd.h, c->invd.n);
    xvbe::SolveOut *h_out;
    b.out = g.slot<xvbe::SolveOut>((size_t)n, &h_out);
    const size_t out_bytes = g.off;
    b.desc = g.put(c->sdesc.h, (size_t)n);
    b.frame_flags = g.put(c->fflags.h, c->fflags.n);
    b.rpe = g.put(c->rpe.h, c->rpe.n);
    b.rot = g.put(c->rot.h, c->rot.n);
    b.pre = g.put(c->pref.h, c->pref.n);
    b.pre_blocks = g.put(c->pblocks.h, c->pblocks.n);
    b.marg_frames = g.put(c->mframes.h, c->mframes.n);
    b.marg = c->d_prior; b.scratch = c->scratch.as<double>();
    b.pre_out = c->pre_out.d;
    CKC(cudaMemcpyAsync(g.d, g.h, g.off, cudaMemcpyHostToDevice, c->st));
    c->stats.h2d_bytes += in_bytes - n * sizeof(xvbe::SolveOut);
    if (c->marg_pending) { CKC(cudaStreamWaitEvent(c->st, c->ev_marg_done, 0)); c->marg_pending = false; }   // priors written by the marginalisation stream
    { ProfScope ps(c, FAM_SOLVE, 1); xvbe::solve(b, n, c->rig_d, c->st); }
    c->stats.kernels_launched += 1;
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(g.h, g.d, out_bytes, cudaMemcpyDeviceToHost, c->st));
    CKC(cudaStreamSynchronize(c->st));
    // the rest of the step reads the results from the arenas' host arrays
    {
        const unsigned char *hp = g.h;
        std::memcpy(c->states.h, hp, c->states.n * sizeof(double)); hp += Stage::padded(c->states.n * sizeof(double));
        if (c->invd.n) std::memcpy(c->invd.h, hp, c->invd.n * sizeof(double));
        hp += Stage::padded(c->invd.n * sizeof(double));
        std::memcpy(c->sout.h, hp, n * sizeof(xvbe::SolveOut));
    }
    c->stats.d2h_bytes += c->states.n * sizeof(double) + c->invd.n * sizeof(double) + n * sizeof(xvbe::SolveOut);
    return XRVIO_OK;
}

void writeback(Lane *c, Session &s, int prob) {
    const xvbe::SolveDesc &d = c->sdesc.h[prob];
    for (int i = 0; i < d.n_active; ++i) std::memcpy(s.prob_frames[i]->st.d, c->states.h + (size_t)(d.frame_off + i) * 16, 16 * sizeof(double));
    for (int k = 0; k < d.n_lm_free; ++k) s.prob_tracks[k]->inv_depth = c->invd.h[d.lm_off + k];
}

}  // namespace

// =================================================================================================
// per-lane implementations of the session API

int xrvio_default_params_impl(xrvio_params *p, int width, int height) {
    if (!p || width <= 0 || height <= 0) { xrvio_set_error("xrvio_default_params: bad argument"); return XRVIO_ERR_ARG; }
    std::memset(p, 0, sizeof(*p));
    p->width = width; p->height = height;
    p->fx = 458.654; p->fy = 457.296; p->cx = width / 2.0; p->cy = height / 2.0;      // configs/euroc_sensor.yaml:43
    const double qbc[4] = {-7.7071797555374275e-03, 1.0499323370587278e-02, 7.0175280029197162e-01, 7.1230146066895372e-01};
    const double pbc[3] = {-0.0216401454975, -0.064676986768, 0.00981073058949};
    for (int k = 0; k < 4; ++k) { p->q_bc[k] = qbc[k]; p->q_bi[k] = (k == 3); p->q_bo[k] = (k == 3); }
    for (int k = 0; k < 3; ++k) p->p_bc[k] = pbc[k];
    for (int k = 0; k < 3; ++k) {
        p->cov_g[4 * k] = 2.8791302399999997e-08; p->cov_a[4 * k] = 4.0e-6; p->cov_bg[4 * k] = 3.7608844899999997e-10; p->cov_ba[4 * k] = 9.0e-6;
    }
    p->keypoint_cov[0] = p->keypoint_cov[1] = 0.5;
    p->sliding_window_size = 10; p->subframe_size = 3; p->force_keyframe_landmarks = 35;     // configs/euroc_slam.yaml:6-9
    p->min_keypoint_distance = 20.0; p->max_keypoint_detection = 200; p->max_init_frames = 60; p->max_frames = 20;
    p->predict_keypoints = 1; p->clahe_clip_limit = 6.0; p->clahe_width = 8; p->clahe_height = 8;
    p->init_keyframe_num = 8; p->init_keyframe_gap = 5; p->solver_iteration_limit = 30;
    p->rotation_misalignment_threshold = 0.02; p->rotation_ransac_threshold = 10.0;
    p->initializer = 1;                                                                       // configs/euroc_slam.yaml:21-28
    p->init_min_matches = 50; p->init_min_parallax = 10.0; p->init_min_triangulation = 20; p->init_min_landmarks = 30;
    p->init_refine_imu = 1; p->random_seed = 648;                                             // config.cpp:66
    return XRVIO_OK;
}

static int lane_destroy(Lane *c);

static int lane_create(const xrvio_params *p, int n_sessions, int device, int host_threads, Lane **out) {
    if (!p || !out || n_sessions <= 0) { xrvio_set_error("xrvio_create: bad argument"); return XRVIO_ERR_ARG; }
    if (p->sliding_window_size + 2 > XRVIO_MAX_ACTIVE || p->max_keypoint_detection > XRVIO_MAX_CORNERS) { xrvio_set_error("xrvio_create: window / detection limit exceeded"); return XRVIO_ERR_LIMIT; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { xrvio_set_error("xrvio_create: no CUDA device (there is no CPU fallback)"); return XRVIO_ERR_CUDA; }
    CKC(cudaSetDevice(device));
    Lane *c = new Lane();
    struct Guard { Lane *l; ~Guard() { if (l) lane_destroy(l); } } guard{c};      // releases everything on an early error return
    c->prm = *p; c->S = n_sessions; c->device = device;
    // The lane's main stream carries the latency chain of the step (single-CTA solves, one-warp filters): highest priority, so that
    // its few CTAs are placed ahead of the big grids of the image kernels (image stream) and the marginalisation of other lanes.
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    CKC(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, prio_hi));
    CKC(cudaStreamCreateWithPriority(&c->st_img, cudaStreamNonBlocking, prio_lo));
    // XRVIO_SHARE_LOW_STREAM=1: marginalisation shares the image stream (two streams per lane, so that 16 lanes still get a
    // hardware queue per stream out of the 32 available); default: its own stream
    { const char *e = std::getenv("XRVIO_SHARE_LOW_STREAM"); c->own_marg_stream = !(e && e[0] == '1'); }
    if (c->own_marg_stream) CKC(cudaStreamCreateWithPriority(&c->st_marg, cudaStreamNonBlocking, prio_lo));
    else c->st_marg = c->st_img;
    CKC(cudaMalloc((void **)&c->d_idx_img, (size_t)2 * n_sessions * sizeof(int)));
    CKC(cudaEventCreateWithFlags(&c->ev_marg_done, cudaEventDisableTiming));
    CKC(cudaEventCreateWithFlags(&c->ev_marg_h2d, cudaEventDisableTiming));
    c->L = xvfe::make_pyr_layout(p->width, p->height);
    c->K[0] = p->fx; c->K[1] = p->fy; c->K[2] = p->cx; c->K[3] = p->cy;
    c->sic[0] = p->fx / std::sqrt(p->keypoint_cov[0]); c->sic[1] = p->fy / std::sqrt(p->keypoint_cov[1]);      // detail.cpp:107-109
    xvbe::RigParams &r = c->rig_h;
    for (int k = 0; k < 4; ++k) { r.cam_q[k] = p->q_bc[k]; r.imu_q[k] = p->q_bi[k]; }
    for (int k = 0; k < 3; ++k) { r.cam_p[k] = p->p_bc[k]; r.imu_p[k] = p->p_bi[k]; }
    r.sic[0] = c->sic[0]; r.sic[1] = c->sic[1];
    for (int k = 0; k < 9; ++k) { r.cov_w[k] = p->cov_g[k]; r.cov_a[k] = p->cov_a[k]; r.cov_bg[k] = p->cov_bg[k]; r.cov_ba[k] = p->cov_ba[k]; }
    c->rig.q_bc = ld4(p->q_bc); c->rig.p_bc = ld3(p->p_bc);
    CKC(cudaMalloc((void **)&c->rig_d, sizeof(xvbe::RigParams)));
    CKC(cudaMemcpy(c->rig_d, &r, sizeof(r), cudaMemcpyHostToDevice));
    const size_t P = (size_t)p->width * p->height;
    const int S = n_sessions;
    CKC(cudaMalloc((void **)&c->d_raw, (size_t)S * kImgQueue * P));
    CKC(cudaMalloc((void **)&c->d_pyr, (size_t)S * kPyrSlots * c->L.total));
    CKC(cudaMalloc((void **)&c->d_lut, (size_t)S * p->clahe_width * p->clahe_height * 256));
    c->klt_maps = xvfe::klt_make_maps(c->d_pyr, S * kPyrSlots, c->L);
    if (p->distortion_model != XRVIO_DISTORTION_NONE) {
        if (p->distortion_model != XRVIO_DISTORTION_RADTAN && p->distortion_model != XRVIO_DISTORTION_EQUIDISTANT) { xrvio_set_error("xrvio_create: unknown distortion_model"); return XRVIO_ERR_ARG; }
        CKC(cudaMalloc((void **)&c->d_und, (size_t)S * P));
        CKC(cudaMalloc((void **)&c->d_und_map, P * sizeof(int2)));
        xvfe::undistort_build_map(xund::make_model(p->distortion_model, c->K, p->distortion), p->width, p->height, c->d_und_map, c->st);
        CKC(cudaGetLastError());
    }
    CKC(cudaMalloc((void **)&c->d_cand, (size_t)S * xvfe::kMaxCand * 8));
    CKC(cudaMalloc((void **)&c->d_cand_count, S * sizeof(int)));
    CKC(cudaMalloc((void **)&c->d_maxval, S * sizeof(unsigned)));
    CKC(c->corners.reserve((size_t)S * xvfe::kMaxCorners * 3)); CKC(c->corner_count.reserve(S));
    CKC(c->klt_prev.reserve((size_t)S * xvfe::kMaxPts * 2)); CKC(c->klt_next.reserve((size_t)S * xvfe::kMaxPts * 2));
    CKC(c->klt_n.reserve(S)); CKC(c->klt_status.reserve((size_t)S * xvfe::kMaxPts));
    CKC(c->poi_pts.reserve((size_t)S * xvfe::kMaxPts * 2)); CKC(c->poi_npre.reserve(S)); CKC(c->poi_ntot.reserve(S));
    CKC(c->poi_veto.reserve((size_t)S * xvfe::kMaxPts)); CKC(c->poi_keep.reserve((size_t)S * xvfe::kMaxPts));
    const size_t n = (size_t)15 * (p->sliding_window_size + 1);      // the window may briefly hold size + 2 keyframes
    c->prior_stride = 16 * (p->sliding_window_size + 1) + 2 * n * n + n + 64;
    CKC(cudaMalloc((void **)&c->d_prior, (size_t)S * 2 * c->prior_stride * sizeof(double)));
    CKC(cudaMemset(c->d_prior, 0, (size_t)S * 2 * c->prior_stride * sizeof(double)));
    // reserve the batch arenas up front: growing them later means cudaFree / cudaHostAlloc, which stall every lane
    CKC(c->pre_jobs.reserve((size_t)S * 16)); CKC(c->imu.reserve((size_t)S * 16 * 64 * 7)); CKC(c->pre_out.reserve((size_t)S * 16 * XRVIO_PRE_BLOCK));
    CKC(c->sdesc.reserve(S)); CKC(c->states.reserve((size_t)S * XRVIO_MAX_FRAMES * 16)); CKC(c->fflags.reserve((size_t)S * XRVIO_MAX_FRAMES));
    CKC(c->invd.reserve((size_t)S * 1024)); CKC(c->rpe.reserve((size_t)S * 4096)); CKC(c->rot.reserve((size_t)S * 256));
    CKC(c->pref.reserve((size_t)S * 16)); CKC(c->pblocks.reserve((size_t)S * 16 * XRVIO_PRE_BLOCK)); CKC(c->mframes.reserve((size_t)S * 16));
    CKC(c->sout.reserve(S)); CKC(c->mdesc.reserve(S)); CKC(c->mrank.reserve(S));
    CKC(c->tjobs.reserve((size_t)S * 512)); CKC(c->tviews.reserve((size_t)S * 256 * 8 * 16)); CKC(c->tout.reserve((size_t)S * 512 * 4));
    CKC(c->scratch.reserve((size_t)S * (xvbe::solve_scratch_doubles(XRVIO_MAX_ACTIVE, 768, 12, XRVIO_MAX_ACTIVE, 4096) + (size_t)2 * 180 * 180 + 8192) * sizeof(double)));
    CKC(c->sbuf.idx_raw.reserve(S)); CKC(c->sbuf.idx_next.reserve(S)); CKC(c->sbuf.idx_prev.reserve(S));
    // the marginalisation stream's own scratch and staging block, sized for a full batch up front (growing them later costs a
    // cudaFree, i.e. a device-wide synchronisation that stalls every lane of the context)
    CKC(c->scratch_marg.reserve((size_t)S * xvbe::marg_scratch_doubles(XRVIO_MAX_ACTIVE, 768, 2048) * sizeof(double)));
    CKC(c->stage_marg.reserve((size_t)S * (sizeof(xvbe::MargDesc) + XRVIO_MAX_ACTIVE * 16 * 8 + 768 * 8 + 2048 * sizeof(xvbe::RpeObs) + XRVIO_PRE_BLOCK * 8 + 1024)));
    CKC(c->stage.reserve((size_t)S * (XRVIO_MAX_FRAMES * 16 * 8 + 1024 * 8 + 4096 * sizeof(xvbe::RpeObs) + 16 * XRVIO_PRE_BLOCK * 8 + 16 * 64 * 7 * 8 + 65536)));
    c->ses.resize(S);
    {
        c->pool = std::make_unique<ThreadPool>(std::max(1, std::min(host_threads, S)));
    }
    for (int i = 0; i < S; ++i) { c->ses[i].idx = i; c->ses[i].ft_map.track_id_counter = &c->ses[i].track_id_counter; }
    guard.l = nullptr;
    *out = c;
    return XRVIO_OK;
}

static int lane_destroy(Lane *c) {
    if (!c) return XRVIO_OK;
    cudaSetDevice(c->device);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->st_marg) { cudaStreamSynchronize(c->st_marg); if (c->own_marg_stream) cudaStreamDestroy(c->st_marg); }
    if (c->st_img) { cudaStreamSynchronize(c->st_img); cudaStreamDestroy(c->st_img); }
#ifndef CUEMU
    for (ImgBatch *b : c->img_all) if (b->gexec) cudaGraphExecDestroy(b->gexec);
#endif
    for (ImgBatch *b : c->img_all) { cudaFreeHost(b->corners_h); cudaFreeHost(b->count_h); cudaFreeHost(b->idx_h); if (b->done) cudaEventDestroy(b->done); delete b; }
    cudaFree(c->d_idx_img);
    if (c->ev_marg_done) cudaEventDestroy(c->ev_marg_done);
    if (c->ev_marg_h2d) cudaEventDestroy(c->ev_marg_h2d);
    cudaFree(c->rig_d); cudaFree(c->d_raw); cudaFree(c->d_pyr); 
    cudaFree(c->d_und); cudaFree(c->d_und_map);
    cudaFree(c->d_lut); cudaFree(c->d_cand); cudaFree(c->d_cand_count); cudaFree(c->d_maxval); cudaFree(c->d_prior);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
    return XRVIO_OK;
}

static int lane_push_bootstrap_state(Lane *c, int session, double t, const double q[4], const double p[3], const double v[3]) {
    if (!c || session < 0 || session >= c->S) { xrvio_set_error("bad session"); return XRVIO_ERR_ARG; }
    State16 s;
    for (int k = 0; k < 4; ++k) s.d[k] = q[k];
    for (int k = 0; k < 3; ++k) { s.d[4 + k] = p[k]; s.d[7 + k] = v[k]; }
    c->ses[session].bootstrap[tkey(t)] = s;
    return XRVIO_OK;
}

// Detail::track_imu (detail.cpp:130-142): IMU samples up to the image time belong to the frame;
// the first later sample releases the frame to the feature tracker.
static int lane_push_imu(Lane *c, const int *session, const xrvio_imu *samples, int n) {
    if (!c || !session || !samples || n < 0) { xrvio_set_error("xrvio_push_imu: bad argument"); return XRVIO_ERR_ARG; }
    for (int i = 0; i < n; ++i) {
        if (session[i] < 0 || session[i] >= c->S) { xrvio_set_error("bad session"); return XRVIO_ERR_ARG; }
        Session &s = c->ses[session[i]];
        Imu m{samples[i].t, ld3(samples[i].w), ld3(samples[i].a)};
        s.frontal_imus.push_back(m);
        s.imus.push_back(m);
        while (!s.imus.empty() && !s.queued.empty()) {
            if (s.imus.front().t <= s.queued.front()->t) { s.queued.front()->imu.push_back(s.imus.front()); s.imus.pop_front(); }
            else { s.released.push_back(std::move(s.queued.front())); s.queued.pop_front(); }
        }
    }
    return XRVIO_OK;
}

static ImgBatch *img_batch_get(Lane *c) {
    if (!c->img_pool.empty()) { ImgBatch *b = c->img_pool.back(); c->img_pool.pop_back(); return b; }
    ImgBatch *b = new ImgBatch();
    b->cap = c->S;
    if (cudaHostAlloc((void **)&b->corners_h, (size_t)c->S * xvfe::kMaxCorners * 3 * sizeof(float), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&b->count_h, (size_t)c->S * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaHostAlloc((void **)&b->idx_h, (size_t)2 * c->S * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaEventCreateWithFlags(&b->done, cudaEventDisableTiming) != cudaSuccess) {
        if (b->corners_h) cudaFreeHost(b->corners_h);
        if (b->count_h) cudaFreeHost(b->count_h);
        if (b->idx_h) cudaFreeHost(b->idx_h);
        delete b;
        return nullptr;
    }
    c->img_all.push_back(b);
    return b;
}
static void img_batch_release(Lane *c, Frame *f) {
    ImgBatch *b = static_cast<ImgBatch *>(f->img_batch);
    f->img_batch = nullptr;
    if (b && --b->pending == 0) c->img_pool.push_back(b);
}

// Detail::track_camera (detail.cpp:103-128) for the n images of one push call that belong to this lane, followed by
// Image::preprocess (CLAHE + pyramid, opencv_image.cpp:156-161) and the detection half of Image::detect_keypoints (GFTT,
// :38-45) on the image stream -- they depend on the image alone (see ImgBatch).
static int lane_push_images(Lane *c, const int *session, const uint8_t *const *data, const int *stride, const double *t, int n, int on_device) {
    if (!c || !session || !data || !t || n < 0) { xrvio_set_error("xrvio_push_images: bad argument"); return XRVIO_ERR_ARG; }
    if (n == 0) return XRVIO_OK;
    if (n > c->S) { xrvio_set_error("xrvio_push_images: more images than sessions in one call"); return XRVIO_ERR_ARG; }
    CKC(cudaSetDevice(c->device));
    PhaseTimer pt_(c, 30);
    const int w = c->prm.width, h = c->prm.height;
    const size_t P = (size_t)w * h;
    for (int i = 0; i < n; ++i) {
        if (session[i] < 0 || session[i] >= c->S) { xrvio_set_error("bad session"); return XRVIO_ERR_ARG; }
        for (int j = 0; j < i; ++j)
            if (session[j] == session[i]) { xrvio_set_error("xrvio_push_images: a session appears twice in one call"); return XRVIO_ERR_ARG; }
        Session &s = c->ses[session[i]];
        if (s.queued.size() + s.released.size() >= (size_t)kImgQueue) { xrvio_set_error("xrvio_push_images: image queue full, call xrvio_step"); return XRVIO_ERR_STATE; }
    }
    ImgBatch *b = img_batch_get(c);
    if (!b) { xrvio_set_error("xrvio_push_images: pinned allocation failed"); return XRVIO_ERR_CUDA; }
    b->pending = n; b->waited = false; b->host_ready = false;
    for (int i = 0; i < n; ++i) {
        Session &s = c->ses[session[i]];
        const int slot = s.img_slot_next; s.img_slot_next = (s.img_slot_next + 1) % kImgQueue;
        uint8_t *dst = c->d_raw + ((size_t)s.idx * kImgQueue + slot) * P;
        const int sp = stride ? stride[i] : w;
        CKC(cudaMemcpy2DAsync(dst, w, data[i], sp, w, h, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->st_img));
        if (!on_device) c->stats.h2d_bytes += P;
        auto f = std::make_unique<Frame>();
        f->id = ++s.frame_id_counter; f->t = t[i];
        f->pyr_slot = s.pyr_next; s.pyr_next = (s.pyr_next + 1) % kPyrSlots;
        f->img_batch = b; f->img_index = i;
        b->idx_h[i] = s.idx * kImgQueue + slot;
        b->idx_h[c->S + i] = s.idx * kPyrSlots + f->pyr_slot;
        s.queued.push_back(std::move(f));
        xrvio_pose op;
        predict_pose(c, s, t[i], op);
        if (t[i] > s.out_t) { s.out_pose = op; s.out_t = t[i]; }
    }
    int *d_raw_idx = c->d_idx_img, *d_pyr_idx = c->d_idx_img + c->S;
    // index upload, (undistortion,) CLAHE, pyramid, GFTT and the read-back of the corner lists: 12-13 stream operations whose
    // arguments depend only on the batch block and the image count -- captured into a CUDA graph the second time a block sees
    // the same count and replayed with ONE launch afterwards (the raw image copies above come from caller pointers and stay
    // outside).  With per-family profiling on (xrvio_set_profiling) the operations are issued one by one so that the events
    // of xrvio_get_kernel_times keep their meaning.
    auto issue = [&]() -> int {
        CKC(cudaMemcpyAsync(c->d_idx_img, b->idx_h, (size_t)2 * c->S * sizeof(int), cudaMemcpyHostToDevice, c->st_img));
        const uint8_t *clahe_src = c->d_raw;
        const int *clahe_idx = d_raw_idx;
        if (c->d_und) {     // the reader's cv::undistort, on the device (row N2); counted with the CLAHE family
            ProfScope ps(c, FAM_CLAHE, 1, c->st_img);
            xvfe::undistort(c->d_raw, d_raw_idx, P, c->d_und, nullptr, P, n, w, h, c->d_und_map, c->st_img);
            clahe_src = c->d_und; clahe_idx = nullptr;
        }
        { ProfScope ps(c, FAM_CLAHE, 2, c->st_img);
          xvfe::clahe(clahe_src, clahe_idx, P, c->d_pyr, d_pyr_idx, c->L.total, c->d_lut, n, w, h, c->prm.clahe_clip_limit,
                      c->prm.clahe_width, c->prm.clahe_height, c->st_img); }
        { ProfScope ps(c, FAM_PYRAMID, 3, c->st_img); xvfe::pyramid(c->d_pyr, d_pyr_idx, n, c->L, c->st_img); }
        { ProfScope ps(c, FAM_GFTT, 2, c->st_img);
          xvfe::gftt_harris(c->d_pyr, d_pyr_idx, c->L.total, n, w, h, c->prm.max_keypoint_detection, 1.0e-3f, 20.0f, c->corners.d,
                            c->corner_count.d, c->d_cand, c->d_cand_count, c->d_maxval, c->st_img); }
        CKC(cudaMemcpyAsync(b->corners_h, c->corners.d, (size_t)n * xvfe::kMaxCorners * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->st_img));
        CKC(cudaMemcpyAsync(b->count_h, c->corner_count.d, n * sizeof(int), cudaMemcpyDeviceToHost, c->st_img));
        return XRVIO_OK;
    };
    bool launched = false;
#ifndef CUEMU
    static const bool graphs_on = [] { const char *e = std::getenv("XRVIO_IMG_GRAPH"); return !(e && e[0] == '0'); }();
    if (graphs_on && !c->prof && b->uses >= 1) {
        if (b->gexec == nullptr || b->graph_n != n) {
            if (b->gexec) { cudaGraphExecDestroy(b->gexec); b->gexec = nullptr; }
            cudaGraph_t graph = nullptr;
            CKC(cudaStreamBeginCapture(c->st_img, cudaStreamCaptureModeThreadLocal));
            const int rc = issue();
            const cudaError_t ce = cudaStreamEndCapture(c->st_img, &graph);
            if (rc != XRVIO_OK || ce != cudaSuccess || graph == nullptr) { if (graph) cudaGraphDestroy(graph); xrvio_set_error("xrvio_push_images: graph capture failed"); return XRVIO_ERR_CUDA; }
            const cudaError_t ie = cudaGraphInstantiate(&b->gexec, graph, 0);
            cudaGraphDestroy(graph);
            if (ie != cudaSuccess) { b->gexec = nullptr; xrvio_set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie)); return XRVIO_ERR_CUDA; }
            b->graph_n = n;
        }
        CKC(cudaGraphLaunch(b->gexec, c->st_img));
        c->stats.graph_launches += 1;
        launched = true;
    }
#endif
    if (!launched) { const int rc = issue(); if (rc) return rc; }
    b->uses++;
    c->stats.kernels_launched += 2 + 3 + 2 + (c->d_und ? 1 : 0);      // the same kernels run whether issued one by one or replayed from the graph
    c->stats.d2h_bytes += (size_t)n * xvfe::kMaxCorners * 3 * sizeof(float) + n * sizeof(int);
    CKC(cudaEventRecord(b->done, c->st_img));
    CKC(cudaGetLastError());
    return XRVIO_OK;
}

static int lane_get_pose(Lane *c, int session, xrvio_pose *pose) {
    if (!c || !pose || session < 0 || session >= c->S) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    *pose = c->ses[session].out_pose;
    pose->valid = pose->valid && c->ses[session].sw_map != nullptr;
    return XRVIO_OK;
}
static int lane_get_latest_state(Lane *c, int session, double *t, double state16[16]) {
    if (!c || session < 0 || session >= c->S) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    Session &s = c->ses[session];
    if (!s.has_swt_latest) return XRVIO_ERR_STATE;
    if (t) *t = s.swt_t;
    if (state16) std::memcpy(state16, s.swt_state.d, sizeof(s.swt_state.d));
    return XRVIO_OK;
}
static int lane_get_state(Lane *c, int session) {
    if (!c || session < 0 || session >= c->S) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    return c->ses[session].sw_map ? XRVIO_STATE_TRACKING : XRVIO_STATE_INITIALIZING;
}
static int lane_get_landmarks(Lane *c, int session, double *xyz, int capacity) {
    if (!c || session < 0 || session >= c->S) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    Session &s = c->ses[session];
    if (!s.sw_map) return 0;
    int n = 0;
    for (auto &t : s.sw_map->tracks)
        if (t->valid && !t->refs.empty()) {
            if (xyz && n < capacity) { V3 x = get_landmark_point(t.get(), c->rig); xyz[3 * n] = x.x; xyz[3 * n + 1] = x.y; xyz[3 * n + 2] = x.z; }
            n++;
        }
    return n;
}
static int lane_set_profiling(Lane *c, int enable) {
    if (!c) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    c->prof = enable != 0;
    return XRVIO_OK;
}
static int lane_get_kernel_times(Lane *c, double *ms, long long *launches, int capacity) {
    if (!c || !ms || !launches) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    cudaStreamSynchronize(c->st);
    cudaStreamSynchronize(c->st_marg);
    cudaStreamSynchronize(c->st_img);
    prof_collect(c);
    for (int k = 0; k < capacity && k < 16; ++k) { ms[k] = c->prof_ms[k]; launches[k] = c->prof_launches[k]; }
    return FAM_COUNT;
}
static int lane_get_phase_times(Lane *c, double *ms, int capacity) {
    if (!c || !ms) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    for (int k = 0; k < capacity && k < 32; ++k) ms[k] = c->phase_ms[k];
    return 32;
}
static int lane_get_stats(Lane *c, xrvio_stats *st) {
    if (!c || !st) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    *st = c->stats;
    return XRVIO_OK;
}

// the step itself lives in session_step.inl to keep this file readable
#include "session_step.inl"

// =================================================================================================
// Public session API: a context is split into independent LANES (disjoint session ranges, each with
// its own CUDA stream, arenas and host worker threads).  xrvio_step advances all lanes concurrently,
// so one lane's host bookkeeping overlaps the other lanes' kernels and small kernels of different
// lanes share the GPU.  Sessions never interact, so the lane split does not change any result.
// =================================================================================================
// One persistent host thread per lane with a command queue.  Synchronous mode (default): xrvio_step hands every lane its step and
// waits for all of them; the push calls run on the caller's thread (the workers are idle then).  Asynchronous mode
// (xrvio_set_async): push_imu / push_images / step are queued per lane and return at once; every lane works through its queue
// at its own pace, so a lane whose step is short (no window solve this frame) starts its next frame instead of waiting for the
// lane with the largest bundle adjustment -- throughput follows the MEAN lane time instead of the maximum per step.  A lane is
// never more than kMaxStepsAhead steps behind the caller (back-pressure in xrvio_step): a window solve that accepts all of its
// 30 trust-region steps makes a lane step 3-4 times longer than the median, and the slack lets the other lanes run on meanwhile.  Errors of queued commands are sticky
// per lane and reported by the next call that waits (xrvio_sync, the getters other than xrvio_get_pose).
constexpr int kMaxStepsAhead = 6;   // < kImgQueue: every queued step may have a frame waiting in the session's image queue
struct LaneWorker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::pair<std::function<int(Lane *)>, bool>> q;      // (command, is_step)
    int inflight = 0, steps_pending = 0;
    bool quit = false;
    int rc = 0;
    std::string err;
    std::mutex pose_mu;                 // output poses are read by xrvio_get_pose without waiting for the queue
};

struct xrvio_ctx {
    std::vector<Lane *> lanes;
    std::vector<std::unique_ptr<LaneWorker>> workers;
    std::vector<std::pair<int, int>> where;     // session -> (lane, local index)
    int S = 0, device = 0;
    bool async = false;
    std::vector<std::vector<int>> tmp_idx;
};

namespace {
void lane_worker_main(xrvio_ctx *ctx, int l) {
    LaneWorker &w = *ctx->workers[l];
    cudaSetDevice(ctx->device);
    std::unique_lock<std::mutex> lk(w.mu);
    while (true) {
        w.cv.wait(lk, [&] { return !w.q.empty() || w.quit; });
        if (w.quit) return;
        std::function<int(Lane *)> job = std::move(w.q.front().first);
        const bool is_step = w.q.front().second;
        w.q.pop_front();
        const bool skip = w.rc != 0;            // after an error the rest of the queue is dropped, the error stays
        lk.unlock();
        int rc = 0;
        std::string err;
        if (!skip) { rc = job(ctx->lanes[l]); if (rc) err = xrvio_last_error(); }
        lk.lock();
        if (rc && !w.rc) { w.rc = rc; w.err = err; }
        w.inflight--;
        if (is_step) w.steps_pending--;
        w.cv.notify_all();
    }
}
void start_workers(xrvio_ctx *ctx) {
    const int L = (int)ctx->lanes.size();
    if (L <= 1) return;
    for (int l = 0; l < L; ++l) ctx->workers.emplace_back(new LaneWorker());
    for (int l = 0; l < L; ++l) ctx->workers[l]->th = std::thread(lane_worker_main, ctx, l);
}
void stop_workers(xrvio_ctx *ctx) {
    for (auto &w : ctx->workers) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; }
        w->cv.notify_all();
        if (w->th.joinable()) w->th.join();
    }
    ctx->workers.clear();
}
void lane_enqueue(xrvio_ctx *ctx, int l, std::function<int(Lane *)> fn, bool is_step) {
    LaneWorker &w = *ctx->workers[l];
    std::unique_lock<std::mutex> lk(w.mu);
    if (is_step) w.cv.wait(lk, [&] { return w.steps_pending < kMaxStepsAhead; });
    w.q.emplace_back(std::move(fn), is_step);
    w.inflight++;
    if (is_step) w.steps_pending++;
    w.cv.notify_all();
}
// waits until lane l has worked through its queue; returns (and clears) its sticky error
int lane_drain(xrvio_ctx *ctx, int l) {
    LaneWorker &w = *ctx->workers[l];
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv.wait(lk, [&] { return w.inflight == 0; });
    const int rc = w.rc;
    if (rc) { xrvio_set_error(w.err); w.rc = 0; w.err.clear(); }
    return rc;
}
int drain_all(xrvio_ctx *ctx) {
    int rc_all = XRVIO_OK;
    std::string first;
    for (int l = 0; l < (int)ctx->workers.size(); ++l) {
        const int rc = lane_drain(ctx, l);
        if (rc && rc_all == XRVIO_OK) { rc_all = rc; first = xrvio_last_error(); }
    }
    if (rc_all) xrvio_set_error(first);
    return rc_all;
}
template <class F> int each_lane(xrvio_ctx *ctx, F fn) {
    const int L = (int)ctx->lanes.size();
    if (L == 1 || ctx->workers.empty()) {
        for (int l = 0; l < L; ++l) { const int rc = fn(ctx->lanes[l]); if (rc) return rc; }
        return XRVIO_OK;
    }
    for (int l = 0; l < L; ++l) lane_enqueue(ctx, l, fn, false);
    return drain_all(ctx);
}
inline bool is_async(const xrvio_ctx *ctx) { return ctx->async && !ctx->workers.empty(); }
}  // namespace

extern "C" {

int xrvio_default_params(xrvio_params *p, int width, int height) { return xrvio_default_params_impl(p, width, height); }

int xrvio_create(const xrvio_params *p, int n_sessions, int device, xrvio_ctx **out) {
    if (!p || !out || n_sessions <= 0) { xrvio_set_error("xrvio_create: bad argument"); return XRVIO_ERR_ARG; }
    // one hardware work queue per stream (8 lanes x 3 streams); only effective when this process has no CUDA context yet -- a host
    // application that initialises CUDA first sets the variable itself
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    // 64 sessions: 10 lanes x 3 streams = 30 of the 32 hardware work queues; the measured optimum (profiles/r2_lane_sweep.txt:
    // 13.2 k frames/s against 10.6 k with 8 lanes, 12.2 k with 12, 12.0 k with 16 lanes sharing the low-priority stream)
    int L = n_sessions >= 64 ? 10 : (n_sessions >= 32 ? 4 : (n_sessions >= 8 ? 2 : 1));
    if (const char *e = std::getenv("XRVIO_LANES")) L = std::max(1, std::min(std::atoi(e), n_sessions));
    unsigned hc = std::thread::hardware_concurrency();
    int nt = (int)std::min<unsigned>(hc ? hc : 4u, 64u);
    if (const char *e = std::getenv("XRVIO_HOST_THREADS")) nt = std::max(1, std::atoi(e));
    const int per_lane_threads = std::max(1, nt / L);
    xrvio_ctx *ctx = new xrvio_ctx();
    ctx->S = n_sessions; ctx->device = device;
    int base = 0;
    for (int l = 0; l < L; ++l) {
        const int cnt = n_sessions / L + (l < n_sessions % L ? 1 : 0);
        Lane *lane = nullptr;
        int rc = lane_create(p, cnt, device, per_lane_threads, &lane);
        if (rc) { for (Lane *x : ctx->lanes) lane_destroy(x); delete ctx; return rc; }
        ctx->lanes.push_back(lane);
        for (int i = 0; i < cnt; ++i) ctx->where.emplace_back(l, i);
        base += cnt;
    }
    ctx->tmp_idx.resize(L);
    start_workers(ctx);
    *out = ctx;
    return XRVIO_OK;
}

int xrvio_destroy(xrvio_ctx *ctx) {
    if (!ctx) return XRVIO_OK;
    if (!ctx->workers.empty()) drain_all(ctx);
    stop_workers(ctx);
    for (Lane *l : ctx->lanes) lane_destroy(l);
    delete ctx;
    return XRVIO_OK;
}

#define LANE_OF(sess, l, i)                                                                         \
    if (!ctx || (sess) < 0 || (sess) >= ctx->S) { xrvio_set_error("bad session index"); return XRVIO_ERR_ARG; } \
    Lane *l = ctx->lanes[ctx->where[(sess)].first];                                                  \
    const int i = ctx->where[(sess)].second;

// getters and one-off setters first wait for the lane's queue in asynchronous mode (a no-op in synchronous mode)
#define LANE_IDLE(sess)                                                                              \
    if (ctx && (sess) >= 0 && (sess) < ctx->S && is_async(ctx)) { const int rc_ = lane_drain(ctx, ctx->where[(sess)].first); if (rc_) return rc_; }

int xrvio_set_async(xrvio_ctx *ctx, int enable) {
    if (!ctx) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    if (!ctx->workers.empty()) { const int rc = drain_all(ctx); if (rc) return rc; }
    ctx->async = enable != 0;
    return XRVIO_OK;
}
int xrvio_sync(xrvio_ctx *ctx) {
    if (!ctx) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    return ctx->workers.empty() ? XRVIO_OK : drain_all(ctx);
}

int xrvio_push_bootstrap_state(xrvio_ctx *ctx, int session, double t, const double q[4], const double p[3], const double v[3]) {
    LANE_IDLE(session);
    LANE_OF(session, l, i);
    return lane_push_bootstrap_state(l, i, t, q, p, v);
}

int xrvio_push_imu(xrvio_ctx *ctx, const int *session, const xrvio_imu *samples, int n) {
    if (!ctx || !session || !samples || n < 0) { xrvio_set_error("xrvio_push_imu: bad argument"); return XRVIO_ERR_ARG; }
    if (is_async(ctx)) {        // one queued command per lane, holding a copy of its samples in arrival order
        const int L = (int)ctx->lanes.size();
        std::vector<std::shared_ptr<std::pair<std::vector<int>, std::vector<xrvio_imu>>>> per(L);
        for (int k = 0; k < n; ++k) {
            if (session[k] < 0 || session[k] >= ctx->S) { xrvio_set_error("bad session index"); return XRVIO_ERR_ARG; }
            const int l = ctx->where[session[k]].first;
            if (!per[l]) per[l] = std::make_shared<std::pair<std::vector<int>, std::vector<xrvio_imu>>>();
            per[l]->first.push_back(ctx->where[session[k]].second);
            per[l]->second.push_back(samples[k]);
        }
        for (int l = 0; l < L; ++l)
            if (per[l]) {
                auto d = per[l];
                lane_enqueue(ctx, l, [d](Lane *lane) { return lane_push_imu(lane, d->first.data(), d->second.data(), (int)d->first.size()); }, false);
            }
        return XRVIO_OK;
    }
    for (int k = 0; k < n; ++k) {
        LANE_OF(session[k], l, i);
        int rc = lane_push_imu(l, &i, samples + k, 1);
        if (rc) return rc;
    }
    return XRVIO_OK;
}

int xrvio_push_images(xrvio_ctx *ctx, const int *session, const uint8_t *const *data, const int *stride, const double *t, int n, int on_device) {
    if (!ctx || !session || !data || !t || n < 0) { xrvio_set_error("xrvio_push_images: bad argument"); return XRVIO_ERR_ARG; }
    // the images of one call are grouped per lane: each group is one batch of that lane's image kernels
    const int L = (int)ctx->lanes.size();
    std::vector<std::vector<int>> loc(L);
    std::vector<std::vector<const uint8_t *>> dat(L);
    std::vector<std::vector<int>> str(L);
    std::vector<std::vector<double>> tt(L);
    for (int k = 0; k < n; ++k) {
        if (session[k] < 0 || session[k] >= ctx->S) { xrvio_set_error("bad session index"); return XRVIO_ERR_ARG; }
        const int l = ctx->where[session[k]].first;
        loc[l].push_back(ctx->where[session[k]].second); dat[l].push_back(data[k]); tt[l].push_back(t[k]);
        if (stride) str[l].push_back(stride[k]);
    }
    for (int l = 0; l < L; ++l) {
        if (loc[l].empty()) continue;
        if (is_async(ctx)) {    // the image buffers themselves must stay valid until the lane has consumed them (xrvio.h)
            struct Cmd { std::vector<int> loc; std::vector<const uint8_t *> dat; std::vector<int> str; std::vector<double> tt; int dev; bool has_str; };
            auto d = std::make_shared<Cmd>(Cmd{loc[l], dat[l], str[l], tt[l], on_device, stride != nullptr});
            LaneWorker *w = ctx->workers[l].get();
            lane_enqueue(ctx, l, [d, w](Lane *lane) {
                std::lock_guard<std::mutex> g(w->pose_mu);          // lane_push_images updates the sessions' output poses
                return lane_push_images(lane, d->loc.data(), d->dat.data(), d->has_str ? d->str.data() : nullptr, d->tt.data(), (int)d->loc.size(), d->dev); }, false);
            continue;
        }
        int rc = lane_push_images(ctx->lanes[l], loc[l].data(), dat[l].data(), stride ? str[l].data() : nullptr, tt[l].data(), (int)loc[l].size(), on_device);
        if (rc) return rc;
    }
    return XRVIO_OK;
}

int xrvio_step(xrvio_ctx *ctx) {
    if (!ctx) { xrvio_set_error("xrvio_step: null context"); return XRVIO_ERR_ARG; }
    if (is_async(ctx)) {
        for (int l = 0; l < (int)ctx->lanes.size(); ++l) lane_enqueue(ctx, l, [](Lane *lane) { return lane_step(lane); }, true);
        return XRVIO_OK;
    }
    return each_lane(ctx, [](Lane *l) { return lane_step(l); });
}

// In asynchronous mode this is the one getter that does not wait: it returns the newest pose the session has produced so far
// (the frame it belongs to is pose->t), like Detail::get_latest_pose of the reference's threaded build.
int xrvio_get_pose(xrvio_ctx *ctx, int session, xrvio_pose *pose) {
    LANE_OF(session, l, i);
    if (is_async(ctx)) {
        std::lock_guard<std::mutex> g(ctx->workers[ctx->where[session].first]->pose_mu);
        return lane_get_pose(l, i, pose);
    }
    return lane_get_pose(l, i, pose);
}
int xrvio_get_latest_state(xrvio_ctx *ctx, int session, double *t, double state16[16]) { LANE_IDLE(session); LANE_OF(session, l, i); return lane_get_latest_state(l, i, t, state16); }
int xrvio_get_state(xrvio_ctx *ctx, int session) { LANE_IDLE(session); LANE_OF(session, l, i); return lane_get_state(l, i); }
int xrvio_get_landmarks(xrvio_ctx *ctx, int session, double *xyz, int capacity) { LANE_IDLE(session); LANE_OF(session, l, i); return lane_get_landmarks(l, i, xyz, capacity); }

int xrvio_get_features(xrvio_ctx *ctx, int session, double *xy, long long *track_id, int *track_len, int capacity) {
    LANE_IDLE(session);
    LANE_OF(session, l, i);
    Session &s = l->ses[i];
    if (s.ft_map.frame_num() == 0) return 0;
    const Frame *f = s.ft_map.frames.back().get();
    const int n = (int)f->keypoint_num();
    for (int k = 0; k < n && k < capacity; ++k) {
        if (xy) apply_k(f->bearings[k], l->K, xy + 2 * k);
        const Track *t = f->tracks[k];
        if (track_id) track_id[k] = t ? t->id : -1;
        if (track_len) track_len[k] = t ? (int)t->keypoint_num() : 0;
    }
    return n;
}

int xrvio_get_init_info(xrvio_ctx *ctx, int session, xrvio_init_info *out) {
    LANE_IDLE(session);
    LANE_OF(session, l, i);
    if (!out) { xrvio_set_error("xrvio_get_init_info: null output"); return XRVIO_ERR_ARG; }
    *out = l->ses[i].init;
    out->essential_rejected = (int)std::min<long long>(l->ses[i].essential_rejected, 2147483647LL);
    return XRVIO_OK;
}

int xrvio_get_stats(xrvio_ctx *ctx, xrvio_stats *st) {
    if (!ctx || !st) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    if (is_async(ctx)) { const int rc = drain_all(ctx); if (rc) return rc; }
    std::memset(st, 0, sizeof(*st));
    for (Lane *l : ctx->lanes) {
        xrvio_stats a;
        lane_get_stats(l, &a);
        st->frames_processed += a.frames_processed; st->kernels_launched += a.kernels_launched; st->h2d_bytes += a.h2d_bytes; st->d2h_bytes += a.d2h_bytes;
        st->window_solves += a.window_solves; st->localize_solves += a.localize_solves; st->subwindow_solves += a.subwindow_solves;
        st->marginalizations += a.marginalizations;
        st->window_iterations += a.window_iterations; st->window_at_limit += a.window_at_limit; st->graph_launches += a.graph_launches;
        if (a.last_window_cost_initial != 0) { st->last_window_cost_initial = a.last_window_cost_initial; st->last_window_cost_final = a.last_window_cost_final; }
    }
    return XRVIO_OK;
}
int xrvio_set_profiling(xrvio_ctx *ctx, int enable) {
    if (!ctx) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    if (is_async(ctx)) { const int rc = drain_all(ctx); if (rc) return rc; }
    for (Lane *l : ctx->lanes) lane_set_profiling(l, enable);
    return XRVIO_OK;
}
int xrvio_get_kernel_times(xrvio_ctx *ctx, double *ms, long long *launches, int capacity) {
    if (!ctx || !ms || !launches) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    if (is_async(ctx)) { const int rc = drain_all(ctx); if (rc) return rc; }
    for (int k = 0; k < capacity; ++k) { ms[k] = 0; launches[k] = 0; }
    int nf = 0;
    for (Lane *l : ctx->lanes) {
        double m[16]; long long ln[16];
        nf = lane_get_kernel_times(l, m, ln, 16);
        for (int k = 0; k < capacity && k < 16; ++k) { ms[k] += m[k]; launches[k] += ln[k]; }
    }
    return nf;
}
int xrvio_debug_lane_busy(xrvio_ctx *ctx, double *ms_per_lane, int capacity) {
    if (!ctx || !ms_per_lane) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    if (is_async(ctx)) { const int rc = drain_all(ctx); if (rc) return rc; }
    const int L = (int)ctx->lanes.size();
    for (int l = 0; l < L && l < capacity; ++l) ms_per_lane[l] = ctx->lanes[l]->phase_ms[31];
    return L;
}
int xrvio_get_phase_times(xrvio_ctx *ctx, double *ms, int capacity) {
    if (!ctx || !ms) { xrvio_set_error("bad argument"); return XRVIO_ERR_ARG; }
    if (is_async(ctx)) { const int rc = drain_all(ctx); if (rc) return rc; }
    for (int k = 0; k < capacity; ++k) ms[k] = 0;
    for (Lane *l : ctx->lanes) {
        double m[32];
        lane_get_phase_times(l, m, 32);
        for (int k = 0; k < capacity && k < 32; ++k) ms[k] += m[k];
    }
    return 32;
}

}  // extern "C"
