// xrvio_step: batched FeatureTracker::work + FrontendWorker::work (included by capi_session.cu).

namespace {

typedef Lane::StepBuffers StepBuffers;
StepBuffers &sb(Lane *c) { return c->sbuf; }

// run body(i) for every active session on the host thread pool; first error wins
template <class F> int for_sessions(Lane *c, std::vector<Session *> &act, F body) {
    for (Session *s : act) s->rc = 0;
    c->pool->parallel_for((int)act.size(), [&](int i) { if (act[i]->rc == 0) { int rc = body(i, *act[i]); if (rc) act[i]->rc = rc; } });
    for (Session *s : act) if (s->rc) { xrvio_set_error(s->err.empty() ? "session error" : s->err); return s->rc; }
    return XRVIO_OK;
}

int run_poisson(Lane *c, int n, bool with_veto) {
    const size_t pb = (size_t)n * xvfe::kMaxPts;
    Stage &g = c->stage;
    CKC(g.reserve(pb * 2 * sizeof(double) + pb + 2 * n * sizeof(int) + 8 * 256));
    g.reset();
    const double *d_pts = g.put(c->poi_pts.h, pb * 2);
    const int *d_npre = g.put(c->poi_npre.h, (size_t)n);
    const int *d_ntot = g.put(c->poi_ntot.h, (size_t)n);
    const uint8_t *d_veto = with_veto ? g.put(c->poi_veto.h, pb) : nullptr;
    CKC(cudaMemcpyAsync(g.d, g.h, g.off, cudaMemcpyHostToDevice, c->st));
    { ProfScope ps(c, FAM_POISSON, 1);
      xvfe::poisson_filter(d_pts, d_npre, d_ntot, d_veto, c->poi_keep.d, n, c->prm.min_keypoint_distance, c->st); }
    c->stats.kernels_launched += 1;
    CKC(cudaMemcpyAsync(c->poi_keep.h, c->poi_keep.d, pb, cudaMemcpyDeviceToHost, c->st));
    CKC(cudaStreamSynchronize(c->st));
    c->stats.h2d_bytes += pb * 2 * sizeof(double) + 2 * n * sizeof(int) + (with_veto ? pb : 0);
    c->stats.d2h_bytes += pb;
    return XRVIO_OK;
}

// ---------------------------------------------------------------------------------------------
// FeatureTracker::work for one released frame of every session in `act`
// ---------------------------------------------------------------------------------------------
int front_end_round(Lane *c, std::vector<Session *> &act) {
    const int n = (int)act.size();
    const int w = c->prm.width, h = c->prm.height;
    const size_t P = (size_t)w * h;
    StepBuffers &B = sb(c);
    CKC(B.idx_raw.reserve(n)); CKC(B.idx_next.reserve(n)); CKC(B.idx_prev.reserve(n));
    PhaseClock pt(c, 0);
    // ---- phase A: preintegration (no Jacobian / covariance).  CLAHE + pyramid + GFTT of the new image were launched on the
    //      image stream when the image was pushed (lane_push_images); this stream only has to be ordered behind them ----
    { int rc = for_sessions(c, act, [&](int i, Session &s) -> int {
        s.stg.clear_jobs();
        s.cur = std::move(s.released.front());
        s.released.pop_front();
        Frame *f = s.cur.get();
        B.idx_next.h[i] = s.idx * kPyrSlots + f->pyr_slot;
        s.fe_has_last = s.ft_map.frame_num() > 0;
        s.last = s.fe_has_last ? s.ft_map.frames.back().get() : nullptr;
        s.pre_job = -1;
        const bool is_init = s.has_swt_latest;
        if (s.fe_has_last) {
            if (is_init) {
                // feature_tracker.cpp:50-76: the last optimised frame takes the back end's state; later
                // frames would be re-propagated (none exist in the synchronous pipeline)
                long idx = s.ft_map.frame_index_by_id(s.swt_frame_id);
                if (idx >= 0) {
                    s.ft_map.frames[idx]->st = s.swt_state;
                    if ((size_t)idx + 1 != s.ft_map.frame_num()) { s.err = "front end ran ahead of the back end (asynchronous mode is not supported)"; return XRVIO_ERR_STATE; }
                } else s.has_latest = false;
            }
            // IMU stitching (feature_tracker.cpp:77-88)
            if (!s.last->imu.empty()) {
                if (f->imu.empty() || (f->imu.front().t - s.last->t > 1.0e-5)) {
                    Imu m = s.last->imu.back();
                    m.t = s.last->t;
                    f->imu.insert(f->imu.begin(), m);
                }
            }
            if (!f->imu.empty()) {
                // a tracking session needs the same preintegration again in the back end, with Jacobians and covariance
                // (mirror_frame, sliding_window_tracker.cpp:76): compute it once, in full, and let phase E reuse the block
                const int fl = s.sw_map ? 3 : 0;
                s.pre_job = add_preint_job(s, f->imu, f->t, s.last->st, fl);
                f->pre_full = fl == 3;
                std::memcpy(f->pre_bias, &s.last->st.d[10], sizeof(f->pre_bias));
            }
        }
        return 0; }); if (rc) return rc; }
    { int rc = merge_preint_jobs(c, act); if (rc) return rc; }
    pt.next(1);
    CKC(cudaMemcpyAsync(B.idx_next.d, B.idx_next.h, n * sizeof(int), cudaMemcpyHostToDevice, c->st));
    { int rc = run_preint_jobs(c); if (rc) return rc; }     // also synchronises the stream (it needs nothing from the image stream)
    if (c->pre_jobs.n == 0) CKC(cudaStreamSynchronize(c->st));
    CKC(cudaGetLastError());
    pt.next(2);
    // ---- phase B: pyramidal LK (previous pyramid -> new pyramid), gyro-predicted initial guess ----
    std::atomic<int> n_klt_a{0};
    { int rc = for_sessions(c, act, [&](int i, Session &s) -> int {
        Frame *f = s.cur.get();
        c->klt_n.h[i] = 0;
        B.idx_prev.h[i] = B.idx_next.h[i];
        if (!s.fe_has_last) return 0;
        if (s.pre_job >= 0) fetch_pre(c, s.pre_base + s.pre_job, f->pre); else f->pre.assign(XRVIO_PRE_BLOCK, 0.0);
        Frame *last = s.last;
        const int np = (int)last->keypoint_num();
        if (np > xvfe::kMaxPts) { s.err = "more than XRVIO_MAX_PTS keypoints in a frame"; return XRVIO_ERR_LIMIT; }
        c->klt_n.h[i] = np;
        B.idx_prev.h[i] = s.idx * kPyrSlots + last->pyr_slot;
        s.klt_next.assign((size_t)np * 2, 0.0);
        // frame.cpp:82-93: delta_key_q = (q_cb^-1 q_ib dq q_ib'^-1 q_cb')^-1
        Q4 dkq = q4(0, 0, 0, 1);
        if (c->prm.predict_keypoints && s.pre_job >= 0) {
            const Q4 qbc = ld4(c->prm.q_bc), qbi = ld4(c->prm.q_bi);
            dkq = qconj(qmul(qmul(qmul(qmul(qconj(qbc), qbi), ld4(&f->pre[xvbe::PB_Q])), qconj(qbi)), qbc));
        }
        for (int k = 0; k < np; ++k) {
            double px[2], pr[2];
            apply_k(last->bearings[k], c->K, px);
            if (c->prm.predict_keypoints) apply_k(qrot(dkq, last->bearings[k]), c->K, pr);
            else { pr[0] = px[0]; pr[1] = px[1]; }
            float *pp = c->klt_prev.h + ((size_t)i * xvfe::kMaxPts + k) * 2;
            float *nn = c->klt_next.h + ((size_t)i * xvfe::kMaxPts + k) * 2;
            pp[0] = (float)px[0]; pp[1] = (float)px[1];
            nn[0] = (float)pr[0]; nn[1] = (float)pr[1];
            s.klt_next[2 * k] = pr[0]; s.klt_next[2 * k + 1] = pr[1];
        }
        n_klt_a += np;
    return 0; }); if (rc) return rc; }
    const int n_klt = n_klt_a.load();
    pt.next(3);
    if (n_klt > 0) {
        const size_t npt = (size_t)n * xvfe::kMaxPts, pb = npt * 2 * sizeof(float);
        Stage &g = c->stage;
        CKC(g.reserve(2 * pb + npt + 2 * n * sizeof(int) + 8 * 256));
        g.reset();
        // [next positions | status] first: they are what comes back
        float *d_next = g.put(c->klt_next.h, npt * 2);
        unsigned char *h_status;
        unsigned char *d_status = g.slot<unsigned char>(npt, &h_status);
        const size_t out_bytes = g.off;
        const float *d_prev = g.put(c->klt_prev.h, npt * 2);
        const int *d_n = g.put(c->klt_n.h, (size_t)n);
        const int *d_idx_prev = g.put(B.idx_prev.h, (size_t)n);
        CKC(cudaMemcpyAsync(g.d, g.h, g.off, cudaMemcpyHostToDevice, c->st));
        for (Session *sp : act) {       // the pyramids KLT reads are complete once the push batches' events have fired
            ImgBatch *b = static_cast<ImgBatch *>(sp->cur->img_batch);
            if (b && !b->waited) { CKC(cudaStreamWaitEvent(c->st, b->done, 0)); b->waited = true; }
        }
        { ProfScope ps(c, FAM_KLT, 1);
          xvfe::klt_track(c->d_pyr, c->d_pyr, d_idx_prev, B.idx_next.d, n, c->L, d_prev, d_next, d_n, d_status, c->st, &c->klt_maps, &c->klt_maps); }
        c->stats.kernels_launched += 1;
        CKC(cudaMemcpyAsync(g.h, g.d, out_bytes, cudaMemcpyDeviceToHost, c->st));
        CKC(cudaStreamSynchronize(c->st));
        std::memcpy(c->klt_next.h, g.h, pb);
        std::memcpy(c->klt_status.h, g.h + Stage::padded(pb), npt);
        c->stats.h2d_bytes += 2 * pb + 2 * n * sizeof(int);
        c->stats.d2h_bytes += pb + (size_t)n * xvfe::kMaxPts;
    }
    pt.next(4);
    // ---- phase C: Frame::track_keypoints after the LK call (frame.cpp:99-173) ----
    std::atomic<int> any_track_a{0};
    { int rc = for_sessions(c, act, [&](int i, Session &s) -> int {
        c->poi_npre.h[i] = 0; c->poi_ntot.h[i] = 0;
        if (!s.fe_has_last) return 0;
        Frame *f = s.cur.get();
        Frame *last = s.last;
        const int np = (int)last->keypoint_num();
        s.klt_status.assign(np, 0);
        for (int k = 0; k < np; ++k) {
            const unsigned char st = c->klt_status.h[(size_t)i * xvfe::kMaxPts + k];
            s.klt_status[k] = st;
            if (st) {
                const float *nn = c->klt_next.h + ((size_t)i * xvfe::kMaxPts + k) * 2;
                s.klt_next[2 * k] = (double)nn[0]; s.klt_next[2 * k + 1] = (double)nn[1];
            }
        }
        s.next_bearings.resize(np);
        for (int k = 0; k < np; ++k) s.next_bearings[k] = remove_k(&s.klt_next[2 * k], c->K);
        // essential-matrix RANSAC mask (frame.cpp:108-114): all ones for n >= 5, all zeros otherwise
        if (np < 5) std::fill(s.klt_status.begin(), s.klt_status.end(), 0);
        else if (c->prm.essential_ransac) {
            std::vector<xvi::V2> ch(np), nh(np);
            for (int k = 0; k < np; ++k) {
                const V3 &a = last->bearings[k], &b = s.next_bearings[k];
                ch[k] = {a.x / a.z, a.y / a.z}; nh[k] = {b.x / b.z, b.y / b.z};
            }
            M3 E = m3_identity();
            std::vector<char> emask;
            xvi::find_essential_matrix(ch, nh, 1.0, 0.999, 1000, 0, E, emask);
            for (int k = 0; k < np; ++k) if (!emask[k]) s.klt_status[k] = 0;
            s.essential_rejected += (long long)std::count(emask.begin(), emask.end(), (char)0);
        }
        std::vector<char> mask;
        M3 R = find_rotation_matrix(last->bearings, s.next_bearings, mask, (M_PI / 180.0) * c->prm.rotation_ransac_threshold);
        std::vector<double> angles;
        for (int k = 0; k < np; ++k)
            if (mask[k]) angles.push_back(std::acos(dot(R * last->bearings[k], s.next_bearings[k])) * 180 / M_PI);
        std::sort(angles.begin(), angles.end());
        const double misalignment = angles.empty() ? 0.0 : angles[angles.size() * 7 / 10];
        if (misalignment < c->prm.rotation_misalignment_threshold) f->no_translation = true;
        // order by track length (frame.cpp:136-150)
        s.order.clear();
        for (int k = 0; k < np; ++k) {
            if (!s.klt_status[k]) continue;
            Track *tr = last->tracks[k];
            if (!tr) continue;
            s.order.emplace_back(k, (int)tr->keypoint_num());
        }
        std::stable_sort(s.order.begin(), s.order.end(), [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.second > b.second; });
        if ((int)s.order.size() > xvfe::kMaxPts) { s.err = "too many tracked keypoints"; return XRVIO_ERR_LIMIT; }
        c->poi_ntot.h[i] = (int)s.order.size();
        for (size_t q = 0; q < s.order.size(); ++q) {
            const int k = s.order[q].first;
            double *pp = c->poi_pts.h + ((size_t)i * xvfe::kMaxPts + q) * 2;
            pp[0] = s.klt_next[2 * k]; pp[1] = s.klt_next[2 * k + 1];
            Track *tr = last->tracks[k];
            c->poi_veto.h[(size_t)i * xvfe::kMaxPts + q] = (tr && tr->trash) ? 1 : 0;
        }
        if (!s.order.empty()) any_track_a = 1;
    return 0; }); if (rc) return rc; }
    const bool any_track = any_track_a.load() != 0;
    pt.next(5);
    if (any_track) { int rc = run_poisson(c, n, true); if (rc) return rc; }
    pt.next(6);
    { int rc = for_sessions(c, act, [&](int i, Session &s) -> int {
        if (!s.fe_has_last) return 0;
        Frame *f = s.cur.get();
        Frame *last = s.last;
        const int np = (int)last->keypoint_num();
        // tracks that were never created for a keypoint (get_track(i) == nullptr at :142) keep their status
        for (size_t q = 0; q < s.order.size(); ++q)
            if (!c->poi_keep.h[(size_t)i * xvfe::kMaxPts + q]) s.klt_status[s.order[q].first] = 0;
        for (int k = 0; k < np; ++k) {
            if (!s.klt_status[k]) continue;
            const int nk = (int)f->keypoint_num();
            f->append_keypoint(s.next_bearings[k]);
            Track *tr = frame_get_track(last, k, nullptr);
            f->map = &s.ft_map;          // add_keypoint only needs the frame; map is attached below
            track_add_keypoint(tr, f, nk);
        }
        if (s.has_swt_latest) {
            // feature_tracker.cpp:93-116
            pre_predict(f->pre, last->st, f->st);
            s.has_latest = true; s.latest_t = f->t; s.latest = f->st;
        }
        return 0; }); if (rc) return rc; }
    pt.next(7);
    // ---- phase D: Frame::detect_keypoints (frame.cpp:55-72, opencv_image.cpp:46-72) ----
    for (Session *sp : act) {           // corner lists of the push batches: read back on the image stream long ago
        ImgBatch *b = static_cast<ImgBatch *>(sp->cur->img_batch);
        if (b && !b->host_ready) { CKC(cudaEventSynchronize(b->done)); b->host_ready = true; }
    }
    { int rc = for_sessions(c, act, [&](int i, Session &s) -> int {
        Frame *f = s.cur.get();
        const int npre = (int)f->keypoint_num();
        const ImgBatch *ib = static_cast<const ImgBatch *>(f->img_batch);
        const int nc = ib ? ib->count_h[f->img_index] : 0;
        if (npre + nc > xvfe::kMaxPts) { s.err = "too many keypoints for the Poisson filter"; return XRVIO_ERR_LIMIT; }
        c->poi_npre.h[i] = npre; c->poi_ntot.h[i] = npre + nc;
        for (int k = 0; k < npre; ++k) apply_k(f->bearings[k], c->K, c->poi_pts.h + ((size_t)i * xvfe::kMaxPts + k) * 2);
        // GFTT corners arrive strongest first; std::sort by response (opencv_image.cpp:47-50) keeps that order
        for (int k = 0; k < nc; ++k) {
            const float *cr = ib->corners_h + ((size_t)f->img_index * xvfe::kMaxCorners + k) * 3;
            double *pp = c->poi_pts.h + ((size_t)i * xvfe::kMaxPts + npre + k) * 2;
            pp[0] = (double)cr[0]; pp[1] = (double)cr[1];
        }
        s.n_corners = nc;
        return 0; }); if (rc) return rc; }
    pt.next(8);
    { int rc = run_poisson(c, n, false); if (rc) return rc; }
    pt.next(9);
    { int rc = for_sessions(c, act, [&](int i, Session &s) -> int {
        Frame *f = s.cur.get();
        const int npre = c->poi_npre.h[i];
        for (int k = 0; k < s.n_corners; ++k) {
            if (!c->poi_keep.h[(size_t)i * xvfe::kMaxPts + npre + k]) continue;
            const double *pp = c->poi_pts.h + ((size_t)i * xvfe::kMaxPts + npre + k) * 2;
            if (pp[0] < 20 || pp[1] < 20 || pp[0] >= w - 20 || pp[1] >= h - 20) continue;
            f->append_keypoint(remove_k(pp, c->K));
        }
        // attach + trim (feature_tracker.cpp:130-135)
        const bool is_init = s.has_swt_latest;
        s.ft_map.attach_frame(std::move(s.cur));
        const size_t max_frames = is_init ? c->prm.max_frames : c->prm.max_init_frames;
        // before initialisation latest_optimized_frame_id is nil() == SIZE_MAX, so the id test is true and the map is capped at
        // max_init_frames (feature_tracker.cpp:46,132-135)
        while (s.ft_map.frame_num() > max_frames && (!is_init || s.ft_map.frames[0]->id < s.swt_frame_id)) {
            map_untrack_frame(&s.ft_map, s.ft_map.frames[0].get(), c->rig);
            s.ft_map.detach_frame(0);
        }
    return 0; }); if (rc) return rc; }
    for (Session *sp : act) img_batch_release(c, sp->ft_map.frames.back().get());      // serial: the pool is not thread-safe
    c->stats.frames_processed += n;
    pt.stop();
    return XRVIO_OK;
}

// ---------------------------------------------------------------------------------------------
// Back end helpers
// ---------------------------------------------------------------------------------------------
int run_triangulation(Lane *c) {
    const int n = (int)c->tjobs.n;
    if (n == 0) return XRVIO_OK;
    CKC(c->tout.reserve((size_t)n * 4));
    Stage &g = c->stage;
    CKC(g.reserve(Stage::padded(n * sizeof(xvbe::TriJob)) + Stage::padded(c->tviews.n * sizeof(double)) + 512));
    g.reset();
    const xvbe::TriJob *d_jobs = g.put(c->tjobs.h, (size_t)n);
    const double *d_views = g.put(c->tviews.h, c->tviews.n);
    CKC(cudaMemcpyAsync(g.d, g.h, g.off, cudaMemcpyHostToDevice, c->st));
    { ProfScope ps(c, FAM_TRI, 1); xvbe::triangulate(d_jobs, n, d_views, c->tout.d, c->st); }
    c->stats.kernels_launched += 1;
    CKC(cudaMemcpyAsync(c->tout.h, c->tout.d, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CKC(cudaStreamSynchronize(c->st));
    c->stats.h2d_bytes += n * sizeof(xvbe::TriJob) + c->tviews.n * sizeof(double);
    c->stats.d2h_bytes += (size_t)n * 4 * sizeof(double);
    return XRVIO_OK;
}
void add_tri_job(Lane *c, Session &s, Track *tr) {          // Track::triangulate input (track.cpp:46-62), staged
    xvbe::TriJob j{(int)(s.stg.tv.size() / 16), (int)tr->refs.size()};
    for (auto &r : tr->refs) {
        Q4 q; V3 p;
        camera_pose(r.first, c->rig, q, p);
        double o[16] = {0};
        st4(o, q); st3(o + 4, p); st3(o + 7, r.first->bearings[r.second]);
        s.stg.tv.insert(s.stg.tv.end(), o, o + 16);
    }
    s.stg.tj.push_back(j);
}
int merge_tri_jobs(Lane *c, std::vector<Session *> &act) {
    c->tjobs.clear(); c->tviews.clear();
    size_t nj = 0, nv = 0;
    for (Session *s : act) { nj += s->stg.tj.size(); nv += s->stg.tv.size(); }
    if (c->tjobs.reserve(nj + 1) != cudaSuccess || c->tviews.reserve(nv + 16) != cudaSuccess) { xrvio_set_error("arena allocation failed"); return XRVIO_ERR_CUDA; }
    for (Session *s : act) {
        s->tri_base = (int)c->tjobs.n;
        const int vb = (int)(c->tviews.n / 16);
        for (xvbe::TriJob j : s->stg.tj) { j.obs_off += vb; c->tjobs.push(j); }
        if (!s->stg.tv.empty()) { std::memcpy(c->tviews.h + c->tviews.n, s->stg.tv.data(), s->stg.tv.size() * sizeof(double)); c->tviews.n += s->stg.tv.size(); }
        s->stg.tj.clear(); s->stg.tv.clear();
    }
    return XRVIO_OK;
}

#include "session_init.inl"

// Initializer::mirror_keyframe_map (initializer.cpp:22-76) + externally supplied states instead of init_sfm / init_imu
bool bootstrap_mirror(Lane *c, Session &s, long long init_frame_id) {
    Map &ft = s.ft_map;
    const long last = ft.frame_index_by_id(init_frame_id);
    const long gap = c->prm.init_keyframe_gap, dist = gap * (c->prm.init_keyframe_num - 1);
    if (last < dist) return false;
    const long first = last - dist;
    for (int i = 0; i < c->prm.init_keyframe_num; ++i)
        if (s.bootstrap.find(tkey(ft.frames[first + i * gap]->t)) == s.bootstrap.end()) return false;
    std::unique_ptr<Map> m = mirror_keyframe_map(c, s, init_frame_id);
    if (!m) return false;
    for (auto &f : m->frames) {
        const State16 &g = s.bootstrap[tkey(f->t)];
        f->st = g;
        for (int k = 10; k < 16; ++k) f->st.d[k] = 0.0;
    }
    s.sw_map = std::move(m);
    return true;
}

// SlidingWindowTracker::mirror_frame up to the preintegration (sliding_window_tracker.cpp:31-75)
bool mirror_frame(Lane *c, Session &s, long long frame_id) {
    Map *m = s.sw_map.get();
    Map &ft = s.ft_map;
    Frame *keyframe = m->frames.back().get();
    Frame *new_i = keyframe->subframes.empty() ? keyframe : keyframe->subframes.back().get();
    const long ii = ft.frame_index_by_id(new_i->id), jj = ft.frame_index_by_id(frame_id);
    if (ii < 0 || jj < 0) return false;
    Frame *old_i = ft.frames[ii].get(), *old_j = ft.frames[jj].get();
    std::unique_ptr<Frame> curr = old_j->clone();
    s.mirror_prepended = jj - 1 > ii;
    for (long index = jj - 1; index > ii; --index) {
        const auto &od = ft.frames[index]->imu;
        curr->imu.insert(curr->imu.begin(), od.begin(), od.end());
    }
    m->attach_frame(std::move(curr));
    Frame *new_j = m->frames.back().get();
    for (size_t ki = 0; ki < old_i->keypoint_num(); ++ki)
        if (Track *tr = old_i->tracks[ki]) {
            int kj = tr->get_keypoint_index(old_j);
            if (kj >= 0) {
                Track *nt = frame_get_track(new_i, (int)ki, m);
                track_add_keypoint(nt, new_j, kj);
                tr->trash = nt->trash && !nt->is_static;
            }
        }
    std::vector<Track *> prune;
    for (auto &t : m->tracks) if (t->trash && !t->is_static) prune.push_back(t.get());
    for (Track *t : prune) map_erase_track(m, t, c->rig);
    s.new_i = new_i; s.new_j = new_j;
    return true;
}

// SlidingWindowTracker::manage_keyframe (sliding_window_tracker.cpp:145-223)
bool manage_keyframe(Lane *c, Session &s) {
    Map *m = s.sw_map.get();
    Frame *ki = m->frames[m->frame_num() - 2].get();
    Frame *nj = m->frames[m->frame_num() - 1].get();
    if (!ki->subframes.empty()) {
        if (ki->subframes.back()->no_translation) {
            if (!nj->no_translation) {
                ki->subframes.back()->keyframe = true;
                std::unique_ptr<Frame> f = std::move(ki->subframes.back());
                ki->subframes.pop_back();
                m->attach_frame(std::move(f), (long)m->frame_num() - 1);
                nj->keyframe = true;
                return true;
            }
        } else {
            if (nj->no_translation) {
                std::unique_ptr<Frame> lifted = std::move(ki->subframes.back());
                ki->subframes.pop_back();
                lifted->keyframe = true;
                lifted->subframes.push_back(m->detach_frame(m->frame_num() - 1));
                m->attach_frame(std::move(lifted));
                return true;
            } else if ((int)ki->subframes.size() >= c->prm.subframe_size) {
                nj->keyframe = true;
                return true;
            }
        }
    }
    int cnt = 0;
    for (size_t k = 0; k < nj->keypoint_num(); ++k)
        if (Track *t = nj->tracks[k]) if (t->valid && t->triangulated && t->is_static) cnt++;
    if (cnt < c->prm.force_keyframe_landmarks) { nj->keyframe = true; return true; }
    ki->subframes.push_back(m->detach_frame(m->frame_num() - 1));
    return false;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
namespace {

int back_end_round(Lane *c, std::vector<Session *> &act) {
    PhaseClock pt(c, 10);
    // ---- sessions that were given no bootstrap state run the reference initializer, one attempt per frame
    //      (frontend_worker.cpp:28-51); on success the map continues below like a bootstrapped one ----
    for (Session *sp : act) {
        Session &s = *sp;
        s.init_now = false;
        if (s.sw_map || !c->prm.initializer || !s.bootstrap.empty()) continue;
        const int rc = initialize_session(c, s, s.ft_map.frames.back()->id);
        if (rc) return rc;
        s.init_now = (bool)s.sw_map;
    }
    // ---- phase E: mirror + preintegration with Jacobians and covariance ----
    { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
        s.stg.clear_jobs();
        s.be_mode = 0; s.pre_jobs.clear(); s.pre_job_frames.clear(); s.prob1 = s.prob2 = -1; s.is_keyframe = false;
        const long long frame_id = s.ft_map.frames.back()->id;
        if (!s.sw_map || s.init_now) {
            if (!s.init_now && !bootstrap_mirror(c, s, frame_id)) return 0;
            s.be_mode = 1;
            Map *m = s.sw_map.get();
            for (size_t j = 1; j < m->frame_num(); ++j) {
                Frame *fi = m->frames[j - 1].get(), *fj = m->frames[j].get();
                s.pre_jobs.push_back(fj->imu.empty() ? -1 : add_preint_job(s, fj->imu, fj->t, fi->st, 3));
            }
        } else {
            s.be_mode = 2;
            if (!mirror_frame(c, s, frame_id)) { s.be_mode = 3; return 0; }
            // -2: the front end already integrated exactly these samples with exactly this bias (the clone carries the block)
            const bool reuse = s.new_j->pre_full && !s.mirror_prepended && s.new_j->pre.size() == (size_t)XRVIO_PRE_BLOCK &&
                               std::memcmp(s.new_j->pre_bias, &s.new_i->st.d[10], sizeof(s.new_j->pre_bias)) == 0;
            s.pre_jobs.push_back(s.new_j->imu.empty() ? -1 : (reuse ? -2 : add_preint_job(s, s.new_j->imu, s.new_j->t, s.new_i->st, 3)));
        }
        return 0; }); if (rc) return rc; }
    { int rc = merge_preint_jobs(c, act); if (rc) return rc; }
    pt.next(11);
    { int rc = run_preint_jobs(c); if (rc) return rc; }
    pt.next(12);
    // ---- phase F: bootstrap triangulation; localize / initial BA problems ----
    bool any_boot = false;
    for (Session *sp : act) any_boot = any_boot || sp->be_mode == 1;
    if (any_boot) {
        { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
            if (s.be_mode != 1) return 0;
            s.tri_tracks.clear();
            for (auto &t : s.sw_map->tracks)
                if (t->keypoint_num() >= 2) { add_tri_job(c, s, t.get()); s.tri_tracks.push_back(t.get()); }
            return 0; }); if (rc) return rc; }
        { int rc = merge_tri_jobs(c, act); if (rc) return rc; }
        { int rc = run_triangulation(c); if (rc) return rc; }
    }
    { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
        s.stg.clear_prob();
        if (s.be_mode == 1) {
            Map *m = s.sw_map.get();
            for (size_t k = 0; k < s.tri_tracks.size(); ++k) {
                const double *o = c->tout.h + (size_t)(s.tri_base + k) * 4;
                if (o[3] > 0) {
                    Track *t = s.tri_tracks[k];
                    t->life = 1;
                    set_landmark_point(t, v3(o[0], o[1], o[2]), c->rig);
                    t->triangulated = true; t->valid = true;
                }
            }
            // initializer.cpp:97-130: frame 0 pose fixed, all landmarks, reprojection + preintegration factors
            m->frames[0]->fix_pose = true;
            Builder b{c, &s};
            for (auto &f : m->frames) b.add_active(f.get());
            for (auto &f : m->frames)
                for (size_t j = 0; j < f->keypoint_num(); ++j) {
                    Track *t = f->tracks[j];
                    if (!t || !(t->valid && t->triangulated)) continue;
                    if (f.get() == t->first_frame()) continue;
                    b.add_rpe(f.get(), t, false, (int)j);
                }
            for (size_t j = 1; j < m->frame_num(); ++j) {
                Frame *fi = m->frames[j - 1].get(), *fj = m->frames[j].get();
                if (s.pre_jobs[j - 1] < 0) continue;
                fetch_pre(c, s.pre_base + s.pre_jobs[j - 1], fj->pre);
                b.add_pre(fi, fj, &fj->pre, false);
            }
            return pack_problem(c, s, b, c->prm.solver_iteration_limit);
        } else if (s.be_mode == 2) {
            // mirror_frame tail (:76-79) + localize_newframe (:119-143)
            Frame *fj = s.new_j, *fi = s.new_i;
            if (s.pre_jobs[0] >= 0) fetch_pre(c, s.pre_base + s.pre_jobs[0], fj->pre);
            if (s.pre_jobs[0] != -1) pre_predict(fj->pre, fi->st, fj->st);
            Builder b{c, &s};
            b.add_active(fj);
            if (s.pre_jobs[0] != -1) b.add_pre(fi, fj, &fj->pre, true);
            for (size_t k = 0; k < fj->keypoint_num(); ++k)
                if (Track *t = fj->tracks[k]) if (t->valid && t->triangulated && t->is_static) b.add_rpe(fj, t, true, (int)k);
            return pack_problem(c, s, b, c->prm.solver_iteration_limit);
        }
        return 0; }); if (rc) return rc; }
    { int rc = merge_problems(c, act, 1); if (rc) return rc; }
    pt.next(13);
    { int rc = run_solves(c); if (rc) return rc; }
    pt.next(14);
    std::atomic<long long> n_win{0}, n_loc{0}, n_sub{0};
    { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
        s.stg.clear_jobs();
        if (s.prob1 < 0) return 0;
        writeback(c, s, s.prob1);
        if (s.be_mode == 1) {
            for (auto &f : s.sw_map->frames) f->keyframe = true;
            n_win++;
        } else {
            n_loc++;
            // ---- manage_keyframe + track_landmark (sliding_window_tracker.cpp:145-245) ----
            s.is_keyframe = manage_keyframe(c, s);
            if (s.is_keyframe) {
                Frame *nj = s.sw_map->frames.back().get();
                s.tri_tracks.clear();
                for (size_t k = 0; k < nj->keypoint_num(); ++k)
                    if (Track *t = nj->tracks[k]) if (!t->triangulated) { add_tri_job(c, s, t); s.tri_tracks.push_back(t); }
            }
        }
        return 0; }); if (rc) return rc; }
    { int rc = merge_tri_jobs(c, act); if (rc) return rc; }
    pt.next(15);
    { int rc = run_triangulation(c); if (rc) return rc; }
    pt.next(16);
    // ---- phase H: preintegration for refine_window / refine_subwindow ----
    { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
        s.stg.clear_jobs();
        if (s.be_mode != 2) return 0;
        Map *m = s.sw_map.get();
        s.pre_jobs.clear(); s.pre_job_frames.clear();
        if (s.is_keyframe) {
            for (size_t k = 0; k < s.tri_tracks.size(); ++k) {
                Track *t = s.tri_tracks[k];
                const double *o = c->tout.h + (size_t)(s.tri_base + k) * 4;
                if (o[3] > 0) { t->life = 1; set_landmark_point(t, v3(o[0], o[1], o[2]), c->rig); t->triangulated = true; t->valid = true; t->is_static = true; }
                else { t->inv_depth = -1.0; t->triangulated = false; t->valid = false; }
            }
            // refine_window: keyframe_preintegration with the subframes' IMU data merged (:298-321)
            for (size_t j = 1; j < m->frame_num(); ++j) {
                Frame *fi = m->frames[j - 1].get(), *fj = m->frames[j].get();
                fj->kf_imu.clear();
                for (auto &sf : fi->subframes) fj->kf_imu.insert(fj->kf_imu.end(), sf->imu.begin(), sf->imu.end());
                fj->kf_imu.insert(fj->kf_imu.end(), fj->imu.begin(), fj->imu.end());
                s.pre_jobs.push_back(fj->kf_imu.empty() ? -1 : add_preint_job(s, fj->kf_imu, fj->t, fi->st, 3));
                s.pre_job_frames.push_back(fj);
            }
        } else {
            Frame *frame = m->frames.back().get();
            if (frame->subframes.empty()) return 0;
            if (frame->subframes[0]->no_translation && frame->subframes.size() >= 9) {
                // compaction of long pure-rotation runs (:374-391)
                for (size_t i = frame->subframes.size() / 3; i > 0; --i) {
                    Frame *tgt = frame->subframes[i * 3 - 1].get();
                    std::vector<Imu> data;
                    for (size_t j = i * 3 - 1; j > (i - 1) * 3; --j) {
                        Frame *src = frame->subframes[j - 1].get();
                        data.insert(data.begin(), src->imu.begin(), src->imu.end());
                        map_untrack_frame(m, src, c->rig);
                        frame->subframes.erase(frame->subframes.begin() + (j - 1));
                    }
                    tgt->imu.insert(tgt->imu.begin(), data.begin(), data.end());
                }
            }
            // every subframe is integrated with its predecessor's bias before the solve (:402-405,:437-441)
            for (size_t i = 0; i < frame->subframes.size(); ++i) {
                Frame *sf = frame->subframes[i].get();
                Frame *prev = i == 0 ? frame : frame->subframes[i - 1].get();
                s.pre_jobs.push_back(sf->imu.empty() ? -1 : add_preint_job(s, sf->imu, sf->t, prev->st, 3));
                s.pre_job_frames.push_back(sf);
            }
        }
        return 0; }); if (rc) return rc; }
    { int rc = merge_preint_jobs(c, act); if (rc) return rc; }
    pt.next(17);
    // queued, not waited for: the host packs the problems while the kernel runs, the solve reads the blocks on the device
    // (PreRef::block < 0) and the host copies (kf_pre / pre, needed from the marginalisation on) are taken after the solve's
    // synchronisation
    { int rc = run_preint_jobs(c, false); if (rc) return rc; }
    pt.next(18);
    // ---- refine_window / refine_subwindow problems ----
    // initial priors (estimation/marginalization_factor.h:18-33): zero information except 1e15 on frame-0 q and p
    for (Session *sp : act) {
        Session &s = *sp;
        if (s.be_mode != 2 || !s.is_keyframe || s.sw_map->marg.present) continue;
        Map *m = s.sw_map.get();
        const int K = (int)m->frame_num() - 1;
        const size_t n = (size_t)15 * K;
        std::vector<double> a(16 * K + 2 * n * n + n + 1, 0.0);
        a.back() = 6.0;                              // rows 0..5 are the only non-zero rows of the initial S
        for (int k = 0; k < K; ++k) std::memcpy(&a[16 * k], m->frames[k]->st.d, 16 * sizeof(double));
        double *S = &a[16 * K], *A0 = S + n * n + n;
        for (int k = 0; k < 6; ++k) { S[k * n + k] = 1.0e15; A0[k * n + k] = 1.0e30; }
        m->marg.frames.clear();
        for (int k = 0; k < K; ++k) m->marg.frames.push_back(m->frames[k].get());
        m->marg.slot = 0; m->marg.present = true;
        CKC(cudaMemcpyAsync(c->d_prior + ((size_t)s.idx * 2) * c->prior_stride, a.data(), a.size() * sizeof(double), cudaMemcpyHostToDevice, c->st));
        CKC(cudaStreamSynchronize(c->st));
        c->stats.h2d_bytes += a.size() * sizeof(double);
    }
    { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
        s.stg.clear_prob();
        if (s.be_mode != 2) return 0;
        Map *m = s.sw_map.get();
        if (s.is_keyframe) {
            Builder b{c, &s};
            for (auto &f : m->frames) b.add_active(f.get());
            b.use_marg = true;
            for (auto &f : m->frames)
                for (size_t j = 0; j < f->keypoint_num(); ++j) {
                    Track *t = f->tracks[j];
                    if (!t || !(t->valid && t->triangulated && t->is_static)) continue;
                    Frame *ff = t->first_frame();
                    if (!ff->keyframe || f.get() == ff) continue;
                    b.add_rpe(f.get(), t, false, (int)j);
                }
            for (size_t j = 1; j < m->frame_num(); ++j) {
                if (s.pre_jobs[j - 1] < 0) continue;
                Frame *fj = s.pre_job_frames[j - 1];
                b.add_pre_dev(m->frames[j - 1].get(), fj, s.pre_base + s.pre_jobs[j - 1], false);
            }
            return pack_problem(c, s, b, c->prm.solver_iteration_limit);
        }
        Frame *frame = m->frames.back().get();
        if (frame->subframes.empty()) return 0;
        Builder b{c, &s};
        frame->fix_pose = frame->fix_motion = true;
        b.add_active(frame);
        const bool rbranch = frame->subframes[0]->no_translation;
        for (size_t i = 0; i < frame->subframes.size(); ++i) {
            Frame *sf = frame->subframes[i].get();
            Frame *prev = i == 0 ? frame : frame->subframes[i - 1].get();
            b.add_active(sf);
            if (s.pre_jobs[i] >= 0) b.add_pre_dev(prev, sf, s.pre_base + s.pre_jobs[i], false);
            if (!rbranch)
                for (size_t k = 0; k < sf->keypoint_num(); ++k)
                    if (Track *t = sf->tracks[k]) if (t->valid && t->triangulated && t->is_static) {
                        if (t->first_frame()->keyframe) b.add_rpe(sf, t, true, (int)k);
                        // else branch (:451-454) is unreachable with the default configuration
                    }
        }
        if (rbranch) {
            Frame *lastsf = frame->subframes.back().get();
            for (size_t k = 0; k < lastsf->keypoint_num(); ++k)
                if (Track *t = lastsf->tracks[k]) if (t->valid) {
                    if (t->triangulated) { if (t->is_static) b.add_rpe(lastsf, t, true, (int)k); }
                    else b.add_rot(lastsf, t);
                }
        }
        const int rc2 = pack_problem(c, s, b, c->prm.solver_iteration_limit);
        frame->fix_pose = frame->fix_motion = false;
        return rc2; }); if (rc) return rc; }
    { int rc = merge_problems(c, act, 2); if (rc) return rc; }
    pt.next(19);
    { int rc = run_solves(c); if (rc) return rc; }
    if (c->sdesc.n == 0 && c->pre_jobs.n > 0) CKC(cudaStreamSynchronize(c->st));      // no solve waited for the preintegration results
    pt.next(20);
    { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
        if (s.be_mode == 2)
            for (size_t k = 0; k < s.pre_jobs.size(); ++k)
                if (s.pre_jobs[k] >= 0) fetch_pre(c, s.pre_base + s.pre_jobs[k], s.is_keyframe ? s.pre_job_frames[k]->kf_pre : s.pre_job_frames[k]->pre);
        if (s.prob2 < 0) return 0;
        writeback(c, s, s.prob2);
        Map *m = s.sw_map.get();
        if (!s.is_keyframe) { n_sub++; return 0; }
        n_win++;
        // validity gates (sliding_window_tracker.cpp:325-357)
        for (auto &tp : m->tracks) {
            Track *t = tp.get();
            if (t->triangulated) {
                bool ok = true;
                const V3 x = get_landmark_point(t, c->rig);
                double rpe = 0.0, cnt = 0.0;
                for (auto &r : t->refs) {
                    Frame *f = r.first;
                    if (!f->keyframe) continue;
                    Q4 q; V3 p;
                    camera_pose(f, c->rig, q, p);
                    const V3 y = qrot(qconj(q), x - p);
                    if (!(y.z > 1.0e-3 && y.z <= 50)) { ok = false; break; }
                    double a[2], bb[2];
                    apply_k(y, c->K, a); apply_k(f->bearings[r.second], c->K, bb);
                    rpe += std::sqrt((a[0] - bb[0]) * (a[0] - bb[0]) + (a[1] - bb[1]) * (a[1] - bb[1]));
                    cnt += 1.0;
                }
                ok = ok && (rpe / std::max(cnt, 1.0) < 3.0);
                t->valid = ok;
            } else t->inv_depth = -1.0;
        }
        for (auto &tp : m->tracks) if (!tp->valid) tp->trash = true;
        return 0; }); if (rc) return rc; }
    for (Session *sp : act)
        if (sp->prob2 >= 0 && sp->is_keyframe) {
            c->stats.last_window_cost_initial = c->sout.h[sp->prob2].initial_cost;
            c->stats.last_window_cost_final = c->sout.h[sp->prob2].final_cost;
            if (c->dbg_on) {
                const xvbe::SolveOut &so = c->sout.h[sp->prob2];
                const xvbe::SolveDesc &sd = c->sdesc.h[sp->prob2];
                c->dbg_line += " [BA it=" + std::to_string(so.iterations) + " ok=" + std::to_string(so.n_successful) + " rej=" + std::to_string(so.n_unsuccessful) +
                               " term=" + std::to_string(so.termination) + " rpe=" + std::to_string(sd.n_rpe) + " lm=" + std::to_string(sd.n_lm_free) +
                               " act=" + std::to_string(sd.n_active) + " c0=" + std::to_string(so.initial_cost).substr(0, 8) + " c1=" + std::to_string(so.final_cost).substr(0, 8) + "]";
            }
            c->stats.window_iterations += c->sout.h[sp->prob2].iterations - 1;       // the trace also holds iteration 0
            c->stats.window_at_limit += c->sout.h[sp->prob2].termination == 0 ? 1 : 0;
        }
    c->stats.window_solves += n_win; c->stats.localize_solves += n_loc; c->stats.subwindow_solves += n_sub;
    pt.next(21);
    // ---- slide_window: marginalise frame 0 while the window is too long (:360-368, map.cpp:51-63) ----
    while (true) {
        { int rc = for_sessions(c, act, [&](int, Session &s) -> int {
            Staged &g = s.stg;
            g.clear_prob();
            if (s.be_mode != 2 || !s.is_keyframe) return 0;
            Map *m = s.sw_map.get();
            if ((int)m->frame_num() <= c->prm.sliding_window_size) return 0;
            Frame *victim = m->frames[0].get();
            for (auto &sf : victim->subframes) map_untrack_frame(m, sf.get(), c->rig);
            const int N = (int)m->frame_num();
            if (N > XRVIO_MAX_ACTIVE) { s.err = "window longer than XRVIO_MAX_ACTIVE at marginalisation"; return XRVIO_ERR_LIMIT; }
            for (int i = 0; i < N; ++i) m->frames[i]->tmp = (i == 0) ? N - 1 : i - 1;       // victim last
            xvbe::MargDesc d{};
            d.n_frames = N;
            g.states.resize((size_t)N * 16);
            for (int i = 1; i < N; ++i) std::memcpy(&g.states[(size_t)(i - 1) * 16], m->frames[i]->st.d, 128);
            std::memcpy(&g.states[(size_t)(N - 1) * 16], victim->st.d, 128);
            // observations of the victim's valid tracks whose first frame is a keyframe (:234-256)
            for (size_t j = 0; j < victim->keypoint_num(); ++j) {
                Track *t = victim->tracks[j];
                if (!t || !t->valid) continue;
                Frame *ref = t->first_frame();
                if (!ref->keyframe || ref->map != m) continue;
                const int lm = (int)g.invd.size();
                bool used = false;
                for (auto &r : t->refs) {
                    Frame *tg = r.first;
                    if (tg == ref || tg->map != m) continue;         // subframes are not in frame_indices
                    xvbe::RpeObs o{};
                    o.tgt = (short)tg->tmp; o.ref = (short)ref->tmp; o.lm = lm; o.prior = 0;
                    st3(o.zt, tg->bearings[r.second]); st3(o.zr, ref->bearings[t->refs.front().second]);
                    g.rpe.push_back(o);
                    used = true;
                }
                if (used) g.invd.push_back(t->inv_depth);
            }
            std::stable_sort(g.rpe.begin(), g.rpe.end(), [](const xvbe::RpeObs &x, const xvbe::RpeObs &y) { return x.tgt != y.tgt ? x.tgt < y.tgt : x.ref < y.ref; });
            d.n_lm = (int)g.invd.size(); d.n_rpe = (int)g.rpe.size(); d.n_pre = 0;
            if (N > 1 && !m->frames[1]->kf_pre.empty()) {
                xvbe::PreRef r{};
                r.i = (short)victim->tmp; r.j = (short)m->frames[1]->tmp; r.prior = 0; r.block = 0;
                g.pblocks = m->frames[1]->kf_pre;
                g.pref.push_back(r);
                d.n_pre = 1;
            }
            d.marg_k = (int)m->marg.frames.size();
            for (Frame *f : m->marg.frames) g.mframes.push_back(f->tmp);
            d.marg_off = (long long)(((size_t)s.idx * 2 + m->marg.slot) * c->prior_stride);
            d.out_off = (long long)(((size_t)s.idx * 2 + (1 - m->marg.slot)) * c->prior_stride);
            d.eigen_form = c->prm.prior_eigen_form;
            for (int i = 0; i < N; ++i) m->frames[i]->tmp = -1;
            g.md = d; g.has_marg = true;
            return 0; }); if (rc) return rc; }
        c->mdesc.clear();
        clear_solve_arenas(c);
        size_t scratch = 0, nst = 0, nl = 0, nr = 0, np = 0, npb = 0, nmf = 0;
        std::vector<Session *> ms;
        for (Session *sp : act) if (sp->stg.has_marg) {
            ms.push_back(sp);
            const Staged &g = sp->stg;
            nst += g.states.size(); nl += g.invd.size(); nr += g.rpe.size(); np += g.pref.size(); npb += g.pblocks.size(); nmf += g.mframes.size();
        }
        if (ms.empty()) break;
        const int nm = (int)ms.size();
        CKC(c->mdesc.reserve(nm)); CKC(c->states.reserve(nst + 16)); CKC(c->invd.reserve(nl + 1)); CKC(c->rpe.reserve(nr + 1));
        CKC(c->pref.reserve(np + 1)); CKC(c->pblocks.reserve(npb + XRVIO_PRE_BLOCK)); CKC(c->mframes.reserve(nmf + 1));
        for (Session *sp : ms) {
            Staged &g = sp->stg;
            xvbe::MargDesc d = g.md;
            d.frame_off = (int)(c->states.n / 16); d.lm_off = (int)c->invd.n; d.rpe_off = (int)c->rpe.n; d.pre_off = (int)c->pref.n;
            d.marg_frames_off = (int)c->mframes.n;
            const int blk_base = (int)(c->pblocks.n / XRVIO_PRE_BLOCK);
            auto app = [](auto &arena, const auto &v) { if (!v.empty()) { std::memcpy(arena.h + arena.n, v.data(), v.size() * sizeof(v[0])); arena.n += v.size(); } };
            app(c->states, g.states); app(c->invd, g.invd); app(c->rpe, g.rpe);
            for (xvbe::PreRef r : g.pref) { r.block += blk_base; c->pref.push(r); }
            app(c->pblocks, g.pblocks); app(c->mframes, g.mframes);
            d.scratch_off = (long long)scratch;
            scratch += xvbe::marg_scratch_doubles(d.n_frames, d.n_lm, d.n_rpe);
            c->mdesc.push(d);
            g.has_marg = false;
        }
        // own stream, staging block and scratch (see Lane::st_marg); a block or scratch that has to grow, and a staging block whose
        // previous upload may not have been consumed yet, are the only reasons to wait for the marginalisation stream
        Stage &g = c->stage_marg;
        const size_t stage_want = nm * sizeof(xvbe::MargDesc) + c->states.n * 8 + c->invd.n * 8 + c->rpe.n * sizeof(xvbe::RpeObs) + c->pref.n * sizeof(xvbe::PreRef) +
                                  c->pblocks.n * 8 + c->mframes.n * sizeof(int) + 16 * 256;
        if (scratch * sizeof(double) > c->scratch_marg.bytes || stage_want > g.cap) CKC(cudaStreamSynchronize(c->st_marg));
        if (c->marg_h2d_pending) { CKC(cudaEventSynchronize(c->ev_marg_h2d)); c->marg_h2d_pending = false; }
        CKC(c->scratch_marg.reserve(scratch * sizeof(double)));
        CKC(c->mrank.reserve(nm));
        CKC(g.reserve(stage_want));
        g.reset();
        xvbe::MargBatch mb;
        mb.desc = g.put(c->mdesc.h, (size_t)nm); mb.states = g.put(c->states.h, c->states.n); mb.inv_depth = g.put(c->invd.h, c->invd.n);
        mb.rpe = g.put(c->rpe.h, c->rpe.n); mb.pre = g.put(c->pref.h, c->pref.n); mb.pre_blocks = g.put(c->pblocks.h, c->pblocks.n);
        mb.marg_frames = g.put(c->mframes.h, c->mframes.n);
        // the window solve that produced the states of this batch has been read back already (the host packed them), and the prior
        // slots read here were written either by an earlier launch on this stream or by a host-synchronous upload: no wait needed
        CKC(cudaMemcpyAsync(g.d, g.h, g.off, cudaMemcpyHostToDevice, c->st_marg));      // one upload for the whole batch of problems
        CKC(cudaEventRecord(c->ev_marg_h2d, c->st_marg)); c->marg_h2d_pending = true;
        c->stats.h2d_bytes += nm * sizeof(xvbe::MargDesc) + c->states.n * 8 + c->invd.n * 8 + c->rpe.n * sizeof(xvbe::RpeObs) + c->pblocks.n * 8;
        mb.marg = c->d_prior; mb.out = c->d_prior; mb.scratch = c->scratch_marg.as<double>(); mb.rank_out = c->mrank.d;
        { ProfScope ps(c, FAM_MARG, 1, c->st_marg); xvbe::marginalize(mb, nm, c->rig_d, c->st_marg); }
        c->stats.kernels_launched += 1;
        CKC(cudaGetLastError());
        CKC(cudaEventRecord(c->ev_marg_done, c->st_marg)); c->marg_pending = true;
        c->pool->parallel_for(nm, [&](int i) {
            Session &s = *ms[i];
            Map *m = s.sw_map.get();
            Frame *victim = m->frames[0].get();
            m->marg.slot = 1 - m->marg.slot;
            m->marg.frames.clear();
            for (size_t k = 1; k < m->frame_num(); ++k) m->marg.frames.push_back(m->frames[k].get());
            for (size_t k = 0; k < victim->keypoint_num(); ++k)
                if (Track *t = victim->tracks[k]) track_remove_keypoint(t, victim, c->rig);
            m->frames.pop_front();
        });
        c->stats.marginalizations += nm;
    }
    pt.next(22);
    // ---- FrontendWorker latest state (frontend_worker.cpp:48-51,72-75) ----
    for (Session *sp : act) {
        Session &s = *sp;
        if (s.be_mode == 0 || !s.sw_map) continue;
        Map *m = s.sw_map.get();
        Frame *f = m->frames.back().get();
        if (!f->subframes.empty()) f = f->subframes.back().get();
        s.has_swt_latest = true; s.swt_t = f->t; s.swt_frame_id = s.ft_map.frames.back()->id; s.swt_state = f->st;
    }
    pt.stop();
    return XRVIO_OK;
}

}  // namespace

static int lane_step(Lane *c) {
    if (!c) { xrvio_set_error("xrvio_step: null context"); return XRVIO_ERR_ARG; }
    PhaseTimer busy_(c, 31);            // slot 31: wall time this lane spent inside its steps (xrvio_debug_lane_busy)
    // diagnostic: XRVIO_SLOW_STEP_MS=<ms> prints the phase breakdown of every lane step that takes longer
    static const double slow_ms = [] { const char *e = std::getenv("XRVIO_SLOW_STEP_MS"); return e ? std::atof(e) : 0.0; }();
    struct SlowDump {
        Lane *c; double thr; double before[32]; std::chrono::steady_clock::time_point t0;
        SlowDump(Lane *c_, double thr_) : c(c_), thr(thr_), t0(std::chrono::steady_clock::now()) { if (thr > 0) std::memcpy(before, c->phase_ms, sizeof(before)); }
        ~SlowDump() {
            if (thr <= 0) return;
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if (ms < thr) return;
            std::string line = "[xrvio] slow lane step " + std::to_string(ms) + " ms:";
            for (int k = 0; k < 31; ++k) { const double d = c->phase_ms[k] - before[k]; if (d > 0.3) line += " p" + std::to_string(k) + "=" + std::to_string(d).substr(0, 5); }
            line += c->dbg_line;
            std::fprintf(stderr, "%s\n", line.c_str());
        }
    } slow_dump_(c, slow_ms);
    c->dbg_on = slow_ms > 0; c->dbg_line.clear();
    CKC(cudaSetDevice(c->device));
    while (true) {
        std::vector<Session *> act;
        for (auto &s : c->ses) if (!s.released.empty()) act.push_back(&s);
        if (act.empty()) break;
        int rc = front_end_round(c, act);
        if (rc) return rc;
        rc = back_end_round(c, act);
        if (rc) return rc;
        if (c->prof) prof_collect(c);
    }
    CKC(cudaStreamSynchronize(c->st));
    return XRVIO_OK;
}
