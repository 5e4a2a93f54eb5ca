// Initializer of one session (row N3 of SURVEY.md 8(f)): Initializer::initialize of the reference
// (core/initializer.cpp:78-571) up to -- not including -- its final visual-inertial BA, which the
// back-end round runs as a regular batched solve right after (be_mode 1).
//
// Host geometry (two-view seed, IMU alignment) is init_host.hpp; everything iterative or per-sample
// runs on the lane's stream: PnP of the middle keyframes and the visual BA through the solve kernel,
// the preintegrations through the preintegration kernel, the N-view triangulations through the
// triangulation kernel.  A session tries once per frame while INITIALIZING, like the reference
// (frontend_worker.cpp:28-51); the attempt is synchronous for that session.
// Included by session_step.inl inside its anonymous namespace.

enum { INIT_OK = 0, INIT_FAIL_MATCHES = 1, INIT_FAIL_PARALLAX = 2, INIT_FAIL_PURE_ROTATION = 3, INIT_FAIL_TRIANGULATION = 4,
       INIT_FAIL_SFM_BA = 5, INIT_FAIL_SCALE = 6, INIT_FAIL_LANDMARKS = 7, INIT_FAIL_NO_IMU = 8, INIT_FAIL_SHORT = 9 };

// Initializer::mirror_keyframe_map (initializer.cpp:22-76)
std::unique_ptr<Map> mirror_keyframe_map(Lane *c, Session &s, long long init_frame_id, std::vector<long> *idxs_out = nullptr) {
    Map &ft = s.ft_map;
    const long last = ft.frame_index_by_id(init_frame_id);
    const long gap = c->prm.init_keyframe_gap, dist = gap * (c->prm.init_keyframe_num - 1);
    if (last < dist) return nullptr;
    const long first = last - dist;
    std::vector<long> idxs;
    for (int i = 0; i < c->prm.init_keyframe_num; ++i) idxs.push_back(first + i * gap);
    auto m = std::make_unique<Map>();
    m->track_id_counter = &s.track_id_counter;
    for (long ix : idxs) m->attach_frame(ft.frames[ix]->clone());
    for (size_t j = 1; j < m->frame_num(); ++j) {
        Frame *oi = ft.frames[idxs[j - 1]].get(), *oj = ft.frames[idxs[j]].get();
        Frame *ni = m->frames[j - 1].get(), *nj = m->frames[j].get();
        for (size_t ki = 0; ki < oi->keypoint_num(); ++ki)
            if (Track *tr = oi->tracks[ki]) {
                int kj = tr->get_keypoint_index(oj);
                if (kj >= 0) track_add_keypoint(frame_get_track(ni, (int)ki, nullptr), nj, kj);
            }
        nj->imu.clear();
        for (long f = idxs[j - 1]; f < idxs[j]; ++f) {
            const auto &od = ft.frames[f + 1]->imu;
            nj->imu.insert(nj->imu.end(), od.begin(), od.end());
        }
    }
    if (idxs_out) *idxs_out = idxs;
    return m;
}

inline void set_camera_pose(Frame *f, const Rig &rig, const Q4 &qc, const V3 &pc) {      // frame.cpp:183-186, sensor = camera
    const Q4 q = qmul(qc, qconj(rig.q_bc));
    st4(f->st.d, q);
    st3(f->st.d + 4, pc - qrot(q, rig.p_bc));
}

// one staged problem of one session, solved right away; out receives the solver summary
int solve_now(Lane *c, Session &s, Builder &b, xvbe::SolveOut &out) {
    int rc = pack_problem(c, s, b, c->prm.solver_iteration_limit);
    if (rc) { xrvio_set_error(s.err); return rc; }
    std::vector<Session *> one{&s};
    rc = merge_problems(c, one, 1);
    if (rc) return rc;
    rc = run_solves(c);
    if (rc) return rc;
    if (s.prob1 < 0) { xrvio_set_error("initializer: problem was not staged"); return XRVIO_ERR_ARG; }
    writeback(c, s, s.prob1);
    out = c->sout.h[s.prob1];
    s.prob1 = -1;
    return XRVIO_OK;
}

// Track::triangulate (track.cpp:46-76) of a list of tracks on the GPU; xyzv[k] = (x, y, z, valid)
int triangulate_now(Lane *c, Session &s, const std::vector<Track *> &tracks, std::vector<std::array<double, 4>> &xyzv) {
    xyzv.clear();
    if (tracks.empty()) return XRVIO_OK;
    s.stg.clear_jobs();
    for (Track *t : tracks) add_tri_job(c, s, t);
    std::vector<Session *> one{&s};
    int rc = merge_tri_jobs(c, one);
    if (rc) return rc;
    rc = run_triangulation(c);
    if (rc) return rc;
    xyzv.resize(tracks.size());
    for (size_t k = 0; k < tracks.size(); ++k) {
        const double *o = c->tout.h + (size_t)(s.tri_base + k) * 4;
        xyzv[k] = {o[0], o[1], o[2], o[3]};
    }
    return XRVIO_OK;
}

// Initializer::preintegrate (initializer.cpp:529-535): every edge with the given bias, Jacobians only
int preintegrate_now(Lane *c, Session &s, Map *m, const V3 &bg, const V3 &ba, std::vector<xvi::AlignDelta> &deltas, bool &has_imu) {
    s.stg.clear_jobs();
    State16 bias;
    st3(bias.d + 10, bg); st3(bias.d + 13, ba);
    has_imu = true;
    for (size_t j = 1; j < m->frame_num(); ++j) {
        Frame *fj = m->frames[j].get();
        if (fj->imu.empty()) { has_imu = false; s.stg.clear_jobs(); return XRVIO_OK; }
        add_preint_job(s, fj->imu, fj->t, bias, 1);
    }
    std::vector<Session *> one{&s};
    int rc = merge_preint_jobs(c, one);
    if (rc) return rc;
    rc = run_preint_jobs(c);
    if (rc) return rc;
    deltas.resize(m->frame_num() - 1);
    for (size_t j = 1; j < m->frame_num(); ++j) {
        const double *b = c->pre_out.h + (size_t)(s.pre_base + (int)j - 1) * XRVIO_PRE_BLOCK;
        xvi::AlignDelta &d = deltas[j - 1];
        d.t = b[xvbe::PB_T]; d.q = ld4(b + xvbe::PB_Q); d.p = ld3(b + xvbe::PB_P); d.v = ld3(b + xvbe::PB_V);
        for (int k = 0; k < 9; ++k) d.dq_dbg.m[k] = b[xvbe::PB_DQ_DBG + k];
    }
    return XRVIO_OK;
}

// Returns XRVIO_OK; s.init.last_fail tells whether s.sw_map now holds an initialised keyframe map.
int initialize_session(Lane *c, Session &s, long long frame_id) {
    xrvio_init_info &info = s.init;
    std::unique_ptr<Map> mp = mirror_keyframe_map(c, s, frame_id);
    if (!mp) { info.last_fail = INIT_FAIL_SHORT; return XRVIO_OK; }
    info.attempts++;
    Map *m = mp.get();
    const Rig &rig = c->rig;
    const size_t N = m->frame_num();
    Frame *fi = m->frames[0].get(), *fj = m->frames[N - 1].get();

    // ---- init_sfm [1]: matches between the first and the last keyframe (initializer.cpp:157-195) ----
    std::vector<xvi::V2> pts_i, pts_j;
    std::vector<double> parallax;
    std::vector<std::pair<int, int>> matches;
    for (size_t ki = 0; ki < fi->keypoint_num(); ++ki) {
        Track *tr = fi->tracks[ki];
        if (!tr) continue;
        const int kj = tr->get_keypoint_index(fj);
        if (kj < 0) continue;
        const V3 &bi = fi->bearings[ki], &bj = fj->bearings[kj];
        pts_i.push_back({bi.x / bi.z, bi.y / bi.z});
        pts_j.push_back({bj.x / bj.z, bj.y / bj.z});
        matches.emplace_back((int)ki, kj);
        double ui[2], uj[2];
        apply_k(bi, c->K, ui); apply_k(bj, c->K, uj);
        parallax.push_back(std::hypot(ui[0] - uj[0], ui[1] - uj[1]));
    }
    info.matches = (int)matches.size();
    xvi::SfmSeedConfig scfg;
    scfg.fx = c->K[0]; scfg.min_matches = c->prm.init_min_matches; scfg.min_parallax = c->prm.init_min_parallax;
    scfg.min_triangulation = c->prm.init_min_triangulation; scfg.seed = c->prm.random_seed;
    const xvi::SfmSeed seed = xvi::sfm_two_view(pts_i, pts_j, parallax, scfg);
    info.parallax = seed.parallax; info.candidate = seed.candidate;
    if (seed.fail) { info.last_fail = seed.fail; return XRVIO_OK; }

    // ---- [2.1] first / last keyframe poses, seed landmarks (:287-303) ----
    set_camera_pose(fi, rig, q4(0, 0, 0, 1), v3(0, 0, 0));
    {
        const M3 Rt = transpose(seed.R);
        set_camera_pose(fj, rig, xvi::mat_to_quat(Rt), -(Rt * seed.T));
    }
    for (size_t k = 0; k < matches.size(); ++k) {
        if (!seed.status[k]) continue;
        Track *tr = fi->tracks[matches[k].first];
        set_landmark_point(tr, seed.points[k], rig);
        tr->valid = true; tr->triangulated = true;
    }

    // ---- [2.2] keyframes in between: chained PnP on the landmarks anchored in keyframe 0 (:305-322) ----
    for (size_t j = 1; j + 1 < N; ++j) {
        Frame *fa = m->frames[j - 1].get(), *fb = m->frames[j].get();
        Q4 qa; V3 pa;
        camera_pose(fa, rig, qa, pa);
        set_camera_pose(fb, rig, qa, pa);
        Builder b{c, &s};
        b.with_motion = false;            // the motion blocks carry no residual here; Ceres drops them
        b.add_active(fb);
        for (size_t k = 0; k < fb->keypoint_num(); ++k) {
            Track *tr = fb->tracks[k];
            if (!tr || tr->find(fi) < 0) continue;
            if (tr->valid && tr->triangulated) b.add_rpe(fb, tr, true, (int)k);
        }
        if (b.rpe.empty()) { fb->tmp = -1; continue; }
        xvbe::SolveOut so;
        const int rc = solve_now(c, s, b, so);
        if (rc) return rc;
    }

    // ---- [2.3] triangulate what is left (:324-334) ----
    {
        std::vector<Track *> todo;
        for (auto &t : m->tracks) if (!t->valid) todo.push_back(t.get());
        std::vector<std::array<double, 4>> r;
        const int rc = triangulate_now(c, s, todo, r);
        if (rc) return rc;
        for (size_t k = 0; k < todo.size(); ++k)
            if (r[k][3] > 0) {
                Track *t = todo[k];
                t->life = 1;
                set_landmark_point(t, v3(r[k][0], r[k][1], r[k][2]), rig);
                t->valid = true; t->triangulated = true;
            }
    }

    // ---- [3.1] visual bundle adjustment: poses only, keyframe 0 fixed (:338-371) ----
    {
        m->frames[0]->fix_pose = true;
        Builder b{c, &s};
        b.with_motion = false;
        for (auto &f : m->frames) b.add_active(f.get());
        for (auto &f : m->frames)
            for (size_t j = 0; j < f->keypoint_num(); ++j) {
                Track *t = f->tracks[j];
                if (!t || !(t->valid && t->triangulated)) continue;
                if (f.get() == t->first_frame()) continue;
                b.add_rpe(f.get(), t, false, (int)j);
            }
        if (b.rpe.empty()) { for (auto &f : m->frames) f->tmp = -1; info.last_fail = INIT_FAIL_SFM_BA; return XRVIO_OK; }
        xvbe::SolveOut so;
        const int rc = solve_now(c, s, b, so);
        if (rc) return rc;
        if (so.termination == 3) { info.last_fail = INIT_FAIL_SFM_BA; return XRVIO_OK; }      // !IsSolutionUsable()
    }
    // ---- [3.2] landmark.reprojection_error is never written by the reference (state.h:41-47): only invalid tracks go ----
    {
        std::vector<Track *> prune;
        for (auto &t : m->tracks) if (!t->valid) prune.push_back(t.get());
        for (Track *t : prune) map_erase_track(m, t, rig);
    }

    // ---- init_imu (:385-396) ----
    xvi::Extr ex;
    ex.q_bc = rig.q_bc; ex.p_bc = rig.p_bc; ex.q_bi = ld4(c->prm.q_bi); ex.p_bi = ld3(c->prm.p_bi);
    std::vector<xvi::AlignFrame> af(N);
    for (size_t k = 0; k < N; ++k) { af[k].q = m->frames[k]->st.q(); af[k].p = m->frames[k]->st.p(); }
    std::vector<xvi::AlignDelta> deltas;
    bool has_imu = true;
    { const int rc = preintegrate_now(c, s, m, v3(0, 0, 0), v3(0, 0, 0), deltas, has_imu); if (rc) return rc; }
    if (!has_imu) { info.last_fail = INIT_FAIL_NO_IMU; return XRVIO_OK; }
    const V3 bg = xvi::solve_gyro_bias(af, deltas, ex);                                   // :398-423
    { const int rc = preintegrate_now(c, s, m, bg, v3(0, 0, 0), deltas, has_imu); if (rc) return rc; }
    xvi::AlignResult ar = xvi::solve_gravity_scale_velocity(af, deltas, ex);               // :425-466
    info.scale = ar.scale;
    if (!(ar.scale >= 0.001 && ar.scale <= 1.0)) { info.last_fail = INIT_FAIL_SCALE; return XRVIO_OK; }
    if (c->prm.init_refine_imu) {
        // the reference preintegrates again with the same bias (:471); the result is the one already held
        ar = xvi::refine_scale_velocity_via_gravity(af, deltas, ex, ar.gravity);           // :468-520
        info.scale = ar.scale;
        if (!(ar.scale >= 0.001 && ar.scale <= 1.0)) { info.last_fail = INIT_FAIL_SCALE; return XRVIO_OK; }
    }

    // ---- apply_init (:537-571) ----
    {
        const Q4 q = xvi::from_two_vectors(ar.gravity, v3(0, 0, -9.80665));
        for (size_t k = 0; k < N; ++k) {
            Frame *f = m->frames[k].get();
            const Q4 qi = qmul(f->st.q(), ex.q_bi);
            const V3 pi = f->st.p() + qrot(f->st.q(), ex.p_bi);
            const Q4 qn = qmul(q, qi);
            const V3 pn = ar.scale * qrot(q, pi);
            const Q4 qb = qmul(qn, qconj(ex.q_bi));
            st4(f->st.d, qb);
            st3(f->st.d + 4, pn - qrot(qb, ex.p_bi));
            st3(f->st.d + 7, qrot(q, ar.velocities[k]));
            st3(f->st.d + 10, bg);
            st3(f->st.d + 13, v3(0, 0, 0));
        }
        std::vector<Track *> all;
        for (auto &t : m->tracks) all.push_back(t.get());
        std::vector<std::array<double, 4>> r;
        const int rc = triangulate_now(c, s, all, r);
        if (rc) return rc;
        int final_points = 0;
        for (size_t k = 0; k < all.size(); ++k) {
            Track *t = all[k];
            if (r[k][3] > 0) {
                t->life = 1;
                set_landmark_point(t, v3(r[k][0], r[k][1], r[k][2]), rig);
                t->valid = true; t->triangulated = true;
                final_points++;
            } else t->valid = false;
        }
        info.landmarks = final_points;
        if (final_points < c->prm.init_min_landmarks) { info.last_fail = INIT_FAIL_LANDMARKS; return XRVIO_OK; }
    }
    info.last_fail = INIT_OK;
    s.sw_map = std::move(mp);
    return XRVIO_OK;
}
