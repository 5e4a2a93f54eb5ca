// CPU-side probe of the session layer's host logic (xrslam_b200/csrc/session_host.hpp): compiled by
// tests/test_host_logic.py with g++ and compared against the oracle.  Test infrastructure only.
//   stdin : "ransac n thr" followed by n rows of 6 doubles (bearing1, bearing2)
//           "tracks" -> exercises Map / Track bookkeeping and prints a digest
//           eig / lstsq / e5 / findE / findH / decompH / decompE / sfm / gyro / gsv / refine -> init_host.hpp
#include "../../xrslam_b200/csrc/session_host.hpp"
#include "../../xrslam_b200/csrc/init_host.hpp"
#include <cstdio>
#include <iostream>
#include <string>
using namespace xvs;

static void print_m3(const xv::M3 &A) { for (int k = 0; k < 9; ++k) std::printf("%.17g ", A.m[k]); }
static void print_v3(const xv::V3 &a) { std::printf("%.17g %.17g %.17g ", a.x, a.y, a.z); }
static xv::M3 read_m3() { xv::M3 A; for (int k = 0; k < 9; ++k) std::cin >> A.m[k]; return A; }
static void read_pairs(int n, std::vector<xvi::V2> &a, std::vector<xvi::V2> &b) {
    a.resize(n); b.resize(n);
    for (int i = 0; i < n; ++i) std::cin >> a[i].x >> a[i].y >> b[i].x >> b[i].y;
}
static void read_align(int N, xvi::Extr &x, std::vector<xvi::AlignFrame> &f, std::vector<xvi::AlignDelta> &d) {
    std::cin >> x.q_bc.x >> x.q_bc.y >> x.q_bc.z >> x.q_bc.w >> x.p_bc.x >> x.p_bc.y >> x.p_bc.z;
    x.q_bi = xv::q4(0, 0, 0, 1); x.p_bi = xv::v3(0, 0, 0);
    f.resize(N); d.resize(N - 1);
    for (auto &a : f) std::cin >> a.q.x >> a.q.y >> a.q.z >> a.q.w >> a.p.x >> a.p.y >> a.p.z;
    for (auto &a : d) {
        std::cin >> a.t >> a.q.x >> a.q.y >> a.q.z >> a.q.w >> a.p.x >> a.p.y >> a.p.z >> a.v.x >> a.v.y >> a.v.z;
        a.dq_dbg = read_m3();
    }
}
static void print_align(const xvi::AlignResult &r) {
    print_v3(r.gravity); std::printf("%.17g ", r.scale);
    for (const auto &v : r.velocities) print_v3(v);
    std::printf("\n");
}

int main() {
    std::string cmd;
    while (std::cin >> cmd) {
        if (cmd == "ransac") {
            int n; double thr;
            std::cin >> n >> thr;
            std::vector<V3> a(n), b(n);
            for (int i = 0; i < n; ++i) std::cin >> a[i].x >> a[i].y >> a[i].z >> b[i].x >> b[i].y >> b[i].z;
            std::vector<char> mask;
            M3 R = find_rotation_matrix(a, b, mask, thr);
            for (int k = 0; k < 9; ++k) std::printf("%.17g ", R.m[k]);
            std::printf("\n");
            for (int i = 0; i < n; ++i) std::printf("%d", (int)mask[i]);
            std::printf("\n");
        } else if (cmd == "k") {
            double K[4], px[2];
            std::cin >> K[0] >> K[1] >> K[2] >> K[3] >> px[0] >> px[1];
            V3 bv = remove_k(px, K);
            double back[2];
            apply_k(bv, K, back);
            std::printf("%.17g %.17g %.17g %.17g %.17g\n", bv.x, bv.y, bv.z, back[0], back[1]);
        } else if (cmd == "eig") {
            int n; std::cin >> n;
            xvi::Mat A(n, n);
            for (double &v : A.a) std::cin >> v;
            std::vector<double> wr, wi;
            const bool ok = xvi::eig_values(A, wr, wi);
            std::printf("%d ", (int)ok);
            for (int i = 0; i < n; ++i) std::printf("%.17g %.17g ", wr[i], wi[i]);
            std::printf("\n");
            for (int i = 0; i < n; ++i) if (wi[i] == 0.0) { auto v = xvi::eig_vector(A, wr[i]); for (double x : v) std::printf("%.17g ", x); std::printf("\n"); }
        } else if (cmd == "lstsq") {
            int m, n; std::cin >> m >> n;
            xvi::Mat A(m, n); std::vector<double> b(m);
            for (double &v : A.a) std::cin >> v;
            for (double &v : b) std::cin >> v;
            for (double x : xvi::lstsq(A, b)) std::printf("%.17g ", x);
            std::printf("\n");
        } else if (cmd == "svd") {
            int m, n; std::cin >> m >> n;
            xvi::Mat A(m, n);
            for (double &v : A.a) std::cin >> v;
            std::vector<double> s; xvi::Mat V;
            xvi::svd_jacobi(A, s, V);
            for (double x : s) std::printf("%.17g ", x);
            std::printf("\n");
            for (double x : V.a) std::printf("%.17g ", x);
            std::printf("\n");
        } else if (cmd == "e5") {
            std::array<xvi::V2, 5> a, b;
            for (int i = 0; i < 5; ++i) std::cin >> a[i].x >> a[i].y >> b[i].x >> b[i].y;
            auto sols = xvi::solve_essential_5pt(a, b);
            std::printf("%zu\n", sols.size());
            for (auto &E : sols) { print_m3(E); std::printf("\n"); }
        } else if (cmd == "findE" || cmd == "findH") {
            int n, seed; double thr;
            std::cin >> n >> thr >> seed;
            std::vector<xvi::V2> a, b;
            read_pairs(n, a, b);
            xv::M3 M = xv::m3_identity(); std::vector<char> mask;
            const bool ok = cmd == "findE" ? xvi::find_essential_matrix(a, b, thr, 0.999, 1000, seed, M, mask)
                                           : xvi::find_homography_matrix(a, b, thr, 0.999, 1000, seed, M, mask);
            std::printf("%d ", (int)ok); print_m3(M); std::printf("\n");
            for (int i = 0; i < n; ++i) std::printf("%d", (int)mask[i]);
            std::printf("\n");
        } else if (cmd == "decompH") {
            xv::M3 H = read_m3(), R1, R2; xv::V3 T1, T2, n1, n2;
            const bool ok = xvi::decompose_homography(H, R1, R2, T1, T2, n1, n2);
            std::printf("%d ", (int)ok); print_m3(R1); print_m3(R2); print_v3(T1); print_v3(T2); print_v3(n1); print_v3(n2); std::printf("\n");
        } else if (cmd == "decompE") {
            xv::M3 E = read_m3(), R1, R2; xv::V3 T;
            xvi::decompose_essential(E, R1, R2, T);
            print_m3(R1); print_m3(R2); print_v3(T); std::printf("\n");
        } else if (cmd == "sfm") {
            int n; xvi::SfmSeedConfig cfg;
            std::cin >> n >> cfg.fx >> cfg.min_matches >> cfg.min_parallax >> cfg.min_triangulation >> cfg.seed;
            std::vector<xvi::V2> a(n), b(n); std::vector<double> par(n);
            for (int i = 0; i < n; ++i) std::cin >> a[i].x >> a[i].y >> b[i].x >> b[i].y >> par[i];
            const xvi::SfmSeed s = xvi::sfm_two_view(a, b, par, cfg);
            std::printf("%d %d %.17g ", s.fail, s.candidate, s.parallax);
            for (int k = 0; k < 8; ++k) std::printf("%d ", s.counts[k]);
            std::printf("\n");
            if (!s.fail) {
                print_m3(s.R); print_v3(s.T); std::printf("\n");
                for (size_t i = 0; i < s.status.size(); ++i) std::printf("%d", (int)s.status[i]);
                std::printf("\n");
                for (size_t i = 0; i < s.points.size(); ++i) print_v3(s.points[i]);
                std::printf("\n");
            }
        } else if (cmd == "gyro" || cmd == "gsv" || cmd == "refine") {
            int N; std::cin >> N;
            xv::V3 g = xv::v3(0, 0, 0);
            if (cmd == "refine") std::cin >> g.x >> g.y >> g.z;
            xvi::Extr x; std::vector<xvi::AlignFrame> f; std::vector<xvi::AlignDelta> d;
            read_align(N, x, f, d);
            if (cmd == "gyro") { print_v3(xvi::solve_gyro_bias(f, d, x)); std::printf("\n"); }
            else if (cmd == "gsv") print_align(xvi::solve_gravity_scale_velocity(f, d, x));
            else print_align(xvi::refine_scale_velocity_via_gravity(f, d, x, g));
        } else if (cmd == "quat") {
            xv::M3 R = read_m3(); xv::V3 a, b;
            std::cin >> a.x >> a.y >> a.z >> b.x >> b.y >> b.z;
            const xv::Q4 q = xvi::mat_to_quat(R), r = xvi::from_two_vectors(a, b);
            std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", q.x, q.y, q.z, q.w, r.x, r.y, r.z, r.w);
        } else if (cmd == "tracks") {
            // two frames with 3 keypoints each; a track over both; removing the first frame re-anchors the landmark
            long long tid = 0;
            Map m; m.track_id_counter = &tid;
            Rig rig; rig.q_bc = q4(0, 0, 0, 1); rig.p_bc = v3(0, 0, 0);
            auto f0 = std::make_unique<Frame>(); f0->id = 1;
            auto f1 = std::make_unique<Frame>(); f1->id = 2;
            f1->st.d[4] = 1.0;                                   // moved 1 m along x
            for (int k = 0; k < 3; ++k) { f0->append_keypoint(normalized(v3(0.1 * k, 0, 1))); f1->append_keypoint(normalized(v3(0.1 * k - 0.25, 0, 1))); }
            Frame *p0 = f0.get(), *p1 = f1.get();
            m.attach_frame(std::move(f0)); m.attach_frame(std::move(f1));
            Track *t = frame_get_track(p0, 1, nullptr);
            track_add_keypoint(t, p1, 1);
            set_landmark_point(t, v3(0.4, 0, 4.0), rig);
            t->triangulated = t->valid = true;
            const V3 x0 = get_landmark_point(t, rig);
            std::printf("%zu %zu %d %.12g %.12g %.12g %.12g\n", m.tracks.size(), t->keypoint_num(), t->life, t->inv_depth, x0.x, x0.y, x0.z);
            track_remove_keypoint(t, p0, rig);                   // first frame leaves: 3-D point is preserved (track.cpp:25-37)
            const V3 x1 = get_landmark_point(t, rig);
            std::printf("%zu %zu %lld %.12g %.12g %.12g %.12g\n", m.tracks.size(), t->keypoint_num(), t->first_frame()->id, t->inv_depth, x1.x, x1.y, x1.z);
            track_remove_keypoint(t, p1, rig);                   // last observation gone: the track recycles itself
            std::printf("%zu\n", m.tracks.size());
        }
    }
    return 0;
}
