// CPU-side probe of the session layer's host logic (xrslam_b200/csrc/session_host.hpp): compiled by
// tests/test_host_logic.py with g++ and compared against the oracle.  Test infrastructure only.
//   stdin : "ransac n thr" followed by n rows of 6 doubles (bearing1, bearing2)
//           "tracks" -> exercises Map / Track bookkeeping and prints a digest
#include "../../xrslam_b200/csrc/session_host.hpp"
#include <cstdio>
#include <iostream>
#include <string>
using namespace xvs;

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
