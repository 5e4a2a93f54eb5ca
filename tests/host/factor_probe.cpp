// CPU probe of the device factor code (xrslam_b200/csrc/factors.cuh compiled for the host): reprojection, rotation
// prior and preintegration residuals / Jacobians, marginalisation error.  Driven by tests/test_host_factors.py.
// Test infrastructure only.
#include "../../xrslam_b200/csrc/factors.cuh"
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>
using namespace xvbe;

static void rd(double *p, int n) { for (int i = 0; i < n; ++i) std::cin >> p[i]; }
static void pr(const double *p, int n) { for (int i = 0; i < n; ++i) std::printf("%.17g ", p[i]); }

int main() {
    std::string cmd;
    while (std::cin >> cmd) {
        RigParams rg{};
        rd(rg.cam_q, 4); rd(rg.cam_p, 3); rd(rg.imu_q, 4); rd(rg.imu_p, 3); rd(rg.sic, 2);
        if (cmd == "rpe") {
            double qt[4], pt[3], qr[4], pr_[3], invd, zt[3], zr[3];
            rd(qt, 4); rd(pt, 3); rd(qr, 4); rd(pr_, 3); std::cin >> invd; rd(zt, 3); rd(zr, 3);
            const M3 T = local_tangent(ld3(zt));
            const double tb[6] = {T(0, 0), T(1, 0), T(2, 0), T(0, 1), T(1, 1), T(2, 1)};
            RpeEval e;
            rpe_eval(ld4(qt), ld3(pt), ld4(qr), ld3(pr_), invd, ld3(zt), ld3(zr), tb, rg, true, true, e);
            pr(e.r, 2); pr(e.Jt, 12); pr(e.Jr, 12); pr(e.Jl, 2); std::printf("\n");
        } else if (cmd == "rot") {
            double qt[4], qr[4], zt[3], zr[3], r[2], J[6];
            rd(qt, 4); rd(qr, 4); rd(zt, 3); rd(zr, 3);
            rot_eval(ld4(qt), ld4(qr), ld3(zt), ld3(zr), rg, true, r, J);
            pr(r, 2); pr(J, 6); std::printf("\n");
        } else if (cmd == "pre") {
            double si[16], sj[16], bias0[6];
            std::vector<double> pb(XRVIO_PRE_BLOCK, 0.0);
            rd(si, 16); rd(sj, 16); rd(bias0, 6);
            std::cin >> pb[PB_T]; rd(&pb[PB_Q], 4); rd(&pb[PB_P], 3); rd(&pb[PB_V], 3);
            rd(&pb[PB_DQ_DBG], 9); rd(&pb[PB_DP_DBG], 9); rd(&pb[PB_DP_DBA], 9); rd(&pb[PB_DV_DBG], 9); rd(&pb[PB_DV_DBA], 9);
            double r[15], Ji[225], Jj[225];
            pre_eval(si, sj, bias0, pb.data(), rg, true, r, Ji, Jj);
            pr(r, 15); pr(Ji, 225); pr(Jj, 225); std::printf("\n");
        } else if (cmd == "marg") {
            double s[16], lin[16], dx[15];
            rd(s, 16); rd(lin, 16);
            M3 Jri;
            marg_dx(s, lin, dx, &Jri);
            pr(dx, 15); pr(Jri.m, 9); std::printf("\n");
        }
    }
    return 0;
}
