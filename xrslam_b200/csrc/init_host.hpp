// Host-side geometry of the initializer: the parts of
// xrslam/src/xrslam/core/initializer.cpp that are small dense algebra on <= 200 matches and 8
// keyframes and run once per session.  The iterative solves of the initializer (PnP of the middle
// keyframes, visual BA, final visual-inertial BA), its preintegrations and its N-view
// triangulations are NOT here: they go through the CUDA solve / preintegration / triangulation
// kernels (session_init.inl).
//
//   generic RANSAC                       utility/ransac.h:26-83 (sampling: LotBox, session_host.hpp)
//   4-point homography + decomposition   geometry/homography.cpp:5-158
//   5-point essential (action matrix)    geometry/essential.cpp:105-218,285-297, decomposition :262-283
//   RANSAC wrappers / error measures     geometry/stereo.cpp:40-69,99-126, essential.h:15-20, homography.h:17-21
//   two-view DLT                         geometry/stereo.h:72-82
//   two-view SfM seed                    core/initializer.cpp:157-285
//   gyro bias / gravity-scale-velocity   core/initializer.cpp:398-520
//
// The reference leans on Eigen (JacobiSVD, EigenSolver, fullPivHouseholderQr); this file carries
// its own small routines: one-sided Jacobi SVD, Hessenberg + double-shift QR eigenvalues with
// inverse iteration, pivoted Householder least squares.
#pragma once
#include "session_host.hpp"
#include <array>
#include <cfloat>
#include <functional>

namespace xvi {
using namespace xv;
using xvs::LotBox;

struct V2 { double x, y; };

// ---------------------------------------------------------------------------------------------
// dense helpers (row-major)
// ---------------------------------------------------------------------------------------------
struct Mat {
    int r = 0, c = 0;
    std::vector<double> a;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return a[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};

// One-sided Jacobi (Hestenes): A (m x n) = U diag(s) V^T with the singular values sorted
// descending.  V is n x n orthogonal, so for m < n or rank-deficient A its trailing columns span
// the null space.  W returns A V (columns = s_i u_i).
inline void svd_jacobi(const Mat &A, std::vector<double> &s, Mat &V, Mat *W_out = nullptr) {
    const int m = A.r, n = A.c;
    Mat W = A;
    V = Mat(n, n);
    for (int i = 0; i < n; ++i) V(i, i) = 1.0;
    for (int sweep = 0; sweep < 80; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double al = 0, be = 0, ga = 0;
                for (int k = 0; k < m; ++k) { const double x = W(k, p), y = W(k, q); al += x * x; be += y * y; ga += x * y; }
                if (ga == 0.0 || std::fabs(ga) <= 1e-15 * std::sqrt(al * be)) continue;
                rotated = true;
                const double zeta = (be - al) / (2.0 * ga);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int k = 0; k < m; ++k) { const double x = W(k, p), y = W(k, q); W(k, p) = cs * x - sn * y; W(k, q) = sn * x + cs * y; }
                for (int k = 0; k < n; ++k) { const double x = V(k, p), y = V(k, q); V(k, p) = cs * x - sn * y; V(k, q) = sn * x + cs * y; }
            }
        if (!rotated) break;
    }
    std::vector<double> nrm(n);
    for (int j = 0; j < n; ++j) { double t = 0; for (int k = 0; k < m; ++k) t += W(k, j) * W(k, j); nrm[j] = std::sqrt(t); }
    std::vector<int> ord(n);
    for (int j = 0; j < n; ++j) ord[j] = j;
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return nrm[x] > nrm[y]; });
    Mat Vs(n, n), Ws(m, n);
    s.resize(n);
    for (int j = 0; j < n; ++j) {
        s[j] = nrm[ord[j]];
        for (int k = 0; k < n; ++k) Vs(k, j) = V(k, ord[j]);
        for (int k = 0; k < m; ++k) Ws(k, j) = W(k, ord[j]);
    }
    V = Vs;
    if (W_out) *W_out = Ws;
}

// 3x3 SVD with proper handling of a vanishing third singular value: U, V orthogonal (any determinant sign)
inline void svd3(const M3 &A, M3 &U, double s[3], M3 &V) {
    Mat a(3, 3);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a(i, j) = A(i, j);
    std::vector<double> sv; Mat v, w;
    svd_jacobi(a, sv, v, &w);
    V3 u[3], vv[3];
    for (int j = 0; j < 3; ++j) { vv[j] = v3(v(0, j), v(1, j), v(2, j)); s[j] = sv[j]; }
    for (int j = 0; j < 3; ++j) u[j] = v3(w(0, j), w(1, j), w(2, j));
    u[0] = s[0] > 0 ? (1.0 / s[0]) * u[0] : v3(1, 0, 0);
    if (s[1] > 1e-14 * s[0]) u[1] = normalized(u[1] - dot(u[1], u[0]) * u[0]);
    else { V3 e = std::fabs(u[0].x) < 0.9 ? v3(1, 0, 0) : v3(0, 1, 0); u[1] = normalized(cross(u[0], e)); }
    if (s[2] > 1e-10 * s[0]) { u[2] = (1.0 / s[2]) * u[2]; u[2] = normalized(u[2] - dot(u[2], u[0]) * u[0] - dot(u[2], u[1]) * u[1]); }
    else u[2] = cross(u[0], u[1]);
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) { U(i, j) = u[j][i]; V(i, j) = vv[j][i]; }
}
inline double det3(const M3 &A) {
    return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
           A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
}
inline M3 outer(const V3 &a, const V3 &b) { M3 C; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C(i, j) = a[i] * b[j]; return C; }
inline bool finite3(const M3 &A) { for (int i = 0; i < 9; ++i) if (!std::isfinite(A.m[i])) return false; return true; }

// Real eigenvalues of a general n x n matrix (n <= 16): Householder reduction to Hessenberg form,
// then implicit double-shift QR with deflation.  wi[k] == 0 exactly for the real ones.
inline bool eig_values(Mat A, std::vector<double> &wr, std::vector<double> &wi) {
    const int n = A.r;
    wr.assign(n, 0.0); wi.assign(n, 0.0);
    // Hessenberg: zero column k below the subdiagonal with a reflector applied from both sides
    for (int k = 0; k + 2 < n; ++k) {
        double alpha = 0;
        for (int i = k + 1; i < n; ++i) alpha += A(i, k) * A(i, k);
        alpha = std::sqrt(alpha);
        if (alpha == 0.0) continue;
        if (A(k + 1, k) > 0) alpha = -alpha;
        std::vector<double> v(n, 0.0);
        v[k + 1] = A(k + 1, k) - alpha;
        for (int i = k + 2; i < n; ++i) v[i] = A(i, k);
        double vn = 0;
        for (int i = k + 1; i < n; ++i) vn += v[i] * v[i];
        if (vn == 0.0) continue;
        for (int j = 0; j < n; ++j) {                 // A <- (I - 2 v v^T / vn) A
            double d = 0;
            for (int i = k + 1; i < n; ++i) d += v[i] * A(i, j);
            d = 2.0 * d / vn;
            for (int i = k + 1; i < n; ++i) A(i, j) -= d * v[i];
        }
        for (int i = 0; i < n; ++i) {                 // A <- A (I - 2 v v^T / vn)
            double d = 0;
            for (int j = k + 1; j < n; ++j) d += A(i, j) * v[j];
            d = 2.0 * d / vn;
            for (int j = k + 1; j < n; ++j) A(i, j) -= d * v[j];
        }
        for (int i = k + 2; i < n; ++i) A(i, k) = 0.0;
    }
    double anorm = 0;
    for (int i = 0; i < n; ++i) for (int j = std::max(i - 1, 0); j < n; ++j) anorm += std::fabs(A(i, j));
    int hi = n - 1;
    double shift_acc = 0.0;
    int its = 0;
    while (hi >= 0) {
        int lo = hi;
        for (; lo >= 1; --lo) {                       // deflation point
            double sc = std::fabs(A(lo - 1, lo - 1)) + std::fabs(A(lo, lo));
            if (sc == 0.0) sc = anorm;
            if (std::fabs(A(lo, lo - 1)) + sc == sc) { A(lo, lo - 1) = 0.0; break; }
        }
        double x = A(hi, hi);
        if (lo == hi) { wr[hi] = x + shift_acc; wi[hi] = 0.0; --hi; its = 0; continue; }
        double y = A(hi - 1, hi - 1), w = A(hi, hi - 1) * A(hi - 1, hi);
        if (lo == hi - 1) {                           // 2 x 2 block
            const double p = 0.5 * (y - x), q = p * p + w;
            double z = std::sqrt(std::fabs(q));
            x += shift_acc;
            if (q >= 0.0) {
                z = p + (p >= 0 ? z : -z);
                wr[hi - 1] = wr[hi] = x + z;
                if (z != 0.0) wr[hi] = x - w / z;
                wi[hi - 1] = wi[hi] = 0.0;
            } else { wr[hi - 1] = wr[hi] = x + p; wi[hi - 1] = z; wi[hi] = -z; }
            hi -= 2; its = 0;
            continue;
        }
        if (its >= 60) return false;
        if (its == 10 || its == 20 || its == 40) {    // exceptional shift
            shift_acc += x;
            for (int i = 0; i <= hi; ++i) A(i, i) -= x;
            const double sc = std::fabs(A(hi, hi - 1)) + std::fabs(A(hi - 1, hi - 2));
            x = y = 0.75 * sc; w = -0.4375 * sc * sc;
        }
        ++its;
        int m = hi - 2;
        double p = 0, q = 0, r = 0;
        for (; m >= lo; --m) {                        // start row of the bulge: two consecutive small subdiagonals
            const double z = A(m, m), rr = x - z, ss = y - z;
            p = (rr * ss - w) / A(m + 1, m) + A(m, m + 1);
            q = A(m + 1, m + 1) - z - rr - ss;
            r = A(m + 2, m + 1);
            const double sc = std::fabs(p) + std::fabs(q) + std::fabs(r);
            if (sc != 0.0) { p /= sc; q /= sc; r /= sc; }
            if (m == lo) break;
            const double u = std::fabs(A(m, m - 1)) * (std::fabs(q) + std::fabs(r));
            const double v = std::fabs(p) * (std::fabs(A(m - 1, m - 1)) + std::fabs(z) + std::fabs(A(m + 1, m + 1)));
            if (u + v == v) break;
        }
        for (int i = m + 2; i <= hi; ++i) { A(i, i - 2) = 0.0; if (i != m + 2) A(i, i - 3) = 0.0; }
        for (int k = m; k <= hi - 1; ++k) {           // chase the bulge
            const bool last = (k == hi - 1);
            if (k != m) {
                p = A(k, k - 1); q = A(k + 1, k - 1); r = last ? 0.0 : A(k + 2, k - 1);
                x = std::fabs(p) + std::fabs(q) + std::fabs(r);
                if (x == 0.0) continue;
                p /= x; q /= x; r /= x;
            }
            double sgn = std::sqrt(p * p + q * q + r * r);
            if (p < 0) sgn = -sgn;
            if (sgn == 0.0) continue;
            if (k == m) { if (lo != m) A(k, k - 1) = -A(k, k - 1); }
            else A(k, k - 1) = -sgn * x;
            p += sgn; x = p / sgn; y = q / sgn;
            double z = r / sgn;
            q /= p; r /= p;
            for (int j = k; j <= hi; ++j) {
                double t = A(k, j) + q * A(k + 1, j);
                if (!last) { t += r * A(k + 2, j); A(k + 2, j) -= t * z; }
                A(k + 1, j) -= t * y; A(k, j) -= t * x;
            }
            const int mm = std::min(hi, k + 3);
            for (int i = lo; i <= mm; ++i) {
                double t = x * A(i, k) + y * A(i, k + 1);
                if (!last) { t += z * A(i, k + 2); A(i, k + 2) -= t * r; }
                A(i, k + 1) -= t * q; A(i, k) -= t;
            }
        }
    }
    return true;
}

// LU with partial pivoting; solves A x = b in place (n <= 32).  Zero pivots are nudged so the
// routine can be used for inverse iteration on a (numerically) singular matrix.
inline void lu_solve(Mat A, std::vector<double> &b) {
    const int n = A.r;
    double scale = 0;
    for (double v : A.a) scale = std::max(scale, std::fabs(v));
    const double tiny = (scale > 0 ? scale : 1.0) * DBL_EPSILON;
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i) if (std::fabs(A(i, k)) > std::fabs(A(piv, k))) piv = i;
        if (piv != k) { for (int j = 0; j < n; ++j) std::swap(A(k, j), A(piv, j)); std::swap(b[k], b[piv]); }
        if (std::fabs(A(k, k)) < tiny) A(k, k) = (A(k, k) < 0 ? -tiny : tiny);
        for (int i = k + 1; i < n; ++i) {
            const double f = A(i, k) / A(k, k);
            if (f == 0.0) continue;
            for (int j = k; j < n; ++j) A(i, j) -= f * A(k, j);
            b[i] -= f * b[k];
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double t = b[i];
        for (int j = i + 1; j < n; ++j) t -= A(i, j) * b[j];
        b[i] = t / A(i, i);
    }
}

// eigenvector of A for the (real) eigenvalue lambda by inverse iteration
inline std::vector<double> eig_vector(const Mat &A, double lambda) {
    const int n = A.r;
    Mat B = A;
    for (int i = 0; i < n; ++i) B(i, i) -= lambda;
    std::vector<double> x(n);
    for (int i = 0; i < n; ++i) x[i] = 1.0 / std::sqrt((double)n) * ((i % 2) ? 0.7 : 1.0);
    for (int it = 0; it < 3; ++it) {
        lu_solve(B, x);
        double nn = 0;
        for (double v : x) nn += v * v;
        nn = std::sqrt(nn);
        if (!(nn > 0) || !std::isfinite(nn)) break;
        for (double &v : x) v /= nn;
    }
    return x;
}

// min ||A x - b||: Householder QR with column pivoting, rank-revealing (fullPivHouseholderQr().solve
// of the reference gives the same x for the full-column-rank systems of the initializer)
inline std::vector<double> lstsq(Mat A, std::vector<double> b) {
    const int m = A.r, n = A.c;
    std::vector<int> perm(n);
    for (int j = 0; j < n; ++j) perm[j] = j;
    std::vector<double> cn(n);
    int rank = 0;
    double first = 0;
    const int steps = std::min(m, n);
    for (int k = 0; k < steps; ++k) {
        int best = k; double bn = -1;
        for (int j = k; j < n; ++j) {
            double t = 0;
            for (int i = k; i < m; ++i) t += A(i, j) * A(i, j);
            cn[j] = t;
            if (t > bn) { bn = t; best = j; }
        }
        if (k == 0) first = bn;
        if (!(bn > first * 1e-26)) break;
        if (best != k) { for (int i = 0; i < m; ++i) std::swap(A(i, k), A(i, best)); std::swap(perm[k], perm[best]); }
        double alpha = std::sqrt(bn);
        if (A(k, k) > 0) alpha = -alpha;
        std::vector<double> v(m, 0.0);
        v[k] = A(k, k) - alpha;
        for (int i = k + 1; i < m; ++i) v[i] = A(i, k);
        double vn = 0;
        for (int i = k; i < m; ++i) vn += v[i] * v[i];
        if (vn > 0) {
            for (int j = k; j < n; ++j) {
                double d = 0;
                for (int i = k; i < m; ++i) d += v[i] * A(i, j);
                d = 2.0 * d / vn;
                for (int i = k; i < m; ++i) A(i, j) -= d * v[i];
            }
            double d = 0;
            for (int i = k; i < m; ++i) d += v[i] * b[i];
            d = 2.0 * d / vn;
            for (int i = k; i < m; ++i) b[i] -= d * v[i];
        }
        rank = k + 1;
    }
    std::vector<double> y(n, 0.0), x(n, 0.0);
    for (int i = rank - 1; i >= 0; --i) {
        double t = b[i];
        for (int j = i + 1; j < rank; ++j) t -= A(i, j) * y[j];
        y[i] = t / A(i, i);
    }
    for (int j = 0; j < n; ++j) x[perm[j]] = y[j];
    return x;
}

// symmetric 3x3 least-squares solve through the eigen-decomposition (JacobiSVD<3x3>::solve)
inline V3 solve_sym3(const M3 &A, const V3 &b) {
    double a[3][3], v[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i][j] = A(i, j);
    xvs::jacobi_eig3(a, v);
    double mx = std::max(std::fabs(a[0][0]), std::max(std::fabs(a[1][1]), std::fabs(a[2][2])));
    V3 x = v3(0, 0, 0);
    for (int k = 0; k < 3; ++k) {
        if (!(std::fabs(a[k][k]) > mx * 3 * DBL_EPSILON)) continue;
        const V3 e = v3(v[0][k], v[1][k], v[2][k]);
        x = x + (dot(e, b) / a[k][k]) * e;
    }
    return x;
}

// ---------------------------------------------------------------------------------------------
// generic RANSAC (utility/ransac.h:26-83); the model is a 3x3 matrix
// ---------------------------------------------------------------------------------------------
template <int DoF, class Solver, class MakeEval>
inline bool ransac(const std::vector<V2> &d1, const std::vector<V2> &d2, double threshold, double confidence, size_t max_iteration, int seed,
                   Solver solver, MakeEval make_eval, M3 &model, std::vector<char> &mask) {
    const size_t n = d1.size();
    mask.assign(n, 0);
    bool have = false;
    if (n < (size_t)DoF) return false;                      // the reference returns an uninitialised model here
    LotBox lot(n);
    lot.seed((unsigned)seed);
    const double K = std::log(std::max(1 - confidence, 1.0e-5));
    size_t best = 0, iter_max = max_iteration;
    std::vector<char> cur(n);
    for (size_t it = 0; it < iter_max; ++it) {
        lot.refill_all();
        std::array<V2, DoF> s1, s2;
        for (int k = 0; k < DoF; ++k) { const size_t i = lot.draw_without_replacement(); s1[k] = d1[i]; s2[k] = d2[i]; }
        std::vector<M3> models = solver(s1, s2);
        for (const M3 &cand : models) {
            auto eval = make_eval(cand);
            size_t cnt = 0;
            for (size_t i = 0; i < n; ++i) {
                const double e = eval(d1[i], d2[i]);
                cur[i] = (e <= threshold) ? 1 : 0;          // NaN compares false
                cnt += cur[i];
            }
            if (cnt > best) {
                model = cand; best = cnt; mask = cur; have = true;
                const double ratio = best / (double)n;
                const double N = K / std::log(1 - std::pow(ratio, 5));
                if (N < (double)iter_max) iter_max = (size_t)std::ceil(N);
            }
        }
    }
    return have;
}

// ---------------------------------------------------------------------------------------------
// homography (geometry/homography.cpp)
// ---------------------------------------------------------------------------------------------
inline M3 segments_to_matrix(const double *v) {            // homography.cpp:82-86: the three segments are COLUMNS
    M3 A;
    for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) A(k, j) = v[j * 3 + k];
    return A;
}

inline M3 solve_homography_4pt(const std::array<V2, 4> &pa, const std::array<V2, 4> &pb) {
    V2 ma{0, 0}, mb{0, 0};
    for (int i = 0; i < 4; ++i) { ma.x += pa[i].x; ma.y += pa[i].y; mb.x += pb[i].x; mb.y += pb[i].y; }
    ma.x /= 4; ma.y /= 4; mb.x /= 4; mb.y /= 4;
    double sa = 0, sb = 0;
    for (int i = 0; i < 4; ++i) { sa += std::hypot(pa[i].x - ma.x, pa[i].y - ma.y); sb += std::hypot(pb[i].x - mb.x, pb[i].y - mb.y); }
    sa = 1.0 / (std::sqrt(2.0) * sa); sb = 1.0 / (std::sqrt(2.0) * sb);
    Mat A(8, 9);
    bool ok = std::isfinite(sa) && std::isfinite(sb);
    for (int i = 0; i < 4; ++i) {
        const double a0 = (pa[i].x - ma.x) * sa, a1 = (pa[i].y - ma.y) * sa, b0 = (pb[i].x - mb.x) * sb, b1 = (pb[i].y - mb.y) * sb;
        A(2 * i, 1) = -a0; A(2 * i, 2) = a0 * b1; A(2 * i, 4) = -a1; A(2 * i, 5) = a1 * b1; A(2 * i, 7) = -1; A(2 * i, 8) = b1;
        A(2 * i + 1, 0) = a0; A(2 * i + 1, 2) = -a0 * b0; A(2 * i + 1, 3) = a1; A(2 * i + 1, 5) = -a1 * b0; A(2 * i + 1, 6) = 1; A(2 * i + 1, 8) = -b0;
    }
    if (!ok) { M3 bad; for (int i = 0; i < 9; ++i) bad.m[i] = NAN; return bad; }
    std::vector<double> s; Mat V;
    svd_jacobi(A, s, V);
    double h[9];
    for (int k = 0; k < 9; ++k) h[k] = V(k, 8);
    const M3 NH = segments_to_matrix(h);
    M3 Na = m3_identity(), Nb = m3_identity();
    Nb(0, 0) = 1 / sb; Nb(0, 2) = mb.x; Nb(1, 1) = 1 / sb; Nb(1, 2) = mb.y;
    Na(0, 0) = sa; Na(0, 2) = -sa * ma.x; Na(1, 1) = sa; Na(1, 2) = -sa * ma.y;
    return Nb * NH * Na;
}

inline double homography_geometric_error(const M3 &H, const V2 &p1, const V2 &p2) {      // homography.h:17-21
    const V3 q = H * v3(p1.x, p1.y, 1.0);
    const double dx = p2.x - q.x / q.z, dy = p2.y - q.y / q.z;
    return dx * dx + dy * dy;
}

inline bool find_homography_matrix(const std::vector<V2> &p1, const std::vector<V2> &p2, double threshold, double confidence, size_t max_iteration,
                                   int seed, M3 &H, std::vector<char> &mask) {               // stereo.cpp:99-126
    auto solver = [](const std::array<V2, 4> &a, const std::array<V2, 4> &b) { return std::vector<M3>{solve_homography_4pt(a, b)}; };
    auto make_eval = [](const M3 &Hc) {
        const M3 Hinv = inverse(Hc);
        return [Hc, Hinv](const V2 &a, const V2 &b) { return homography_geometric_error(Hc, a, b) + homography_geometric_error(Hinv, b, a); };
    };
    return ransac<4>(p1, p2, 2.0 * 5.99 * threshold * threshold, confidence, max_iteration, seed, solver, make_eval, H, mask);
}

// homography.cpp:5-78; returns false for a pure rotation
inline bool decompose_homography(const M3 &H, M3 &R1, M3 &R2, V3 &T1, V3 &T2, V3 &n1, V3 &n2) {
    M3 U, V; double sv[3];
    svd3(H, U, sv, V);
    const M3 Hn = (1.0 / sv[1]) * H;
    const M3 S = transpose(Hn) * Hn - m3_identity();
    bool pure = true;
    for (int i = 0; i < 9; ++i) if (std::fabs(S.m[i]) > 1e-3) pure = false;
    if (pure) {
        R1 = U * transpose(V);
        if (det3(R1) < 0) R1 = -R1;
        R2 = R1;
        T1 = T2 = n1 = n2 = v3(0, 0, 0);
        return false;
    }
    const double Ms00 = S(1, 2) * S(1, 2) - S(1, 1) * S(2, 2);
    const double Ms11 = S(0, 2) * S(0, 2) - S(0, 0) * S(2, 2);
    const double Ms22 = S(0, 1) * S(0, 1) - S(0, 0) * S(1, 1);
    const double s00 = std::sqrt(Ms00), s11 = std::sqrt(Ms11), s22 = std::sqrt(Ms22);
    const double tr = S(0, 0) + S(1, 1) + S(2, 2);
    const double nu = 2.0 * std::sqrt(1 + tr - Ms00 - Ms11 - Ms22);
    const double tenormsq = 2 + tr - nu;
    V3 t1, t2;
    if (S(0, 0) > S(1, 1) && S(0, 0) > S(2, 2)) {
        const double e = ((S(0, 1) * S(0, 2) - S(0, 0) * S(1, 2)) < 0) ? -1 : 1;
        n1 = v3(S(0, 0), S(0, 1) + s22, S(0, 2) + e * s11);
        n2 = v3(S(0, 0), S(0, 1) - s22, S(0, 2) - e * s11);
        t1 = (norm(n1) / S(0, 0)) * n2;
        t2 = (norm(n2) / S(0, 0)) * n1;
    } else if (S(1, 1) > S(0, 0) && S(1, 1) > S(2, 2)) {
        const double e = ((S(1, 1) * S(0, 2) - S(0, 1) * S(1, 2)) < 0) ? -1 : 1;
        n1 = v3(S(0, 1) + s22, S(1, 1), S(1, 2) - e * s00);
        n2 = v3(S(0, 1) - s22, S(1, 1), S(1, 2) + e * s00);
        t2 = (norm(n2) / S(1, 1)) * n1;
        t1 = (norm(n1) / S(1, 1)) * n2;
    } else {
        const double e = ((S(1, 2) * S(0, 2) - S(0, 1) * S(2, 2)) < 0) ? -1 : 1;
        n1 = v3(S(0, 2) + e * s11, S(1, 2) + s00, S(2, 2));
        n2 = v3(S(0, 2) - e * s11, S(1, 2) - s00, S(2, 2));
        t1 = (norm(n1) / S(2, 2)) * n2;
        t2 = (norm(n2) / S(2, 2)) * n1;
    }
    n1 = normalized(n1); n2 = normalized(n2);
    t1 = t1 - tenormsq * n1; t2 = t2 - tenormsq * n2;
    R1 = Hn * (m3_identity() - outer((1.0 / nu) * t1, n1));
    R2 = Hn * (m3_identity() - outer((1.0 / nu) * t2, n2));
    T1 = R1 * (0.5 * t1);
    T2 = R2 * (0.5 * t2);
    return true;
}

// ---------------------------------------------------------------------------------------------
// essential matrix, 5-point (geometry/essential.cpp)
// ---------------------------------------------------------------------------------------------
// graded reverse lexicographic monomials up to degree 3 in (x, y, z), essential.cpp:13-16
enum { XXX = 0, XXY, XYY, YYY, XXZ, XYZ, YYZ, XZZ, YZZ, ZZZ, XX, XY, YY, XZ, YZ, ZZ, MX, MY, MZ, MI, NMONO };

struct MonoTable {
    int ex[NMONO][3];
    int prod[NMONO][NMONO];        // index of the product monomial or -1 when its degree exceeds 3
    MonoTable() {
        static const int e[NMONO][3] = {{3, 0, 0}, {2, 1, 0}, {1, 2, 0}, {0, 3, 0}, {2, 0, 1}, {1, 1, 1}, {0, 2, 1}, {1, 0, 2}, {0, 1, 2}, {0, 0, 3},
                                        {2, 0, 0}, {1, 1, 0}, {0, 2, 0}, {1, 0, 1}, {0, 1, 1}, {0, 0, 2}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
        for (int i = 0; i < NMONO; ++i) for (int k = 0; k < 3; ++k) ex[i][k] = e[i][k];
        for (int i = 0; i < NMONO; ++i)
            for (int j = 0; j < NMONO; ++j) {
                prod[i][j] = -1;
                const int a = ex[i][0] + ex[j][0], b = ex[i][1] + ex[j][1], c = ex[i][2] + ex[j][2];
                if (a + b + c > 3) continue;
                for (int k = 0; k < NMONO; ++k) if (ex[k][0] == a && ex[k][1] == b && ex[k][2] == c) prod[i][j] = k;
            }
    }
};
inline const MonoTable &mono_table() { static const MonoTable t; return t; }

struct Poly {
    double v[NMONO];
    Poly() { for (double &x : v) x = 0; }
    static Poly linear(double x, double y, double z, double w) { Poly p; p.v[MX] = x; p.v[MY] = y; p.v[MZ] = z; p.v[MI] = w; return p; }
    Poly operator+(const Poly &b) const { Poly r; for (int i = 0; i < NMONO; ++i) r.v[i] = v[i] + b.v[i]; return r; }
    Poly operator-(const Poly &b) const { Poly r; for (int i = 0; i < NMONO; ++i) r.v[i] = v[i] - b.v[i]; return r; }
    Poly operator*(const Poly &b) const {             // truncated at total degree 3 (essential.cpp:50-93)
        const MonoTable &t = mono_table();
        Poly r;
        for (int i = 0; i < NMONO; ++i) {
            if (v[i] == 0.0) continue;
            for (int j = 0; j < NMONO; ++j) { const int k = t.prod[i][j]; if (k >= 0) r.v[k] += v[i] * b.v[j]; }
        }
        return r;
    }
    Poly scaled(double s) const { Poly r; for (int i = 0; i < NMONO; ++i) r.v[i] = s * v[i]; return r; }
};

inline std::vector<M3> solve_essential_5pt(const std::array<V2, 5> &p1, const std::array<V2, 5> &p2) {
    // null space of the five epipolar constraints (essential.cpp:105-117)
    Mat A(5, 9);
    for (int i = 0; i < 5; ++i) {
        const double a[3] = {p1[i].x, p1[i].y, 1.0}, b[3] = {p2[i].x, p2[i].y, 1.0};
        for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) A(i, j * 3 + k) = a[j] * b[k];
    }
    for (double x : A.a) if (!std::isfinite(x)) return {};
    std::vector<double> sv; Mat V;
    svd_jacobi(A, sv, V);
    double basis[4][9];
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 9; ++k) basis[c][k] = V(k, 5 + c);
    M3 Eb[4];
    for (int c = 0; c < 4; ++c) Eb[c] = segments_to_matrix(basis[c]);
    // ten cubic constraints (essential.cpp:119-149)
    Poly E[3][3], Et[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { E[i][j] = Poly::linear(Eb[0](i, j), Eb[1](i, j), Eb[2](i, j), Eb[3](i, j)); Et[j][i] = E[i][j]; }
    auto matmul = [](const Poly X[3][3], const Poly Y[3][3], Poly Z[3][3]) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { Poly acc; for (int k = 0; k < 3; ++k) acc = acc + X[i][k] * Y[k][j]; Z[i][j] = acc; }
    };
    Poly EEt[3][3], EEtE[3][3];
    matmul(E, Et, EEt);
    matmul(EEt, E, EEtE);
    const Poly half_trace = (EEt[0][0] + EEt[1][1] + EEt[2][2]).scaled(0.5);
    Mat P(10, 20);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { const Poly c = EEtE[i][j] - half_trace * E[i][j]; for (int k = 0; k < 20; ++k) P(i * 3 + j, k) = c.v[k]; }
    {
        const Poly det = E[0][0] * (E[1][1] * E[2][2] - E[1][2] * E[2][1]) - E[0][1] * (E[1][0] * E[2][2] - E[1][2] * E[2][0]) +
                         E[0][2] * (E[1][0] * E[2][1] - E[1][1] * E[2][0]);
        for (int k = 0; k < 20; ++k) P(9, k) = det.v[k];
    }
    // Gauss-Jordan on the cubic monomials with the reference's pivot search (essential.cpp:151-176)
    int perm[10];
    for (int i = 0; i < 10; ++i) perm[i] = i;
    for (int i = 0; i < 10; ++i) {
        for (int j = i + 1; j < 10; ++j) if (std::fabs(P(perm[i], i)) < std::fabs(P(perm[j], i))) std::swap(perm[i], perm[j]);
        if (P(perm[i], i) == 0) continue;
        const double d = P(perm[i], i);
        for (int k = 0; k < 20; ++k) P(perm[i], k) /= d;
        for (int j = i + 1; j < 10; ++j) { const double f = P(perm[j], i); for (int k = 0; k < 20; ++k) P(perm[j], k) -= P(perm[i], k) * f; }
    }
    for (int i = 9; i > 0; --i)
        for (int j = 0; j < i; ++j) { const double f = P(perm[j], i); for (int k = 0; k < 20; ++k) P(perm[j], k) -= P(perm[i], k) * f; }
    // action matrix of multiplication by x on the quotient basis (essential.cpp:178-199)
    Mat act(10, 10);
    static const int rows[6] = {XXX, XXY, XYY, XXZ, XYZ, XZZ};
    for (int r = 0; r < 6; ++r) for (int k = 0; k < 10; ++k) act(r, k) = -P(perm[rows[r]], XX + k);
    act(6, XX - XX) = 1; act(7, XY - XX) = 1; act(8, XZ - XX) = 1; act(9, MX - XX) = 1;
    for (double x : act.a) if (!std::isfinite(x)) return {};
    std::vector<double> wr, wi;
    std::vector<M3> out;
    if (!eig_values(act, wr, wi)) return out;
    for (int i = 0; i < 10; ++i) {
        if (std::fabs(wi[i]) >= 1.0e-10) continue;
        const std::vector<double> h = eig_vector(act, wr[i]);
        const double w = h[MI - XX];
        const double x = h[MX - XX] / w, y = h[MY - XX] / w, z = h[MZ - XX] / w;
        double e[9];
        for (int k = 0; k < 9; ++k) e[k] = basis[0][k] * x + basis[1][k] * y + basis[2][k] * z + basis[3][k];
        out.push_back(segments_to_matrix(e));
    }
    return out;
}

inline double essential_geometric_error(const M3 &E, const V2 &p1, const V2 &p2) {        // essential.h:15-20
    const V3 Ep1 = E * v3(p1.x, p1.y, 1.0);
    const double r = p2.x * Ep1.x + p2.y * Ep1.y + Ep1.z;
    return r * r / (Ep1.x * Ep1.x + Ep1.y * Ep1.y);
}

inline bool find_essential_matrix(const std::vector<V2> &p1, const std::vector<V2> &p2, double threshold, double confidence, size_t max_iteration,
                                  int seed, M3 &E, std::vector<char> &mask) {                // stereo.cpp:40-69
    auto solver = [](const std::array<V2, 5> &a, const std::array<V2, 5> &b) { return solve_essential_5pt(a, b); };
    auto make_eval = [](const M3 &Ec) {
        const M3 Et = transpose(Ec);
        return [Ec, Et](const V2 &a, const V2 &b) { return essential_geometric_error(Ec, a, b) + essential_geometric_error(Et, b, a); };
    };
    return ransac<5>(p1, p2, 2.0 * 3.84 * threshold * threshold, confidence, max_iteration, seed, solver, make_eval, E, mask);
}

inline void decompose_essential(const M3 &E, M3 &R1, M3 &R2, V3 &T) {                       // essential.cpp:262-283
    M3 U, V; double s[3];
    svd3(E, U, s, V);
    M3 Vt = transpose(V);
    if (det3(U) < 0) U = -U;
    if (det3(Vt) < 0) Vt = -Vt;
    M3 W; for (int i = 0; i < 9; ++i) W.m[i] = 0;
    W(0, 1) = 1; W(1, 0) = -1; W(2, 2) = 1;
    R1 = U * W * Vt;
    R2 = U * transpose(W) * Vt;
    T = v3(U(0, 2), U(1, 2), U(2, 2));
}

// two-view DLT (stereo.h:72-82); P = [R | T] row-major 3x4
inline void triangulate_point_2v(const double P1[12], const double P2[12], const V3 &x1, const V3 &x2, double q[4]) {
    Mat A(4, 4);
    for (int k = 0; k < 4; ++k) {
        A(0, k) = x1.x * P1[8 + k] - x1.z * P1[k];
        A(1, k) = x1.y * P1[8 + k] - x1.z * P1[4 + k];
        A(2, k) = x2.x * P2[8 + k] - x2.z * P2[k];
        A(3, k) = x2.y * P2[8 + k] - x2.z * P2[4 + k];
    }
    std::vector<double> s; Mat V;
    svd_jacobi(A, s, V);
    for (int k = 0; k < 4; ++k) q[k] = V(k, 3);
}

// ---------------------------------------------------------------------------------------------
// Eigen conventions needed when states are written back
// ---------------------------------------------------------------------------------------------
inline Q4 mat_to_quat(const M3 &R) {                       // Eigen::Quaterniond(Matrix3d)
    double t = R(0, 0) + R(1, 1) + R(2, 2);
    double q[4];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[3] = 0.5 * t; t = 0.5 / t;
        q[0] = (R(2, 1) - R(1, 2)) * t; q[1] = (R(0, 2) - R(2, 0)) * t; q[2] = (R(1, 0) - R(0, 1)) * t;
    } else {
        int i = 0;
        if (R(1, 1) > R(0, 0)) i = 1;
        if (R(2, 2) > R(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
        q[i] = 0.5 * t; t = 0.5 / t;
        q[3] = (R(k, j) - R(j, k)) * t; q[j] = (R(j, i) + R(i, j)) * t; q[k] = (R(k, i) + R(i, k)) * t;
    }
    return q4(q[0], q[1], q[2], q[3]);
}
inline Q4 from_two_vectors(const V3 &a, const V3 &b) {     // Eigen::Quaterniond::FromTwoVectors
    const V3 v0 = normalized(a), v1 = normalized(b);
    double c = dot(v1, v0);
    if (c < -1.0 + 1e-12) {
        c = std::max(c, -1.0);
        Mat m(2, 3);
        for (int k = 0; k < 3; ++k) { m(0, k) = v0[k]; m(1, k) = v1[k]; }
        std::vector<double> s; Mat V;
        svd_jacobi(m, s, V);
        const V3 axis = v3(V(0, 2), V(1, 2), V(2, 2));
        const double w2 = (1.0 + c) * 0.5;
        const double k = std::sqrt(1.0 - w2);
        return q4(axis.x * k, axis.y * k, axis.z * k, std::sqrt(w2));
    }
    const V3 axis = cross(v0, v1);
    const double s = std::sqrt((1.0 + c) * 2.0);
    return q4(axis.x / s, axis.y / s, axis.z / s, s * 0.5);
}

// ---------------------------------------------------------------------------------------------
// two-view SfM seed (initializer.cpp:157-285)
// ---------------------------------------------------------------------------------------------
struct SfmSeedConfig {
    double fx = 458.654;
    int min_matches = 50, min_triangulation = 20, seed = 648;
    double min_parallax = 10.0;
};
struct SfmSeed {
    int fail = 0;                 // 0 ok, 1 matches, 2 parallax, 3 pure rotation, 4 triangulation
    double parallax = 0;
    int candidate = -1;
    int counts[8] = {0};
    M3 R; V3 T;
    std::vector<V3> points;
    std::vector<char> status;
};

// pts_i / pts_j: hnormalized keypoints of the first / last keyframe; parallax_px: per-match pixel distance
inline SfmSeed sfm_two_view(const std::vector<V2> &pts_i, const std::vector<V2> &pts_j, const std::vector<double> &parallax_px, const SfmSeedConfig &cfg) {
    SfmSeed out;
    const size_t n = pts_i.size();
    if ((int)n < cfg.min_matches) { out.fail = 1; return out; }
    double total = 0;
    for (double p : parallax_px) total += p;
    total /= std::max<size_t>(n, 1);
    out.parallax = total;
    if (total < cfg.min_parallax) { out.fail = 2; return out; }
    std::vector<M3> Rs; std::vector<V3> Ts;
    std::vector<char> mask;
    M3 H = m3_identity();
    find_homography_matrix(pts_i, pts_j, 0.7 / cfg.fx, 0.999, 1000, cfg.seed, H, mask);
    M3 RH1, RH2; V3 TH1, TH2, nH1, nH2;
    if (!decompose_homography(H, RH1, RH2, TH1, TH2, nH1, nH2)) { out.fail = 3; return out; }
    TH1 = normalized(TH1); TH2 = normalized(TH2);
    Rs.insert(Rs.end(), {RH1, RH1, RH2, RH2});
    Ts.insert(Ts.end(), {TH1, -TH1, TH2, -TH2});
    M3 E = m3_identity();
    find_essential_matrix(pts_i, pts_j, 0.7 / cfg.fx, 0.999, 1000, cfg.seed, E, mask);
    M3 RE1, RE2; V3 TE;
    decompose_essential(E, RE1, RE2, TE);
    TE = normalized(TE);
    Rs.insert(Rs.end(), {RE1, RE1, RE2, RE2});
    Ts.insert(Ts.end(), {TE, -TE, TE, -TE});

    std::vector<std::vector<V3>> cpts(8);
    std::vector<std::vector<char>> cst(8);
    double scores[8];
    size_t best = 0;
    const double P1[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    for (size_t c = 0; c < 8; ++c) {
        double P2[12];
        for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) P2[4 * r + k] = Rs[c](r, k); P2[4 * r + 3] = Ts[c][r]; }
        cpts[c].assign(n, v3(0, 0, 0)); cst[c].assign(n, 0);
        int count = 0; double score = 0;
        for (size_t j = 0; j < n; ++j) {
            double q[4];
            triangulate_point_2v(P1, P2, v3(pts_i[j].x, pts_i[j].y, 1.0), v3(pts_j[j].x, pts_j[j].y, 1.0), q);
            const V3 q1 = v3(q[0], q[1], q[2]);
            const V3 q2 = v3(P2[0] * q[0] + P2[1] * q[1] + P2[2] * q[2] + P2[3] * q[3], P2[4] * q[0] + P2[5] * q[1] + P2[6] * q[2] + P2[7] * q[3],
                             P2[8] * q[0] + P2[9] * q[1] + P2[10] * q[2] + P2[11] * q[3]);
            if (q1.z * q[3] > 0 && q2.z * q[3] > 0 && q1.z / q[3] < 100 && q2.z / q[3] < 100) {
                cpts[c][j] = v3(q[0] / q[3], q[1] / q[3], q[2] / q[3]);
                cst[c][j] = 1;
                count++;
                const double ax = q1.x / q1.z - pts_i[j].x, ay = q1.y / q1.z - pts_i[j].y, bx = q2.x / q2.z - pts_j[j].x, by = q2.y / q2.z - pts_j[j].y;
                score += 0.5 * (ax * ax + ay * ay + bx * bx + by * by);
            }
        }
        out.counts[c] = count; scores[c] = score;
        if (out.counts[c] > cfg.min_triangulation && scores[c] < scores[best]) best = c;      // selection rule of :262-268
        else if (out.counts[c] > out.counts[best]) best = c;
    }
    out.candidate = (int)best;
    out.R = Rs[best]; out.T = Ts[best];
    out.points.swap(cpts[best]); out.status.swap(cst[best]);
    if (out.counts[best] < cfg.min_triangulation) { out.fail = 4; return out; }
    return out;
}

// ---------------------------------------------------------------------------------------------
// visual-inertial alignment (initializer.cpp:398-520).  Inputs per keyframe k: body pose (q, p) of the
// up-to-scale SfM; per edge k-1 -> k: preintegrated delta (t, q, p, v) and dq_dbg.
// ---------------------------------------------------------------------------------------------
struct AlignFrame { Q4 q; V3 p; };
struct AlignDelta { double t; Q4 q; V3 p, v; M3 dq_dbg; };
struct Extr { Q4 q_bc; V3 p_bc; Q4 q_bi; V3 p_bi; };

inline V3 solve_gyro_bias(const std::vector<AlignFrame> &f, const std::vector<AlignDelta> &d, const Extr &x) {     // :398-423
    M3 A; for (int i = 0; i < 9; ++i) A.m[i] = 0;
    V3 b = v3(0, 0, 0);
    for (size_t j = 1; j < f.size(); ++j) {
        const Q4 qi = qmul(f[j - 1].q, x.q_bi), qj = qmul(f[j].q, x.q_bi);
        const M3 &J = d[j - 1].dq_dbg;
        A = A + transpose(J) * J;
        b = b + transpose(J) * logmap(qmul(qconj(qmul(qi, d[j - 1].q)), qj));
    }
    return solve_sym3(A, b);
}

struct AlignResult { V3 gravity; double scale; std::vector<V3> velocities; };

inline AlignResult solve_gravity_scale_velocity(const std::vector<AlignFrame> &f, const std::vector<AlignDelta> &d, const Extr &x) {   // :425-466
    const int N = (int)f.size();
    Mat A((N - 1) * 6, 3 + 1 + 3 * N);
    std::vector<double> b((N - 1) * 6, 0.0);
    for (int j = 1; j < N; ++j) {
        const int i = j - 1;
        const AlignDelta &dl = d[i];
        const V3 pci = f[i].p + qrot(f[i].q, x.p_bc), pcj = f[j].p + qrot(f[j].q, x.p_bc);
        const V3 dp = pcj - pci;
        const V3 b0 = qrot(f[i].q, dl.p) + (qrot(f[j].q, x.p_bc) - qrot(f[i].q, x.p_bc));
        const V3 b1 = qrot(f[i].q, dl.v);
        for (int k = 0; k < 3; ++k) {
            A(i * 6 + k, k) = -0.5 * dl.t * dl.t;
            A(i * 6 + k, 3) = dp[k];
            A(i * 6 + k, 4 + i * 3 + k) = -dl.t;
            b[i * 6 + k] = b0[k];
            A(i * 6 + 3 + k, k) = -dl.t;
            A(i * 6 + 3 + k, 4 + i * 3 + k) = -1;
            A(i * 6 + 3 + k, 4 + j * 3 + k) = 1;
            b[i * 6 + 3 + k] = b1[k];
        }
    }
    const std::vector<double> s = lstsq(A, b);
    AlignResult r;
    r.gravity = 9.80665 * normalized(v3(s[0], s[1], s[2]));
    r.scale = s[3];
    for (int i = 0; i < N; ++i) r.velocities.push_back(v3(s[4 + 3 * i], s[5 + 3 * i], s[6 + 3 * i]));
    return r;
}

inline AlignResult refine_scale_velocity_via_gravity(const std::vector<AlignFrame> &f, const std::vector<AlignDelta> &d, const Extr &x, const V3 &gravity) {   // :468-520
    const double damp = 0.1;
    const int N = (int)f.size();
    Mat A((N - 1) * 6, 2 + 1 + 3 * N);
    std::vector<double> b((N - 1) * 6, 0.0);
    const M3 Tg = local_tangent(gravity);          // columns 0, 1 = s2_tangential_basis(gravity)
    for (int j = 1; j < N; ++j) {
        const int i = j - 1;
        const AlignDelta &dl = d[i];
        const V3 pci = f[i].p + qrot(f[i].q, x.p_bc), pcj = f[j].p + qrot(f[j].q, x.p_bc);
        const V3 dp = pcj - pci;
        const V3 b0 = (0.5 * dl.t * dl.t) * gravity + qrot(f[i].q, dl.p) + (qrot(f[j].q, x.p_bc) - qrot(f[i].q, x.p_bc));
        const V3 b1 = dl.t * gravity + qrot(f[i].q, dl.v);
        for (int k = 0; k < 3; ++k) {
            A(i * 6 + k, 0) = -0.5 * dl.t * dl.t * Tg(k, 0);
            A(i * 6 + k, 1) = -0.5 * dl.t * dl.t * Tg(k, 1);
            A(i * 6 + k, 2) = dp[k];
            A(i * 6 + k, 3 + i * 3 + k) = -dl.t;
            b[i * 6 + k] = b0[k];
            A(i * 6 + 3 + k, 0) = -dl.t * Tg(k, 0);
            A(i * 6 + 3 + k, 1) = -dl.t * Tg(k, 1);
            A(i * 6 + 3 + k, 3 + i * 3 + k) = -1;
            A(i * 6 + 3 + k, 3 + j * 3 + k) = 1;
            b[i * 6 + 3 + k] = b1[k];
        }
    }
    const std::vector<double> s = lstsq(A, b);
    AlignResult r;
    const V3 dg = v3(Tg(0, 0) * s[0] + Tg(0, 1) * s[1], Tg(1, 0) * s[0] + Tg(1, 1) * s[1], Tg(2, 0) * s[0] + Tg(2, 1) * s[1]);
    r.gravity = 9.80665 * normalized(gravity + damp * dg);
    r.scale = s[2];
    for (int i = 0; i < N; ++i) r.velocities.push_back(v3(s[3 + 3 * i], s[4 + 3 * i], s[5 + 3 * i]));
    return r;
}

}  // namespace xvi
