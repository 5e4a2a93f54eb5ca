// Small fixed-size double-precision linear algebra and SO(3) helpers used by the back-end
// kernels.  Conventions follow the reference (quaternions stored x,y,z,w;
// xrslam/src/xrslam/geometry/lie_algebra.{h,cpp}); the code itself is written for one CUDA
// thread holding everything in registers.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

#define XV_HD __host__ __device__ __forceinline__

namespace xv {

struct V3 {
    double x, y, z;
    XV_HD double &operator[](int i) { return (&x)[i]; }
    XV_HD const double &operator[](int i) const { return (&x)[i]; }
};
struct Q4 { double x, y, z, w; };
struct M3 { double m[9]; XV_HD double &operator()(int r, int c) { return m[3 * r + c]; } XV_HD const double &operator()(int r, int c) const { return m[3 * r + c]; } };

XV_HD V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
XV_HD V3 operator+(const V3 &a, const V3 &b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
XV_HD V3 operator-(const V3 &a, const V3 &b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
XV_HD V3 operator-(const V3 &a) { return v3(-a.x, -a.y, -a.z); }
XV_HD V3 operator*(double s, const V3 &a) { return v3(s * a.x, s * a.y, s * a.z); }
XV_HD V3 operator*(const V3 &a, double s) { return v3(s * a.x, s * a.y, s * a.z); }
XV_HD double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
XV_HD V3 cross(const V3 &a, const V3 &b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
XV_HD double norm(const V3 &a) { return sqrt(dot(a, a)); }
XV_HD V3 normalized(const V3 &a) { double n = norm(a); return v3(a.x / n, a.y / n, a.z / n); }

XV_HD Q4 q4(double x, double y, double z, double w) { Q4 q; q.x = x; q.y = y; q.z = z; q.w = w; return q; }
XV_HD Q4 qconj(const Q4 &q) { return q4(-q.x, -q.y, -q.z, q.w); }
XV_HD Q4 qmul(const Q4 &a, const Q4 &b) {
    return q4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
              a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
XV_HD Q4 qnormalized(const Q4 &q) {
    double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return q4(q.x / n, q.y / n, q.z / n, q.w / n);
}
XV_HD V3 qrot(const Q4 &q, const V3 &v) {
    V3 u = v3(q.x, q.y, q.z);
    V3 t = cross(u, v);
    t = t + t;
    return v + q.w * t + cross(u, t);
}
XV_HD M3 qmat(const Q4 &q) {
    M3 R;
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R(0, 0) = 1 - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = 1 - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = 1 - (txx + tyy);
    return R;
}
XV_HD M3 hat(const V3 &w) {
    M3 H;
    H(0, 0) = 0; H(0, 1) = -w.z; H(0, 2) = w.y;
    H(1, 0) = w.z; H(1, 1) = 0; H(1, 2) = -w.x;
    H(2, 0) = -w.y; H(2, 1) = w.x; H(2, 2) = 0;
    return H;
}
XV_HD M3 m3_identity() { M3 I; for (int i = 0; i < 9; ++i) I.m[i] = 0; I.m[0] = I.m[4] = I.m[8] = 1; return I; }
XV_HD M3 operator*(const M3 &A, const M3 &B) {
    M3 C;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C(r, c) = A(r, 0) * B(0, c) + A(r, 1) * B(1, c) + A(r, 2) * B(2, c);
    return C;
}
XV_HD V3 operator*(const M3 &A, const V3 &v) {
    return v3(A(0, 0) * v.x + A(0, 1) * v.y + A(0, 2) * v.z, A(1, 0) * v.x + A(1, 1) * v.y + A(1, 2) * v.z,
              A(2, 0) * v.x + A(2, 1) * v.y + A(2, 2) * v.z);
}
XV_HD M3 operator*(double s, const M3 &A) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = s * A.m[i]; return C; }
XV_HD M3 operator+(const M3 &A, const M3 &B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] + B.m[i]; return C; }
XV_HD M3 operator-(const M3 &A, const M3 &B) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = A.m[i] - B.m[i]; return C; }
XV_HD M3 operator-(const M3 &A) { M3 C; for (int i = 0; i < 9; ++i) C.m[i] = -A.m[i]; return C; }
XV_HD M3 transpose(const M3 &A) { M3 C; for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) C(r, c) = A(c, r); return C; }
XV_HD M3 inverse(const M3 &A) {
    M3 C;
    double c00 = A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1);
    double c01 = A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2);
    double c02 = A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0);
    double det = A(0, 0) * c00 + A(0, 1) * c01 + A(0, 2) * c02;
    double id = 1.0 / det;
    C(0, 0) = c00 * id; C(1, 0) = c01 * id; C(2, 0) = c02 * id;
    C(0, 1) = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * id;
    C(1, 1) = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * id;
    C(2, 1) = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * id;
    C(0, 2) = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * id;
    C(1, 2) = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * id;
    C(2, 2) = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * id;
    return C;
}

// lie_algebra.h:13-18  expmap(w) = AngleAxis(|w|, stableNormalized(w))
XV_HD Q4 expmap(const V3 &w) {
    double ax = fabs(w.x), ay = fabs(w.y), az = fabs(w.z);
    double m = fmax(ax, fmax(ay, az));
    double angle = norm(w);
    V3 axis = w;
    if (m > 0.0) {
        V3 s = v3(w.x / m, w.y / m, w.z / m);
        double z = dot(s, s);
        if (z > 0.0) { double d = sqrt(z) * m; axis = v3(w.x / d, w.y / d, w.z / d); }
    }
    double sh = sin(0.5 * angle), ch = cos(0.5 * angle);
    return q4(sh * axis.x, sh * axis.y, sh * axis.z, ch);
}
// lie_algebra.h:20-23  logmap(q) = AngleAxis(q).angle * axis  (angle in [0, pi], no shortest-path flip)
XV_HD V3 logmap(const Q4 &q) {
    V3 v = v3(q.x, q.y, q.z);
    double n = norm(v);
    if (n == 0.0) return v3(0, 0, 0);
    double angle = 2.0 * atan2(n, fabs(q.w));
    if (q.w < 0) n = -n;
    double s = angle / n;
    // (angle * (v / n)) in the reference; keep its operation order
    V3 axis = v3(v.x / n, v.y / n, v.z / n);
    (void)s;
    return v3(angle * axis.x, angle * axis.y, angle * axis.z);
}
// lie_algebra.cpp:5-45
XV_HD M3 right_jacobian(const V3 &w) {
    const double root2_eps = 1.4901161193847656e-08;   // sqrt(DBL_EPSILON)
    const double root4_eps = 1.2207031250000000e-04;   // sqrt(sqrt(DBL_EPSILON))
    const double qdrt720 = 5.1800401282227027;          // 720^(1/4)
    const double qdrt5040 = 8.4257318612210419;         // 5040^(1/4)
    const double sqrt24 = 4.8989794855663560;
    const double sqrt120 = 10.954451150103322;
    double angle = norm(w);
    double cangle = cos(angle), sangle = sin(angle);
    double angle2 = angle * angle;
    double cos_term, sin_term;
    if (angle > root4_eps * qdrt720) cos_term = (1 - cangle) / angle2;
    else { cos_term = 0.5; if (angle > root2_eps * sqrt24) cos_term -= angle2 / 24.0; }
    if (angle > root4_eps * qdrt5040) sin_term = (angle - sangle) / (angle * angle2);
    else { sin_term = 1.0 / 6.0; if (angle > root2_eps * sqrt120) sin_term -= angle2 / 120.0; }
    M3 H = hat(w);
    return m3_identity() - cos_term * H + sin_term * (H * H);
}
// lie_algebra.cpp:47-56 ; returns T = [b1 b2 z] (columns)
XV_HD M3 local_tangent(const V3 &z) {
    int d = 0;
    if (fabs(z[1]) > fabs(z[d])) d = 1;
    if (fabs(z[2]) > fabs(z[d])) d = 2;
    V3 e = v3(0, 0, 0);
    e[(d + 1) % 3] = 1.0;
    V3 b1 = normalized(cross(z, e));
    V3 b2 = normalized(cross(z, b1));
    M3 T;
    T(0, 0) = b1.x; T(1, 0) = b1.y; T(2, 0) = b1.z;
    T(0, 1) = b2.x; T(1, 1) = b2.y; T(2, 1) = b2.z;
    T(0, 2) = z.x; T(1, 2) = z.y; T(2, 2) = z.z;
    return T;
}

XV_HD V3 ld3(const double *p) { return v3(p[0], p[1], p[2]); }
XV_HD Q4 ld4(const double *p) { return q4(p[0], p[1], p[2], p[3]); }
XV_HD void st3(double *p, const V3 &v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
XV_HD void st4(double *p, const Q4 &q) { p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w; }

}  // namespace xv
