"""Lie-group helpers, float64.  Quaternions are stored x, y, z, w (Eigen coeffs()
order, reference xrslam/src/xrslam/estimation/solver.cpp:85-87).

Follows xrslam/src/xrslam/geometry/lie_algebra.h:8-23 and lie_algebra.cpp:5-56.
Oracle code: test infrastructure only.
"""
import numpy as np

EPS = np.finfo(np.float64).eps
_ROOT2_EPS = np.sqrt(EPS)
_ROOT4_EPS = np.sqrt(_ROOT2_EPS)
_QDRT720 = np.sqrt(np.sqrt(720.0))
_QDRT5040 = np.sqrt(np.sqrt(5040.0))
_SQRT24 = np.sqrt(24.0)
_SQRT120 = np.sqrt(120.0)


def hat(w):
    # lie_algebra.h:8-11
    return np.array([[0.0, -w[2], w[1]], [w[2], 0.0, -w[0]], [-w[1], w[0], 0.0]])


def qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz,
    ])


def qconj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def qnormalized(q):
    return q / np.linalg.norm(q)


def qmat(q):
    """Rotation matrix of a (unit) quaternion, Eigen's toRotationMatrix formula."""
    x, y, z, w = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([
        [1 - (tyy + tzz), txy - twz, txz + twy],
        [txy + twz, 1 - (txx + tzz), tyz - twx],
        [txz - twy, tyz + twx, 1 - (txx + tyy)],
    ])


def qrot(q, v):
    """q * v (Eigen's quaternion-vector product: v + w*t + u x t, t = 2 u x v)."""
    u = q[:3]
    t = 2.0 * np.cross(u, v)
    return v + q[3] * t + np.cross(u, t)


def stable_normalized(w):
    # Eigen stableNormalized: divide by max-abs first, zero stays zero.
    m = np.max(np.abs(w))
    if m == 0.0:
        return w.copy()
    z = np.sum((w / m) ** 2)
    if z > 0:
        return w / (np.sqrt(z) * m)
    return w.copy()


def expmap(w):
    # lie_algebra.h:13-18: AngleAxis(|w|, stableNormalized(w)) -> quaternion
    w = np.asarray(w, dtype=np.float64)
    angle = np.linalg.norm(w)
    axis = stable_normalized(w)
    s = np.sin(0.5 * angle)
    return np.array([s * axis[0], s * axis[1], s * axis[2], np.cos(0.5 * angle)])


def logmap(q):
    # lie_algebra.h:20-23 via Eigen AngleAxis(const Quaternion&):
    #   n = vec.norm(); if n < tiny: n = vec.stableNorm()
    #   if n != 0: angle = 2 atan2(n, |w|); if w < 0: n = -n; axis = vec / n
    #   else angle = 0, axis = (1,0,0)
    v = np.asarray(q[:3], dtype=np.float64)
    n = np.linalg.norm(v)
    if n == 0.0:
        return np.zeros(3)
    angle = 2.0 * np.arctan2(n, abs(q[3]))
    if q[3] < 0:
        n = -n
    return angle * (v / n)


def right_jacobian(w):
    # lie_algebra.cpp:5-45
    w = np.asarray(w, dtype=np.float64)
    angle = np.linalg.norm(w)
    cangle, sangle = np.cos(angle), np.sin(angle)
    angle2 = angle * angle
    if angle > _ROOT4_EPS * _QDRT720:
        cos_term = (1 - cangle) / angle2
    else:
        cos_term = 0.5
        if angle > _ROOT2_EPS * _SQRT24:
            cos_term -= angle2 / 24.0
    if angle > _ROOT4_EPS * _QDRT5040:
        sin_term = (angle - sangle) / (angle * angle2)
    else:
        sin_term = 1.0 / 6.0
        if angle > _ROOT2_EPS * _SQRT120:
            sin_term -= angle2 / 120.0
    H = hat(w)
    return np.eye(3) - cos_term * H + sin_term * (H @ H)


def s2_tangential_basis(x):
    # lie_algebra.cpp:47-56
    d = 0
    for i in (1, 2):
        if abs(x[i]) > abs(x[d]):
            d = i
    e = np.zeros(3)
    e[(d + 1) % 3] = 1.0
    b1 = np.cross(x, e)
    b1 = b1 / np.linalg.norm(b1)
    b2 = np.cross(x, b1)
    b2 = b2 / np.linalg.norm(b2)
    return np.stack([b1, b2], axis=1)  # 3x2


def local_tangent(z):
    """[b1 b2 z] as in reprojection_factor.h:17-23."""
    T = np.empty((3, 3))
    T[:, :2] = s2_tangential_basis(z)
    T[:, 2] = z
    return T
