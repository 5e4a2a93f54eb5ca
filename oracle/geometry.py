"""Two-view geometry used by the initializer and the per-frame verification (N1):

  * generic RANSAC                               xrslam/src/xrslam/utility/ransac.h:26-83
  * 4-point homography + decomposition           xrslam/src/xrslam/geometry/homography.cpp:5-158
  * 5-point essential (Groebner / action matrix) xrslam/src/xrslam/geometry/essential.cpp:105-218,285-297
  * essential decomposition (SVD branch)         xrslam/src/xrslam/geometry/essential.cpp:262-283
  * RANSAC wrappers and error measures           xrslam/src/xrslam/geometry/stereo.cpp:40-69,99-126,
                                                 essential.h:15-20, homography.h:17-21
  * two-view DLT triangulation                   xrslam/src/xrslam/geometry/stereo.h:72-82

Where the reference calls Eigen (JacobiSVD null spaces, EigenSolver) numpy's LAPACK routines are used.
A null-space basis and the order of eigenvalues are not unique, so candidate models can come out in a
different ORDER than in the reference build; RANSAC only notices on exact inlier-count ties.
Pinned against OpenCV (cv2.findEssentialMat / findHomography / decomposeHomographyMat) on synthetic
two-view data in tests/test_oracle_geometry.py.
Oracle code: test infrastructure only.
"""
import math
import numpy as np

from .ransac import LotBox


# ------------------------------------------------------------------ generic RANSAC (ransac.h:26-83)
def ransac(dof, solver, make_evaluator, threshold, data1, data2, confidence=0.999, max_iteration=1000, seed=0):
    """solver(sample1, sample2) -> list of models; make_evaluator(model) -> f(data1, data2) -> errors (n,).
    Returns model (None when size < dof: the reference returns an uninitialised matrix), mask."""
    n = len(data1)
    mask = np.zeros(n, np.uint8)
    model = None
    if n < dof:
        return model, mask
    lot = LotBox(n, seed)
    K = math.log(max(1 - confidence, 1.0e-5))
    best = 0
    iter_max = max_iteration
    it = 0
    while it < iter_max:
        lot.refill_all()
        idx = [lot.draw_without_replacement() for _ in range(dof)]
        models = solver(data1[idx], data2[idx])
        for cand in models:
            with np.errstate(all='ignore'):
                err = make_evaluator(cand)(data1, data2)
            cur = err <= threshold                      # NaN compares false, as in C++
            cnt = int(cur.sum())
            if cnt > best:
                model, best, mask = cand, cnt, cur.astype(np.uint8)
                ratio = best / float(n)
                denom = math.log(1 - ratio ** 5) if ratio < 1 else -math.inf
                N = K / denom if denom != 0 else math.inf
                if N < float(iter_max):
                    iter_max = int(math.ceil(N))
        it += 1
    return model, mask


def _hom(p):
    return np.concatenate([p, np.ones((len(p), 1))], axis=1)


# ------------------------------------------------------------------ homography (homography.cpp)
def _to_matrix(v):
    # homography.cpp:82-86 / essential.cpp:99-103: the three segments become COLUMNS
    return np.array(v, dtype=np.float64).reshape(3, 3).T.copy()


def solve_homography_4pt(pa, pb):
    # homography.cpp:117-156 (Hartley normalisation with the reference's scale) + :88-113 (DLT null vector)
    pa = np.asarray(pa, np.float64)
    pb = np.asarray(pb, np.float64)
    ma, mb = pa.mean(axis=0), pb.mean(axis=0)
    sa = np.linalg.norm(pa - ma, axis=1).sum()
    sb = np.linalg.norm(pb - mb, axis=1).sum()
    with np.errstate(divide='ignore'):
        sa = 1.0 / (math.sqrt(2.0) * sa)
        sb = 1.0 / (math.sqrt(2.0) * sb)
    na = (pa - ma) * sa
    nb = (pb - mb) * sb
    A = np.zeros((8, 9))
    for i in range(4):
        a, b = na[i], nb[i]
        A[2 * i, 1] = -a[0]
        A[2 * i, 2] = a[0] * b[1]
        A[2 * i, 4] = -a[1]
        A[2 * i, 5] = a[1] * b[1]
        A[2 * i, 7] = -1
        A[2 * i, 8] = b[1]
        A[2 * i + 1, 0] = a[0]
        A[2 * i + 1, 2] = -a[0] * b[0]
        A[2 * i + 1, 3] = a[1]
        A[2 * i + 1, 5] = -a[1] * b[0]
        A[2 * i + 1, 6] = 1
        A[2 * i + 1, 8] = -b[0]
    if not np.all(np.isfinite(A)):
        return np.full((3, 3), np.nan)
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    NH = _to_matrix(Vt[8])
    with np.errstate(divide='ignore', invalid='ignore'):
        Nb = np.array([[1 / sb, 0, mb[0]], [0, 1 / sb, mb[1]], [0, 0, 1]])
        Na = np.array([[sa, 0, -sa * ma[0]], [0, sa, -sa * ma[1]], [0, 0, 1]])
        return Nb @ NH @ Na


def homography_geometric_error(H, p1, p2):
    q = _hom(p1) @ H.T
    return np.sum((p2 - q[:, :2] / q[:, 2:3]) ** 2, axis=1)


def find_homography_matrix(p1, p2, threshold=1.0, confidence=0.999, max_iteration=1000, seed=0):
    # stereo.cpp:99-126
    p1 = np.asarray(p1, np.float64).reshape(-1, 2)
    p2 = np.asarray(p2, np.float64).reshape(-1, 2)

    def make_eval(H):
        with np.errstate(all='ignore'):
            try:
                Hinv = np.linalg.inv(H)
            except np.linalg.LinAlgError:
                Hinv = np.full((3, 3), np.nan)
        return lambda a, b: homography_geometric_error(H, a, b) + homography_geometric_error(Hinv, b, a)

    return ransac(4, lambda a, b: [solve_homography_4pt(a, b)], make_eval, 2.0 * 5.99 * threshold * threshold,
                  p1, p2, confidence, max_iteration, seed)


def decompose_homography(H):
    """homography.cpp:5-78.  Returns ok(not pure rotation), R1, R2, T1, T2, n1, n2."""
    sv = np.linalg.svd(H, compute_uv=False)
    Hn = H / sv[1]
    S = Hn.T @ Hn - np.eye(3)
    if np.all(np.abs(S) <= 1e-3):
        U, _, Vt = np.linalg.svd(H)
        R1 = U @ Vt
        if np.linalg.det(R1) < 0:
            R1 = -R1
        z = np.zeros(3)
        return False, R1, R1.copy(), z, z.copy(), z.copy(), z.copy()
    Ms00 = S[1, 2] * S[1, 2] - S[1, 1] * S[2, 2]
    Ms11 = S[0, 2] * S[0, 2] - S[0, 0] * S[2, 2]
    Ms22 = S[0, 1] * S[0, 1] - S[0, 0] * S[1, 1]
    with np.errstate(invalid='ignore'):
        s00, s11, s22 = np.sqrt(Ms00), np.sqrt(Ms11), np.sqrt(Ms22)
        nu = 2.0 * np.sqrt(1 + np.trace(S) - Ms00 - Ms11 - Ms22)
    tenormsq = 2 + np.trace(S) - nu
    if S[0, 0] > S[1, 1] and S[0, 0] > S[2, 2]:
        e = -1.0 if (S[0, 1] * S[0, 2] - S[0, 0] * S[1, 2]) < 0 else 1.0
        n1 = np.array([S[0, 0], S[0, 1] + s22, S[0, 2] + e * s11])
        n2 = np.array([S[0, 0], S[0, 1] - s22, S[0, 2] - e * s11])
        t1 = np.linalg.norm(n1) * n2 / S[0, 0]
        t2 = np.linalg.norm(n2) * n1 / S[0, 0]
    elif S[1, 1] > S[0, 0] and S[1, 1] > S[2, 2]:
        e = -1.0 if (S[1, 1] * S[0, 2] - S[0, 1] * S[1, 2]) < 0 else 1.0
        n1 = np.array([S[0, 1] + s22, S[1, 1], S[1, 2] - e * s00])
        n2 = np.array([S[0, 1] - s22, S[1, 1], S[1, 2] + e * s00])
        t2 = np.linalg.norm(n2) * n1 / S[1, 1]
        t1 = np.linalg.norm(n1) * n2 / S[1, 1]
    else:
        e = -1.0 if (S[1, 2] * S[0, 2] - S[0, 1] * S[2, 2]) < 0 else 1.0
        n1 = np.array([S[0, 2] + e * s11, S[1, 2] + s00, S[2, 2]])
        n2 = np.array([S[0, 2] - e * s11, S[1, 2] - s00, S[2, 2]])
        t1 = np.linalg.norm(n1) * n2 / S[2, 2]
        t2 = np.linalg.norm(n2) * n1 / S[2, 2]
    n1 = n1 / np.linalg.norm(n1)
    n2 = n2 / np.linalg.norm(n2)
    t1 = t1 - tenormsq * n1
    t2 = t2 - tenormsq * n2
    R1 = Hn @ (np.eye(3) - np.outer(t1 / nu, n1))
    R2 = Hn @ (np.eye(3) - np.outer(t2 / nu, n2))
    t1 = t1 * 0.5
    t2 = t2 * 0.5
    return True, R1, R2, R1 @ t1, R2 @ t2, n1, n2


# ------------------------------------------------------------------ essential (essential.cpp)
# GRevLex monomial order of the reference (essential.cpp:13-16)
XXX, XXY, XYY, YYY, XXZ, XYZ, YYZ, XZZ, YZZ, ZZZ, XX, XY, YY, XZ, YZ, ZZ, X, Y, Z, I = range(20)
_EXP = {XXX: (3, 0, 0), XXY: (2, 1, 0), XYY: (1, 2, 0), YYY: (0, 3, 0), XXZ: (2, 0, 1), XYZ: (1, 1, 1),
        YYZ: (0, 2, 1), XZZ: (1, 0, 2), YZZ: (0, 1, 2), ZZZ: (0, 0, 3), XX: (2, 0, 0), XY: (1, 1, 0),
        YY: (0, 2, 0), XZ: (1, 0, 1), YZ: (0, 1, 1), ZZ: (0, 0, 2), X: (1, 0, 0), Y: (0, 1, 0), Z: (0, 0, 1),
        I: (0, 0, 0)}
_IDX = {e: k for k, e in _EXP.items()}
_MUL = [(a, b, _IDX[tuple(np.add(_EXP[a], _EXP[b]))]) for a in range(20) for b in range(20)
        if sum(np.add(_EXP[a], _EXP[b])) <= 3]


def _pmul(a, b):
    """Product of two polynomials in x, y, z truncated at total degree 3 (essential.cpp:50-93)."""
    r = np.zeros(20)
    for (i, j, k) in _MUL:
        r[k] += a[i] * b[j]
    return r


def generate_nullspace_basis(p1, p2):
    # essential.cpp:105-117: rows vec(p1h p2h^T), the last four right singular vectors
    A = np.zeros((5, 9))
    for i in range(5):
        h = np.outer(np.append(p1[i], 1.0), np.append(p2[i], 1.0))
        A[i] = h.reshape(9)
    _, _, Vt = np.linalg.svd(A, full_matrices=True)
    return Vt[5:9].T.copy()                                      # 9x4


def generate_polynomials(basis):
    # essential.cpp:119-149
    Ex, Ey, Ez, Ew = (_to_matrix(basis[:, k]) for k in range(4))
    E = np.zeros((3, 3, 20))
    E[:, :, X], E[:, :, Y], E[:, :, Z], E[:, :, I] = Ex, Ey, Ez, Ew

    def matmul(A, B):
        C = np.zeros((3, 3, 20))
        for i in range(3):
            for j in range(3):
                for k in range(3):
                    C[i, j] += _pmul(A[i, k], B[k, j])
        return C

    Et = np.transpose(E, (1, 0, 2))
    EEt = matmul(E, Et)
    tr = 0.5 * (EEt[0, 0] + EEt[1, 1] + EEt[2, 2])
    c = matmul(EEt, E)
    polys = np.zeros((10, 20))
    for i in range(3):
        for j in range(3):
            polys[i * 3 + j] = c[i, j] - _pmul(tr, E[i, j])
    det = (_pmul(E[0, 0], _pmul(E[1, 1], E[2, 2]) - _pmul(E[1, 2], E[2, 1]))
           - _pmul(E[0, 1], _pmul(E[1, 0], E[2, 2]) - _pmul(E[1, 2], E[2, 0]))
           + _pmul(E[0, 2], _pmul(E[1, 0], E[2, 1]) - _pmul(E[1, 1], E[2, 0])))
    polys[9] = det
    return polys


def generate_action_matrix(polys):
    # essential.cpp:151-199: Gauss-Jordan on the cubic monomials with the reference's pivot search
    P = polys.copy()
    perm = list(range(10))
    for i in range(10):
        for j in range(i + 1, 10):
            if abs(P[perm[i], i]) < abs(P[perm[j], i]):
                perm[i], perm[j] = perm[j], perm[i]
        if P[perm[i], i] == 0:
            continue
        P[perm[i]] = P[perm[i]] / P[perm[i], i]
        for j in range(i + 1, 10):
            P[perm[j]] = P[perm[j]] - P[perm[i]] * P[perm[j], i]
    for i in range(9, 0, -1):
        for j in range(i):
            P[perm[j]] = P[perm[j]] - P[perm[i]] * P[perm[j], i]
    A = np.zeros((10, 10))
    for r, mono in enumerate((XXX, XXY, XYY, XXZ, XYZ, XZZ)):
        A[r] = -P[perm[mono], XX:XX + 10]
    A[6, XX - XX] = 1
    A[7, XY - XX] = 1
    A[8, XZ - XX] = 1
    A[9, X - XX] = 1
    return A


def solve_grobner_system(action):
    # essential.cpp:201-218
    if not np.all(np.isfinite(action)):
        return []
    w, V = np.linalg.eig(action)
    out = []
    for i in range(10):
        if abs(w[i].imag) < 1.0e-10:
            h = V[:, i].real
            with np.errstate(divide='ignore', invalid='ignore'):
                out.append(np.array([h[X - XX], h[Y - XX], h[Z - XX]]) / h[I - XX])
    return out


def solve_essential_5pt(p1, p2):
    # essential.cpp:285-297
    basis = generate_nullspace_basis(p1, p2)
    action = generate_action_matrix(generate_polynomials(basis))
    return [_to_matrix(basis @ np.append(s, 1.0)) for s in solve_grobner_system(action)]


def essential_geometric_error(E, p1, p2):
    # essential.h:15-20
    Ep1 = _hom(p1) @ E.T
    r = np.sum(_hom(p2) * Ep1, axis=1)
    return r * r / np.sum(Ep1[:, :2] ** 2, axis=1)


def find_essential_matrix(p1, p2, threshold=1.0, confidence=0.999, max_iteration=1000, seed=0):
    # stereo.cpp:40-69
    p1 = np.asarray(p1, np.float64).reshape(-1, 2)
    p2 = np.asarray(p2, np.float64).reshape(-1, 2)

    def make_eval(E):
        Et = E.T
        return lambda a, b: essential_geometric_error(E, a, b) + essential_geometric_error(Et, b, a)

    return ransac(5, solve_essential_5pt, make_eval, 2.0 * 3.84 * threshold * threshold, p1, p2,
                  confidence, max_iteration, seed)


def decompose_essential(E):
    # essential.cpp:262-283 (SVD branch; DECOMPOSE_ESSENTIAL_HORN is not defined)
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0.0, 1, 0], [-1, 0, 0], [0, 0, 1]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2].copy()


# ------------------------------------------------------------------ two-view DLT (stereo.h:72-82)
def triangulate_point_2v(P1, P2, x1, x2):
    A = np.stack([x1[0] * P1[2] - x1[2] * P1[0], x1[1] * P1[2] - x1[2] * P1[1],
                  x2[0] * P2[2] - x2[2] * P2[0], x2[1] * P2[2] - x2[2] * P2[1]])
    _, _, Vt = np.linalg.svd(A)
    return Vt[3]
