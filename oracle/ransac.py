"""Per-frame geometric verification helpers (D3): 2-point rotation RANSAC
(xrslam/src/xrslam/geometry/wahba.h:9-27, geometry/stereo.cpp:71-97, utility/ransac.h:26-83)
with the reference's sampling: LotBox over std::default_random_engine (= minstd_rand0) seeded
with 0 and libstdc++ uniform_int_distribution<size_t> (utility/random.h:80-126).

The 5-point essential RANSAC (geometry/stereo.cpp:40-69) is NOT restated: its threshold is
2*3.84 in squared normalised image units (an epipolar distance of ~2.8 focal lengths), so on
any realistic data every point is an inlier of the first hypothesis and the mask is all ones;
the only case that matters -- fewer than 5 points -> all-zero mask (ransac.h:46-50) -- is kept.
That simplification is SURVEY.md 8(f) row N1 ("next") and is stated in DESIGN.md.
Oracle code: test infrastructure only.
"""
import math
import numpy as np


class MinStdRand0:
    """std::minstd_rand0: x <- 16807 x mod (2^31-1); seed 0 -> 1."""
    M = 2147483647

    def __init__(self, seed=0):
        s = seed % self.M
        self.x = 1 if s == 0 else s

    def next(self):
        self.x = (self.x * 16807) % self.M
        return self.x


def uniform_int(engine, a, b):
    """libstdc++ uniform_int_distribution<size_t>(a,b) over minstd_rand0 (range not a power of
    two -> classic down-scaling with rejection)."""
    urngmin, urngmax = 1, 2147483646
    urngrange = urngmax - urngmin
    urange = b - a
    if urngrange > urange:
        uerange = urange + 1
        scaling = urngrange // uerange
        past = uerange * scaling
        while True:
            ret = engine.next() - urngmin
            if ret < past:
                break
        ret //= scaling
    elif urngrange < urange:
        raise NotImplementedError
    else:
        ret = engine.next() - urngmin
    return ret + a


class LotBox:
    def __init__(self, size, seed=0):
        self.cap = 0
        self.lots = list(range(size))
        self.engine = MinStdRand0(seed)

    def refill_all(self):
        self.cap = 0

    def draw_without_replacement(self):
        rem = len(self.lots) - self.cap
        if rem > 1:
            k = uniform_int(self.engine, self.cap, len(self.lots) - 1)
            self.lots[self.cap], self.lots[k] = self.lots[k], self.lots[self.cap]
            r = self.lots[self.cap]
            self.cap += 1
            return r
        elif rem == 1:
            self.cap += 1
            return self.lots[-1]
        return -1


def solve_rotation_2pt(p1, p2):
    # wahba.h:9-27
    cov = np.zeros((3, 3))
    for a, b in zip(p1, p2):
        cov += np.outer(a, b)
    cov *= 0.5
    U, s, Vt = np.linalg.svd(cov)
    V = Vt.T
    E = np.eye(3)
    E[2, 2] = 1.0 if np.linalg.det(V @ U.T) >= 0.0 else -1.0
    return V @ E @ U.T


def find_rotation_matrix(b1, b2, threshold_rad, confidence=0.999, max_iteration=1000, seed=0):
    """Returns R, inlier_mask (uint8).  b1, b2: (n,3) unit bearings."""
    n = len(b1)
    thr = 5.99 * threshold_rad * threshold_rad
    mask = np.zeros(n, np.uint8)
    model = np.zeros((3, 3))     # value-initialised Eigen matrix is uninitialised in C++; size<2 is unused
    if n < 2:
        return model, mask
    lot = LotBox(n, seed)
    K = math.log(max(1 - confidence, 1.0e-5))
    best = 0
    iter_max = max_iteration
    it = 0
    while it < iter_max:
        lot.refill_all()
        i0 = lot.draw_without_replacement()
        i1 = lot.draw_without_replacement()
        R = solve_rotation_2pt([b1[i0], b1[i1]], [b2[i0], b2[i1]])
        c = np.einsum('ij,nj->ni', R, b1)
        with np.errstate(invalid='ignore'):
            err = np.arccos(np.sum(c * b2, axis=1))
        cur = err <= thr            # NaN (dot > 1) compares false, as in C++
        cnt = int(cur.sum())
        if cnt > best:
            model, best, mask = R, cnt, cur.astype(np.uint8)
            ratio = best / float(n)
            with np.errstate(divide='ignore'):
                denom = math.log(1 - ratio ** 5) if ratio < 1 else -math.inf
            N = K / denom if denom != 0 else math.inf
            if N < float(iter_max):
                iter_max = int(math.ceil(N))
        it += 1
    return model, mask
