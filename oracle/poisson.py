"""PoissonDiskFilter<2> restated from
xrslam-extra/include/xrslam/extra/poisson_disk_filter.h:9-114 (identical copy at
xrslam/src/xrslam/utility/poisson_disk_filter.h), including its scan-order quirk in
test_point (:81-93): the first cell of the 5x5 neighbourhood is skipped and one extra cell
(ibegin.x, iend.y + 1) is visited; a grid cell remembers only the LAST point stored in it.
Oracle code: test infrastructure only.
"""
import math


class PoissonDiskFilter2:
    def __init__(self, radius):
        self.radius = float(radius)
        self.radius_squared = self.radius * self.radius
        self.grid_size = self.radius / math.sqrt(2.0)
        self.grid_span = int(math.ceil(math.sqrt(2.0)))  # 2
        self.points = []
        self.grid = {}

    def _index(self, p):
        return (int(math.floor(p[0] / self.grid_size)), int(math.floor(p[1] / self.grid_size)))

    def preset_point(self, p):
        self.grid[self._index(p)] = len(self.points)
        self.points.append((float(p[0]), float(p[1])))

    def preset_points(self, pts):
        for p in pts:
            self.preset_point(p)

    def _test(self, p):
        ix, iy = self._index(p)
        bx, by = ix - self.grid_span, iy - self.grid_span
        ex, ey = ix + self.grid_span, iy + self.grid_span
        cx, cy = bx, by
        while cy <= ey:
            cx += 1
            if cx > ex:
                cx = bx
                cy += 1
            k = self.grid.get((cx, cy))
            if k is not None:
                q = self.points[k]
                dx, dy = p[0] - q[0], p[1] - q[1]
                if dx * dx + dy * dy < self.radius_squared:
                    return False, (ix, iy)
        return True, (ix, iy)

    def permit_point(self, p):
        return self._test(p)[0]

    def insert_point(self, p):
        ok, idx = self._test(p)
        if ok:
            self.grid[idx] = len(self.points)
            self.points.append((float(p[0]), float(p[1])))
        return ok

    def insert_points(self, candidates):
        """Returns the accepted subset, in order (poisson_disk_filter.h:52-62)."""
        out = []
        for p in candidates:
            if self.insert_point(p):
                out.append(p)
        return out
