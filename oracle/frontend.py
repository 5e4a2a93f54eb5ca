"""Front-end image operators of the reference (K1-K5), executed with the SAME third-party
library the reference calls (OpenCV; here cv2 4.13.0, the only OpenCV on this box) and the same
parameters.  Follows xrslam-extra/src/xrslam/extra/opencv_image.cpp:38-161 line by line.
The reference does not pin an OpenCV version for the PC build (cmake/external/opencv.cmake:2).
Oracle code: test infrastructure only.
"""
import numpy as np
import cv2
from .poisson import PoissonDiskFilter2

LK_CRITERIA = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
LEVEL_NUM = 3  # opencv_image.h:20


class OracleImage:
    """Mirror of xrslam::extra::OpenCvImage for the four operator entry points."""

    _clahe = None   # function-local static in the reference (opencv_image.cpp:179-182)
    _gftt = None    # (opencv_image.cpp:184-188)

    def __init__(self, t, gray_u8):
        self.t = float(t)
        self.image = np.ascontiguousarray(gray_u8)
        self.pyramid = None

    # opencv_image.cpp:156-161
    def preprocess(self, clip_limit=6.0, width=8, height=8):
        if OracleImage._clahe is None:
            OracleImage._clahe = cv2.createCLAHE(clip_limit, (width, height))
        self.image = OracleImage._clahe.apply(self.image)
        _, self.pyramid = cv2.buildOpticalFlowPyramid(self.image, (21, 21), LEVEL_NUM, None, True)

    # opencv_image.cpp:38-73
    def detect_keypoints(self, keypoints, max_points=200, keypoint_distance=20.0):
        """keypoints: list of (x,y) already present (pre-seeded); returns the extended list."""
        if OracleImage._gftt is None:
            OracleImage._gftt = cv2.GFTTDetector_create(int(max_points), 1.0e-3, 20, 3, True)
        cvk = OracleImage._gftt.detect(self.image)
        keypoints = list(keypoints)
        if len(cvk) > 0:
            # std::sort by response, descending.  cv2 4.13 fills response with the corner quality and
            # returns corners already in that order; Python's sort is stable so ties keep GFTT order.
            cvk = sorted(cvk, key=lambda k: -k.response)
            new_kps = [(float(k.pt[0]), float(k.pt[1])) for k in cvk]
            flt = PoissonDiskFilter2(keypoint_distance)
            flt.preset_points(keypoints)
            new_kps = flt.insert_points(new_kps)
            rows, cols = self.image.shape
            new_kps = [p for p in new_kps if not (p[0] < 20 or p[1] < 20 or p[0] >= cols - 20 or p[1] >= rows - 20)]
            keypoints.extend(new_kps)
        return keypoints

    def gftt_raw(self):
        """The raw GFTT corners (x, y, response) before the reference's Poisson/border filter."""
        if OracleImage._gftt is None:
            OracleImage._gftt = cv2.GFTTDetector_create(200, 1.0e-3, 20, 3, True)
        return [(k.pt[0], k.pt[1], k.response) for k in OracleImage._gftt.detect(self.image)]

    # opencv_image.cpp:75-154
    def track_keypoints(self, next_image, curr_keypoints, next_keypoints):
        """curr_keypoints (n,2) f64; next_keypoints (n,2) f64 predictions or None.
        Returns next_keypoints (n,2) f64 (updated where status) and status (n,) uint8."""
        n = len(curr_keypoints)
        curr = np.asarray(curr_keypoints, dtype=np.float64).reshape(n, 2)
        curr_cv = curr.astype(np.float32)
        if next_keypoints is not None and len(next_keypoints) > 0:
            nxt = np.asarray(next_keypoints, dtype=np.float64).reshape(n, 2).copy()
            next_cv = nxt.astype(np.float32)
        else:
            nxt = np.zeros((n, 2))
            next_cv = curr_cv.copy()
        status = np.zeros(n, np.uint8)
        if n > 0:
            rows, cols = self.image.shape
            next_cv, st, _ = cv2.calcOpticalFlowPyrLK(self.image, next_image.image, curr_cv, next_cv.copy(),
                                                     winSize=(21, 21), maxLevel=LEVEL_NUM, criteria=LK_CRITERIA,
                                                     flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
            status = st.ravel().astype(np.uint8).copy()
            oob = (next_cv[:, 0] < 20) | (next_cv[:, 0] >= cols - 20) | (next_cv[:, 1] < 20) | (next_cv[:, 1] >= rows - 20)
            status[oob] = 0
            d = (next_cv - curr_cv).astype(np.float64)       # Point2f difference, then vector<2> (double) norm
            flow = np.sqrt(d[:, 0] ** 2 + d[:, 1] ** 2)
            status[(status != 0) & (flow > rows // 4)] = 0    # image.rows / 4 is integer division
            rev, rst, _ = cv2.calcOpticalFlowPyrLK(next_image.image, self.image, next_cv, curr_cv.copy(),
                                                  winSize=(21, 21), maxLevel=LEVEL_NUM, criteria=LK_CRITERIA,
                                                  flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
            rst = rst.ravel()
            dd = (curr_cv - rev).astype(np.float64)
            back = np.sqrt(dd[:, 0] ** 2 + dd[:, 1] ** 2)     # cv::norm(Point2f) is computed in double
            status[(status != 0) & ((rst == 0) | (back > 0.5))] = 0
            ok = status != 0
            nxt[ok] = next_cv[ok].astype(np.float64)
        return nxt, status
