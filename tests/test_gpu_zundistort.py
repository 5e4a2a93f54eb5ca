"""Undistortion kernel (row N2): xrvio_op_undistort against cv2.undistort (radtan, as the EuRoC reader calls it) and against
cv2.convertMaps + cv2.remap on the ImageUndistorter map (equidistant, TUM-VI reader).  The arithmetic is shared with the
player's CPU reader (xrslam_b200/csrc/undistort_math.h), which tests/test_player_io.py checks without a GPU.

Written after the round-1 GPU budget was spent; checked with the SIMT emulator (tests/cuemu, XRVIO_EMU=1) before its
first run on a B200."""
import os
import numpy as np
import pytest
import cv2

pytestmark = pytest.mark.gpu


def _texture(rng, h, w):
    img = cv2.GaussianBlur(rng.integers(0, 256, size=(h, w), dtype=np.uint8), (0, 0), 2.0)
    cv2.rectangle(img, (w // 4, h // 4), (w // 2, h // 2), 255, -1)
    return img


def test_undistort_radtan_matches_opencv(xrvio):
    rng = np.random.default_rng(1)
    w, h = 752, 480
    imgs = np.stack([_texture(rng, h, w) for _ in range(3)])
    K = [458.654, 457.296, 367.215, 248.375]
    D = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]
    got = xrvio.op_undistort(imgs, 1, K, D)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    for i in range(3):
        ref = cv2.undistort(imgs[i], Km, np.array(D, np.float32).reshape(4, 1))
        diff = np.abs(got[i].astype(int) - ref.astype(int))
        assert (diff == 0).mean() > 0.999 and diff.max() <= 12


def test_undistort_equidistant_matches_opencv(xrvio):
    rng = np.random.default_rng(2)
    wt = ht = 512
    img = _texture(rng, ht, wt)
    Kt = [190.978, 190.973, 254.93, 256.897]
    Dt = [0.0034823894, 0.0007150348, -0.0020532361, 0.00020293673]
    got = xrvio.op_undistort(img[None], 2, Kt, Dt)[0]
    u, v = np.meshgrid(np.arange(wt, dtype=np.float64), np.arange(ht, dtype=np.float64))
    x, y = (u - Kt[2]) / Kt[0], (v - Kt[3]) / Kt[1]
    r = np.sqrt(x * x + y * y)
    th = np.arctan(r)
    thd = th * (1 + Dt[0] * th ** 2 + Dt[1] * th ** 4 + Dt[2] * th ** 6 + Dt[3] * th ** 8)
    sc = np.where(r > 1e-8, thd / np.maximum(r, 1e-300), 1.0)
    mx = np.where(r < 1e-10, u, Kt[0] * x * sc + Kt[2]).astype(np.float32)
    my = np.where(r < 1e-10, v, Kt[1] * y * sc + Kt[3]).astype(np.float32)
    m1, m2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
    ref = cv2.remap(img, m1, m2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert (diff == 0).mean() > 0.999 and diff.max() <= 12
