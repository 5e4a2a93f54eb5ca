"""Host logic of the session layer (xrslam_b200/csrc/session_host.hpp) compiled for the CPU with g++ and
checked against the oracle: 2-point rotation RANSAC with the reference's libstdc++ sampling, camera model,
track bookkeeping.  No GPU needed."""
import os
import subprocess
import numpy as np
import pytest

from oracle.ransac import find_rotation_matrix
from oracle.lie import expmap, qmat
from oracle.pipeline import remove_k, apply_k

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp('host') / 'host_probe'
    cuda_inc = '/usr/local/cuda/include'
    subprocess.check_call(['g++', '-O1', '-std=c++17', '-I', cuda_inc, os.path.join(ROOT, 'tests', 'host', 'host_probe.cpp'), '-o', str(out)])

    def run(text):
        return subprocess.run([str(out)], input=text, capture_output=True, text=True, check=True).stdout.strip().split('\n')
    return run


def _case(rng, n, noise, outliers=0):
    b1 = rng.normal(size=(n, 3)) + np.array([0, 0, 4.0])
    b1 /= np.linalg.norm(b1, axis=1, keepdims=True)
    R = qmat(expmap(rng.normal(size=3) * 0.03))
    b2 = b1 @ R.T + rng.normal(size=(n, 3)) * noise
    b2 /= np.linalg.norm(b2, axis=1, keepdims=True)
    for k in range(outliers):
        b2[k] = -b2[k]
    return b1, b2


def test_rotation_ransac_matches_oracle(probe):
    rng = np.random.default_rng(9)
    thr = np.deg2rad(10.0)
    for (n, noise, outl) in ((150, 1e-4, 0), (40, 2e-3, 6), (2, 1e-4, 0), (1, 1e-4, 0), (200, 1e-3, 40)):
        b1, b2 = _case(rng, n, noise, outl)
        txt = 'ransac %d %.17g\n' % (n, thr) + '\n'.join(' '.join('%.17g' % v for v in np.concatenate([b1[i], b2[i]])) for i in range(n)) + '\n'
        out = probe(txt)
        R = np.array([float(x) for x in out[0].split()]).reshape(3, 3)
        mask = np.array([int(c) for c in out[1]], np.uint8) if n > 0 and len(out) > 1 else np.zeros(0, np.uint8)
        Ro, mo = find_rotation_matrix(b1, b2, thr)
        assert np.array_equal(mask, mo), (n, noise, outl)
        if n >= 2:
            assert np.abs(R - Ro).max() < 1e-9


def test_camera_model_roundtrip(probe):
    K = (458.654, 457.296, 320.0, 240.0)
    for px in ((100.25, 50.5), (639.0, 479.0), (320.0, 240.0)):
        out = probe('k %r %r %r %r %r %r\n' % (K + px))
        v = [float(x) for x in out[0].split()]
        b = remove_k(np.array(px), np.array(K))
        assert np.allclose(v[:3], b, atol=1e-15)
        assert np.allclose(v[3:], apply_k(b, np.array(K)), atol=1e-12) and np.allclose(v[3:], px, atol=1e-9)


def test_track_bookkeeping(probe):
    out = probe('tracks\n')
    a = out[0].split(); b = out[1].split(); c = out[2].split()
    assert a[0] == '1' and a[1] == '2' and a[2] == '1'
    # landmark stays put when its first frame is removed (track.cpp:25-37); inverse depth is re-anchored in frame 2
    x0 = np.array([float(v) for v in a[4:7]]); x1 = np.array([float(v) for v in b[4:7]])
    assert np.allclose(x0, [0.4, 0.0, 4.0], atol=1e-12)
    assert b[0] == '1' and b[1] == '1' and b[2] == '2'
    c2 = np.array([1.0, 0.0, 0.0])                                    # frame 2 body == camera centre
    assert abs(float(b[3]) - 1.0 / np.linalg.norm(x0 - c2)) < 1e-12
    # the point now lies on frame 2's stored bearing at the preserved range
    bearing = np.array([-0.15, 0.0, 1.0]); bearing /= np.linalg.norm(bearing)
    assert np.allclose(x1, c2 + bearing * np.linalg.norm(x0 - c2), atol=1e-12)
    assert c[0] == '0'
