"""I/O of the headless player (player/*.hpp): PNG / PGM decoding, EuRoC and TUM-VI csv
readers with the reference's event order, the fixed-point undistortion against cv2, the trajectory writers'
line formats, and that the player binary refuses to run without a GPU.  CPU only."""
import os
import subprocess
import numpy as np
import pytest
import cv2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def probe(tmp_path_factory):
    out = tmp_path_factory.mktemp('player') / 'player_probe'
    subprocess.check_call(['g++', '-O1', '-std=c++17', os.path.join(ROOT, 'tests', 'host', 'player_probe.cpp'), '-o', str(out), '-lz'])

    def run(*args, check=True):
        r = subprocess.run([str(out)] + [str(a) for a in args], capture_output=True, text=True)
        if check:
            assert r.returncode == 0, (r.stdout, r.stderr)
        return r
    return run


def _texture(rng, h, w):
    img = cv2.GaussianBlur(rng.integers(0, 256, size=(h, w), dtype=np.uint8), (0, 0), 2.0)
    cv2.rectangle(img, (w // 4, h // 4), (w // 2, h // 2), 255, -1)
    cv2.line(img, (0, h - 1), (w - 1, 0), 0, 3)
    return img


def test_png_and_pgm_decoding(probe, tmp_path):
    rng = np.random.default_rng(0)
    gray = _texture(rng, 97, 131)
    bgr = np.stack([_texture(rng, 97, 131) for _ in range(3)], axis=2)
    cases = {'gray.png': gray, 'bgr.png': bgr, 'gray16.png': (gray.astype(np.uint16) << 8) | 0x5a, 'bgra.png': np.dstack([bgr, gray])}
    for name, img in cases.items():
        path = str(tmp_path / name)
        assert cv2.imwrite(path, img, [cv2.IMWRITE_PNG_COMPRESSION, 6])
        r = probe('png', path, tmp_path / 'o.raw')
        w, h = [int(x) for x in r.stdout.split()]
        got = np.fromfile(str(tmp_path / 'o.raw'), np.uint8).reshape(h, w)
        ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)                       # euroc_dataset_reader.cpp:60,76-77
        if ref.dtype == np.uint16:
            ref = (ref >> 8).astype(np.uint8)                                # documented: the high byte of 16-bit samples
        if ref.ndim == 3:
            ref = cv2.cvtColor(ref, cv2.COLOR_BGR2GRAY if ref.shape[2] == 3 else cv2.COLOR_BGRA2GRAY)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), name
    path = str(tmp_path / 'g.pgm')
    cv2.imwrite(path, gray)
    r = probe('png', path, tmp_path / 'o.raw')
    assert np.array_equal(np.fromfile(str(tmp_path / 'o.raw'), np.uint8).reshape(gray.shape), gray)
    (tmp_path / 'bad.png').write_bytes(b'\x89PNG\r\n\x1a\n' + b'\0' * 10)
    assert probe('png', tmp_path / 'bad.png', tmp_path / 'o.raw', check=False).returncode == 1


def test_undistortion_matches_opencv(probe, tmp_path):
    rng = np.random.default_rng(1)
    w, h = 752, 480
    img = _texture(rng, h, w)
    img.tofile(str(tmp_path / 'in.raw'))
    K = [458.654, 457.296, 367.215, 248.375]
    D = [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05]                      # configs/euroc_sensor.yaml:44-46
    probe('undistort', 'radtan', 1, w, h, *K, *D, tmp_path / 'in.raw', tmp_path / 'out.raw')
    got = np.fromfile(str(tmp_path / 'out.raw'), np.uint8).reshape(h, w)
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1]], np.float32)
    ref = cv2.undistort(img, Km, np.array(D, np.float32).reshape(4, 1))              # euroc_dataset_reader.cpp:62-72
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert (diff == 0).mean() > 0.999 and diff.max() <= 12, ((diff == 0).mean(), diff.max())
    # TUM-VI: equidistant map as float32, convertMaps + remap (image_undistorter.h:13-42)
    Kt = [190.978, 190.973, 254.93, 256.897]
    Dt = [0.0034823894, 0.0007150348, -0.0020532361, 0.00020293673]
    wt = ht = 512
    img2 = _texture(rng, ht, wt)
    img2.tofile(str(tmp_path / 'in2.raw'))
    probe('undistort', 'equidistant', 0, wt, ht, *Kt, *Dt, tmp_path / 'in2.raw', tmp_path / 'out2.raw')
    got2 = np.fromfile(str(tmp_path / 'out2.raw'), np.uint8).reshape(ht, wt)
    u, v = np.meshgrid(np.arange(wt, dtype=np.float64), np.arange(ht, dtype=np.float64))
    x, y = (u - Kt[2]) / Kt[0], (v - Kt[3]) / Kt[1]
    r = np.sqrt(x * x + y * y)
    th = np.arctan(r)
    thd = th * (1 + Dt[0] * th ** 2 + Dt[1] * th ** 4 + Dt[2] * th ** 6 + Dt[3] * th ** 8)
    sc = np.where(r > 1e-8, thd / np.maximum(r, 1e-300), 1.0)
    mx = np.where(r < 1e-10, u, Kt[0] * x * sc + Kt[2]).astype(np.float32)
    my = np.where(r < 1e-10, v, Kt[1] * y * sc + Kt[3]).astype(np.float32)
    m1, m2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
    ref2 = cv2.remap(img2, m1, m2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    diff2 = np.abs(got2.astype(int) - ref2.astype(int))
    assert (diff2 == 0).mean() > 0.999 and diff2.max() <= 12, ((diff2 == 0).mean(), diff2.max())


def _make_tree(root, crlf, n_img=4, n_imu=25):
    rng = np.random.default_rng(5)
    os.makedirs(os.path.join(root, 'cam0', 'data'))
    os.makedirs(os.path.join(root, 'imu0'))
    nl = '\r\n' if crlf else '\n'
    t0 = 1403636579763555584
    cam_t = [t0 + k * 50000000 for k in range(n_img)]
    imu_t = [t0 + k * 5000000 for k in range(n_imu)]
    imgs = []
    with open(os.path.join(root, 'cam0', 'data.csv'), 'w', newline='') as f:
        f.write('#timestamp [ns],filename' + nl)
        for t in cam_t:
            f.write('%d,%d.png%s' % (t, t, nl))
            img = rng.integers(0, 256, size=(48, 64), dtype=np.uint8)
            cv2.imwrite(os.path.join(root, 'cam0', 'data', '%d.png' % t), img)
            imgs.append(img)
    imu = rng.normal(size=(n_imu, 6))
    with open(os.path.join(root, 'imu0', 'data.csv'), 'w', newline='') as f:
        f.write('#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y,w_RS_S_z,a_RS_S_x [m s^-2],a_RS_S_y,a_RS_S_z' + nl)
        for t, r in zip(imu_t, imu):
            f.write('%d,%s%s' % (t, ','.join('%.17g' % v for v in r), nl))
    return cam_t, imu_t, imu, imgs


@pytest.mark.parametrize('scheme,crlf', [('euroc', True), ('euroc', False), ('tum', False)])
def test_dataset_reader_event_stream(probe, tmp_path, scheme, crlf):
    root = str(tmp_path / 'mav0')
    cam_t, imu_t, imu, imgs = _make_tree(root, crlf)
    yaml = tmp_path / 'sensor.yaml'
    yaml.write_text('%YAML:1.0\nimu:\n  noise: 1\ncam0:\n  resolution: [64, 48]\n  distortion_model: radtan\n  intrinsics: [60.0, 60.0, 32.0, 24.0] # fu\n'
                    '  camera_distortion_flag: 0     # off\n  distortion: [-0.1, 0.01, 0.0, 0.0]\n  time_offset: 0.002\n')
    out = probe('events', '%s://%s' % (scheme, root), yaml).stdout.strip().split('\n')
    # expected order: stable sort by time of [cameras..., (gyro, acc) per imu row]
    ev = [(t * 1e-9 + 0.002, 0, 'C', i) for i, t in enumerate(cam_t)]
    for i, t in enumerate(imu_t):
        ev.append((t * 1e-9, 1, 'G', i)); ev.append((t * 1e-9, 1, 'A', i))
    order = sorted(range(len(ev)), key=lambda k: ev[k][0])            # Python's sort is stable, like std::stable_sort
    assert len(out) == len(ev)
    for line, k in zip(out, order):
        t, _, kind, i = ev[k]
        f = line.split()
        assert f[0] == kind
        assert float(f[1]) == t
        if kind == 'C':
            assert int(f[2]) == 64 and int(f[3]) == 48 and int(f[4]) == int(imgs[i].sum())
        else:
            vals = imu[i, 0:3] if kind == 'G' else imu[i, 3:6]
            assert np.array_equal(np.array([float(x) for x in f[2:5]]), vals) and float(f[5]) == t
    assert probe('events', 'bag://' + root, yaml, check=False).returncode == 1


def test_trajectory_writer_formats(probe, tmp_path):
    probe('write', 'tum', tmp_path / 't.txt')
    lines = (tmp_path / 't.txt').read_text().strip().split('\n')
    t, q, p = 1403636580.838555574, (0.1234567891234, -0.5, 0.25, 0.8), (1.5, -2.25, 3.000000001)
    assert lines[0] == '%.18e %.9e %.9e %.9e %.7e %.7e %.7e %.7e' % (t, p[0], p[1], p[2], q[0], q[1], q[2], q[3])      # trajectory_writer.h:69-75
    assert len(lines) == 2
    probe('write', 'csv', tmp_path / 't.csv')
    assert (tmp_path / 't.csv').read_text().strip() == '%.18e,%.9e,%.9e,%.9e,%.7e,%.7e,%.7e,%.7e' % (t, p[0], p[1], p[2], q[0], q[1], q[2], q[3])
    a = np.loadtxt(str(tmp_path / 't.txt'))
    assert a.shape == (2, 8)


def test_player_binary_fails_loudly_without_gpu(tmp_path):
    from xrslam_b200 import capi, build
    exe = build.build_player()
    assert os.path.exists(exe)
    if capi.lib.xrvio_device_count() > 0:
        pytest.skip('a GPU is present')
    gold = os.path.join(ROOT, 'tests', 'golden')
    r = subprocess.run([exe, '-sc', os.path.join(gold, 'synth_slam.yaml'), '-dc', os.path.join(gold, 'synth_sensor.yaml'), 'euroc://' + str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode != 0 and 'create SLAM success: 0' in r.stdout
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and 'usage' in r.stderr
