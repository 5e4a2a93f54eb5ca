"""BASELINE config 1 plumbing check: the headless player binary (player/main.cpp) plays a
synthetic EuRoC-layout tree through libxrvio.so's XRSLAM C ABI -- csv readers, PNG decoding, gyroscope /
accelerometer pairing in the shim, the initializer, tracking, TUM writer -- and the written trajectory is compared
with the oracle's (tests/golden/stream7_init_traj.npz).

Written after the round-1 GPU budget was spent; checked with the SIMT emulator (tests/cuemu, XRVIO_EMU=1) before its
first run on a B200."""
import os
import subprocess
import numpy as np
import pytest
import cv2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ate(a, b):
    a, b = np.asarray(a), np.asarray(b)
    ca, cb = a.mean(0), b.mean(0)
    H = (b - cb).T @ (a - ca)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    R = (U @ D @ Vt).T
    al = (R @ (b - cb).T).T + ca
    return float(np.sqrt(np.mean(np.sum((al - a) ** 2, axis=1))))


def write_mav0(root, stream, n_frames):
    """EuRoC ASL layout with CRLF line ends, integer-nanosecond stamps (euroc_dataset_reader.h:36-47,86-97)."""
    os.makedirs(os.path.join(root, 'cam0', 'data'))
    os.makedirs(os.path.join(root, 'imu0'))
    with open(os.path.join(root, 'cam0', 'data.csv'), 'w', newline='') as f:
        f.write('#timestamp [ns],filename\r\n')
        for k in range(n_frames):
            ns = int(round(stream.frame_time(k) * 1e9))
            f.write('%d,%d.png\r\n' % (ns, ns))
            cv2.imwrite(os.path.join(root, 'cam0', 'data', '%d.png' % ns), stream.frame(k))
    imu = stream.imu(n_frames * stream.imu_per_cam + 2)
    with open(os.path.join(root, 'imu0', 'data.csv'), 'w', newline='') as f:
        f.write('#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y [rad s^-1],w_RS_S_z [rad s^-1],a_RS_S_x [m s^-2],a_RS_S_y [m s^-2],a_RS_S_z [m s^-2]\r\n')
        for r in imu:
            if r[0] > stream.frame_time(n_frames - 1) + 1e-9:
                break
            f.write('%d,%s\r\n' % (int(round(r[0] * 1e9)), ','.join('%.17g' % v for v in r[1:7])))


def test_player_runs_synthetic_euroc_tree(xrvio, golden_dir, tmp_path):
    from oracle.synth import Stream
    from xrslam_b200 import build
    exe = build.PLAYER if os.path.exists(build.PLAYER) else build.build_player()      # built by __graft_entry__.build(); no rebuild on the GPU box
    exe = os.path.abspath(exe)
    s = Stream(7)
    root = str(tmp_path / 'mav0')
    write_mav0(root, s, 70)
    out = str(tmp_path / 'traj.txt')
    env = dict(os.environ)
    if os.environ.get('XRVIO_EMU') == '1':          # emulator run (tests/cuemu): the binary resolves libxrvio.so to the emulated build
        from xrslam_b200 import capi
        d = os.path.join(os.path.dirname(capi.LIB_PATH), 'as_libxrvio')
        os.makedirs(d, exist_ok=True)
        link = os.path.join(d, 'libxrvio.so')
        if not os.path.exists(link):
            os.symlink(capi.LIB_PATH, link)
        env['LD_LIBRARY_PATH'] = d + ':' + env.get('LD_LIBRARY_PATH', '')
    r = subprocess.run([exe, '-sc', os.path.join(golden_dir, 'synth_slam.yaml'), '-dc', os.path.join(golden_dir, 'synth_sensor.yaml'),
                        '--tum', out, 'euroc://' + root], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert 'create SLAM success: 1' in r.stdout
    traj = np.loadtxt(out)
    ref = np.load(os.path.join(golden_dir, 'stream7_init_traj.npz'))['traj']
    assert len(traj) >= len(ref) - 2
    gt = np.array([s.gt_state(t)[1] for t in traj[:, 0]])
    assert _ate(gt, traj[:, 1:4]) < 0.02
    # same frames as the oracle run (time stamps go through integer nanoseconds: compare on rounded times)
    common = [(i, j) for i, t in enumerate(traj[:, 0]) for j in [int(np.argmin(np.abs(ref[:, 0] - t)))] if abs(ref[j, 0] - t) < 1e-6]
    assert len(common) >= len(ref) - 2
    a = np.array([traj[i, 1:4] for i, _ in common]); b = np.array([ref[j, 5:8] for _, j in common])
    assert _ate(b, a) < 5e-3
