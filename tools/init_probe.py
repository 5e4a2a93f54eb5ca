"""One stream through the CUDA session WITHOUT bootstrap states: the session runs the reference initializer
(session_init.inl).  Prints the initializer outcome and the trajectory error against the committed oracle
trajectory (tests/golden/stream7_init_traj.npz) and against the generator's ground truth.
Usage: python tools/init_probe.py [seed] [frames]"""
import os
import sys
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.synth import Stream          # noqa: E402
from oracle import drive                 # noqa: E402
from xrslam_b200 import capi             # noqa: E402


def ate(a, b):
    a, b = np.asarray(a), np.asarray(b)
    ca, cb = a.mean(0), b.mean(0)
    H = (b - cb).T @ (a - ca)
    U, _, Vt = np.linalg.svd(H)
    D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    R = (U @ D @ Vt).T
    al = (R @ (b - cb).T).T + ca
    return float(np.sqrt(np.mean(np.sum((al - a) ** 2, axis=1))))


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 7
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 70
    s = Stream(seed)
    t0 = time.time()
    ctx = capi.Context(drive.params_for(s), 1)
    imu = s.imu(n * s.imu_per_cam + 2)
    k_imu = 0
    outs = []
    last = None
    for k in range(n):
        t_img = s.frame_time(k)
        rows = []
        while k_imu < len(imu) and imu[k_imu, 0] <= t_img + 1e-9:
            rows.append(imu[k_imu]); k_imu += 1
        if rows:
            ctx.push_imu(*drive.imu_batch([0] * len(rows), rows))
        ctx.step()
        im = np.ascontiguousarray(s.frame(k))
        ctx.push_images([0], [im.ctypes.data], [t_img])
        info = ctx.get_init_info(0)
        key = (info['attempts'], info['last_fail'])
        if key != last:
            print('frame', k, 'state', ctx.get_state(0), info, flush=True)
            last = key
        po = ctx.get_pose(0)
        if po['valid']:
            outs.append([float(po['t'])] + list(po['q']) + list(po['p']))
    got = np.array(outs)
    print('poses', len(got), 'wall %.1f s' % (time.time() - t0), dict(zip(ctx.stats().dtype.names, ctx.stats().tolist())), flush=True)
    if len(got):
        gt = np.array([s.gt_state(t)[1] for t in got[:, 0]])
        print('ATE vs ground truth %.5f m' % ate(gt, got[:, 5:8]))
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'stream%d_init_traj.npz' % seed)
        if os.path.exists(path):
            g = np.load(path)
            ref = g['traj']
            print('oracle: poses', len(ref), 'first t', ref[0, 0], 'info', g['info'], '| cuda first t', got[0, 0])
            m = min(len(ref), len(got))
            if np.allclose(ref[:m, 0], got[:m, 0]):
                print('ATE vs oracle %.5f m, raw max %.5f m' % (ate(ref[:m, 5:8], got[:m, 5:8]), np.linalg.norm(ref[:m, 5:8] - got[:m, 5:8], axis=1).max()))
    ctx.close()


if __name__ == '__main__':
    main()
