"""Drives synthetic streams through the CUDA session API (xrslam_b200.capi.Context) the way
xrslam-pc/player drives XRSLAM (IMU samples in time order, then the image, main.cpp:116-165).
Test / bench infrastructure only."""
import numpy as np
from xrslam_b200 import capi


def params_for(stream):
    p = capi.default_params(stream.w, stream.h)
    p['fx'], p['fy'], p['cx'], p['cy'] = stream.K
    return p


def imu_batch(sessions, rows):
    a = np.zeros(len(rows), capi.imu_dtype)
    a['t'] = [r[0] for r in rows]
    a['w'] = [r[1:4] for r in rows]
    a['a'] = [r[4:7] for r in rows]
    return np.asarray(sessions, np.int32), a


def run_streams_cuda(streams, n_frames, frames=None, ctx=None, collect=True, bootstrap=True):
    """Runs len(streams) sessions for n_frames frames.  frames[s][k]: optional pre-rendered u8 images.
    Returns (ctx, outputs) with outputs[s] = list of (t, q, p) poses as the player would write them."""
    S = len(streams)
    ctx = ctx or capi.Context(params_for(streams[0]), S)
    imus = [st.imu(n_frames * st.imu_per_cam + 2) for st in streams]
    k_imu = [0] * S
    outputs = [[] for _ in range(S)]
    for s, st in enumerate(streams):
        if not bootstrap:
            break                           # the sessions run the reference initializer instead
        for k in range(min(n_frames, 40)):
            t = st.frame_time(k)
            q, p, v = st.gt_state(t)
            ctx.push_bootstrap_state(s, t, q, p, v)
    for k in range(n_frames):
        sess, rows = [], []
        for s, st in enumerate(streams):
            t_img = st.frame_time(k)
            while k_imu[s] < len(imus[s]) and imus[s][k_imu[s], 0] <= t_img + 1e-9:
                sess.append(s); rows.append(imus[s][k_imu[s]]); k_imu[s] += 1
        if rows:
            ctx.push_imu(*imu_batch(sess, rows))
        ctx.step()
        imgs = [np.ascontiguousarray(frames[s][k] if frames is not None else st.frame(k)) for s, st in enumerate(streams)]
        ctx.push_images(list(range(S)), [im.ctypes.data for im in imgs], [st.frame_time(k) for st in streams])
        if collect:
            for s in range(S):
                po = ctx.get_pose(s)
                if po['valid']:
                    outputs[s].append((float(po['t']), po['q'].copy(), po['p'].copy()))
    return ctx, outputs
