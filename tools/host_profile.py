"""Profile of the session layer's HOST glue without a GPU (test infrastructure, uses tests/cuemu):
  1. python tools/host_profile.py record   -- 8 sessions x 76 frames on the SIMT emulator (~10 min), every device->host copy
                                              is stored in /tmp/xrvio_host_profile.rec
  2. python tools/host_profile.py replay   -- re-runs the same steps 30 times with the kernels skipped and the recorded
                                              copies served instead (the host glue is deterministic given those), sampling the
                                              instruction pointer at 1 kHz; prints the host time per lane-step and the hottest
                                              functions / source lines."""
import collections
import os
import subprocess
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'cuemu'))
REC = '/tmp/xrvio_host_profile.rec'
mode = sys.argv[1] if len(sys.argv) > 1 else 'replay'
os.environ['XRVIO_LANES'] = '1'
os.environ.setdefault('XRVIO_HOST_THREADS', '1')
os.environ['XRVIO_EMU_DEBUGINFO'] = '1'
os.environ['CUEMU_RECORD' if mode == 'record' else 'CUEMU_REPLAY'] = REC
import build_emu                      # noqa: E402
from xrslam_b200 import capi          # noqa: E402
capi.LIB_PATH = build_emu.build()
from oracle.synth import Stream       # noqa: E402
from oracle import drive              # noqa: E402

S, N = 8, 76
streams = [Stream(1000 + i) for i in range(S)]
imus = [st.imu(N * st.imu_per_cam + 2) for st in streams]
zero = np.zeros((480, 640), np.uint8)
frames = [[np.ascontiguousarray(st.frame(k)) for k in range(N)] for st in streams] if mode == 'record' else None
dll = capi.lib.load()
HOST = (0, 2, 4, 6, 7, 9, 10, 12, 14, 16, 18, 20, 22)


def run():
    ctx = capi.Context(drive.params_for(streams[0]), S)
    for s, st in enumerate(streams):
        for k in range(40):
            t = st.frame_time(k)
            q, p, v = st.gt_state(t)
            ctx.push_bootstrap_state(s, t, q, p, v)
    ki = [0] * S
    for k in range(N):
        sess, rows = [], []
        for s, st in enumerate(streams):
            while ki[s] < len(imus[s]) and imus[s][ki[s], 0] <= st.frame_time(k) + 1e-9:
                sess.append(s); rows.append(imus[s][ki[s]]); ki[s] += 1
        ctx.push_imu(*drive.imu_batch(sess, rows))
        ctx.step()
        ctx.push_images(list(range(S)), [(frames[s][k] if frames else zero).ctypes.data for s in range(S)], [st.frame_time(k) for st in streams])
    ph = ctx.phase_times().copy()
    ctx.close()
    return ph


if mode == 'record':
    run()
    print('recorded', os.path.getsize(REC), 'bytes')
else:
    R = 30
    dll.cuemu_prof_start(1000)
    t0 = time.time()
    for _ in range(R):
        dll.cuemu_replay_rewind()
        ph = run()
    dll.cuemu_prof_dump(b'/tmp/xrvio_host_profile.samples')
    print('host glue: %.3f ms per lane-step of %d sessions (%d runs in %.1f s)' % (sum(ph[i] for i in HOST) / (N - 1), S, R, time.time() - t0))
    lines = open('/tmp/xrvio_host_profile.samples').read().split()
    base = int(lines[0], 16)
    cnt = collections.Counter(int(x, 16) - base for x in lines[1:])
    size = os.path.getsize(capi.LIB_PATH)
    cnt = {a: c for a, c in cnt.items() if 0 <= a < 16 * size}
    addrs = ['0x%x' % a for a in cnt]
    out = subprocess.run(['addr2line', '-e', capi.LIB_PATH, '-f', '-C'] + addrs, capture_output=True, text=True).stdout.split('\n')
    byfunc, byline = collections.Counter(), collections.Counter()
    for i, a in enumerate(cnt):
        byfunc[out[2 * i][:110]] += cnt[a]
        byline[os.path.basename(out[2 * i + 1])] += cnt[a]
    tot = sum(cnt.values())
    print('samples inside the library:', tot)
    for f, c in byfunc.most_common(20):
        print('%5.1f%% %s' % (100 * c / tot, f))
    print('--- by source line')
    for f, c in byline.most_common(20):
        print('%5.1f%% %s' % (100 * c / tot, f))
