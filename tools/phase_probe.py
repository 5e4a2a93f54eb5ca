"""Diagnostic: host wall time per step phase for the bench workload (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xrslam_b200 import capi
from oracle.synth_torch import TorchRenderer
from oracle import drive
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bench.N_STREAMS = S
PRIME = int(sys.argv[2]) if len(sys.argv) > 2 else 190
n_frames = PRIME + 30 + 1
streams, imus = bench.build_streams(0, n_frames)
dev = torch.device('cuda', 0)
fr = torch.empty((S, n_frames, 480, 640), dtype=torch.uint8, device=dev)
for s, st in enumerate(streams):
    TorchRenderer(st, dev).render_sequence(n_frames, out=fr[s])
torch.cuda.synchronize()
batches = bench.make_step_inputs(streams, imus, n_frames)
ctx = capi.Context(drive.params_for(streams[0]), S)
for s, st in enumerate(streams):
    for k in range(40):
        q, p, v = st.gt_state(st.frame_time(k)); ctx.push_bootstrap_state(s, st.frame_time(k), q, p, v)
P = 640 * 480
tp = [0, 0, 0]
lane_hist = []
import ctypes as _ct
for k in range(n_frames - 1):
    if k == PRIME:
        _arr = (_ct.c_ulonglong * 32)(); capi.lib.xrvio_debug_kprof(_arr, 1)
    if k == PRIME:
        base = ctx.phase_times().copy(); tp = [0, 0, 0]; t00 = time.perf_counter()
    t0 = time.perf_counter(); ctx.push_imu(*batches[k]); t1 = time.perf_counter(); ctx.step(); t2 = time.perf_counter()
    if k >= PRIME:
        _b = (_ct.c_double * 64)(); _nl = capi.lib.xrvio_debug_lane_busy(ctx._h, _b, 64)
        lane_hist.append([_b[i] for i in range(_nl)])
    ctx.push_images(list(range(S)), [fr.data_ptr() + (s * n_frames + k) * P for s in range(S)], [st.frame_time(k) for st in streams], on_device=True)
    t3 = time.perf_counter(); tp[0] += t1 - t0; tp[1] += t2 - t1; tp[2] += t3 - t2
torch.cuda.synchronize()
tot = time.perf_counter() - t00
ph = ctx.phase_times() - base
n = n_frames - 1 - PRIME
print('steps', n, 'total ms/step', tot / n * 1e3, 'python push_imu/step/push_images ms', [x / n * 1e3 for x in tp])
if len(lane_hist) > 2:
    lh = np.diff(np.array(lane_hist), axis=0)            # per step, per lane: wall time inside lane_step
    print('lane-step ms: mean %.2f  p50 %.2f  p90 %.2f  p99 %.2f  max %.2f | per-step max over lanes: mean %.2f p50 %.2f max %.2f | lane means %s'
          % (lh.mean(), np.percentile(lh, 50), np.percentile(lh, 90), np.percentile(lh, 99), lh.max(), lh.max(axis=1).mean(),
             np.percentile(lh.max(axis=1), 50), lh.max(axis=1).max(), np.round(lh.mean(axis=0), 2).tolist()))
    print('slowest lane-steps (step, lane, ms):', sorted([(int(i), int(j), round(float(lh[i, j]), 2)) for i in range(lh.shape[0]) for j in range(lh.shape[1])], key=lambda x: -x[2])[:8])
names = {0: 'A host prep', 1: 'A launch+preint sync', 2: 'B host klt prep', 3: 'B klt launch+sync', 4: 'C host ransac/order', 5: 'C poisson', 6: 'C host apply', 7: 'D host', 8: 'D poisson', 9: 'D host attach',
         10: 'E mirror', 11: 'E preint run', 12: 'F tri+build', 13: 'F solve run', 14: 'G writeback/manage', 15: 'G tri run', 16: 'H host preint build', 17: 'H preint run', 18: 'H build problems', 19: 'H solve run', 20: 'I gates', 21: 'I marg', 22: 'latest', 30: 'push_images'}
for i in range(32):
    if ph[i] > 0:
        print('%2d %-24s %8.3f ms/step' % (i, names.get(i, ''), ph[i] / n))
import ctypes
arr = (ctypes.c_ulonglong * 32)()
capi.lib.xrvio_debug_kprof(arr, 0)
a = list(arr)
npb = max(a[31], 1)
lab = {0: 'J zero', 1: 'J rpe', 2: 'J rot', 3: 'J pre simt', 4: 'J pre coop', 5: 'J marg', 6: 'J reduce', 8: 'C zero/sync', 9: 'C rpe', 10: 'C rot', 11: 'C pre simt', 12: 'C pre coop', 13: 'C marg',
       14: 'C reduce', 17: 'loop pre-cand', 18: 'D/alpha', 19: 'gauss_newton', 20: 'dogleg', 21: 'cand+ambient', 22: 'post-accept eval etc', 23: 'apply_scaling', 24: 'GN schur', 25: 'GN cholesky', 26: 'GN substitution', 27: 'GN calls', 28: 'GN cholesky failures', 29: 'GN setup (mu D^2)'}
print('window problems', npb)
for i in range(31):
    if a[i]:
        if i in (27, 28):
            print('%2d %-22s %10.3f per problem' % (i, lab.get(i, ''), a[i] / npb))
        else:
            print('%2d %-22s %10.1f kcycles/problem' % (i, lab.get(i, ''), a[i] / npb / 1e3))
