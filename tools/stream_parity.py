#!/usr/bin/env python
"""Live parity of the session path against the oracle pipeline on synthetic streams other than the two the golden trajectories
were recorded from (seeds 3 and 7): the oracle runs the stream on the CPU, the library runs it on the GPU (or on the SIMT emulator
with XRVIO_EMU=1), and the trajectories are compared.  Test infrastructure, like tests/.

  python tools/stream_parity.py [n_frames] [seed ...]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

if os.environ.get('XRVIO_EMU') == '1':
    sys.path.insert(0, os.path.join(ROOT, 'tests', 'cuemu'))
    import build_emu
    from xrslam_b200 import capi
    capi.LIB_PATH = build_emu.build()
    os.environ.setdefault('XRVIO_LANES', '1')
from oracle.synth import Stream
from oracle.pipeline import run_stream
from oracle import drive


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 72
    seeds = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 9]
    worst = 0.0
    for seed in seeds:
        t0 = time.time()
        s = Stream(seed)
        ref = np.array([[t] + list(q) + list(p) for (t, q, p) in run_stream(s, n).outputs])
        ctx, outs = drive.run_streams_cuda([Stream(seed)], n)
        got = np.array([[t] + list(q) + list(p) for (t, q, p) in outs[0]])
        st = ctx.stats()
        ctx.close()
        same = got.shape == ref.shape and np.allclose(got[:, 0], ref[:, 0])
        dp = np.linalg.norm(got[:, 5:8] - ref[:, 5:8], axis=1).max() if same else float('nan')
        worst = max(worst, dp) if same else float('inf')
        print('seed %d: %d poses, same frames %s, max |dp| vs oracle %.3e m, window solves %d, marginalisations %d  (%.0f s)'
              % (seed, len(got), same, dp, int(st['window_solves']), int(st['marginalizations']), time.time() - t0), flush=True)
    print('worst %.3e m' % worst)
    return 0 if worst <= 1e-3 else 1


if __name__ == '__main__':
    sys.exit(main())
