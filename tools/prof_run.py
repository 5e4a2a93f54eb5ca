"""Profiling driver: 64 sessions, primes `prime` frames, then runs `steps` more (run under ncu on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from xrslam_b200 import capi
from oracle.synth_torch import TorchRenderer
from oracle import drive
prime = int(sys.argv[1]) if len(sys.argv) > 1 else 60
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
S = 64
n_frames = prime + steps + 1
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
for k in range(n_frames - 1):
    ctx.push_imu(*batches[k]); ctx.step()
    ctx.push_images(list(range(S)), [fr.data_ptr() + (s * n_frames + k) * P for s in range(S)], [st.frame_time(k) for st in streams], on_device=True)
print('done', ctx.stats())
