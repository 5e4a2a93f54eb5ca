#!/usr/bin/env python
"""bench.py -- VIO frames/sec over batched synthetic EuRoC-shaped streams (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (libxrvio.so, session C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the restated reference CPU path (oracle/)
                                                            timed on the box's host cores

Workload (config.workload "64x640x480_streams"): 64 concurrent synthetic streams per GPU, 640x480 u8
at 20 Hz, IMU at 200 Hz, ~150 tracks, 10-keyframe window (BASELINE configs[2]).  A step is one
camera frame for every session: 10 IMU samples + 1 image per session pushed through the C ABI and one
xrvio_step (front end + localize solve + window BA / marginalisation or sub-window solve).  All
sessions are brought to TRACKING before anything is timed (bootstrap priming is setup, not warm-up).
`value` is measured with the frames already resident in HBM (device pointers); `e2e` through the same
C ABI with pinned HOST frames (H2D inside the timed region) and a per-step read-back of every pose.
The timed region is exactly K steps = K frames per session through the whole path (`ms_per_step` = time / K).
Every step consumes 64 frames never touched before (19.7 MB of fresh input; the pre-rendered frame
set is far larger than L2), so no L2 flush is needed: config.l2 = "inputs_larger_than_L2".
"""
import argparse
import json
import os
# One hardware work queue per CUDA stream: the library drives 8 lanes x 3 streams, and with the default of 8 connections a lane's
# 20 us filter kernel can sit behind another lane's multi-millisecond bundle adjustment that happens to share its queue (measured:
# the short launches' wait times are 2-3x longer with 8 connections once the lanes run out of phase).  Must be set before the
# CUDA context exists, i.e. before torch / the library touch the device.
os.environ.setdefault('CUDA_DEVICE_MAX_CONNECTIONS', '32')
import sys as _sys
if '--impl' in _sys.argv and 'reference' in _sys.argv:
    # one single-threaded process per stream: stop BLAS / OpenMP pools from oversubscribing the host
    for _v in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
        os.environ[_v] = '1'
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STREAMS = 64
W_IMG, H_IMG = 640, 480
# Frames before the warm-up steps.  The sessions bootstrap at frame 36; the timed window starts >= 150 frames (7.5 s of stream
# time, ~38 window BAs and marginalisations per session) after that, i.e. in steady state: full 10-keyframe windows, full-rank
# priors, track lengths and landmark counts at their stationary level (frames 45-65 were the young-map transient).
PRIME_FRAMES = int(os.environ.get('XRVIO_BENCH_PRIME', '190'))      # (the override exists for profiler runs only: fewer frames to render under ncu)
# algorithmic HBM bytes per kernel family and frame (DESIGN.md "Roofline")
P_IMG = W_IMG * H_IMG
ALGO_BYTES_PER_FRAME = {'clahe': 3.0 * P_IMG, 'pyramid': 1.641 * P_IMG, 'gftt': 1.0 * P_IMG, 'klt': 2.656 * P_IMG}


def committed_profile():
    try:
        with open(os.path.join(ROOT, 'profiles', 'r2_kernels.json')) as f:
            return json.load(f)
    except Exception:
        return None


def read_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)['hbm_gbs'], 'measured'
    except Exception:
        return 6650.0, 'fallback'


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region.  Uses NVML in-process (cheap); falls back to nvidia-smi."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for name, bit in (('hw_slowdown', n.nvmlClocksThrottleReasonHwSlowdown), ('hw_thermal_slowdown', n.nvmlClocksThrottleReasonHwThermalSlowdown),
                          ('sw_thermal_slowdown', n.nvmlClocksThrottleReasonSwThermalSlowdown), ('sw_power_cap', n.nvmlClocksThrottleReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        out = subprocess.run(['nvidia-smi', '-i', str(self.idx), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                             capture_output=True, text=True, timeout=5).stdout.strip().split(',')
        self.samples.append(float(out[0]))
        self.max_mhz = float(out[1])
        for nm, v in zip(names, out[2:]):
            if 'Active' in v and 'Not' not in v:
                self.reasons.add(nm)

    def run(self):
        while not self._halt.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._halt.wait(0.1 if self._nvml is not None else 0.5)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None, 'sm_max_mhz': self.max_mhz,
                'reasons': sorted(self.reasons), 'samples': len(self.samples)}


def pin_to_gpu_numa_node(local_rank, world):
    """Binds this rank (and the library threads it will create) to its share of the CPU cores that are local to its GPU.
    On the 8-GPU boxes GPUs 0-3 and 4-7 hang off different NUMA nodes; an unpinned rank's lane threads and pinned staging
    buffers end up on the far node half of the time (weak-scaling efficiency 0.65 at N = 8).  The ranks that share
    a node split its cores evenly.  Returns a description for the JSON line (None if nothing was done)."""
    if world <= 1 or not hasattr(os, 'sched_setaffinity'):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        n_words = (os.cpu_count() + 63) // 64
        masks = []
        for g in range(world):
            words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(g), n_words)
            masks.append(frozenset(64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1))
        mine = masks[local_rank]
        sharers = [g for g in range(world) if masks[g] == mine]
        allowed = sorted(mine & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        # keep hyper-thread siblings together: split the sorted list into contiguous shares per sharer, interleaved over the two
        # halves of the node's list (cores c and c + ncores/2 are siblings on these hosts)
        k, m = sharers.index(local_rank), len(sharers)
        half = len(allowed) // 2
        lo, hi = allowed[:half], allowed[half:]
        share = lo[k * len(lo) // m:(k + 1) * len(lo) // m] + hi[k * len(hi) // m:(k + 1) * len(hi) // m]
        if not share:
            share = allowed
        os.sched_setaffinity(0, share)
        return {'cpus': len(share), 'first': share[0], 'last': share[-1], 'numa_sharers': m}
    except Exception as e:          # pinning is an optimisation, never a reason to fail the run
        return {'error': str(e)[:80]}


def stream_seeds(rank, n_streams=None):
    """Session -> synthetic stream seed; ranks own disjoint seed ranges (seeds = 1000*rank + i)."""
    return [1000 * rank + i for i in range(n_streams or N_STREAMS)]


def reduce_max_ms(ms, device=None):
    """Whole-job time = max over ranks (the only collective of the harness)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([ms], device=device if device is not None else 'cpu', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def build_streams(rank, n_frames):
    from oracle.synth import Stream
    streams = [Stream(seed, W_IMG, H_IMG) for seed in stream_seeds(rank)]
    imus = [s.imu(n_frames * s.imu_per_cam + 2) for s in streams]
    return streams, imus


def make_step_inputs(streams, imus, n_frames):
    """Per-step IMU batches (session ids + samples), precomputed outside the timed region."""
    from xrslam_b200 import capi
    batches = []
    k_imu = [0] * len(streams)
    for k in range(n_frames):
        sess, rows = [], []
        for s, st in enumerate(streams):
            t_img = st.frame_time(k)
            while k_imu[s] < len(imus[s]) and imus[s][k_imu[s], 0] <= t_img + 1e-9:
                sess.append(s); rows.append(imus[s][k_imu[s]]); k_imu[s] += 1
        a = np.zeros(len(rows), capi.imu_dtype)
        if rows:
            r = np.array(rows)
            a['t'], a['w'], a['a'] = r[:, 0], r[:, 1:4], r[:, 4:7]
        batches.append((np.asarray(sess, np.int32), a))
    return batches


def run_cuda_arm(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from xrslam_b200 import capi
    from oracle.synth_torch import TorchRenderer
    from oracle import drive
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    pin = pin_to_gpu_numa_node(local_rank, world)
    # several ranks share the host: give each its share of cores for the library's bookkeeping threads
    n_cpu = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 8)
    os.environ.setdefault('XRVIO_HOST_THREADS', str(max(8, min(64, n_cpu if pin and 'cpus' in pin else n_cpu // max(world, 1)))))
    K, Wm = args.steps, max(args.warmup, 3)
    n_frames = PRIME_FRAMES + Wm + K + 1
    streams, imus = build_streams(rank, n_frames)
    frames_dev = torch.empty((N_STREAMS, n_frames, H_IMG, W_IMG), dtype=torch.uint8, device=dev)
    for s, st in enumerate(streams):
        TorchRenderer(st, dev).render_sequence(n_frames, out=frames_dev[s])
    torch.cuda.synchronize()
    frames_host = torch.empty(frames_dev.shape, dtype=torch.uint8, pin_memory=True)
    frames_host.copy_(frames_dev)
    batches = make_step_inputs(streams, imus, n_frames)
    times = [[st.frame_time(k) for st in streams] for k in range(n_frames)]
    sess_ids = list(range(N_STREAMS))
    P = W_IMG * H_IMG
    dev_ptrs = [[frames_dev.data_ptr() + (s * n_frames + k) * P for s in sess_ids] for k in range(n_frames)]
    host_ptrs = [[frames_host.data_ptr() + (s * n_frames + k) * P for s in sess_ids] for k in range(n_frames)]

    def lane_busy(ctx):
        import ctypes as _ct
        a = (_ct.c_double * 64)()
        n = capi.lib.xrvio_debug_lane_busy(ctx._h, a, 64)
        return [a[i] for i in range(max(n, 0))]

    def one_run(ptrs, on_device, read_back, profile, sample_clocks=False):
        ctx = capi.Context(drive.params_for(streams[0]), N_STREAMS, device=local_rank)
        for s, st in enumerate(streams):
            for k in range(PRIME_FRAMES):
                q, p, v = st.gt_state(st.frame_time(k))
                ctx.push_bootstrap_state(s, st.frame_time(k), q, p, v)

        def step(k):
            ctx.push_imu(*batches[k])
            ctx.step()
            ctx.push_images(sess_ids, ptrs[k], times[k], on_device=on_device)
            if read_back:
                return [ctx.get_pose(s) for s in sess_ids]

        for k in range(PRIME_FRAMES + Wm):
            step(k)
        assert all(ctx.get_state(s) == 1 for s in sess_ids), 'sessions did not reach TRACKING'
        st0 = ctx.stats()
        busy0 = lane_busy(ctx)
        ph0 = ctx.phase_times().copy()
        if profile:
            ctx.set_profiling(True)
            kt0 = ctx.kernel_times()
        # timed region in the library's asynchronous mode (include/xrvio.h): the calls queue work per lane of sessions and the
        # lanes run through it at their own pace, at most six steps behind this loop; xrvio_sync at the end waits for all of it
        ctx.set_async(not args.sync_steps)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank) if (rank == 0 and sample_clocks) else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        # every step processes the frame pushed by the step before it and pushes the next one: K steps = K frames per session
        # through the whole path and K image pushes inside the timed region
        for k in range(PRIME_FRAMES + Wm, PRIME_FRAMES + Wm + K):
            step(k)
        ctx.sync()
        torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        ctx.set_async(False)
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        st1 = ctx.stats()
        kt = None
        if profile:
            kt1 = ctx.kernel_times()
            kt = {k: (kt1[k][0] - kt0[k][0], kt1[k][1] - kt0[k][1]) for k in kt1}
        busy1 = lane_busy(ctx)
        ph1 = ctx.phase_times()
        res = dict(ms=max(ms, wall * 1e3), stats0=st0, stats1=st1, kt=kt, clocks=clocks, lane_busy_ms=[round(b - a, 1) for a, b in zip(busy0, busy1)],
                   phase_ms={int(i): round(float(ph1[i] - ph0[i]), 1) for i in range(31) if ph1[i] - ph0[i] > 0.05})
        ctx.close()
        return res

    # Three passes over the same K steps: (1) with per-launch CUDA events on the library's streams (xrvio_set_profiling) for the kernel
    # times of the roofline -- the events and their collection cost throughput, so the metric is not taken from this pass; (2) the same
    # device-resident pass without them, SM clocks sampled during it: `value`; (3) the end-to-end pass from host buffers: `e2e`.
    # The roofline entry prints the step time of pass (1) next to that of pass (2).
    if args.dev_only:           # profiler runs: one pass is enough, the line is not a bench value
        r_dev = one_run(dev_ptrs, True, False, False)
        if rank == 0:
            print(json.dumps({'dev_only': True, 'ms_per_step': r_dev['ms'] / K}))
        return
    r_prof = one_run(dev_ptrs, True, False, True)
    r_dev = one_run(dev_ptrs, True, False, False, sample_clocks=True)
    r_e2e = one_run(host_ptrs, False, True, False)

    ms_dev, ms_e2e = reduce_max_ms(r_dev['ms'], dev), reduce_max_ms(r_e2e['ms'], dev)
    if rank != 0:
        return
    s0, s1 = r_dev['stats0'], r_dev['stats1']
    e0s, e1s = r_e2e['stats0'], r_e2e['stats1']
    # frames that went through the whole path inside the timed region (identical on every rank: weak scaling)
    total_frames = int(s1['frames_processed'] - s0['frames_processed']) * world
    assert total_frames == K * N_STREAMS * world, (total_frames, K)
    value = total_frames / (ms_dev * 1e-3)
    e2e = int(e1s['frames_processed'] - e0s['frames_processed']) * world / (ms_e2e * 1e-3)
    launches = int(s1['kernels_launched'] - s0['kernels_launched'])
    # roofline of the dominant kernel family (device time from CUDA events on the library's stream)
    kt = r_prof['kt']
    dom = max(kt, key=lambda k: kt[k][0])
    peak, peak_src = read_peaks()
    frames_rank = int(s1['frames_processed'] - s0['frames_processed'])
    fam_ms = kt[dom][0]
    n_launch = max(kt[dom][1], 1)
    window_solves = int(s1['window_solves'] - s0['window_solves'])
    assert int(r_prof['stats1']['window_solves'] - r_prof['stats0']['window_solves']) == window_solves      # same work in both passes
    if dom in ALGO_BYTES_PER_FRAME:
        algo_bytes = ALGO_BYTES_PER_FRAME[dom] * frames_rank
    elif dom == 'solve':
        # per frame: localize ~15 KB; per window BA ~193 KB per Jacobian evaluation x 2 + 28 cost-only passes of ~60 KB
        algo_bytes = 15e3 * frames_rank + window_solves * (2 * 193e3 + 28 * 60e3)
    elif dom == 'marginalize':
        algo_bytes = int(s1['marginalizations'] - s0['marginalizations']) * 326e3
    else:
        algo_bytes = 0.0
    achieved = algo_bytes / (fam_ms * 1e-3) / 1e9 if fam_ms > 0 else 0.0
    # DRAM traffic per launch and pipe utilisations of the dominant kernel: from the committed ncu --set full capture of the same
    # command (profiles/r2_kernels.json, written by tools/profile_final.sh + tools/summarize_ncu.py), never a literal
    prof = committed_profile()
    pk = {'solve': 'solve_kernel[window]', 'marginalize': 'marg_kernel', 'gftt': 'corner_select_kernel', 'klt': 'klt_kernel', 'poisson': 'poisson_kernel',
          'preintegrate': 'preint_kernel', 'clahe': 'clahe_apply_kernel', 'pyramid': 'pyrdown_kernel', 'triangulate': 'tri_kernel'}.get(dom)
    pe = prof.get(pk, {}) if prof else {}
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': pe.get('dram_bytes'), 'traffic_source': ('profiles/r2_kernels.json:' + pk) if pe else None,
                'traffic_capture': prof.get('_capture') if prof else None,      # which library state the committed ncu capture belongs to
                'launches': n_launch, 'avg_launch_ms': fam_ms / n_launch,
                'pass': 'per-launch events in a separate pass over the same K steps: %.3f ms/step with them, %.3f in the metric pass' % (r_prof['ms'] / K, ms_dev / K),
                'share_of_kernel_time': fam_ms / max(sum(v[0] for v in kt.values()), 1e-9),
                # the path moves 2.57 MB per frame: it is latency / FP64-issue bound, not bandwidth bound -- the HBM fraction is reported
                # because the contract asks for it, the pipe utilisations of the same capture say where the kernel really stands
                'fp64_pipe_pct': pe.get('fp64_pipe_pct'), 'dmma_pipe_pct': pe.get('dmma_pipe_pct'), 'issue_active_pct': pe.get('issue_active_pct'),
                'family_ms_note': 'CUDA-event time per kernel family summed over the lanes\' streams, which overlap: the sum exceeds the step time',
                'family_ms': {k: round(v[0], 3) for k, v in kt.items()}}
    out = {
        'metric': 'VIO frames/sec (batched streams)', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': K,
        'warmup': Wm, 'ms_per_step': ms_dev / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64 back end / u8+f32 front end', 'data': 'synthetic',
        'config': {'workload': '64x640x480_streams', 'streams_per_gpu': N_STREAMS, 'image': [W_IMG, H_IMG], 'cam_hz': 20, 'imu_hz': 200,
                   'window_keyframes': 10, 'max_detections': 200, 'prime_frames': PRIME_FRAMES, 'l2': 'inputs_larger_than_L2',
                   'parallelism': 'sessions sharded per GPU, no data-path collective'},
        'stepping': 'synchronous' if args.sync_steps else 'asynchronous lanes (<= 6 steps ahead), poses polled every step',
        'e2e': {'value': e2e, 'unit': 'frames/s', 'h2d_bytes_per_step': int((e1s['h2d_bytes'] - e0s['h2d_bytes']) / K),
                'd2h_bytes_per_step': int((e1s['d2h_bytes'] - e0s['d2h_bytes']) / K) + N_STREAMS * 72, 'ms_per_step': ms_e2e / K},
        'gpu_launches': launches,
        'lane_busy_ms': r_dev['lane_busy_ms'],         # host wall time each lane spent inside its K steps (diagnostic)
        'phase_ms': r_dev['phase_ms'],                 # host wall time per step phase summed over lanes and steps (tools/phase_probe.py names them)
        'host_cpus': len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count(),
        'clocks': r_dev['clocks'],
        'roofline': roofline,
        'work': {'window_solves': window_solves, 'localize_solves': int(s1['localize_solves'] - s0['localize_solves']),
                 'subwindow_solves': int(s1['subwindow_solves'] - s0['subwindow_solves']),
                 'marginalizations': int(s1['marginalizations'] - s0['marginalizations'])},
    }
    if world == 1 and not args.no_cpu_baseline:
        out['cpu_baseline'] = cpu_baseline_sample(n_tracked=30)
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------
def _oracle_worker(seed, n_total, n_timed, q=None):
    """Runs one stream through the oracle pipeline; returns seconds spent on the last n_timed frames
    (image rendering excluded)."""
    import cv2
    cv2.setNumThreads(1)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    from oracle.synth import Stream
    from oracle.pipeline import Session, Config
    st = Stream(seed, W_IMG, H_IMG)
    frames = [st.frame(k) for k in range(n_total)]
    imu = st.imu(n_total * st.imu_per_cam + 2)
    sess = Session(Config(st.w, st.h, st.K), gt_fn=st.gt_state)
    k_imu = 0
    t_timed = 0.0
    for k in range(n_total):
        t0 = time.perf_counter()
        t_img = st.frame_time(k)
        while k_imu < len(imu) and imu[k_imu, 0] <= t_img + 1e-9:
            sess.track_imu(imu[k_imu, 0], imu[k_imu, 1:4], imu[k_imu, 4:7]); k_imu += 1
        sess.track_camera(t_img, frames[k])
        if k >= n_total - n_timed:
            t_timed += time.perf_counter() - t0
    if q is not None:
        q.put(t_timed)
    return t_timed


def cpu_baseline_sample(n_tracked=30):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    t = _oracle_worker(1000, PRIME_FRAMES + n_tracked, n_tracked)
    return {'value': n_tracked / t, 'unit': 'frames/s', 'cores': 1, 'kind': 'port',
            'sample': '%d tracked frames of one 640x480 stream after the %d-frame bootstrap, restated reference CPU path '
                      '(cv2 4.13 front end, numpy float64 back end), 1 thread' % (n_tracked, PRIME_FRAMES)}


def run_reference_arm(args, rank, world):
    """The restated reference CPU path on all host cores: one process per stream (the reference pipeline is
    single-threaded, XRSLAM_ENABLE_THREADING off and Ceres num_threads = 1)."""
    if rank != 0:
        return
    import multiprocessing as mp
    # one single-threaded stream per core, at most the job's stream count (64 per GPU of the launch)
    cores = min(os.cpu_count() or 1, N_STREAMS * max(world, 1))
    K, Wm = args.steps, max(args.warmup, 3)
    n_timed = max(1, min(K, 20))                     # bounded sample: at most 20 tracked frames per stream
    ctx = mp.get_context('fork')
    q = ctx.Queue()
    t0 = time.perf_counter()
    procs = [ctx.Process(target=_oracle_worker, args=(1000 + i, PRIME_FRAMES + Wm + n_timed, n_timed, q)) for i in range(cores)]
    for p in procs:
        p.start()
    secs = [q.get() for _ in procs]
    for p in procs:
        p.join()
    value = cores * n_timed / max(secs)
    line = {
        'impl': 'reference', 'metric': 'VIO frames/sec (batched streams)', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
        'steps': K, 'warmup': Wm, 'ms_per_step': max(secs) / n_timed * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64 back end / u8+f32 front end', 'data': 'synthetic',
        'config': {'workload': '64x640x480_streams', 'streams_per_gpu': N_STREAMS, 'image': [W_IMG, H_IMG], 'cam_hz': 20, 'imu_hz': 200,
                   'window_keyframes': 10, 'max_detections': 200, 'prime_frames': PRIME_FRAMES, 'l2': 'inputs_larger_than_L2',
                   'parallelism': 'sessions sharded per GPU, no data-path collective'},
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d streams (one per host core) x %d tracked frames each, restated reference CPU path '
                                   '(cv2 4.13 + numpy), not the original Ceres/Eigen binary' % (cores, n_timed)},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t0,
    }
    print(json.dumps(line))


def run_ba_stress(args):
    """Secondary line (not the driver's): BASELINE config 5, a batch of 30-keyframe x 2000-landmark window solves, 10 trust-region
    iterations each, through xrvio_op_solve (host buffers in, host buffers out = e2e); `value` uses the device time of the solve
    kernel alone (xrvio_debug_last_solve_ms, CUDA events on the launch stream).  Flop count per iteration: SURVEY.md 8(d),
    ~0.3 GFLOP (Schur SYRK 130 M, pose-block J^T J 33 M, prior 110 M, Cholesky 30 M); FP64 peak is not in MEASURED_PEAKS.json:
    it is measured in the same run (xrvio_debug_fp64_peak), the datasheet's 40 TFLOP/s only if the probe fails."""
    import numpy as np
    from xrslam_b200 import capi
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import stress, bridge                     # problem generator + packer only (test infrastructure, not timed)
    pbs = [stress.make_stress_problem(100 + i, 30, 2000) for i in range(args.problems)]
    pk = bridge.PackedProblems(pbs)
    rig = bridge.rig_from_problem(pbs[0])
    st0, inv0 = pk.states.copy(), pk.inv_depth.copy()
    kern_ms, wall_ms, iters = [], [], 0
    for k in range(args.warmup + args.steps):
        pk.states[:], pk.inv_depth[:] = st0, inv0
        t0 = time.perf_counter()
        out = pk.solve(rig)
        t1 = time.perf_counter()
        if k >= args.warmup:
            kern_ms.append(capi.last_solve_ms())
            wall_ms.append((t1 - t0) * 1e3)
            iters = int(sum(int(o['iterations']) - 1 for o in out))
    km, wm = float(np.mean(kern_ms)), float(np.mean(wall_ms))
    # clock64 sections of the solve kernel (only a -DXRVIO_KPROF diagnostic build fills them; all zero in the product build)
    import ctypes as _ct
    _arr = (_ct.c_ulonglong * 32)()
    capi.lib.xrvio_debug_kprof(_arr, 1)
    _names = {0: 'J zero', 1: 'J rpe', 3: 'J pre simt', 4: 'J pre coop', 5: 'J marg', 9: 'C evaluations', 17: 'make_candidate', 18: 'D/alpha',
              20: 'dogleg', 21: 'ambient', 22: 'post-accept', 23: 'apply_scaling', 24: 'GN schur', 25: 'GN cholesky', 26: 'GN substitution', 29: 'GN setup'}
    _np = max(int(_arr[31]), 1)
    kprof = {_names.get(i, str(i)): round(_arr[i] / _np / 1e3, 1) for i in range(31) if _arr[i] and i not in (27, 28)}
    try:
        dfma_peak, dmma_peak = capi.fp64_peak()           # measured on this GPU: DFMA loop / mma.sync.m8n8k4.f64 loop, TFLOP/s
        peak, peak_source = max(dfma_peak, dmma_peak), 'measured (xrvio_debug_fp64_peak: DFMA %.1f, DMMA %.1f TFLOP/s)' % (dfma_peak, dmma_peak)
    except capi.XrvioError:
        peak, peak_source = 40.0, 'datasheet FP64 (probe unavailable)'
    flops = 0.3e9 * iters
    h2d = int(pk.states.nbytes + pk.inv_depth.nbytes + pk.rpe.nbytes + pk.pre_blocks.nbytes + pk.marg.nbytes)
    print(json.dumps({
        'metric': 'BA iterations/sec (30 keyframes x 2000 landmarks)', 'value': iters / (km * 1e-3), 'unit': 'iterations/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': km, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': 'ba_stress_30kf_2000lm', 'problems': args.problems, 'iterations': 10,
                                                          'residuals_per_problem': 58000, 'l2': 'inputs_larger_than_L2' if h2d > 126e6 else 'not flushed'},
        'e2e': {'value': iters / (wm * 1e-3), 'unit': 'iterations/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': int(pk.states.nbytes + pk.inv_depth.nbytes),
                'ms_per_step': wm},
        'gpu_launches': args.steps * (1 + args.problems),
        'roofline': {'bound': 'tensor', 'kernel': 'solve_large', 'achieved': flops / (km * 1e-3) / 1e12, 'peak': peak, 'peak_source': peak_source,
                     'unit': 'TFLOP/s', 'frac': flops / (km * 1e-3) / 1e12 / peak, 'traffic': None},
        'final_cost_mean': float(np.mean([o['final_cost'] for o in out])),
        'kprof_kcycles_per_problem': kprof or None}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='streams', choices=['streams', 'ba_stress'],
                    help='streams: BASELINE config 3 (the metric line); ba_stress: config 5, batch of large-window solves')
    ap.add_argument('--problems', type=int, default=64)
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--dev-only', action='store_true', help='profiler runs: only the device-resident pass, no JSON metric line')
    ap.add_argument('--sync-steps', action='store_true', help='every xrvio_step waits for all lanes (A/B against the asynchronous mode)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.workload == 'ba_stress':
        if rank == 0:
            run_ba_stress(args)
        return
    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    if world > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'        # keep stdout to the single JSON line
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    run_cuda_arm(args, rank, world, local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
