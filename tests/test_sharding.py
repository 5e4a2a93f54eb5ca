"""N > 1 harness logic on CPU with the gloo backend (world_size 2): disjoint session seeds per rank and the
max-over-ranks timing reduction that bench.py uses (the only collective: the VIO path itself has none)."""
import os
import sys
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys, json
    sys.path.insert(0, %r)
    import torch.distributed as dist
    import bench
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    seeds = bench.stream_seeds(rank, 4)
    ms = bench.reduce_max_ms(10.0 + 5.0 * rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, seeds)
    if rank == 0:
        print(json.dumps({'ms': ms, 'seeds': gathered}))
    dist.barrier()
    dist.destroy_process_group()
''') % ROOT


def test_two_rank_gloo_harness(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29577')
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                          '--master-port', '29577', str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith('{')][-1]
    d = json.loads(line)
    assert d['ms'] == 15.0                                  # max over ranks
    assert d['seeds'] == [[0, 1, 2, 3], [1000, 1001, 1002, 1003]]
    assert not set(d['seeds'][0]) & set(d['seeds'][1])      # disjoint shards


def test_reference_arm_runs_on_rank0_only():
    """bench.py --impl reference: ranks other than 0 exit without work."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ''
