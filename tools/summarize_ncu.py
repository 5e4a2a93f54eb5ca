#!/usr/bin/env python
"""Turns the ncu raw pages brought back by tools/profile_final.sh into the committed summaries under profiles/:
   python tools/summarize_ncu.py r2
reads gpurun_out/<tag>_full_raw.csv.gz, <tag>_solve_src_raw.csv.gz, <tag>_large_raw.csv.gz, <tag>_launches.csv.gz and writes
profiles/<tag>_kernels.json (what bench.py reads for roofline.traffic and the pipe utilisations), profiles/<tag>_kernels.md
(the same as a table) and profiles/<tag>_launch_summary.txt (shares of the step from the launch list)."""
import collections
import csv
import gzip
import io
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r2'
OUT = os.path.join(ROOT, 'gpurun_out')
PROF = os.path.join(ROOT, 'profiles')

WANT = {
    'gpu__time_duration.sum': 'duration_us',
    'dram__bytes_read.sum': 'dram_read_bytes',
    'dram__bytes_write.sum': 'dram_write_bytes',
    'lts__t_sectors.sum': 'l2_sectors',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
    'l1tex__t_sector_hit_rate.pct': 'l1_hit_pct',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_throughput_pct',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_throughput_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct',
    'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active': 'dmma_pipe_pct',
    'launch__registers_per_thread': 'registers',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'launch__shared_mem_per_block_dynamic': 'dyn_smem',
    'smsp__sass_inst_executed_op_local_ld.sum': 'local_ld_inst',
    'smsp__sass_inst_executed_op_local_st.sum': 'local_st_inst',
}
UNIT_SCALE = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'nsecond': 1e-3, 'ms': 1e3, 'msecond': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def short(name):
    m = re.search(r'(\w+_kernel|\w+)\s*[<(]', name)
    n = m.group(1) if m else name
    return n.split('::')[-1]


def read_raw(path):
    if not os.path.exists(path):
        return []
    with gzip.open(path, 'rt') as f:
        rows = list(csv.reader(f))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = {'kernel': short(r[idx['Kernel Name']])}
        for m, nm in WANT.items():
            if m in idx and r[idx[m]] not in ('', 'n/a'):
                try:
                    v = float(r[idx[m]].replace(',', ''))
                except ValueError:
                    continue
                k[nm] = v * UNIT_SCALE.get(units[idx[m]], 1.0)
        st = {}
        for i, h in stall_cols:
            try:
                st[h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = float(r[i])
            except ValueError:
                pass
        tot = sum(st.values()) or 1.0
        k['stalls'] = {a: round(b / tot, 3) for a, b in sorted(st.items(), key=lambda x: -x[1])[:4]}
        out.append(k)
    return out


def aggregate(launches):
    by = collections.defaultdict(list)
    for k in launches:
        key = k['kernel']
        if key == 'solve_kernel':      # localize / sub-window batches vs window-BA batches are different animals
            key = 'solve_kernel[window]' if k.get('duration_us', 0) > 800 else 'solve_kernel[localize,subwindow]'
        by[key].append(k)
    res = {}
    for key, ls in by.items():
        a = {'launches_captured': len(ls)}
        for nm in set(WANT.values()):
            vs = [l[nm] for l in ls if nm in l]
            if vs:
                a[nm] = sum(vs) / len(vs)
        a['dram_bytes'] = a.get('dram_read_bytes', 0) + a.get('dram_write_bytes', 0)
        st = collections.Counter()
        for l in ls:
            for s_, v in l.get('stalls', {}).items():
                st[s_] += v / len(ls)
        a['stalls'] = {s_: round(v, 3) for s_, v in st.most_common(4)}
        res[key] = a
    return res


def main():
    os.makedirs(PROF, exist_ok=True)
    launches = read_raw(os.path.join(OUT, TAG + '_full_raw.csv.gz'))
    res = aggregate(launches)
    e = read_raw(os.path.join(OUT, TAG + '_large_raw.csv.gz'))
    if e:
        res['solve_large_kernel'] = list(aggregate(e).values())[0]
    # --set full + sources of a few consecutive launches of solve_kernel: the longest one is the window-BA batch
    e = read_raw(os.path.join(OUT, TAG + '_solve_src_raw.csv.gz'))
    if e:
        longest = max(e, key=lambda k: k.get('duration_us', 0))
        a = aggregate([dict(longest, kernel='x')])
        res['solve_kernel[longest launch of the --set full capture]'] = list(a.values())[0]
        if 'solve_kernel[window]' not in res and longest.get('duration_us', 0) > 800:
            res['solve_kernel[window]'] = list(a.values())[0]
    with open(os.path.join(PROF, TAG + '_kernels.json'), 'w') as f:
        json.dump(res, f, indent=1, sort_keys=True)
    cols = ['duration_us', 'grid', 'block', 'registers', 'dram_bytes', 'l2_hit_pct', 'sm_throughput_pct', 'issue_active_pct', 'warps_active_pct', 'fp64_pipe_pct', 'dmma_pipe_pct', 'local_ld_inst']
    with open(os.path.join(PROF, TAG + '_kernels.md'), 'w') as f:
        f.write('ncu, averages over the captured launches of each kernel (tools/profile_ncu.sh: the metrics below over 400 consecutive launches; '
                '--set full for solve_kernel and solve_large_kernel; tools/summarize_ncu.py)\n\n')
        f.write('| kernel | n | ' + ' | '.join(cols) + ' | top stalls |\n|---|---|' + '---|' * (len(cols) + 1) + '\n')
        for key in sorted(res):
            a = res[key]
            f.write('| %s | %d | ' % (key, a['launches_captured']) + ' | '.join(('%.4g' % a[c]) if c in a else '-' for c in cols) +
                    ' | ' + ', '.join('%s %.0f%%' % (s_, 100 * v) for s_, v in a['stalls'].items()) + ' |\n')
    # launch list -> shares
    lp = os.path.join(OUT, TAG + '_launches.csv.gz')
    if os.path.exists(lp):
        with gzip.open(lp, 'rt') as f:
            rows = list(csv.reader(l for l in f if l.startswith('"')))
        hdr = rows[0]
        i_name, i_val, i_unit = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
        agg = collections.defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            v = float(r[i_val].replace(',', '')) * UNIT_SCALE.get(r[i_unit], 1.0)
            a = agg[short(r[i_name])]
            a[0] += 1; a[1] += v
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(PROF, TAG + '_launch_summary.txt'), 'w') as f:
            f.write('ncu --metrics gpu__time_duration.sum over %d launches of the repo\'s kernels in the tracking region of bench.py (serialised, cold cache: shares, not absolutes)\n' % (len(rows) - 1))
            for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write('%-28s n=%5d  total %10.1f us  avg %8.1f us  share %5.1f %%\n' % (n, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    print(json.dumps({k: {c: round(v[c], 2) for c in ('duration_us', 'dram_bytes', 'fp64_pipe_pct', 'dmma_pipe_pct') if c in v} for k, v in res.items()}, indent=1))


if __name__ == '__main__':
    main()
