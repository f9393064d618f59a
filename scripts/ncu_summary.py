"""Summarise `ncu --page raw --csv` exports of --set full captures: one row per launch with the metrics the roofline
needs.  python scripts/ncu_summary.py raw.csv [raw2.csv ...] [--json out.json] [--md out.md]"""
import csv, json, re, sys

WANT = [
    ('time_us', 'gpu__time_duration.sum', 1e-3),
    ('dram_read_MB', 'dram__bytes_read.sum', None),
    ('dram_write_MB', 'dram__bytes_write.sum', None),
    ('dram_pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1),
    ('l2_hit_pct', 'lts__t_sector_hit_rate.pct', 1),
    ('sm_busy_pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 1),
    ('fp64_pipe_pct', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 1),
    ('dmma_pipe_pct', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 1),
    ('warps_active_pct', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1),
    ('regs', 'launch__registers_per_thread', 1),
    ('smem_dyn_KB', 'launch__shared_mem_per_block_dynamic', None),
    ('occ_limit_regs', 'launch__occupancy_limit_registers', 1),
    ('occ_limit_smem', 'launch__occupancy_limit_shared_mem', 1),
    ('dfma', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 1),
    ('dmul', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 1),
    ('dadd', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 1),
]
STALL = re.compile(r'smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio')


def unit_scale(unit, want_mb):
    u = unit.lower()
    if want_mb:
        return {'byte': 1e-6, 'kbyte': 1e-3, 'mbyte': 1.0, 'gbyte': 1e3}.get(u, 1e-6)
    return 1.0


def load(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith('==')))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    out = []
    for r in body:
        name = re.sub(r'\(.*', '', r[col['Kernel Name']]).replace('void ', '').replace('sb::<unnamed>::', '').replace('unnamed>::', '')
        e = {'kernel': name, 'grid': r[col['Grid Size']], 'block': r[col['Block Size']]}
        for key, metric, scale in WANT:
            c = next((i for h, i in col.items() if h == metric or h.endswith('.' + metric)), None)
            if c is None or r[c] in ('', 'n/a'):
                continue
            v = float(r[c].replace(',', ''))
            u = units[c]
            if key == 'time_us':
                v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1e-3)
            elif scale is None:
                if 'KB' in key:
                    v *= {'byte': 1 / 1024, 'kbyte': 1.0, 'mbyte': 1024.0}.get(u.lower(), 1 / 1024)
                else:
                    v *= unit_scale(u, True)
            e[key] = v
        stalls = {}
        for h, i in col.items():
            m = STALL.search(h)
            if m and r[i] not in ('', 'n/a'):
                stalls[m.group(1) or m.group(2)] = max(stalls.get(m.group(1) or m.group(2), 0.0), float(r[i].replace(',', '')))
        e['top_stalls'] = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
        if 'dfma' in e:
            fl = 2 * e.get('dfma', 0) + e.get('dmul', 0) + e.get('dadd', 0)
            e['fp64_alu_gflop'] = fl / 1e9
            if e.get('time_us'):
                e['fp64_alu_tflops'] = fl / (e['time_us'] * 1e-6) / 1e12
        if e.get('time_us') and 'dram_read_MB' in e:
            e['dram_GBps'] = (e['dram_read_MB'] + e.get('dram_write_MB', 0.0)) * 1e-3 / (e['time_us'] * 1e-6)
        out.append(e)
    return out


if __name__ == '__main__':
    args = sys.argv[1:]
    jpath = mpath = None
    if '--json' in args:
        jpath = args[args.index('--json') + 1]; args = [a for a in args if a not in ('--json', jpath)]
    if '--md' in args:
        mpath = args[args.index('--md') + 1]; args = [a for a in args if a not in ('--md', mpath)]
    rows = [e for p in args for e in load(p)]
    keys = ['kernel', 'grid', 'block', 'time_us', 'dram_read_MB', 'dram_write_MB', 'dram_GBps', 'dram_pct', 'l2_hit_pct',
            'fp64_pipe_pct', 'dmma_pipe_pct', 'fp64_alu_tflops', 'warps_active_pct', 'regs', 'smem_dyn_KB', 'top_stalls']
    lines = ['| ' + ' | '.join(keys) + ' |', '|' + '---|' * len(keys)]
    for e in rows:
        cells = []
        for k in keys:
            v = e.get(k, '')
            if k == 'top_stalls':
                v = ', '.join('%s %.2f' % kv for kv in v)
            elif isinstance(v, float):
                v = '%.3g' % v if abs(v) < 1000 else '%.0f' % v
            cells.append(str(v))
        lines.append('| ' + ' | '.join(cells) + ' |')
    text = '\n'.join(lines)
    print(text)
    if mpath:
        open(mpath, 'w').write(text + '\n')
    if jpath:
        json.dump(rows, open(jpath, 'w'), indent=1)
