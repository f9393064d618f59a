"""Per-kernel DRAM traffic of the LAST C2 NLL+grad step from an ncu csv holding gpu__time_duration.sum,
dram__bytes_read.sum and dram__bytes_write.sum for every launch of scripts/prof_paths.py nll.
usage: dram_summary.py launches.csv out.json"""
import collections, csv, json, re, sys


def main(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    launches = collections.OrderedDict()
    for r in csv.DictReader(lines):
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('void sb::<unnamed>::', '').replace('sb::<unnamed>::', '').replace('void ', '')
        name = re.sub(r'<.*', '', name.replace('unnamed>::', ''))
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        m = r['Metric Name']
        if m.startswith('gpu__time'):
            v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)          # -> us
        else:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        launches.setdefault(r['ID'], {'name': name})[m] = v
    seq = list(launches.values())
    # the last step starts at the last gp_assemble_kernel launch
    start = max(i for i, l in enumerate(seq) if l['name'].endswith('gp_assemble_kernel'))
    agg = collections.OrderedDict()
    for l in seq[start:]:
        if l['name'].startswith('at::'):
            continue
        a = agg.setdefault(l['name'], {'launches': 0, 'us': 0.0, 'dram_read_bytes': 0.0, 'dram_write_bytes': 0.0})
        a['launches'] += 1
        a['us'] += l.get('gpu__time_duration.sum', 0.0)
        a['dram_read_bytes'] += l.get('dram__bytes_read.sum', 0.0)
        a['dram_write_bytes'] += l.get('dram__bytes_write.sum', 0.0)
    doc = {'step': 'C2 NLL+grad, k=20, n=2000 (scripts/prof_paths.py nll, last step)', 'kernels': agg,
           'how': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none'}
    json.dump(doc, open(out, 'w'), indent=1)
    for k, v in agg.items():
        print('%-28s n=%3d %9.1f us  read %8.1f MB  write %8.1f MB' % (k[-28:], v['launches'], v['us'], v['dram_read_bytes'] / 1e6, v['dram_write_bytes'] / 1e6))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
