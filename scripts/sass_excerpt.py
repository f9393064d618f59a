"""SASS evidence of the Blackwell-native tile movement: for every kernel of libsurmise_b200.so that contains DMMA, count
DMMA / UTMALDG / SYNCS (mbarrier) / LDGSTS (cp.async) / BAR instructions, and print the main-loop excerpt of the
TMA-staged GEMM and of the task-graph Cholesky.  python scripts/sass_excerpt.py > profiles/sass_dgemm_r2.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'surmise_b200', 'lib', 'libsurmise_b200.so')
txt = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur:
        funcs[cur].append(line)


def demangle(n):
    try:
        return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip().replace('sb::(anonymous namespace)::', '')
    except OSError:
        return n


def op(line):
    m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)', line)
    return m.group(1) if m else None


print('# cuobjdump -sass surmise_b200/lib/libsurmise_b200.so (sm_100a), kernels that contain DMMA\n')
print('%-110s %6s %8s %6s %7s %5s' % ('kernel', 'DMMA', 'UTMALDG', 'SYNCS', 'LDGSTS', 'BAR'))
for name, lines in sorted(funcs.items(), key=lambda kv: demangle(kv[0])):
    ops = [op(l) for l in lines]
    c = lambda p: sum(1 for o in ops if o and o.startswith(p))
    if c('DMMA'):
        print('%-110s %6d %8d %6d %7d %5d' % (demangle(name)[:110], c('DMMA'), c('UTMALDG'), c('SYNCS'), c('LDGSTS'), c('BAR')))

for pat, title in (('dgemm_tma_kernelILi64ELb1ELb0', 'dgemm_tma_kernel<64, k-contiguous A, row-contiguous B, 3 stages>'),
                   ('potrf_dag_kernelILb1', 'potrf_dag_kernel<TMA>')):
    name = next(n for n in funcs if pat in n)
    lines = funcs[name]
    idx = [i for i, l in enumerate(lines) if op(l) and op(l).startswith('UTMALDG')]
    print('\n\n## %s -- producer: wait for the free slot, post the byte count, issue the bulk-tensor loads\n' % title)
    first_block_end = next((idx[i] for i in range(len(idx) - 1) if idx[i + 1] - idx[i] > 40), idx[-1])
    lo, hi = max(0, idx[0] - 30), min(len(lines), first_block_end + 6)
    for l in lines[lo:hi]:
        if op(l):
            print(re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l.rstrip()))
    # consumer: the first TRYWAIT after the producer block up to the ARRIVE that releases the slot
    print('\n## %s -- consumer: wait for the bytes (SYNCS.PHASECHK.TRYWAIT), LDS.64 fragments, DMMA.8x8x4, release (SYNCS.ARRIVE)\n' % title)
    start = next(i for i in range(idx[-1], len(lines)) if op(lines[i]) and 'TRYWAIT' in op(lines[i]))
    end = next(i for i in range(start, len(lines)) if op(lines[i]) and op(lines[i]).startswith('SYNCS.ARRIVE'))
    shown = 0
    for l in lines[start:end + 2]:
        o = op(l)
        if not o:
            continue
        shown += 1
        if shown <= 70 or o.startswith('SYNCS'):
            print(re.sub(r'\s*/\* 0x[0-9a-f]+ \*/\s*$', '', l.rstrip()))
        elif shown == 71:
            print('        ... (%d more instructions of the same LDS.64 / DMMA.8x8x4 pattern)' %
                  sum(1 for x in lines[start:end] if op(x)))
