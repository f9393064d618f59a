import csv, re, sys, collections
def load(path):
    lines=[l for l in open(path) if not l.startswith('==')]
    rows=[]
    for r in csv.DictReader(lines):
        name=re.sub(r'\(.*','',r['Kernel Name']).replace('void sb::<unnamed>::','').replace('sb::<unnamed>::','').replace('void ','')
        v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
        v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
        rows.append((name, r['Grid Size'], v))
    return rows
if __name__=='__main__':
    rows=load(sys.argv[1])
    lo=int(sys.argv[2]) if len(sys.argv)>2 else 0
    hi=int(sys.argv[3]) if len(sys.argv)>3 else len(rows)
    rows=rows[lo:hi]
    # bench.py's FP64-peak probe (cuBLAS DGEMM 8192^3, 'cutlass_80_tensorop_d884gemm') can fall inside the capture
    # window: it is not part of a step, so it is listed but kept out of the shares
    probe=[r for r in rows if 'cutlass' in r[0]]
    rows=[r for r in rows if 'cutlass' not in r[0]]
    if probe: print('left out: %d cuBLAS peak-probe launches, %.1f us'%(len(probe),sum(r[2] for r in probe)))
    agg=collections.defaultdict(lambda:[0,0.0])
    for n,g,v in rows: agg[n][0]+=1; agg[n][1]+=v
    tot=sum(v[1] for v in agg.values())
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1]):
        print('%-70s n=%4d %10.1f us %5.1f%%'%(k[:70],v[0],v[1],100*v[1]/tot))
    print('total %.1f us over %d launches'%(tot,len(rows)))
