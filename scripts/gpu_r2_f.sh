#!/bin/bash
# which switch moves the event-bracketed GEMM time of the roofline leg?
for cfg in "1 1" "0 1" "1 0" "1 1"; do
  set -- $cfg
  echo "== SB200_TRMV_SIDE=$1 SB200_DAG_YIELD=$2"
  SB200_TRMV_SIDE=$1 SB200_DAG_YIELD=$2 timeout 200 python bench.py --skip-secondary 2>gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        r = d['roofline']
        print('ms_per_step %.4f  gemm_ms %.4f frac %.4f  potrf_ms %.4f  whole %.3f' % (d['ms_per_step'], r['gemm_ms_per_step'], r['frac'], d['roofline_kernels'][1]['ms_per_step'], r['whole_step_frac']))
"
done
