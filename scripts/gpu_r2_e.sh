#!/bin/bash
# A/B: update tasks of the task-graph Cholesky yield the FP64 pipe to a 64x64 factorisation running on the same SM
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_blocks.py -q -m gpu -k "potrf or factorisation or trtri" 2>&1 | tail -3
for y in 0 1 0 1; do
  echo "== SB200_DAG_YIELD=$y"
  SB200_DAG_YIELD=$y timeout 100 python scripts/dag_stats.py 2000 20 2>&1 | tail -2
  SB200_DAG_YIELD=$y timeout 100 python scripts/dag_stats.py 8000 8 2>&1 | tail -1
  SB200_DAG_YIELD=$y timeout 200 python bench.py --skip-secondary --steps 20 --warmup 5 2>gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('ms_per_step %.4f  e2e %.4f  sustained %.1f evals/s  whole %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['sustained']['value'], d['roofline']['whole_step_frac']))
"
done
