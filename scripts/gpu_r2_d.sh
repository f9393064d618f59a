#!/bin/bash
# unguarded per-dimension loops (gradient contraction, assembly, cross-covariance, predict epilogue): parity + A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py tests/test_gpu_arbiter.py tests/test_gpu_sampler.py -q -m gpu -x 2>&1 | tail -4
for occ in 3 2 3 2; do
  echo "== SB200_CONTRACT_OCC=$occ"
  SB200_CONTRACT_OCC=$occ timeout 200 python bench.py --skip-secondary --steps 20 --warmup 5 2>gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('ms_per_step %.4f  e2e %.4f  sustained %.1f evals/s  gemm frac %.3f  whole %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['sustained']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac']))
"
done
timeout 300 python scripts/prof_c5_step.py 2>&1 | tail -15
