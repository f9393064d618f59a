#!/bin/bash
# A/B of the two step-level changes (no clearing of the inverse-factor buffer; alpha = Linv' z beside the lauum product)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_path.py tests/test_gpu_fullsize.py -q -m gpu -k "nll or c2" 2>&1 | tail -4
for cfg in "1 0" "0 0" "0 1" "1 0" "0 1"; do
  set -- $cfg
  echo "== SB200_LINV_MEMSET=$1 SB200_TRMV_SIDE=$2"
  SB200_LINV_MEMSET=$1 SB200_TRMV_SIDE=$2 timeout 200 python bench.py --skip-secondary --steps 20 --warmup 5 2>gpurun_out/ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('ms_per_step %.4f  e2e %.4f  sustained %.1f evals/s  gemm frac %.3f  whole %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['sustained']['value'], d['roofline']['frac'], d['roofline']['whole_step_frac']))
"
done
