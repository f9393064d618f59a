#!/bin/bash
# A/B of the fused leaf-pair panel (SB200_PANEL_FUSE) on the C2 NLL+grad step.  Run under gpurun.
set -x
python -m pytest tests/test_gpu_blocks.py -m gpu -x -q 2>&1 | tail -5
SB200_PANEL_FUSE=0 timeout 120 python bench.py --steps 10 --warmup 3 --skip-secondary > gpurun_out/bench_nofuse.json 2>gpurun_out/bench_nofuse.err
timeout 120 python bench.py --steps 10 --warmup 3 --skip-secondary > gpurun_out/bench_fuse.json 2>gpurun_out/bench_fuse.err
SB200_PANEL_FUSE=0 python scripts/panel_timing.py
python scripts/panel_timing.py
python - <<'P'
import json
for f in ["gpurun_out/bench_nofuse.json", "gpurun_out/bench_fuse.json"]:
    d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    print(f, d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["roofline"]["gemm_ms_per_step"], d["cpu_baseline"]["parity_vs_gpu"])
P
