#!/bin/bash
# last GPU call of round 2: the generic-emulator likelihood (new kernel), the Woodbury tests, the reference's own test files
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_path.py -q -m gpu -k "woodbury or loglik or closure" 2>&1 | tail -6
timeout 150 python scripts/run_reference_suite.py > gpurun_out/ref_suite.txt 2>&1; tail -4 gpurun_out/ref_suite.txt
