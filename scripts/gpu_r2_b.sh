#!/bin/bash
# Round 2, second GPU call (one B200): GPU tests incl. the device-resident sampler, b200 arm.
set -x
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.txt
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
