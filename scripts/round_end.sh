#!/bin/bash
# End-of-round evidence on one B200 (run under gpurun): GPU tests, bench, ncu launch list, DRAM traffic, one --set full capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; tail -c 300 gpurun_out/bench_reference.json
timeout 150 python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/bench_short.json 2>> gpurun_out/bench.err &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/ncu.log 2>&1
python scripts/prof_paths.py nll > /dev/null 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/dram.csv python scripts/prof_paths.py nll > gpurun_out/ncu_dram.log 2>&1
python scripts/prof_gemm.py &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dgemm_dmma -s 3 -c 3 -o gpurun_out/prof_gemm_64 -f python scripts/prof_gemm.py > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_gemm_64.ncu-rep --page raw --csv > gpurun_out/prof_gemm_64_raw.csv 2>/dev/null
ls -la gpurun_out
