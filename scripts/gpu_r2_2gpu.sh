#!/bin/bash
# Round 2: N-GPU call (gpurun --gpus N): sharded-path checks and the strong-scaling bench line.
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check_${N}gpu.log 2>&1; tail -12 gpurun_out/dist_check_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -c 5000 gpurun_out/bench_${N}gpu.json; tail -8 gpurun_out/bench_${N}gpu.err
