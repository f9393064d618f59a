#!/bin/bash
# Round 2 ncu evidence (one B200): launch list of the bench step, then --set full captures of the non-GEMM kernels.
# Every ncu command runs only after the same command line exited 0 without ncu.
set -x
mkdir -p gpurun_out
timeout 200 python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/ncu_launches.log 2>&1
python scripts/prof_paths.py nll > gpurun_out/plain_nll.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'potrf_dag|gp_grad_contract|gp_assemble|trmv_lower|nll_finalize|gp_grad_finalize' -s 7 -c 7 -o gpurun_out/prof_nll_r2 -f python scripts/prof_paths.py nll > gpurun_out/ncu_nll.log 2>&1
ncu -i gpurun_out/prof_nll_r2.ncu-rep --page raw --csv > gpurun_out/prof_nll_r2_raw.csv 2>/dev/null
python scripts/prof_paths.py loglik > gpurun_out/plain_loglik.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'predict|woodbury|matern|cross|syevj|col_partial|standardize_kernel|nll_small|sweep' -c 24 -o gpurun_out/prof_loglik_r2 -f python scripts/prof_paths.py loglik > gpurun_out/ncu_loglik.log 2>&1
ncu -i gpurun_out/prof_loglik_r2.ncu-rep --page raw --csv > gpurun_out/prof_loglik_r2_raw.csv 2>/dev/null
ls -la gpurun_out | tail -15
