#!/bin/bash
# Round 2 ncu evidence (one B200): launch list of the bench step, then --set full captures of the kernels of the
# NLL+grad step and of one PT-LMC step.  Every ncu command runs only after the same command line exited 0 without ncu.
set -x
mkdir -p gpurun_out
timeout 200 python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --skip-secondary > gpurun_out/ncu_launches.log 2>&1
python scripts/prof_paths.py nll > gpurun_out/plain_nll.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'potrf_dag|dgemm_tma|dgemm_dmma|gp_grad_contract|gp_assemble|trmv_lower' -s 17 -c 17 -o gpurun_out/prof_nll_r2 -f python scripts/prof_paths.py nll > gpurun_out/ncu_nll.log 2>&1
ncu -i gpurun_out/prof_nll_r2.ncu-rep --page raw --csv > gpurun_out/prof_nll_r2_raw.csv 2>/dev/null
python scripts/prof_c5_chain.py 40 > gpurun_out/plain_c5.log 2>&1 &&
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'cross_cov|predict_partial|predict_reduce|woodbury_loglik|ptlmc_step' -s 200 -c 10 -o gpurun_out/prof_c5_r2 -f python scripts/prof_c5_chain.py 40 > gpurun_out/ncu_c5.log 2>&1
ncu -i gpurun_out/prof_c5_r2.ncu-rep --page raw --csv > gpurun_out/prof_c5_r2_raw.csv 2>/dev/null
python scripts/pca_profile.py C4 > gpurun_out/plain_pca.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'syevj|col_partial|standardize_kernel' -c 6 -o gpurun_out/prof_pca_r2 -f python scripts/pca_profile.py C4 > gpurun_out/ncu_pca.log 2>&1
ncu -i gpurun_out/prof_pca_r2.ncu-rep --page raw --csv > gpurun_out/prof_pca_r2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_c5_r2.ncu-rep gpurun_out/prof_pca_r2.ncu-rep
ls -la gpurun_out | tail -12
