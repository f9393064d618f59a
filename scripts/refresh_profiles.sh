#!/bin/bash
# After scripts/gpu_r2_ncu.sh and scripts/round_end.sh have run under gpurun: copy / summarise their outputs from
# gpurun_out/ (scratch) into profiles/ (tracked).  Runs here, no GPU.
set -e
cd "$(dirname "$0")/.."
G=gpurun_out
cp $G/launches_r2.csv profiles/launches_r2.csv
python scripts/summarize_launches.py profiles/launches_r2.csv > profiles/launch_summary_r2.txt
python scripts/ncu_summary.py $G/prof_nll_r2_raw.csv --md profiles/ncu_nll_step_r2.md > /dev/null
python scripts/ncu_summary.py $G/prof_c5_r2_raw.csv $G/prof_pca_r2_raw.csv --md profiles/ncu_c5_pca_r2.md > /dev/null
python scripts/make_ncu_kernels.py $G/prof_nll_r2_raw.csv $G/prof_c5_r2_raw.csv $G/prof_pca_r2_raw.csv $G/prof_covmat_raw.csv
python scripts/ncu_summary.py $G/prof_covmat_raw.csv --md profiles/ncu_covmat_r2.md > /dev/null
python scripts/sass_excerpt.py > profiles/sass_dgemm_r2.txt
grep '^{' $G/bench.json | tail -1 > profiles/bench_r2_1gpu.json
grep '^{' $G/bench_reference.json | tail -1 > profiles/bench_r2_reference_arm.json
cp $G/parity_report.txt profiles/parity_report_r2.txt
tail -3 $G/pytest_gpu.log
