"""profiles/ncu_kernels_r2.json (the FP64-ALU / HBM / latency-bound kernels bench.py lists in `roofline_kernels`) and
profiles/dram_traffic_r2.json from the raw exports of scripts/gpu_r2_ncu.sh:
    python scripts/make_ncu_kernels.py gpurun_out/prof_nll_r2_raw.csv gpurun_out/prof_c5_r2_raw.csv gpurun_out/prof_pca_r2_raw.csv \
        gpurun_out/prof_covmat_raw.csv
One representative launch per kernel (the last captured); the tensor kernels are timed live by bench.py and only feed
the DRAM-traffic file here."""
import json, os, re, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HBM = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbps_burst', 6556.5) \
    if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else 6556.5

# kernel (prefix of the ncu name) -> (what, bound)
TABLE = [
    ('trmv_lower_t_kernel', 'alpha = Linv^T y: streams the lower triangle of Linv once (side stream, beside the lauum product)', 'hbm'),
    ('gp_grad_contract_kernel', 'sum_ab W_ab dR_j(a,b) for the d+3 hyper-parameters, dR recomputed per pair', 'fp64'),
    ('gp_assemble_kernel', 'lower triangle of R = (1-nu) R0 + nu I + e^hv diag(gvar), g row', 'fp64'),
    ('matern_covmat_kernel<0>', 'standalone covariance, 8000 x 8000, d = 10, R only (0.51 GB written): one exp + 5d operations per pair', 'fp64'),
    ('matern_covmat_kernel<1>', 'standalone covariance with the 11 hyper-parameter gradient planes (6.1 GB written)', 'hbm'),
    ('cross_cov_t_kernel', 'cross-covariance of the proposals with the design, theta-batch contiguous', 'fp64'),
    ('predict_partial_factor_kernel', 'predict epilogue per factor: T1^2, r pw, T2 dr_i, dr_i pw partial sums', 'fp64'),
    ('woodbury_loglik_kernel', 'one CTA per theta: m-space residual, k x k Cholesky / inverse in shared memory, gradient', 'latency'),
    ('ptlmc_step_kernel', 'single-CTA PT-LMC step (accept, 5B sequential exchange proposals, adapt/save, proposal)', 'latency'),
    ('predict_reduce_kernel', 'fixed-order reduction of the predict partials', 'latency'),
    ('col_partial_kernel<0>', 'column sums / counts of f (standardisation pass 1), C4: 8000 x 2000', 'hbm'),
    ('col_partial_kernel<1>', 'centred sums of squares (standardisation pass 2)', 'hbm'),
    ('standardize_kernel', 'fs = (f - offset) / scale, read + write', 'hbm'),
    ('syevj_kernel', 'Jacobi eigensolver of the p x p Rayleigh-Ritz / orthonormalisation matrices (p = 72 at C4)', 'latency'),
]


def main(paths):
    rows = [e for p in paths if os.path.exists(p) for e in load(p)]
    out = []
    for prefix, what, bound in TABLE:
        hits = [e for e in rows if e['kernel'].replace('(bool)', '').replace('(int)', '').startswith(prefix)]
        if not hits:
            continue
        e = hits[-1] if prefix != 'syevj_kernel' else max(hits, key=lambda h: h.get('time_us', 0))
        d = {'kernel': e['kernel'], 'what': what, 'grid': e['grid'], 'time_us': round(e.get('time_us', 0.0), 1),
             'dram_read_MB': round(e.get('dram_read_MB', 0.0), 2), 'dram_write_MB': round(e.get('dram_write_MB', 0.0), 2),
             'fp64_pipe_pct': round(e.get('fp64_pipe_pct', 0.0), 1), 'warps_active_pct': round(e.get('warps_active_pct', 0.0), 1),
             'regs': e.get('regs'), 'top_stalls': [[k, v] for k, v in e.get('top_stalls', [])]}
        if bound == 'hbm':
            ach = (e.get('dram_read_MB', 0.0) + e.get('dram_write_MB', 0.0)) * 1e-3 / (e['time_us'] * 1e-6)
            d.update({'bound': 'hbm', 'achieved': round(ach, 0), 'peak': HBM, 'unit': 'GB/s', 'frac': round(ach / HBM, 3)})
        elif bound == 'fp64':
            d.update({'bound': 'fp64 alu (shares the 64 FMA/clk/SM pipe with DMMA)', 'achieved': d['fp64_pipe_pct'], 'peak': 100.0,
                      'unit': '% of FP64 pipe (ncu sm__inst_executed_pipe_fp64)', 'frac': round(d['fp64_pipe_pct'] / 100.0, 3)})
        else:
            d.update({'bound': 'latency (one CTA or one CTA per theta; neither HBM nor a pipe is loaded)', 'achieved': None,
                      'peak': None, 'unit': None, 'frac': None})
        out.append(d)
    doc = {'source': 'ncu --set full --clock-control none (scripts/gpu_r2_ncu.sh), B200, round 2 final build; one representative '
                     'launch per kernel (scripts/make_ncu_kernels.py)', 'hbm_peak_GBps': HBM, 'kernels': out}
    json.dump(doc, open(os.path.join(ROOT, 'profiles', 'ncu_kernels_r2.json'), 'w'), indent=1)
    for d in out:
        print('%-48s %8.1f us  fp64 %5.1f %%  frac %s' % (d['kernel'][:48], d['time_us'], d['fp64_pipe_pct'], d['frac']))

    # DRAM traffic of the tensor kernels of one step (the nll capture holds one step's launches)
    nll = [e for e in load(paths[0])] if paths and os.path.exists(paths[0]) else []
    # the capture starts after an assembly + factorisation and runs into the next step: one step's products are the
    # dgemm launches before the next gp_assemble launch
    cut = next((i for i, e in enumerate(nll) if e['kernel'].startswith('gp_assemble')), len(nll))
    gem = [e for e in nll[:cut] if e['kernel'].startswith('dgemm_')]
    pot = [e for e in nll if e['kernel'].startswith('potrf_dag_kernel')]
    if gem and pot:
        n_tma = sum(e['kernel'].startswith('dgemm_tma') for e in gem)
        traffic = {
            'source': 'ncu --set full --clock-control none, scripts/gpu_r2_ncu.sh (prof_paths.py nll): one C2 NLL+grad step, round 2 final build',
            'kernels': {
                'dgemm (dgemm_tma_kernel x%d + dgemm_dmma_kernel x%d: trtri levels and lauum)' % (n_tma, len(gem) - n_tma): {
                    'launches': len(gem), 'dram_read_bytes': sum(e.get('dram_read_MB', 0) for e in gem) * 1e6,
                    'dram_write_bytes': sum(e.get('dram_write_MB', 0) for e in gem) * 1e6,
                    'time_us': sum(e.get('time_us', 0) for e in gem)},
                'potrf_dag_kernel': {
                    'launches': len(pot), 'dram_read_bytes': sum(e.get('dram_read_MB', 0) for e in pot) * 1e6,
                    'dram_write_bytes': sum(e.get('dram_write_MB', 0) for e in pot) * 1e6,
                    'time_us': sum(e.get('time_us', 0) for e in pot),
                    'algorithmic_bytes': '20 x (2001 x 2000 x 8 B lower half read + written) ~ 0.64 GB; the left-looking '
                                         'tiles re-read their operand rows'}}}
        json.dump(traffic, open(os.path.join(ROOT, 'profiles', 'dram_traffic_r2.json'), 'w'), indent=1)
        for k, v in traffic['kernels'].items():
            print('%-40s n=%d  %.2f GB read  %.2f GB written  %.0f us' % (k[:40], v['launches'], v['dram_read_bytes'] / 1e9,
                                                                          v['dram_write_bytes'] / 1e9, v['time_us']))


if __name__ == '__main__':
    main(sys.argv[1:])
