"""ctypes binding of libsurmise_b200.so (the C ABI in include/surmise_b200.h).

There is deliberately no CPU fallback: importing this module without the built library, or
calling into it without a CUDA device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('SURMISE_B200_LIB') or os.path.join(_HERE, 'lib', 'libsurmise_b200.so')   # env: experiments only

if not os.path.exists(LIB_PATH):
    raise ImportError(
        'surmise_b200: %s is missing -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
        '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i = C.c_int
_l = C.c_long
_d = C.c_double
_z = C.c_size_t

SIGNATURES = {
    'sb200_version': (_i, []),
    'sb200_last_error': (C.c_char_p, []),
    'sb200_launch_count': (_l, []),
    'sb200_tma_gemm_count': (_l, []),
    'sb200_gemm_profile_enable': (None, [_i]),
    'sb200_gemm_profile_collect': (_i, [_p, _p]),
    'sb200_kernel_profile_collect': (_i, [_i, _p, _p]),
    'sb200_gemm_profile_dump': (_i, [_p, _p, _i]),
    'sb200_matern_covmat': (_i, [_p, _i, _p, _i, _i, _p, _p, _l, _p, _i, _p]),
    'sb200_gp_nll_small_max_n': (_i, []),
    'sb200_spd_solve_small': (_i, [_i, _i, _p, _l, _l, _p, _p, _p]),
    'sb200_gp_force_blocked_path': (None, [_i]),
    'sb200_gp_nll_small_profile': (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _z, _p]),
    'sb200_gp_nll_grad_workspace': (_z, [_i, _i, _i, _i]),
    'sb200_gp_nll_grad': (_i, [_i, _i, _i, _p, _l, _p, _l, _p, _l, _p, _p, _p, _d, _i, _p, _p, _p, _p, _z, _p]),
    'sb200_gp_factorize_workspace': (_z, [_i, _i]),
    'sb200_gp_factorize': (_i, [_i, _i, _i, _p, _l, _p, _l, _p, _l, _p, _d, _p, _l, _l, _p, _p, _p, _p, _p, _z, _p]),
    'sb200_gp_weights': (_i, [_i, _i, _p, _l, _l, _p, _p, _d, _p, _p, _p, _p]),
    'sb200_gp_predict_workspace': (_z, [_i, _i, _i, _i, _i, _i, _i]),
    'sb200_gp_predict': (_i, [_i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _p, _i, _l,
                              _p, _p, _p, _p, _p, _z, _p]),
    'sb200_pc_backproject': (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    'sb200_woodbury_loglik': (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _l, _p, _p]),
    'sb200_woodbury_workspace': (_z, [_i, _i, _i]),
    'sb200_woodbury_loglik_ws': (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _l, _p,
                                      _p, _z, _p]),
    'sb200_woodbury_force_global': (None, [_i]),
    'sb200_woodbury_loglik_mspace': (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _i, _p, _p, _p]),
    'sb200_woodbury_trigger': (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _i, _l, _p, _p]),
    'sb200_ptlmc_step': (_i, [_i, _i, _i, _i, _i, _i, C.c_ulonglong, _d] + [_p] * 17 + [_p]),
    'sb200_host_tempexchange': (None, [_p, _p, _i, _p, _p, _i, _p]),
    'sb200_dgemm': (_i, [_i, _i, _i, _i, _i, _d, _p, _l, _l, _p, _l, _l, _d, _p, _l, _l, _i, _i, _i, _p]),
    'sb200_potrf_batched': (_i, [_i, _i, _i, _p, _l, _l, _p, _l, _l, _p, _p]),
    'sb200_potrf_info_ints': (_z, [_i, _i]),
    'sb200_trtri_tmp_elems': (_z, [_i]),
    'sb200_debug_panel_stamps': (None, [_p]),
    'sb200_debug_dag_stats': (None, [_p]),
    'sb200_trtri_batched': (_i, [_i, _i, _p, _l, _l, _p, _l, _l, _p, _z, _p]),
    'sb200_lauum_batched': (_i, [_i, _i, _p, _l, _l, _p, _l, _l, _p]),
    'sb200_pca_col_stats_workspace': (_z, [_i, _i]),
    'sb200_pca_col_stats': (_i, [_i, _i, _p, _l, _p, _p, _p, _p, _z, _p]),
    'sb200_pca_standardize': (_i, [_i, _i, _p, _l, _p, _p, _p, _l, _p, _p, _p]),
    'sb200_pca_sumsq': (_i, [_l, _p, _p, _p, _p]),
    'sb200_peer_allgather_post': (_i, [_p, _l, _p, _l, _l, _i, _i, _p, _p, _p]),
    'sb200_peer_allgather_wait': (_i, [_p, _i, _i, _p, _p, _p]),
    'sb200_syevj_max_p': (_i, []),
    'sb200_syevj_batched': (_i, [_i, _i, _p, _l, _l, _p, _l, _p, _l, _l, _i, _p, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


class SurmiseB200Error(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        msg = lib.sb200_last_error().decode('utf-8', 'replace')
        raise SurmiseB200Error('%s failed (rc=%d): %s' % (what, rc, msg))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise SurmiseB200Error('surmise_b200 needs a CUDA device (built for sm_100a / B200); '
                               'there is no CPU fallback.')
    return torch
