// Device-side pieces of the Matern-3/2 GP pair arithmetic shared by covmat.cu (assembly, gradient contraction) and by
// the GEMM epilogue that contracts the R^-1 tiles with dR as they are produced (dgemm_tma.cu).
#pragma once
#include "common.cuh"

namespace sb {
namespace gpdev {

struct PairConsts {
    double c;        // 1 / (1 + exp(gamma_d))
    double onemc;    // exp(gamma_d) / (1 + exp(gamma_d))
};

__device__ __forceinline__ PairConsts pair_consts(double gamma_d) {
    double e = exp(gamma_d);
    return PairConsts{1.0 / (1.0 + e), e / (1.0 + e)};
}

struct GpConsts {
    double ell[MAXD];
    double c, onemc, nu, ev;
};

__device__ __forceinline__ void load_gp_consts(GpConsts* gc, const double* __restrict__ hyp, int d) {
    if (threadIdx.x < d) gc->ell[threadIdx.x] = exp(hyp[threadIdx.x]);
    if (threadIdx.x == 32) {
        PairConsts pc = pair_consts(hyp[d]);
        gc->c = pc.c;
        gc->onemc = pc.onemc;
        double e = exp(hyp[d + 2]);
        gc->nu = e / (1.0 + e);                                    // PCGPwM.py:853
        gc->ev = exp(hyp[d + 1]);
    }
}

// Accumulators of sum_ab W_ab dR_j(a,b): d length-scales | gamma_d | h_v | h_nu      (PCGPwM.py:871-907)
template <int DT>
struct GradAcc {
    double ls[DT], g, v, nn;
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < DT; k++) ls[k] = 0.0;
        g = v = nn = 0.0;
    }
};

// One off-diagonal pair (a > b) with weight w = W_ab (the caller has NOT doubled it): a, b are the pair's coordinates
// divided by the length-scales, PADDED WITH ZEROS up to DT dimensions.  dR is recomputed, never stored.
// The loops run over all DT dimensions without a guard: a padded dimension has S = 0, so it multiplies the product
// by exactly 1, adds exactly 0 to the sum and contributes exactly 0 -- the results do not depend on the padding, and
// the DT reciprocal chains stay independent straight-line code (guarded by `k < d` each became a branch region of
// its own and ran one after the other: FP64 pipe 41 % in ncu).
template <int DT>
__device__ __forceinline__ void contract_offdiag(GradAcc<DT>& A, double w, const double* a, const double* b,
                                                 const GpConsts& gc, double omn) {
    w *= 2.0;                                                      // (a,b) and (b,a)
    double prod = gc.c, ssum = 0.0, wsum[DT];
#pragma unroll
    for (int k = 0; k < DT; k++) {
        double S = fabs(a[k] - b[k]);
        double op = 1.0 + S;
        prod *= op;
        ssum -= S;
        // dR0/dgamma_k / R_pre.  (One reciprocal of the product plus leave-one-out products was measured
        // slower: the loop is FP64-pipe bound and the reciprocal is only a few DFMAs.)
        wsum[k] = (S * S) * rcp_ge1(op);
    }
    double rp = prod * exp(ssum);
    double wr = w * omn * rp;
#pragma unroll
    for (int k = 0; k < DT; k++) A.ls[k] += wr * wsum[k];          // (1-nu) dR0/dgamma_k
    A.g += w * omn * gc.onemc * (gc.c - rp);                       // (1-nu) dR0/dgamma_d
    A.nn -= w * gc.nu * omn * (rp + gc.onemc);                     // nu (1-nu) (I - R0), PCGPwM.py:878
}

}  // namespace gpdev
}  // namespace sb
