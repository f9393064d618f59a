// Matern-3/2 covariance kernels (replacement for surmise's matern_covmat C extension) and the
// element-wise GP kernels built on the same pair function.
#pragma once
#include "common.cuh"

namespace sb {

// R (n1, ldr) [+ dR as (D, n1, n2) planes: grad_mode 1 -> D = d+1 (d/dgamma), 2 -> D = d (d/dx1)]
int matern_covmat(const double* x1, int n1, const double* x2, int n2, int d, const double* gammav, double* R,
                  long ldr, double* dR, int grad_mode, cudaStream_t stream);

// Batched GP matrix, lower triangle + diagonal, plus g' as row n (the augmented right-hand side):
//   A[z][a][b] = (1-nu) R0(a,b) + [a==b] (nu + exp(h_v) gvar[a]),  a >= b;   A[z][n][b] = g[z][b]
// hyp: (batch, d+3) = [gamma_0..gamma_{d-1}, gamma_d, h_v, h_nu]
int gp_assemble(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                const double* gvar, long strideGvar, const double* hyp, double* A, long lda, long strideA,
                cudaStream_t stream);

// grad[z][j] = sum_ab W_ab dR_j(a,b) + prior_j,  W = 1/2 Rinv - 1/2 n/((n+s0) sig2hat) alpha alpha'
// Rinv: lower triangle valid.  partial: (batch, ntiles, MAXP) scratch, ntiles = grad_contract_tiles(n)
int gp_grad_contract(int batch, int n, int d, const double* theta, long strideTheta, const double* gvar,
                     long strideGvar, const double* hyp, const double* regmean, const double* regstd,
                     const double* Rinv, long ldr, long strideR, const double* alpha, long strideAlpha,
                     const double* sig2hat, double s0, double* partial, double* grad, cudaStream_t stream);
int grad_contract_tiles(int n);
// the fixed-order sum over the tile partials plus the prior term (second half of gp_grad_contract; also finishes the
// partials written by the GEMM's contraction epilogue, dgemm.cuh GradContractEpilogue)
int gp_grad_finalize(int batch, int d, int ntiles, const double* partial, const double* hyp, const double* regmean,
                     const double* regstd, double* grad, cudaStream_t stream);

// rT[u][j][b] = R0(thetanew_b, theta_j; hypcov_u)   (b contiguous, leading dimension ldb)
int cross_cov_t(int nu, int n, int d, int B, const double* theta, const double* thetanew, const double* hypcov,
                double* rT, long ldb, long strideU, cudaStream_t stream);

}  // namespace sb
