// GP-level entry points (see gp.cu / woodbury.cu); the C ABI in capi.cu forwards to these.
#pragma once
#include "common.cuh"

namespace sb {

size_t gp_nll_grad_workspace(int batch, int n, int d, int want_grad);
int gp_nll_grad(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                const double* gvar, long strideGvar, const double* hyp, const double* regmean, const double* regstd,
                double s0, int want_grad, double* nll, double* grad, int* info, void* workspace,
                size_t workspace_bytes, cudaStream_t stream);

void gp_force_blocked_path(int on);
// single-CTA fused path for n <= gp_nll_small_max_n() (nll_small.cu); gp_nll_grad dispatches to it
int gp_nll_small_max_n();
int spd_solve_small(int batch, int n, double* A, long lda, long strideA, double* logdet, int* info,
                    cudaStream_t stream);
int gp_nll_grad_small(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                      const double* gvar, long strideGvar, const double* hyp, const double* regmean,
                      const double* regstd, double s0, int want_grad, double* nll, double* grad, int* info,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream, long long* stamps = nullptr);
size_t gp_nll_small_workspace(int batch, int n, int want_grad);

size_t gp_factorize_workspace(int batch, int n);
int gp_factorize(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                 const double* gvar, long strideGvar, const double* hyp, double s0, double* Linv, long ldl,
                 long strideL, double* pw, double* sig2, double* logdet_half, int* info, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream);

// pw (K,n), sig2 (K) for PC scores g (K,n) through factors Linv[fidx[j]]; ytmp: (K,n) scratch
int gp_weights(int K, int n, const double* Linv, long ldl, long strideL, const int* fidx, const double* g, double s0,
               double* pw, double* sig2, double* ytmp, cudaStream_t stream);

size_t gp_predict_workspace(int K, int nf, int nu, int n, int d, int B, int want_grad);
int gp_predict(int K, int nf, int nu, int n, int d, int B, const double* theta, const double* thetanew,
               const double* hypcov, const int* fu, const int* fidx, const double* nug, const double* sig2,
               const double* Linv, long ldl, long strideL, const double* pw, int want_grad, long out_ld,
               double* predvecs, double* predvars, double* dpv, double* dpvar, void* workspace, size_t workspace_bytes,
               cudaStream_t stream);

int pc_backproject(int m, int K, int B, int d, const double* phi, const double* offset, const double* extravar,
                   const double* predvecs, const double* predvars, const double* dpv, const double* dpvar,
                   double* mean, double* var, double* covxhalf, double* mean_grad, double* covxhalf_grad,
                   cudaStream_t stream);

// sinv: (2, m) = 1/yvar and 1/(yvar+|extravar|);  G: (2, K, K) = Phi' diag(sinv_v) Phi;  fired: device int
int woodbury_loglik(int B, int K, int m, int d, const double* predvecs, const double* predvars, const double* dpv,
                    const double* dpvar, const double* phiT, const double* resid0, const double* extravar,
                    const double* sinv, const double* G, int want_grad, int variant, double* ll, double* dll,
                    int* fired, const int* row_valid, int segk, long segstride, const int* seg_pos, void* ws,
                    size_t ws_bytes, cudaStream_t stream);
size_t woodbury_workspace(int B, int K, int m);      // 0 while the k x k systems fit shared memory (k <= 112 at m <= 3000)
int woodbury_trigger(int B, int K, int m, const double* predvars, const double* phiT, const double* extravar,
                     int* fired, const int* row_valid, int segk, long segstride, const int* seg_pos, cudaStream_t stream);
void woodbury_force_global(int on);
// generic form for an emulator outside the PCGPwM family: mean (B,m), var (B,m), covxhalf (B,m,K) supplied by the caller;
// want_trigger = 1: only the batch-wide variance test into *fired;  0: ll (B) with Sigma^-1 = sinv (m)
int woodbury_loglik_mspace(int B, int K, int m, const double* meanT, const double* varT, const double* cxhT,
                           const double* y, const double* sinv, int want_trigger, double* ll, int* fired,
                           cudaStream_t stream);      // tests: k x k matrices in global scratch at any k (the k > 112 path)


// sampler.cu: one step of the device-resident PT-LMC sampler (surmise utilitiesmethods/PTLMC.py:203-273)
int ptlmc_step(int B, int d, int numtemps, int tune, int nsave, int mode, unsigned long long seed, double taracc,
               double* thetac, double* fval, double* dfval, double* thetap, double* z, double* adjrho, const double* ll,
               const double* dll, const double* temps, const double* lo, const double* hi, const double* hc,
               const double* cov0, double* scal, long long* kstep, double* save, int* row_valid, cudaStream_t stream);

// peer.cu: all-gather of per-rank blocks through NVLink peer memory (symmetric buffers), see the file header
int peer_allgather_post(const double* block, long count, const unsigned long long* peer_bufs, long slot_off, long flag_off,
                        int rank, int world, const unsigned long long* epoch, int* arrive, cudaStream_t stream);
int peer_allgather_wait(const unsigned long long* my_flags, int rank, int world, unsigned long long* epoch, int* error,
                        cudaStream_t stream);

}  // namespace sb
