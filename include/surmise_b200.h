/* surmise_b200 -- C ABI of the B200 (sm_100a) hot path for bandframework/surmise.
 *
 * Scope: the PCGPwM emulator's covariance / likelihood / factorisation / prediction and the
 * directbayeswoodbury calibration log-likelihood.  Each entry point names the reference code it
 * replaces (paths relative to the surmise source tree).
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer to fp64 (or int32 where typed so), row-major;
 *   - the caller owns all memory, including workspaces (size them with the *_workspace queries);
 *     the library never allocates or frees device memory.  The compute entry points keep no mutable state of their
 *     own: what exists process-wide is (a) a thread-local last-error string, (b) the one-time per-device set-up
 *     (kernel attributes and, for sb200_gp_nll_grad, one non-blocking side stream with two events per device on which a
 *     bandwidth-bound product runs beside a tensor-bound one; forked from and joined to the caller's stream inside the
 *     call, capturable; mutex-guarded, safe for concurrent first calls from several host threads), and (c) the
 *     instrumentation / testing switches declared below -- an atomic launch counter, the optional GEMM timing record
 *     (mutex-guarded), sb200_gp_force_blocked_path, sb200_woodbury_force_global and sb200_debug_panel_stamps (atomics; set them only while no
 *     call is in flight).  None of (c) changes a numerical result;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); outputs are valid
 *     after the stream is synchronised;
 *   - return value: 0 success, < 0 invalid argument, > 0 a cudaError_t; sb200_last_error() explains;
 *   - numerical breakdown (a non-positive pivot) is reported per problem in `info` (LAPACK
 *     convention: j > 0 = leading minor j not positive definite), never by aborting;
 *   - parameter dimension d <= 16 (the per-dimension loops live in registers).
 */
#ifndef SURMISE_B200_H
#define SURMISE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int sb200_version(void);
const char* sb200_last_error(void);

/* Instrumentation used by bench.py: number of kernels this library has launched in the process so far, and an
 * optional per-launch CUDA-event timing of the DMMA GEMM (enable, run, collect -> launches, summed ms, summed
 * algorithmic flop = 2 M N K with exact triangular ranges). */
long sb200_launch_count(void);
long sb200_tma_gemm_count(void);    /* DMMA GEMM launches that staged their operands by TMA (cp.async.bulk.tensor) */
void sb200_gemm_profile_enable(int on);
int sb200_gemm_profile_collect(double* total_ms, double* total_flops);
int sb200_kernel_profile_collect(int kind, double* total_ms, double* total_flops);   /* kind 0 = DMMA GEMM launches,
                                                   1 = task-graph Cholesky launches (algorithmic flop n^3/3 + rhs n^2) */
int sb200_gemm_profile_dump(double* ms, double* flops, int cap);   /* per launch, host arrays */

/* Replaces surmise.emulationsupport.matern_covmat.covmat (emulationsupport/matern_covmat.pyx:27-64,
 * map4 at :13-24).  R[i][j] (n1 x n2, leading dimension ldr) is the separable Matern-3/2 correlation
 * between x1[i] and x2[j] (both (.,d) contiguous) for gammav (d+1).
 * grad_mode 0: R only.  1: also dR = d/dgammav as (d+1) planes of n1 x n2 (the reference returns the
 * (1,2,0)-transpose view of exactly such a buffer, pyx:52).  2: dR = d/dx1 as d planes of n1 x n2. */
int sb200_matern_covmat(const double* x1, int n1, const double* x2, int n2, int d, const double* gammav,
                        double* R, long ldr, double* dR, int grad_mode, void* stream);

/* Replaces PCGPwM.__negloglik and PCGPwM.__negloglikgrad (emulationmethods/PCGPwM.py:850-868, 871-907)
 * for `batch` problems at once (principal components and/or candidate hyper-parameter vectors).
 * theta (n,d), g (n), gvar (n, may be NULL) with per-problem strides in elements (0 = shared);
 * hyp (batch, d+3) = [gamma_0..gamma_d, log var-const, logit nugget]; regmean/regstd (d+3) shared.
 * Outputs nll (batch), grad (batch, d+3; only if want_grad), info (batch).
 * Problems with n <= sb200_gp_nll_small_max_n() (207) take three launches (assemble, register-resident sweep
 * operator with one CTA per problem, gradient contraction); larger ones use the blocked DMMA path.  sb200_gp_force_blocked_path(1) sends everything through
 * the blocked path (testing). */
int sb200_gp_nll_small_max_n(void);
void sb200_gp_force_blocked_path(int on);
/* instrumentation: as sb200_gp_nll_grad on shared theta/g without gvar, plus clock64() stamps (8 slots) of CTA 0 */
int sb200_gp_nll_small_profile(int batch, int n, int d, const double* theta, const double* g, const double* hyp,
                               const double* regmean, const double* regstd, double* nll, double* grad, int* info,
                               long long* stamps, void* workspace, size_t workspace_bytes, void* stream);
/* Batched symmetric-positive-definite solve on the same register-resident sweep kernel, n <= sb200_gp_nll_small_max_n().
 * A (batch, n+1, lda): lower triangle of S in rows 0..n-1 and the right-hand side z' as row n.  On return rows 0..n-1
 * hold the lower triangle of S^-1 and row n holds (S^-1 z)'; logdet (batch, may be NULL) = log|S|; info per problem
 * (LAPACK convention).  Replaces the per-row ridge solves of the imputation sweeps, PCGPwM.__standardizef
 * (emulationmethods/PCGPwM.py:573-587), in their |missing| x |missing| Schur form. */
int sb200_spd_solve_small(int batch, int n, double* A, long lda, long stride_a, double* logdet, int* info,
                          void* stream);
size_t sb200_gp_nll_grad_workspace(int batch, int n, int d, int want_grad);
int sb200_gp_nll_grad(int batch, int n, int d, const double* theta, long stride_theta, const double* g,
                      long stride_g, const double* gvar, long stride_gvar, const double* hyp,
                      const double* regmean, const double* regstd, double sig2ofconst, int want_grad,
                      double* nll, double* grad, int* info, void* workspace, size_t workspace_bytes, void* stream);

/* Replaces the full-data stage of PCGPwM.__fitGP1d (PCGPwM.py:810-846): for each problem builds
 * R = (1-nug) R0 + nug I + exp(h_v) diag(gvar), factors it and returns
 *   Linv (n x n lower triangular, leading dimension ldl, zero above the diagonal) with Linv' Linv = R^-1
 *        -- stored instead of the reference's R, Rinv and Vh;
 *   pw = R^-1 g (n), sig2 = (g' R^-1 g + sig2ofconst)/(n + sig2ofconst), logdet_half = 1/2 log|R|. */
size_t sb200_gp_factorize_workspace(int batch, int n);
int sb200_gp_factorize(int batch, int n, int d, const double* theta, long stride_theta, const double* g,
                       long stride_g, const double* gvar, long stride_gvar, const double* hyp, double sig2ofconst,
                       double* Linv, long ldl, long stride_linv, double* pw, double* sig2, double* logdet_half,
                       int* info, void* workspace, size_t workspace_bytes, void* stream);

/* Companion of sb200_gp_factorize for principal components that share a factor (identical hyper-parameters
 * and gvar, PCGPwM.py:808-833): pw_j = R^-1 g_j and sig2_j (PCGPwM.py:823-846) for scores g (k,n) through
 * Linv[fidx[j]] (fidx NULL = identity).  ytmp: (k,n) scratch. */
int sb200_gp_weights(int k, int n, const double* Linv, long ldl, long stride_linv, const int* fidx, const double* g,
                     double sig2ofconst, double* pw, double* sig2, double* ytmp, void* stream);

/* Replaces the per-PC loop of PCGPwM.predict (PCGPwM.py:219-270): predvecs/predvars (B,k) and, if
 * want_grad, their gradients (B,k,d) at thetanew (B,d).  want_grad: 0 none, 1 gradients, 2 gradients for a caller whose
 * WHOLE population is one theta with d == 1 -- there the reference applies (1-nug) to predvecs_gradtheta once instead
 * of twice (PCGPwM.py:236 and 262-268); the caller decides this from its full population, so chunks and per-rank
 * slices of a larger population (B == 1 locally) keep the B > 1 behaviour.
 *   hypcov (nu, d+1): the distinct covariance hyper-parameter sets;  fu (nf): set used by factor f;
 *   Linv (nf, n, ldl): distinct factors;  fidx (k): factor used by PC j;  nug, sig2 (k);  pw (k, n).
 *   out_ld: row length of the outputs in components, (B, out_ld) and (B, out_ld, d); 0 or k = dense.  A caller that
 *   assembles the components of several ranks side by side (PC-sharded prediction) passes the padded
 *   width and this call fills its k columns of every row. */
size_t sb200_gp_predict_workspace(int k, int nf, int nu, int n, int d, int B, int want_grad);
int sb200_gp_predict(int k, int nf, int nu, int n, int d, int B, const double* theta, const double* thetanew,
                     const double* hypcov, const int* fu, const int* fidx, const double* nug, const double* sig2,
                     const double* Linv, long ldl, long stride_linv, const double* pw, int want_grad, long out_ld,
                     double* predvecs, double* predvars, double* predvecs_grad, double* predvars_grad,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Replaces the back-projection of PCGPwM.predict (PCGPwM.py:273-327).  phi (m,k) = pcti * scale;
 * outputs mean, var (m,B); optional (NULL to skip) covxhalf (m,B,k), mean_grad (m,B,d),
 * covxhalf_grad (m,B,k,d). */
int sb200_pc_backproject(int m, int k, int B, int d, const double* phi, const double* offset,
                         const double* extravar, const double* predvecs, const double* predvars,
                         const double* predvecs_grad, const double* predvars_grad, double* mean, double* var,
                         double* covxhalf, double* mean_grad, double* covxhalf_grad, void* stream);

/* Replaces directbayeswoodbury.loglik / loglik_grad (calibrationmethods/directbayeswoodbury.py:261-306,
 * 309-383) for a whole batch of theta in one launch, in principal-component space.
 *   phiT (k,m); resid0 (m) = y - offset; extravar (m); sinv (2,m) = 1/yvar, 1/(yvar+|extravar|);
 *   G (2,k,k) = Phi' diag(sinv[v]) Phi.  Outputs ll (B), dll (B,d; if want_grad) and *fired (device int):
 *   1 when the batch-wide unmodelled-variance test (dbw.py:277-284) selected the second variant.
 *   variant < 0: run that test over this batch; 0 / 1: use the given variant (a theta population sharded over
 *   several GPUs evaluates sb200_woodbury_trigger per shard, OR-reduces the flags and passes the result).
 *   row_valid (B ints, may be NULL): rows with 0 take no part in the batch-wide test -- the reference only ever
 *   evaluates the proposals with finite prior (dbw.py:100-129), so a caller that keeps the whole population on the
 *   device masks the others here; their ll / dll are still written and are the caller's to discard.
 *   seg_k, seg_stride: 0, 0 for dense (B,k) / (B,k,d) inputs.  Otherwise the k components arrive in k / seg_k groups,
 *   each a dense (B, seg_k[, d]) block, consecutive groups seg_stride doubles apart -- exactly what an all-gather of the
 *   per-rank blocks of a PC-sharded prediction (sb200_gp_predict with out_ld = seg_k) leaves in memory.  seg_pos NULL:
 *   k counts the padded slots too and phiT / G are given in that order, with zero rows for the padded components;
 *   seg_pos (k ints, device): k counts the REAL components only, seg_pos[a] is the padded slot of component a, and
 *   phiT / G are the ordinary (k, m) / (2, k, k) arrays -- the k x k work then does not grow with the number of ranks. */
int sb200_woodbury_loglik(int B, int k, int m, int d, const double* predvecs, const double* predvars,
                          const double* predvecs_grad, const double* predvars_grad, const double* phiT,
                          const double* resid0, const double* extravar, const double* sinv, const double* G,
                          int want_grad, int variant, double* ll, double* dll, int* fired, const int* row_valid,
                          int seg_k, long seg_stride, const int* seg_pos, void* stream);
int sb200_woodbury_trigger(int B, int k, int m, const double* predvars, const double* phiT, const double* extravar,
                           int* fired, const int* row_valid, int seg_k, long seg_stride, const int* seg_pos,
                           void* stream);
/* The kernel keeps the two k x k matrices of a proposal (I + D G D, its factor and inverse) in shared memory while
 * (2 k^2 + m + 7 k + 32) doubles fit 227 KB -- k <= 112 at m <= 3000; sb200_woodbury_workspace then returns 0 and
 * sb200_woodbury_loglik needs no scratch.  Beyond that each CTA works on a slice of a caller-provided global buffer
 * (same code, L2-resident): size it with sb200_woodbury_workspace(B, k, m) (at most 1 GiB; the proposals are then taken
 * in chunks) and call sb200_woodbury_loglik_ws; sb200_woodbury_loglik itself reports such a shape as an argument
 * error.  sb200_woodbury_force_global(1) sends every shape through the scratch path (testing switch, as (c) above). */
size_t sb200_woodbury_workspace(int B, int k, int m);
int sb200_woodbury_loglik_ws(int B, int k, int m, int d, const double* predvecs, const double* predvars,
                             const double* predvecs_grad, const double* predvars_grad, const double* phiT,
                             const double* resid0, const double* extravar, const double* sinv, const double* G,
                             int want_grad, int variant, double* ll, double* dll, int* fired, const int* row_valid,
                             int seg_k, long seg_stride, const int* seg_pos, void* workspace, size_t workspace_bytes,
                             void* stream);
void sb200_woodbury_force_global(int on);

/* directbayeswoodbury.loglik (calibrationmethods/directbayeswoodbury.py:261-306) as written, for an emulator OUTSIDE the
 * PCGPwM family (no principal-component factors on the device): the caller evaluates that emulator's own predict and
 * passes its outputs at the B proposals, proposal-major: mean (B,m), var (B,m), covxhalf (B,m,k).
 *   want_trigger = 1: only the batch-wide unmodelled-variance test (dbw.py:277-279) -> *fired (device int); ll unused.
 *   want_trigger = 0: ll (B) for observation precision sinv (m) = 1 / obsvar (the caller adds the unmodelled variance
 *   to obsvar when the test fired, dbw.py:280-284); one CTA per proposal, I + J'J (k x k) factored in shared memory
 *   ((k^2 + m + k + 32) doubles <= 227 KB).  No gradient in this form (the samplers that need one take a PCGPwM
 *   emulator). */
int sb200_woodbury_loglik_mspace(int B, int k, int m, const double* mean, const double* var, const double* covxhalf,
                                 const double* y, const double* sinv, int want_trigger, double* ll, int* fired,
                                 void* stream);

/* One step of the device-resident PT-LMC sampler: what surmise's PTLMC (utilitiesmethods/PTLMC.py:203-273) does between
 * two evaluations of the log-posterior, for a population of B = numtemps + numchain chains kept on the device.
 *   mode bit 0: accept / reject the pending proposal thetap given its log-likelihood ll (B) and gradient dll (B,d)
 *               (:214-237), 5 B temperature-exchange proposals (:239-246, 259-273), then while k < tune and k % 10 == 0
 *               the step-size adaptation (:247-254), or for k >= tune save thetac[numtemps:] as draw k - tune (:255-257);
 *   mode bit 1: draw the Langevin proposal of the next step into thetap (:205-213) and mark in row_valid (B ints)
 *               which proposals lie inside the prior box [lo, hi] (d) -- the others are never accepted.
 *   First call: mode 2; then gp_predict + woodbury_loglik(thetap, row_valid) and mode 3, repeatedly; last call: mode 1.
 * State (device, caller-owned): thetac (B,d), fval (B) = log-posterior / temperature, dfval (B,d) likewise, z (B,d),
 * adjrho (B), scal (8): [0] tau [1] accepted fraction since the last adaptation [2] accepted moves [3] accepted swaps
 * [4] rho [5] proposals outside the box; kstep (1, 64-bit): index of the pending step; save (numchain, nsave, d).
 * Constants: temps (B) as PTLMC.py:75-82, hc / cov0 (d,d) the proposal factors (:192-201).  Randomness: Philox4x32-10
 * keyed by seed with counter (step, stream, index), so replicated ranks that run the same launches hold identical
 * chains.  The prior is uniform on the box; a constant log-density cancels in every acceptance ratio. */
int sb200_ptlmc_step(int B, int d, int numtemps, int tune, int nsave, int mode, unsigned long long seed, double taracc,
                     double* thetac, double* fval, double* dfval, double* thetap, double* z, double* adjrho,
                     const double* ll, const double* dll, const double* temps, const double* lo, const double* hi,
                     const double* hc, const double* cov0, double* scal, long long* kstep, double* save,
                     int* row_valid, void* stream);

/* Host-side helper (no device work) for the PTLMC_B200 sampler plugin: the temperature-exchange sweep of
 * utilitiesmethods/PTLMC.py:259-273 (tempexchange), with the random draws supplied by the caller.  lpost, temps (B);
 * rtv, logu (iters*B); order (B) out. */
void sb200_host_tempexchange(const double* lpost, const double* temps, int B, const int* rtv, const double* logu,
                             int iters, int* order);

/* Building blocks, exported for testing and reuse (no reference counterpart: the reference calls LAPACK).
 * sb200_dgemm: C[m][n] = alpha sum_k Aop[m][k] Bop[n][k] + beta C[m][n] on the FP64 tensor cores (DMMA);
 *   a_kc != 0: Aop[m][k] = A[m*lda+k] else A[k*lda+m]; same for B.  flags: 1 A lower, 2 A upper,
 *   4 B lower, 8 B upper (triangular operands must hold zeros in the empty triangle), 16 lower-C only.
 * sb200_potrf_batched: in-place lower Cholesky of `batch` matrices (rows n..mrows-1 are solved as
 *   right-hand sides X <- X L^-T); Linv receives the inverted diagonal blocks and is zeroed elsewhere;
 *   `info` must hold sb200_potrf_info_ints(batch, mrows) ints (the first `batch` = result, the rest = scratch: arrival
 *   counters, and the row-progress counters / ticket of the task-graph kernel that factors the whole batch in ONE
 *   persistent launch -- left-looking 64 x 64 tile tasks, dependencies tracked per tile row).
 * sb200_trtri_batched: completes Linv = L^-1 from the output of sb200_potrf_batched.
 * sb200_lauum_batched: lower triangle of C = Linv' Linv. */
int sb200_dgemm(int a_kc, int b_kc, int M, int N, int K, double alpha, const double* A, long lda, long stride_a,
                const double* B, long ldb, long stride_b, double beta, double* C, long ldc, long stride_c,
                int batch, int flags, int tile, void* stream);
int sb200_potrf_batched(int batch, int n, int mrows, double* A, long lda, long stride_a, double* Linv, long ldl,
                        long stride_linv, int* info, void* stream);
size_t sb200_potrf_info_ints(int batch, int mrows);
size_t sb200_trtri_tmp_elems(int n);
void sb200_debug_panel_stamps(long long* device_ptr);   /* instrumentation: clock64 stamps of the panel kernel's phases */
void sb200_debug_dag_stats(long long* device_ptr);      /* instrumentation: per CTA of the task-graph Cholesky, 8 slots
                                                           [cycles total, waiting on dependencies, in diagonal tasks'
                                                           factorisation, in off-diagonal tasks' solve, tasks, 64x64 factor +
                                                           inverse, solve of the rows under the block, diagonal tasks]; NULL = off */
int sb200_trtri_batched(int batch, int n, const double* L, long lda, long stride_a, double* Linv, long ldl,
                        long stride_linv, double* tmp, size_t tmp_elems_per_matrix, void* stream);
int sb200_lauum_batched(int batch, int n, const double* Linv, long ldl, long stride_linv, double* C, long ldc,
                        long stride_c, void* stream);

/* Standardisation and PCA stage of PCGPwM (emulationmethods/PCGPwM.py:533-605 __standardizef, 608-654 __PCs).
 * sb200_pca_col_stats: per column of f (n x m, leading dimension ldf; non-finite cells are missing) offset = nanmean,
 *   scale = nanstd / sqrt(1 - missing fraction) (PCGPwM.py:553-557) and nmiss = number of missing cells.
 * sb200_pca_standardize: fs = (f - offset) / scale; with miss != NULL (n x m bytes) the missing mask is written and
 *   the missing cells get the reference's alternating +2 / -2 seed minus its mean, in row order (PCGPwM.py:565-571).
 * sb200_pca_sumsq: out[0] = sum of squares of `count` doubles (deterministic); scratch holds 1024 doubles.
 * sb200_syevj_batched: eigen-decomposition of `batch` symmetric p x p matrices, p <= sb200_syevj_max_p() (112), by
 *   cyclic Jacobi with one CTA per matrix (matrix and vectors in shared memory): W (p) descending, the columns of
 *   Z (p x p, ldz) the eigenvectors; sweeps (batch, may be NULL) = sweeps used.  With sb200_dgemm it forms the block
 *   subspace iteration (orthonormalisation + Rayleigh-Ritz) that replaces the reference's full SVD of fs', of which
 *   only the components with S^2 > epsilonPC are ever used (PCGPwM.py:575-577, 589-591, 620-622). */
size_t sb200_pca_col_stats_workspace(int n, int m);
int sb200_pca_col_stats(int n, int m, const double* f, long ldf, double* offset, double* scale, int* nmiss,
                        void* workspace, size_t workspace_bytes, void* stream);
int sb200_pca_standardize(int n, int m, const double* f, long ldf, const double* offset, const double* scale,
                          double* fs, long ldfs, unsigned char* miss, const int* nmiss, void* stream);
int sb200_pca_sumsq(long count, const double* x, double* out, double* scratch, void* stream);
int sb200_syevj_max_p(void);
int sb200_syevj_batched(int batch, int p, const double* H, long ldh, long stride_h, double* W, long stride_w,
                        double* Z, long ldz, long stride_z, int max_sweeps, int* sweeps, void* stream);

/* All-gather of per-rank blocks through NVLink peer memory, for the PC-sharded device-resident sampler (the exchange
 * step between predict and the replicated Woodbury kernel; SURVEY 8e "allgather of the results").  Every rank owns a
 * buffer that is mapped into all peers (symmetric memory); peer_bufs is a DEVICE array of `world` base addresses of
 * those buffers as seen from this rank.  post: the `count` doubles at `block` are stored into every peer's buffer at
 * slot_offset (doubles), then the flag word (peer buffer + flag_offset doubles, 64-bit words, index = rank) is raised to
 * epoch[0] + 1 with release semantics at system scope.  wait: one CTA spins on this rank's own flags (my_flags =
 * own buffer + flag_offset) until every peer has reached epoch[0] + 1, then stores that epoch; a peer that does not
 * arrive within ~2 s sets *error instead of hanging.  epoch, arrive (world ints, zero) and error are device words
 * owned by the caller; no host value changes between steps, so the pair can be replayed inside a CUDA graph. */
int sb200_peer_allgather_post(const double* block, long count, const unsigned long long* peer_bufs, long slot_offset,
                              long flag_offset, int rank, int world, const unsigned long long* epoch, int* arrive,
                              void* stream);
int sb200_peer_allgather_wait(const unsigned long long* my_flags, int rank, int world, unsigned long long* epoch,
                              int* error, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SURMISE_B200_H */
