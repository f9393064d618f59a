// extern "C" boundary: plain pointers and sizes only (see include/surmise_b200.h).
#include "../../include/surmise_b200.h"
#include "chol.cuh"
#include "covmat.cuh"
#include "dgemm.cuh"
#include "gp.cuh"
#include "pca.cuh"
#include <atomic>
#include <cstdarg>
#include <mutex>
#include <vector>

namespace sb {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct GemmSample { cudaEvent_t a, b; double flops; int kind; };
static std::mutex g_prof_mu;
static std::vector<GemmSample> g_prof;
static std::atomic<bool> g_prof_on{false};
static std::vector<cudaEvent_t> g_event_pool;      // guarded by g_prof_mu
bool gemm_profile_enabled() { return g_prof_on.load(std::memory_order_relaxed); }
// The events of a collected sample go back to a pool: creating two events per launch kept the host behind the device
// on the small launches of the roofline leg, and the gaps ended up inside the event brackets (GEMM time 4.06 -> 4.25 ms
// on a slow host).
int profile_event_pair(cudaEvent_t* a, cudaEvent_t* b) {
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (g_event_pool.size() >= 2) {
            *a = g_event_pool.back(); g_event_pool.pop_back();
            *b = g_event_pool.back(); g_event_pool.pop_back();
            return 0;
        }
    }
    SB_CUDA(cudaEventCreate(a));
    SB_CUDA(cudaEventCreate(b));
    return 0;
}
void gemm_profile_add(cudaEvent_t a, cudaEvent_t b, double flops) { kernel_profile_add(0, a, b, flops); }
void kernel_profile_add(int kind, cudaEvent_t a, cudaEvent_t b, double flops) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(GemmSample{a, b, flops, kind});
}
}  // namespace sb

using namespace sb;
#define ST(s) ((cudaStream_t)(s))

extern "C" {

int sb200_version(void) { return 100; }

// Host helper of the PTLMC_B200 sampler: the parallel-tempering exchange sweep of surmise's PTLMC
// (utilitiesmethods/PTLMC.py:259-273) -- `iters` passes of B random adjacent-pair proposals, each accepted when
// (lpost[order[rt]] - lpost[order[rt-1]]) * (1/T[rt-1] - 1/T[rt]) > log u.  Inherently sequential (every swap changes
// what the next proposal sees), 8.8 ms per MCMC step as a Python loop, microseconds here.  Pure host code:
// rtv (iters*B draws from 1..B-1) and logu (iters*B log-uniforms) are supplied by the caller's RNG.
void sb200_host_tempexchange(const double* lpost, const double* temps, int B, const int* rtv, const double* logu,
                             int iters, int* order) {
    for (int i = 0; i < B; i++) order[i] = i;
    for (int it = 0; it < iters; it++)
        for (int j = 0; j < B; j++) {
            const int rt = rtv[it * B + j];
            if (rt < 1 || rt >= B) continue;
            const double rhoh = 1.0 / temps[rt - 1] - 1.0 / temps[rt];
            if ((lpost[order[rt]] - lpost[order[rt - 1]]) * rhoh > logu[it * B + j]) {
                const int t = order[rt - 1];
                order[rt - 1] = order[rt];
                order[rt] = t;
            }
        }
}

long sb200_launch_count(void) { return sb::g_launches.load(); }
long sb200_tma_gemm_count(void) { return sb::dgemm_tma_launch_count(); }

void sb200_gemm_profile_enable(int on) { sb::g_prof_on.store(on != 0); }

// Per-launch variant of sb200_gemm_profile_collect: fills ms[i], flops[i] for up to cap launches (in launch order),
// returns the number of launches recorded; clears the record.
int sb200_gemm_profile_dump(double* ms, double* flops, int cap) {
    std::lock_guard<std::mutex> lk(sb::g_prof_mu);
    int n = 0;
    for (auto& smp : sb::g_prof) {
        float t = 0.f;
        bool ok = smp.kind == 0 && cudaEventSynchronize(smp.b) == cudaSuccess &&
                  cudaEventElapsedTime(&t, smp.a, smp.b) == cudaSuccess;
        if (ok && n < cap) {
            ms[n] = t;
            flops[n] = smp.flops;
            n++;
        }
        sb::g_event_pool.push_back(smp.a);
        sb::g_event_pool.push_back(smp.b);
    }
    sb::g_prof.clear();
    return n;
}

// Synchronises the recorded events of one kind (0 = DMMA GEMM launches, 1 = task-graph Cholesky launches), returns how
// many were booked since the last call and their summed device time (ms) and algorithmic flop; clears those records.
int sb200_kernel_profile_collect(int kind, double* total_ms, double* total_flops) {
    std::lock_guard<std::mutex> lk(sb::g_prof_mu);
    double ms = 0.0, fl = 0.0;
    int n = 0;
    std::vector<sb::GemmSample> keep;
    for (auto& smp : sb::g_prof) {
        if (smp.kind != kind) {
            keep.push_back(smp);
            continue;
        }
        float t = 0.f;
        if (cudaEventSynchronize(smp.b) == cudaSuccess && cudaEventElapsedTime(&t, smp.a, smp.b) == cudaSuccess) {
            ms += t;
            fl += smp.flops;
            n++;
        }
        sb::g_event_pool.push_back(smp.a);
        sb::g_event_pool.push_back(smp.b);
    }
    sb::g_prof.swap(keep);
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    return n;
}
int sb200_gemm_profile_collect(double* total_ms, double* total_flops) {
    return sb200_kernel_profile_collect(0, total_ms, total_flops);
}
const char* sb200_last_error(void) { return sb::g_err; }

int sb200_matern_covmat(const double* x1, int n1, const double* x2, int n2, int d, const double* gammav,
                        double* R, long ldr, double* dR, int grad_mode, void* stream) {
    SB_CHECK_ARG(x1 && x2 && gammav && R, "matern_covmat: NULL pointer");
    SB_CHECK_ARG(n1 >= 0 && n2 >= 0 && ldr >= n2, "matern_covmat: bad shape n1=%d n2=%d ldr=%ld", n1, n2, ldr);
    return matern_covmat(x1, n1, x2, n2, d, gammav, R, ldr, dR, grad_mode, ST(stream));
}

size_t sb200_gp_nll_grad_workspace(int batch, int n, int d, int want_grad) {
    return gp_nll_grad_workspace(batch, n, d, want_grad);
}
int sb200_gp_nll_grad(int batch, int n, int d, const double* theta, long stride_theta, const double* g,
                      long stride_g, const double* gvar, long stride_gvar, const double* hyp,
                      const double* regmean, const double* regstd, double sig2ofconst, int want_grad,
                      double* nll, double* grad, int* info, void* workspace, size_t workspace_bytes, void* stream) {
    SB_CHECK_ARG(theta && g && hyp && regmean && regstd && nll && info && workspace, "gp_nll_grad: NULL pointer");
    return gp_nll_grad(batch, n, d, theta, stride_theta, g, stride_g, gvar, stride_gvar, hyp, regmean, regstd,
                       sig2ofconst, want_grad, nll, grad, info, workspace, workspace_bytes, ST(stream));
}

void sb200_gp_force_blocked_path(int on) { gp_force_blocked_path(on); }
int sb200_gp_nll_small_max_n(void) { return gp_nll_small_max_n(); }

int sb200_spd_solve_small(int batch, int n, double* A, long lda, long strideA, double* logdet, int* info,
                          void* stream) {
    SB_CHECK_ARG(A && info, "spd_solve_small: NULL pointer");
    return spd_solve_small(batch, n, A, lda, strideA, logdet, info, ST(stream));
}

int sb200_gp_nll_small_profile(int batch, int n, int d, const double* theta, const double* g, const double* hyp,
                               const double* regmean, const double* regstd, double* nll, double* grad, int* info,
                               long long* stamps, void* workspace, size_t workspace_bytes, void* stream) {
    return gp_nll_grad_small(batch, n, d, theta, 0, g, 0, nullptr, 0, hyp, regmean, regstd, 0.01, 1, nll, grad, info,
                             workspace, workspace_bytes, ST(stream), stamps);
}

size_t sb200_gp_factorize_workspace(int batch, int n) { return gp_factorize_workspace(batch, n); }
int sb200_gp_factorize(int batch, int n, int d, const double* theta, long stride_theta, const double* g,
                       long stride_g, const double* gvar, long stride_gvar, const double* hyp, double sig2ofconst,
                       double* Linv, long ldl, long stride_linv, double* pw, double* sig2, double* logdet_half,
                       int* info, void* workspace, size_t workspace_bytes, void* stream) {
    SB_CHECK_ARG(theta && g && hyp && Linv && pw && sig2 && info && workspace, "gp_factorize: NULL pointer");
    return gp_factorize(batch, n, d, theta, stride_theta, g, stride_g, gvar, stride_gvar, hyp, sig2ofconst, Linv, ldl,
                        stride_linv, pw, sig2, logdet_half, info, workspace, workspace_bytes, ST(stream));
}

int sb200_gp_weights(int k, int n, const double* Linv, long ldl, long stride_linv, const int* fidx, const double* g,
                     double sig2ofconst, double* pw, double* sig2, double* ytmp, void* stream) {
    SB_CHECK_ARG(Linv && g && pw && sig2 && ytmp, "gp_weights: NULL pointer");
    return gp_weights(k, n, Linv, ldl, stride_linv, fidx, g, sig2ofconst, pw, sig2, ytmp, ST(stream));
}

size_t sb200_gp_predict_workspace(int k, int nf, int nu, int n, int d, int B, int want_grad) {
    return gp_predict_workspace(k, nf, nu, n, d, B, want_grad);
}
int sb200_gp_predict(int k, int nf, int nu, int n, int d, int B, const double* theta, const double* thetanew,
                     const double* hypcov, const int* fu, const int* fidx, const double* nug, const double* sig2,
                     const double* Linv, long ldl, long stride_linv, const double* pw, int want_grad, long out_ld,
                     double* predvecs, double* predvars, double* predvecs_grad, double* predvars_grad,
                     void* workspace, size_t workspace_bytes, void* stream) {
    SB_CHECK_ARG(theta && thetanew && hypcov && fu && fidx && nug && sig2 && Linv && pw && predvecs && predvars &&
                     workspace, "gp_predict: NULL pointer");
    return gp_predict(k, nf, nu, n, d, B, theta, thetanew, hypcov, fu, fidx, nug, sig2, Linv, ldl, stride_linv, pw,
                      want_grad, out_ld, predvecs, predvars, predvecs_grad, predvars_grad, workspace, workspace_bytes,
                      ST(stream));
}

int sb200_pc_backproject(int m, int k, int B, int d, const double* phi, const double* offset,
                         const double* extravar, const double* predvecs, const double* predvars,
                         const double* predvecs_grad, const double* predvars_grad, double* mean, double* var,
                         double* covxhalf, double* mean_grad, double* covxhalf_grad, void* stream) {
    SB_CHECK_ARG(phi && offset && extravar && predvecs && predvars && mean && var, "pc_backproject: NULL pointer");
    return pc_backproject(m, k, B, d, phi, offset, extravar, predvecs, predvars, predvecs_grad, predvars_grad, mean,
                          var, covxhalf, mean_grad, covxhalf_grad, ST(stream));
}

int sb200_woodbury_loglik(int B, int k, int m, int d, const double* predvecs, const double* predvars,
                          const double* predvecs_grad, const double* predvars_grad, const double* phiT,
                          const double* resid0, const double* extravar, const double* sinv, const double* G,
                          int want_grad, int variant, double* ll, double* dll, int* fired, const int* row_valid,
                          int seg_k, long seg_stride, const int* seg_pos, void* stream) {
    SB_CHECK_ARG(predvecs && predvars && phiT && resid0 && extravar && sinv && G && ll && fired,
                 "woodbury_loglik: NULL pointer");
    return woodbury_loglik(B, k, m, d, predvecs, predvars, predvecs_grad, predvars_grad, phiT, resid0, extravar, sinv,
                           G, want_grad, variant, ll, dll, fired, row_valid, seg_k, seg_stride, seg_pos, nullptr, 0,
                           ST(stream));
}

size_t sb200_woodbury_workspace(int B, int k, int m) { return woodbury_workspace(B, k, m); }

int sb200_woodbury_loglik_ws(int B, int k, int m, int d, const double* predvecs, const double* predvars,
                             const double* predvecs_grad, const double* predvars_grad, const double* phiT,
                             const double* resid0, const double* extravar, const double* sinv, const double* G,
                             int want_grad, int variant, double* ll, double* dll, int* fired, const int* row_valid,
                             int seg_k, long seg_stride, const int* seg_pos, void* workspace, size_t workspace_bytes,
                             void* stream) {
    SB_CHECK_ARG(predvecs && predvars && phiT && resid0 && extravar && sinv && G && ll && fired,
                 "woodbury_loglik: NULL pointer");
    return woodbury_loglik(B, k, m, d, predvecs, predvars, predvecs_grad, predvars_grad, phiT, resid0, extravar, sinv,
                           G, want_grad, variant, ll, dll, fired, row_valid, seg_k, seg_stride, seg_pos, workspace,
                           workspace_bytes, ST(stream));
}

void sb200_woodbury_force_global(int on) { woodbury_force_global(on); }

int sb200_woodbury_loglik_mspace(int B, int k, int m, const double* mean, const double* var, const double* covxhalf,
                                 const double* y, const double* sinv, int want_trigger, double* ll, int* fired,
                                 void* stream) {
    return woodbury_loglik_mspace(B, k, m, mean, var, covxhalf, y, sinv, want_trigger, ll, fired, ST(stream));
}

int sb200_woodbury_trigger(int B, int k, int m, const double* predvars, const double* phiT, const double* extravar,
                           int* fired, const int* row_valid, int seg_k, long seg_stride, const int* seg_pos, void* stream) {
    SB_CHECK_ARG(predvars && phiT && extravar && fired, "woodbury_trigger: NULL pointer");
    return woodbury_trigger(B, k, m, predvars, phiT, extravar, fired, row_valid, seg_k, seg_stride, seg_pos, ST(stream));
}

int sb200_ptlmc_step(int B, int d, int numtemps, int tune, int nsave, int mode, unsigned long long seed, double taracc,
                     double* thetac, double* fval, double* dfval, double* thetap, double* z, double* adjrho,
                     const double* ll, const double* dll, const double* temps, const double* lo, const double* hi,
                     const double* hc, const double* cov0, double* scal, long long* kstep, double* save,
                     int* row_valid, void* stream) {
    SB_CHECK_ARG(thetac && fval && dfval && thetap && z && adjrho && temps && lo && hi && hc && cov0 && scal && kstep &&
                     row_valid, "ptlmc_step: NULL pointer");
    SB_CHECK_ARG(!(mode & 1) || (ll && dll), "ptlmc_step: ll / dll are NULL");
    SB_CHECK_ARG(nsave == 0 || save, "ptlmc_step: save is NULL");
    return ptlmc_step(B, d, numtemps, tune, nsave, mode, seed, taracc, thetac, fval, dfval, thetap, z, adjrho, ll, dll,
                      temps, lo, hi, hc, cov0, scal, kstep, save, row_valid, ST(stream));
}

int sb200_dgemm(int a_kc, int b_kc, int M, int N, int K, double alpha, const double* A, long lda, long stride_a,
                const double* B, long ldb, long stride_b, double beta, double* C, long ldc, long stride_c,
                int batch, int flags, int tile, void* stream) {
    SB_CHECK_ARG(A && B && C, "dgemm: NULL pointer");
    SB_CHECK_ARG(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "dgemm: bad shape");
    GemmArgs g{};
    g.M = M; g.N = N; g.K = K;
    g.A = A; g.lda = lda; g.strideA = stride_a; g.a_kc = a_kc != 0;
    g.B = B; g.ldb = ldb; g.strideB = stride_b; g.b_kc = b_kc != 0;
    g.C = C; g.ldc = ldc; g.strideC = stride_c;
    g.alpha = alpha; g.beta = beta; g.flags = flags;
    return launch_dgemm(g, batch, tile, ST(stream));
}

int sb200_potrf_batched(int batch, int n, int mrows, double* A, long lda, long stride_a, double* Linv, long ldl,
                        long stride_linv, int* info, void* stream) {
    SB_CHECK_ARG(A && Linv && info, "potrf: NULL pointer");
    SB_CHECK_ARG(ldl >= n && stride_linv >= (long)n * ldl, "potrf: bad Linv layout");
    SB_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * 2 * batch, ST(stream)));      // [info | panel counters | task graph]
    SB_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * ((size_t)(batch - 1) * stride_linv + (size_t)n * ldl), ST(stream)));
    CholBatch c{};
    c.batch = batch; c.n = n; c.mrows = mrows; c.A = A; c.lda = lda; c.strideA = stride_a; c.info = info;
    c.counters = info + batch;
    c.dag_scratch = info + 2 * batch;
    c.Dinv = Linv; c.ldd = ldl; c.strideD = stride_linv; c.blockD = (long)CHOL_NB * ldl + CHOL_NB;
    return potrf_batched(c, ST(stream));
}

size_t sb200_potrf_info_ints(int batch, int mrows) { return 2 * (size_t)batch + potrf_dag_scratch_ints(batch, mrows); }
size_t sb200_trtri_tmp_elems(int n) { return trtri_tmp_elems(n); }
void sb200_debug_panel_stamps(long long* device_ptr) { debug_panel_stamps(device_ptr); }
void sb200_debug_dag_stats(long long* device_ptr) { debug_dag_stats(device_ptr); }

int sb200_trtri_batched(int batch, int n, const double* L, long lda, long stride_a, double* Linv, long ldl,
                        long stride_linv, double* tmp, size_t tmp_elems_per_matrix, void* stream) {
    SB_CHECK_ARG(L && Linv, "trtri: NULL pointer");
    return trtri_batched(batch, n, L, lda, stride_a, Linv, ldl, stride_linv, tmp, (long)tmp_elems_per_matrix, ST(stream));
}

int sb200_lauum_batched(int batch, int n, const double* Linv, long ldl, long stride_linv, double* C, long ldc,
                        long stride_c, void* stream) {
    SB_CHECK_ARG(Linv && C, "lauum: NULL pointer");
    return lauum_batched(batch, n, Linv, ldl, stride_linv, C, ldc, stride_c, ST(stream));
}

size_t sb200_pca_col_stats_workspace(int n, int m) { return pca_col_stats_workspace(n, m); }

int sb200_pca_col_stats(int n, int m, const double* f, long ldf, double* offset, double* scale, int* nmiss,
                        void* workspace, size_t workspace_bytes, void* stream) {
    SB_CHECK_ARG(f && offset && scale && nmiss && workspace, "pca_col_stats: NULL pointer");
    return pca_col_stats(n, m, f, ldf, offset, scale, nmiss, workspace, workspace_bytes, ST(stream));
}

int sb200_pca_standardize(int n, int m, const double* f, long ldf, const double* offset, const double* scale,
                          double* fs, long ldfs, unsigned char* miss, const int* nmiss, void* stream) {
    SB_CHECK_ARG(f && offset && scale && fs, "pca_standardize: NULL pointer");
    return pca_standardize(n, m, f, ldf, offset, scale, fs, ldfs, miss, nmiss, ST(stream));
}

int sb200_pca_sumsq(long count, const double* x, double* out, double* scratch, void* stream) {
    SB_CHECK_ARG(x && out && scratch, "pca_sumsq: NULL pointer");
    return pca_sumsq(count, x, out, scratch, ST(stream));
}

int sb200_syevj_max_p(void) { return SYEVJ_MAXP; }

int sb200_syevj_batched(int batch, int p, const double* H, long ldh, long stride_h, double* W, long stride_w,
                        double* Z, long ldz, long stride_z, int max_sweeps, int* sweeps, void* stream) {
    SB_CHECK_ARG(H && W && Z, "syevj: NULL pointer");
    return syevj_batched(batch, p, H, ldh, stride_h, W, stride_w, Z, ldz, stride_z, max_sweeps, sweeps, ST(stream));
}

int sb200_peer_allgather_post(const double* block, long count, const unsigned long long* peer_bufs, long slot_offset,
                              long flag_offset, int rank, int world, const unsigned long long* epoch, int* arrive,
                              void* stream) {
    SB_CHECK_ARG(block && peer_bufs && epoch && arrive, "peer_allgather_post: NULL pointer");
    return peer_allgather_post(block, count, peer_bufs, slot_offset, flag_offset, rank, world, epoch, arrive, ST(stream));
}

int sb200_peer_allgather_wait(const unsigned long long* my_flags, int rank, int world, unsigned long long* epoch, int* error,
                              void* stream) {
    SB_CHECK_ARG(my_flags && epoch && error, "peer_allgather_wait: NULL pointer");
    return peer_allgather_wait(my_flags, rank, world, epoch, error, ST(stream));
}

}  // extern "C"
