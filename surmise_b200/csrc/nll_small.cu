// NLL (+ gradient) for the optimiser's regime: n <= 207 points (the reference trains hyper-parameters on
// min(20 d, n) points, PCGPwM.py:744).  Three launches instead of the blocked path's seventeen:
//   gp_assemble_kernel        (covmat.cu)  lower triangle of R and g' as row n
//   gp_sweep_small_kernel<R>  (here)       one CTA per problem, matrix resident in REGISTERS
//   gp_grad_contract_kernel   (covmat.cu)  sum_ab W_ab dR_j(a,b), only when the gradient is wanted
//
// Sweep operator on the augmented matrix M = [[R, g], [g', 0]]: sweeping the n pivots of R in order leaves
// M_RR = -R^-1, M_gR = (R^-1 g)' = alpha', M_gg = -g' R^-1 g, and the pivots d_k are the LDL' pivots of R, so
// log|R| = sum log d_k -- everything PCGPwM.__negloglik / __negloglikgrad (PCGPwM.py:850-907) need, n^3/2 FMA.
// 16 x 16 threads; thread (ti, tc) owns elements (ti + 16 ii, tc + 16 cc), cc <= ii, of the lower triangle.
// Each pivot step is a register-resident rank-1 update; the pivot column travels through shared memory.
// The loop is software-pipelined: inside step k the elements that form column k+1 are updated FIRST, published
// (double buffer) together with 1/d_{k+1} computed by the owner thread, and only then the bulk of the update
// runs -- so the barrier, the shared-memory round trip and the reciprocal of step k+1 hide behind FP64 work.
#include "gp.cuh"
#include "covmat.cuh"

namespace sb {
namespace {

constexpr int T16 = 16;

struct SweepArgs {
    int n, d;
    double* A; long lda, strideA;          // in: lower(R) + g row;  out: lower(R^-1) + alpha row
    const double* hyp; const double* regmean; const double* regstd;
    double s0;
    double* nll; double* sig2hat; int* info;
    long long* stamps;
};

template <int R, int KB, int KB1>
__device__ __forceinline__ void sweep_step(double (&a)[R][R], const int kr, const int k, const int n, const int ti,
                                           const int tc, const double* __restrict__ vcur, double* __restrict__ vnext,
                                           double* __restrict__ dinv_s, double* __restrict__ piv, int& bad_at) {
    const double dinv = dinv_s[k & 1];
    double wi[R], vc[R];
#pragma unroll
    for (int ii = 0; ii < R; ii++) wi[ii] = vcur[ti + T16 * ii] * dinv;
#pragma unroll
    for (int cc = 0; cc < R; cc++) vc[cc] = vcur[tc + T16 * cc];
    // The pivot row and column need  a_kc <- a_kc / d  and  a_ik <- a_ik / d  instead of the rank-1 update.  With
    //   w_k := 1 - 1/d   (row k)   and   v_k := d - 1   (column k)
    // the SAME fused multiply-add produces them (a_kc - (1 - 1/d) a_kc = a_kc / d;  a_ik - (a_ik / d)(d - 1) = a_ik / d,
    // relative perturbation <= eps * d), so no per-element selects are needed; only (k,k) <- -1/d is set explicitly.
    const double dk = vcur[k];
    if (ti == kr) wi[KB] = 1.0 - dinv;
    if (tc == kr) vc[KB] = dk - 1.0;
    const int k1 = k + 1, kr1 = k1 & 15;

    // ---- (a) the elements that make up column / row k+1: block column KB1 and block row KB1
#pragma unroll
    for (int ii = 0; ii < R; ii++)
        if (ii >= KB1) a[ii][KB1] -= wi[ii] * vc[KB1];
#pragma unroll
    for (int cc = 0; cc < R; cc++)
        if (cc < KB1) a[KB1][cc] -= wi[KB1] * vc[cc];
    if (KB == KB1 && ti == kr && tc == kr) a[KB][KB] = -dinv;
    // ---- (b) publish column k+1 and 1/d_{k+1} (reciprocal computed branch-free by every thread on its own (KB1,KB1)
    // element so that its dependent chain interleaves with the FMAs of (c); only the owner's value is stored)
    double d1 = a[KB1][KB1];
    const bool okp = d1 > 0.0;
    d1 = okp ? d1 : 1.0;
    const double r1 = __drcp_rn(d1);
    if (k1 < n) {
        if (tc == kr1) {
#pragma unroll
            for (int ii = 0; ii < R; ii++)
                if (ii >= KB1 && (ii > KB1 || ti >= kr1)) vnext[ti + T16 * ii] = a[ii][KB1];
        }
        if (ti == kr1) {
#pragma unroll
            for (int cc = 0; cc < R; cc++)
                if (cc <= KB1 && (cc < KB1 || tc <= kr1)) vnext[tc + T16 * cc] = a[KB1][cc];
        }
    }
    // ---- (c) the rest of the rank-1 update
#pragma unroll
    for (int ii = 0; ii < R; ii++)
#pragma unroll
        for (int cc = 0; cc < R; cc++)
            if (cc <= ii && cc != KB1 && ii != KB1) a[ii][cc] -= wi[ii] * vc[cc];
    if (KB != KB1 && ti == kr && tc == kr) a[KB][KB] = -dinv;
    if (k1 < n && ti == kr1 && tc == kr1) {
        piv[k1] = d1;
        dinv_s[k1 & 1] = r1;
        if (!okp && bad_at == 0) bad_at = k1 + 1;
    }
    __syncthreads();
}

template <int R, int KB>
__device__ __forceinline__ void sweep_block(double (&a)[R][R], const int n, const int ti, const int tc, double* vb,
                                            double* dinv_s, double* piv, int& bad_at) {
    constexpr int NP = T16 * R;
#pragma unroll 1
    for (int kr = 0; kr < T16 - 1; kr++) {
        const int k = KB * T16 + kr;
        if (k >= n) return;
        sweep_step<R, KB, KB>(a, kr, k, n, ti, tc, vb + (k & 1) * NP, vb + ((k + 1) & 1) * NP, dinv_s, piv, bad_at);
    }
    const int k = KB * T16 + T16 - 1;
    if (k >= n) return;
    if constexpr (KB + 1 < R) {
        sweep_step<R, KB, KB + 1>(a, T16 - 1, k, n, ti, tc, vb + (k & 1) * NP, vb + ((k + 1) & 1) * NP, dinv_s, piv,
                                  bad_at);
        sweep_block<R, KB + 1>(a, n, ti, tc, vb, dinv_s, piv, bad_at);
    }
}

template <int R>
__global__ void __launch_bounds__(256, 1) gp_sweep_small_kernel(SweepArgs p) {
    constexpr int NP = T16 * R;                       // padded order (>= n + 1)
    __shared__ double vbuf[2][NP];
    __shared__ double piv[NP];
    __shared__ double dinv_s[2];
    __shared__ double red[32];
    __shared__ int s_bad;
    const int z = blockIdx.x, tid = threadIdx.x, n = p.n, d = p.d;
    const int ti = tid >> 4, tc = tid & 15;
    double* A = p.A + (long)z * p.strideA;
#define SB_STAMP(slot) do { if (p.stamps && z == 0 && tid == 0) p.stamps[slot] = clock64(); } while (0)
    SB_STAMP(0);

    // ---- registers <- lower triangle (+ g row, + identity padding)
    double a[R][R];
#pragma unroll
    for (int ii = 0; ii < R; ii++)
#pragma unroll
        for (int cc = 0; cc < R; cc++)
            if (cc <= ii) {
                int i = ti + T16 * ii, c = tc + T16 * cc;
                double v = 0.0;
                if (c <= i) {
                    if (i <= n && c < n) v = A[(long)i * p.lda + c];
                    else if (i > n && i == c) v = 1.0;
                }
                a[ii][cc] = v;
            }
    if (tid == 0) s_bad = 0x7fffffff;
    int bad_at = 0;
    // prologue: publish column 0 and 1 / d_0
    if (tc == 0) {
#pragma unroll
        for (int ii = 0; ii < R; ii++) vbuf[0][ti + T16 * ii] = a[ii][0];
        if (ti == 0) {
            double d0 = a[0][0];
            if (!(d0 > 0.0)) {
                bad_at = 1;
                d0 = 1.0;
            }
            piv[0] = d0;
            dinv_s[0] = __drcp_rn(d0);
        }
    }
    __syncthreads();
    SB_STAMP(1);

    // ---- sweep the n pivots of R (block index of the pivot is a template parameter: it selects registers)
    sweep_block<R, 0>(a, n, ti, tc, &vbuf[0][0], dinv_s, piv, bad_at);
    SB_STAMP(2);
    if (bad_at != 0) atomicMin(&s_bad, bad_at);        // first non-positive pivot (LAPACK info convention)

    // ---- write back  R^-1 (lower) | alpha' ;  g' R^-1 g from the (n, n) element
    __shared__ double s_qf;
#pragma unroll
    for (int ii = 0; ii < R; ii++)
#pragma unroll
        for (int cc = 0; cc < R; cc++)
            if (cc <= ii) {
                int i = ti + T16 * ii, c = tc + T16 * cc;
                if (c <= i) {
                    if (i < n) A[(long)i * p.lda + c] = -a[ii][cc];
                    else if (i == n && c < n) A[(long)n * p.lda + c] = a[ii][cc];
                    else if (i == n && c == n) s_qf = -a[ii][cc];
                }
            }
    __syncthreads();
    double ld = 0.0;                                               // sum log d_k = log|R|
    for (int k = tid; k < n; k += 256) ld += log(piv[k]);
    const double logdet = block_sum(ld, red);
    if (tid == 0 && p.hyp == nullptr) {                            // plain SPD solve (spd_solve_small): no GP terms
        if (p.nll) p.nll[z] = logdet;
        p.info[z] = (s_bad == 0x7fffffff) ? 0 : s_bad;
    } else if (tid == 0) {
        const double sig2hat = (s_qf + p.s0) / (n + p.s0);         // PCGPwM.py:863
        const int np = d + 3;
        const double* hyp = p.hyp + (long)z * np;
        double pr = 0.0;
        for (int k = 0; k < np; k++) {
            double t = (1e-8 + hyp[k] - p.regmean[k]) / p.regstd[k];
            pr += t * t;
        }
        p.nll[z] = 0.5 * logdet + 0.5 * n * log(sig2hat) + 0.5 * pr;   // PCGPwM.py:864-867
        p.sig2hat[z] = sig2hat;
        p.info[z] = (s_bad == 0x7fffffff) ? 0 : s_bad;
    }
    SB_STAMP(3);
#undef SB_STAMP
}


template <int R>
int launch_sweep(const SweepArgs& a, int batch, cudaStream_t stream) {
    gp_sweep_small_kernel<R><<<batch, 256, 0, stream>>>(a);
    SB_LAUNCH_CHECK();
    return 0;
}

long lda_small(int n) { return round_up(n, 8); }

}  // namespace

int gp_nll_small_max_n() { return T16 * 13 - 1; }

static int launch_sweep_for(const SweepArgs& a, int n, int batch, cudaStream_t stream) {
    const int R = (n + 1 + T16 - 1) / T16;
    if (R <= 4) return launch_sweep<4>(a, batch, stream);
    if (R <= 6) return launch_sweep<6>(a, batch, stream);
    if (R <= 8) return launch_sweep<8>(a, batch, stream);
    if (R <= 10) return launch_sweep<10>(a, batch, stream);
    if (R <= 11) return launch_sweep<11>(a, batch, stream);
    return launch_sweep<13>(a, batch, stream);
}

// Batched SPD solve with the sweep kernel: A (batch, n+1, lda) holds lower(S) and the right-hand side as row n;
// on return lower(S^-1) and (S^-1 z)' in row n.  Used by the device imputation (Schur systems of the missing entries).
int spd_solve_small(int batch, int n, double* A, long lda, long strideA, double* logdet, int* info,
                    cudaStream_t stream) {
    SB_CHECK_ARG(batch > 0, "spd_solve_small: empty batch");
    SB_CHECK_ARG(n >= 1 && n <= gp_nll_small_max_n(), "spd_solve_small: n=%d outside [1,%d]", n, gp_nll_small_max_n());
    SB_CHECK_ARG(lda >= n && strideA >= (long)(n + 1) * lda, "spd_solve_small: bad layout");
    SB_CHECK_ARG(info != nullptr, "spd_solve_small: info is NULL");
    SweepArgs a{n, 0, A, lda, strideA, nullptr, nullptr, nullptr, 0.0, logdet, nullptr, info, nullptr};
    return launch_sweep_for(a, n, batch, stream);
}

size_t gp_nll_small_workspace(int batch, int n, int want_grad) {
    Workspace w(nullptr, 0);
    w.take<double>((size_t)batch * (n + 1) * lda_small(n));
    w.take<double>((size_t)batch);
    if (want_grad) w.take<double>((size_t)batch * grad_contract_tiles(n) * MAXP);
    return w.off;
}

int gp_nll_grad_small(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                      const double* gvar, long strideGvar, const double* hyp, const double* regmean,
                      const double* regstd, double s0, int want_grad, double* nll, double* grad, int* info,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream, long long* stamps) {
    SB_CHECK_ARG(n >= 1 && n <= gp_nll_small_max_n(), "gp_nll_small: n=%d outside [1,%d]", n, gp_nll_small_max_n());
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "gp_nll_small: d=%d outside [1,%d]", d, MAXD);
    SB_CHECK_ARG(workspace_bytes >= gp_nll_small_workspace(batch, n, want_grad), "gp_nll_small: workspace too small");
    Workspace w(workspace, workspace_bytes);
    const long lda = lda_small(n), strideA = (long)(n + 1) * lda;
    double* A = w.take<double>((size_t)batch * strideA);
    double* sig2hat = w.take<double>((size_t)batch);
    SB_TRY(gp_assemble(batch, n, d, theta, strideTheta, g, strideG, gvar, strideGvar, hyp, A, lda, strideA, stream));
    SweepArgs a{n, d, A, lda, strideA, hyp, regmean, regstd, s0, nll, sig2hat, info, stamps};
    SB_TRY(launch_sweep_for(a, n, batch, stream));
    if (!want_grad) return 0;
    double* partial = w.take<double>((size_t)batch * grad_contract_tiles(n) * MAXP);
    return gp_grad_contract(batch, n, d, theta, strideTheta, gvar, strideGvar, hyp, regmean, regstd, A, lda, strideA,
                            A + (long)n * lda, strideA, sig2hat, s0, partial, grad, stream);
}

}  // namespace sb
