// GP-level operations of the PCGPwM hot path, orchestrated over the kernels in covmat.cu / chol.cu /
// dgemm.cu.  Everything is batched over "problems" (principal components or candidate
// hyper-parameter sets) that share n and d.
//
//   gp_nll_grad   : PCGPwM.__negloglik + __negloglikgrad in one pass        (PCGPwM.py:850-907)
//   gp_factorize  : full-n stage of __fitGP1d -> Linv, pw, sig2             (PCGPwM.py:810-846)
//   gp_predict    : PC-space predictive mean/variance (+ d/dtheta)          (PCGPwM.py:219-270)
//   pc_backproject: mean/var/covxhalf (+ gradients) through Phi             (PCGPwM.py:273-327)
#include "gp.cuh"
#include <cstdlib>
#include <mutex>
#include "chol.cuh"
#include "covmat.cuh"
#include "dgemm.cuh"

namespace sb {
namespace {

constexpr int NB = CHOL_NB;

// nll = sum log L_ii + n/2 log sig2hat + 1/2 sum ((1e-8 + hyp - mu)/s)^2 ;  sig2hat = (|y|^2 + s0)/(n + s0)
__global__ void __launch_bounds__(256) nll_finalize_kernel(int n, int d, const double* __restrict__ Ab, long lda,
                                                           long strideA, const double* __restrict__ hypb,
                                                           const double* __restrict__ regmean,
                                                           const double* __restrict__ regstd, double s0,
                                                           double* __restrict__ nll, double* __restrict__ sig2hat,
                                                           double* __restrict__ logdet_half) {
    __shared__ double red[32];
    const int z = blockIdx.x;
    const double* A = Ab + (long)z * strideA;
    double ld = 0.0, yy = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        ld += log(A[(long)i * lda + i]);
        double y = A[(long)n * lda + i];
        yy += y * y;
    }
    ld = block_sum(ld, red);
    yy = block_sum(yy, red);
    if (threadIdx.x == 0) {
        double s2 = (yy + s0) / (n + s0);
        if (sig2hat) sig2hat[z] = s2;
        if (logdet_half) logdet_half[z] = ld;
        if (nll) {
            double pr = 0.0;
            const int p = d + 3;
            for (int k = 0; k < p; k++) {
                double t = (1e-8 + hypb[(long)z * p + k] - regmean[k]) / regstd[k];
                pr += t * t;
            }
            nll[z] = ld + 0.5 * n * log(s2) + 0.5 * pr;
        }
    }
}

struct CovConsts {
    double ell[MAXD], ginv[MAXD];
    double onemc;
};

constexpr int JC = 256;          // design points per CTA in the predict epilogue

// members of every factor: mem_off (nf + 1) and mem_list (K), components in increasing order within a factor
__global__ void __launch_bounds__(256) factor_members_kernel(int K, int nf, const int* __restrict__ fidx,
                                                             int* __restrict__ mem_off, int* __restrict__ mem_list) {
    for (int f = threadIdx.x; f < nf; f += blockDim.x) {
        int c = 0;
        for (int k = 0; k < K; k++) c += (fidx[k] == f);
        mem_off[f + 1] = c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mem_off[0] = 0;
        for (int f = 0; f < nf; f++) mem_off[f + 1] += mem_off[f];
    }
    __syncthreads();
    for (int f = threadIdx.x; f < nf; f += blockDim.x) {
        int o = mem_off[f];
        for (int k = 0; k < K; k++)
            if (fidx[k] == f) mem_list[o++] = k;
    }
}

// Per (FACTOR, 32 new thetas, 256 design points): partial sums over j of T1^2, r pw and (with gradients) T2 dr_i,
// dr_i pw, where dr_i = dR0/dtheta*_i is rebuilt from r -- with everything that depends only on the
// factor computed once for all the components that share it -- T1^2, the derivative dr_i of the cross-covariance
// (the expensive part: a reciprocal per dimension) and T2 dr_i -- and only the two products with the component's own
// weights pw_k inside the loop over its members.  At C2 (20 components on 11-14 factors) this is 0.24 -> 0.14 ms of
// the 1.5 ms a B = 264 evaluation takes.  Members are handled MP at a time (one pass for nearly every factor).
template <bool WANT_GRAD, int DT, int MP>
__global__ void __launch_bounds__(256, 2) predict_partial_factor_kernel(
    int n, int d, int B, int K, const double* __restrict__ theta, const double* __restrict__ thetanew,
    const double* __restrict__ hypcovb, const int* __restrict__ mem_off, const int* __restrict__ mem_list,
    const int* __restrict__ fu, const double* __restrict__ pwb, const double* __restrict__ rTb, long ldb, long strideU,
    const double* __restrict__ T1b, const double* __restrict__ T2b, long strideT, double* __restrict__ partial,
    int nchunks, int Q) {
    constexpr bool want_grad = WANT_GRAD;
    constexpr int XL = DT + 1;                                         // row stride of the staged points (d <= DT)
    __shared__ double xj[JC * XL];
    __shared__ double xb[32 * XL];
    __shared__ double pwj[MP][JC];
    __shared__ double red[8][33];
    __shared__ CovConsts cc;
    const int f = blockIdx.z, chunk = blockIdx.y, b0 = blockIdx.x * 32;
    const int u = fu[f];
    const double* hc = hypcovb + (long)u * (d + 1);
    if (threadIdx.x < DT) {                                            // dimensions d..DT-1 are padding: ginv = 0
        cc.ell[threadIdx.x] = threadIdx.x < d ? exp(hc[threadIdx.x]) : 1.0;
        cc.ginv[threadIdx.x] = threadIdx.x < d ? exp(-hc[threadIdx.x]) : 0.0;
    }
    if (threadIdx.x == 32) {
        double e = exp(hc[d]);
        cc.onemc = e / (1.0 + e);
    }
    const int* s_mem = mem_list + mem_off[f];                          // the components that use this factor, in order
    const int cnt = mem_off[f + 1] - mem_off[f];
    if (cnt == 0) return;
    __syncthreads();
    const int j0 = chunk * JC;
    if (want_grad) {                                                   // the scaled points are only needed for dr_i
        // zero-padded up to DT dimensions: the derivative loop below then runs unguarded over all DT of them (a padded
        // dimension has s = 0 and ginv = 0, it adds exactly zero); guarded by `i < d` every dimension became a branch
        // region of its own and the DT reciprocal chains ran one after the other
        for (int idx = threadIdx.x; idx < JC * DT; idx += blockDim.x) {
            int r = idx / DT, c = idx % DT;
            xj[r * XL + c] = (j0 + r < n && c < d) ? theta[(long)(j0 + r) * d + c] / cc.ell[c] : 0.0;
        }
        for (int idx = threadIdx.x; idx < 32 * DT; idx += blockDim.x) {
            int r = idx / DT, c = idx % DT;
            xb[r * XL + c] = (b0 + r < B && c < d) ? thetanew[(long)(b0 + r) * d + c] / cc.ell[c] : 0.0;
        }
    }
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int b = b0 + tx;
    const double* rT = rTb + (long)u * strideU;
    const double* T1 = T1b + (long)f * strideT;
    const double* T2 = T2b ? T2b + (long)f * strideT : nullptr;
    const int jmax = min(JC, n - j0);

    // fixed-order reduction over the 8 j-lanes of one quantity, written for component k as quantity q
    auto emit = [&](double v, int k, int q) {
        __syncthreads();
        red[ty][tx] = v;
        __syncthreads();
        if (ty == 0 && b < B) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 8; r++) s += red[r][tx];
            partial[((((long)k * nchunks + chunk) * Q) + q) * ldb + b] = s;
        }
    };

    for (int p0 = 0; p0 < cnt; p0 += MP) {
        const int np = min(MP, cnt - p0);
        __syncthreads();                                               // the previous pass is done with pwj
        for (int idx = threadIdx.x; idx < MP * JC; idx += blockDim.x) {
            const int mm = idx / JC, jl = idx % JC;
            pwj[mm][jl] = (mm < np && j0 + jl < n) ? pwb[(long)s_mem[p0 + mm] * n + j0 + jl] : 0.0;
        }
        __syncthreads();
        double s1 = 0.0, dv[DT], s2[MP], dp[MP][DT];
#pragma unroll
        for (int i = 0; i < DT; i++) dv[i] = 0.0;
#pragma unroll
        for (int mm = 0; mm < MP; mm++) {
            s2[mm] = 0.0;
#pragma unroll
            for (int i = 0; i < DT; i++) dp[mm][i] = 0.0;
        }
        if (b < B) {
            // the three global loads of PF design points go out together, ahead of the arithmetic of the first of them:
            // with one or two points per trip the loop waited a full L2 round trip per trip (ncu: 22 % of the samples
            // on the first use of t1)
            constexpr int PF = 4;
#pragma unroll 1
            for (int jb = ty; jb < jmax; jb += 8 * PF) {
                double rr[PF], tt1[PF], tt2[PF];
#pragma unroll
                for (int u2 = 0; u2 < PF; u2++) {
                    const int jl = jb + 8 * u2;
                    const bool ok = jl < jmax;
                    const long off = (long)(j0 + (ok ? jl : jb)) * ldb + b;
                    rr[u2] = ok ? rT[off] : 0.0;
                    tt1[u2] = ok ? T1[off] : 0.0;
                    tt2[u2] = (want_grad && ok) ? T2[off] : 0.0;
                }
#pragma unroll
                for (int u2 = 0; u2 < PF; u2++) {
                    const int jl = min(jb + 8 * u2, JC - 1);             // (beyond jmax: r = t1 = t2 = 0 contribute nothing)
                    const bool ok = jb + 8 * u2 < jmax;
                    const double r = rr[u2], t1 = tt1[u2], t2 = tt2[u2];
                    double pwv[MP];
#pragma unroll
                    for (int mm = 0; mm < MP; mm++) pwv[mm] = ok ? pwj[mm][jl] : 0.0;
                    s1 += t1 * t1;
#pragma unroll
                    for (int mm = 0; mm < MP; mm++) s2[mm] += r * pwv[mm];
                    if (want_grad && ok) {
                        const double rp = r - cc.onemc;                // R_pre
#pragma unroll
                        for (int i = 0; i < DT; i++) {
                            const double s = xb[tx * XL + i] - xj[jl * XL + i];
                            const double dr = rp * (-cc.ginv[i] * s * rcp_ge1(1.0 + fabs(s)));       // matern_covmat.pyx:13-24
                            dv[i] += t2 * dr;
#pragma unroll
                            for (int mm = 0; mm < MP; mm++) dp[mm][i] += dr * pwv[mm];
                        }
                    }
                }
            }
        }
        // the factor-level sums (T1^2, T2 dr_i) are the same for every member: each gets its copy
#pragma unroll
        for (int mm = 0; mm < MP; mm++) {
            if (mm >= np) break;
            const int k = s_mem[p0 + mm];
            emit(s2[mm], k, 1);
            if (want_grad) {
#pragma unroll
                for (int i = 0; i < DT; i++)
                    if (i < d) emit(dp[mm][i], k, 3 + 2 * i);
            }
        }
        {
            __syncthreads();
            red[ty][tx] = s1;
            __syncthreads();
            if (ty == 0 && b < B) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < 8; r++) s += red[r][tx];
                for (int mm = 0; mm < np; mm++)
                    partial[((((long)s_mem[p0 + mm] * nchunks + chunk) * Q) + 0) * ldb + b] = s;
            }
        }
        if (want_grad) {
#pragma unroll
            for (int i = 0; i < DT; i++)
                if (i < d) {
                    __syncthreads();
                    red[ty][tx] = dv[i];
                    __syncthreads();
                    if (ty == 0 && b < B) {
                        double s = 0.0;
#pragma unroll
                        for (int r = 0; r < 8; r++) s += red[r][tx];
                        for (int mm = 0; mm < np; mm++)
                            partial[((((long)s_mem[p0 + mm] * nchunks + chunk) * Q) + 2 + 2 * i) * ldb + b] = s;
                    }
                }
        }
    }
}

__global__ void predict_reduce_kernel(int K, int d, int B, long ldb, int nchunks, int Q, const double* __restrict__ partial,
                                      const double* __restrict__ nug, const double* __restrict__ sig2, int want_grad,
                                      int single_quirk, long ldk, double* __restrict__ predvecs,
                                      double* __restrict__ predvars, double* __restrict__ dpv, double* __restrict__ dpvar) {
    // outputs are (B, ldk[, d]) with ldk >= K: a caller that assembles the components of several ranks side by side
    // (PC-sharded prediction) hands in a wider row than this launch fills
    const int b = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (b >= B) return;
    const double omn = 1.0 - nug[k], s2k = sig2[k];
    for (int q = 0; q < Q; q++) {
        double s = 0.0;
        for (int c = 0; c < nchunks; c++) s += partial[((((long)k * nchunks + c) * Q) + q) * ldb + b];
        if (q == 0) predvars[(long)b * ldk + k] = s2k * fabs(1.0 - omn * omn * s);           // PCGPwM.py:270
        else if (q == 1) predvecs[(long)b * ldk + k] = omn * s;                               // PCGPwM.py:246
        else if (want_grad) {
            int i = (q - 2) >> 1;
            if ((q - 2) & 1) {
                // the reference applies (1-nug) twice here (PCGPwM.py:236,268) except when B == 1 and d == 1
                double extra = single_quirk ? 1.0 : omn;
                dpv[((long)b * ldk + k) * d + i] = extra * omn * s;
            } else {
                dpvar[((long)b * ldk + k) * d + i] = -2.0 * s2k * omn * omn * s;              // PCGPwM.py:269
            }
        }
    }
}

__global__ void backproject_meanvar_kernel(int m, int K, int B, const double* __restrict__ phi,
                                           const double* __restrict__ offset, const double* __restrict__ extravar,
                                           const double* __restrict__ pv, const double* __restrict__ pvar,
                                           double* __restrict__ mean, double* __restrict__ var) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x, x = blockIdx.y;
    if (b >= B) return;
    double sm = 0.0, sv = 0.0;
    for (int k = 0; k < K; k++) {
        double ph = phi[(long)x * K + k];
        sm += pv[(long)b * K + k] * ph;
        sv += pvar[(long)b * K + k] * ph * ph;
    }
    mean[(long)x * B + b] = sm + offset[x];                        // PCGPwM.py:277-278
    var[(long)x * B + b] = extravar[x] + sv;                       // PCGPwM.py:279-280
}

__global__ void backproject_covxhalf_kernel(int m, int K, int B, const double* __restrict__ phi,
                                            const double* __restrict__ pvar, double* __restrict__ cxh) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)m * B * K;
    if (idx >= total) return;
    int k = (int)(idx % K);
    long xb = idx / K;
    int b = (int)(xb % B), x = (int)(xb / B);
    cxh[idx] = sqrt(pvar[(long)b * K + k]) * phi[(long)x * K + k];  // PCGPwM.py:291
}

__global__ void backproject_meangrad_kernel(int m, int K, int B, int d, const double* __restrict__ phi,
                                            const double* __restrict__ dpv, double* __restrict__ mg) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)m * B * d;
    if (idx >= total) return;
    int i = (int)(idx % d);
    long xb = idx / d;
    int b = (int)(xb % B), x = (int)(xb / B);
    double s = 0.0;
    for (int k = 0; k < K; k++) s += dpv[((long)b * K + k) * d + i] * phi[(long)x * K + k];
    mg[idx] = s;                                                    // PCGPwM.py:302-304
}

__global__ void backproject_covxgrad_kernel(int m, int K, int B, int d, const double* __restrict__ phi,
                                            const double* __restrict__ pvar, const double* __restrict__ dpvar,
                                            double* __restrict__ cg) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long total = (long)m * B * K * d;
    if (idx >= total) return;
    int i = (int)(idx % d);
    long r = idx / d;
    int k = (int)(r % K);
    r /= K;
    int b = (int)(r % B), x = (int)(r / B);
    double ds = 0.5 * dpvar[((long)b * K + k) * d + i] / sqrt(pvar[(long)b * K + k]);      // PCGPwM.py:309-310
    cg[idx] = ds * phi[(long)x * K + k];
}

long lda_for(int n) { return round_up(n, 8); }

}  // namespace

// testing hook: route small problems through the blocked (multi-launch) path as well
// Per-device side stream (+ fork / join events) for work that overlaps the main stream inside one entry point; created
// on first use, part of the one-time per-device set-up (include/surmise_b200.h, conventions (b)).
struct SideLane {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    std::mutex mu;
};
static int side_lane(SideLane** out) {
    static SideLane lanes[64];
    static DeviceOnce once;
    static int enabled = -1;
    *out = nullptr;
    if (enabled < 0) {
        const char* env = getenv("SB200_TRMV_SIDE");
        enabled = (env && env[0] == '0') ? 0 : 1;
    }
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (!enabled || dev < 0 || dev >= 64) return 0;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaStreamCreateWithFlags(&lanes[dev].s, cudaStreamNonBlocking));
        SB_CUDA(cudaEventCreateWithFlags(&lanes[dev].fork, cudaEventDisableTiming));
        SB_CUDA(cudaEventCreateWithFlags(&lanes[dev].join, cudaEventDisableTiming));
        return 0;
    }));
    *out = &lanes[dev];
    return 0;
}

static std::atomic<bool> g_force_blocked{false};      // testing switch (sb200_gp_force_blocked_path)
void gp_force_blocked_path(int on) { g_force_blocked.store(on != 0); }

// ------------------------------------------------------------------------------------------------
size_t gp_nll_grad_workspace(int batch, int n, int d, int want_grad) {
    if (n <= gp_nll_small_max_n() && !g_force_blocked) return gp_nll_small_workspace(batch, n, want_grad);
    Workspace w(nullptr, 0);
    long lda = lda_for(n);
    w.take<double>((size_t)batch * (n + 1) * lda);
    w.take<double>((size_t)batch);                                  // sig2hat
    w.take<int>((size_t)batch);                                     // panel arrival counters
    w.take<int>(potrf_dag_scratch_ints(batch, n + 1));
    if (want_grad) {
        w.take<double>((size_t)batch * n * lda);                    // Linv
        w.take<double>((size_t)batch * trtri_tmp_elems(n));
        w.take<double>((size_t)batch * n);                          // alpha
        w.take<double>((size_t)batch * grad_contract_tiles(n) * MAXP);
    } else {
        w.take<double>((size_t)batch * cdiv(n, NB) * NB * NB);      // compact Dinv
    }
    return w.off;
}

int gp_nll_grad(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                const double* gvar, long strideGvar, const double* hyp, const double* regmean, const double* regstd,
                double s0, int want_grad, double* nll, double* grad, int* info, void* workspace,
                size_t workspace_bytes, cudaStream_t stream) {
    SB_CHECK_ARG(batch > 0 && n > 0, "gp_nll_grad: batch=%d n=%d", batch, n);
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "gp_nll_grad: d=%d outside [1,%d]", d, MAXD);
    SB_CHECK_ARG(workspace_bytes >= gp_nll_grad_workspace(batch, n, d, want_grad), "gp_nll_grad: workspace too small");
    SB_CHECK_ARG(!want_grad || grad != nullptr, "gp_nll_grad: grad is NULL");
    if (n <= gp_nll_small_max_n() && !g_force_blocked)
        return gp_nll_grad_small(batch, n, d, theta, strideTheta, g, strideG, gvar, strideGvar, hyp, regmean, regstd,
                                 s0, want_grad, nll, grad, info, workspace, workspace_bytes, stream);
    Workspace w(workspace, workspace_bytes);
    const long lda = lda_for(n);
    const long strideA = (long)(n + 1) * lda;
    double* A = w.take<double>((size_t)batch * strideA);
    double* sig2hat = w.take<double>((size_t)batch);
    int* counters = w.take<int>((size_t)batch);
    int* dag_scratch = w.take<int>(potrf_dag_scratch_ints(batch, n + 1));
    SB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * batch, stream));
    SB_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, stream));
    SB_TRY(gp_assemble(batch, n, d, theta, strideTheta, g, strideG, gvar, strideGvar, hyp, A, lda, strideA, stream));

    CholBatch c{};
    c.batch = batch; c.n = n; c.mrows = n + 1; c.A = A; c.lda = lda; c.strideA = strideA; c.info = info;
    c.counters = counters; c.dag_scratch = dag_scratch;
    double *Linv = nullptr, *tmp = nullptr, *alpha = nullptr, *partial = nullptr;
    const long strideL = (long)n * lda;
    if (want_grad) {
        Linv = w.take<double>((size_t)batch * strideL);
        tmp = w.take<double>((size_t)batch * trtri_tmp_elems(n));
        alpha = w.take<double>((size_t)batch * n);
        partial = w.take<double>((size_t)batch * grad_contract_tiles(n) * MAXP);
        // No clearing of Linv (640 MB per C2 step): the factorisation writes every diagonal 64 x 64 block in full (zeros
        // above the diagonal included), the trtri levels every block below them, and the products that follow restrict
        // their k-ranges to those tiles -- the strictly upper tiles are never read.  tests/test_gpu_blocks.py poisons
        // the workspace with NaN to hold this; SB200_LINV_MEMSET=1 restores the memset (A/B experiments).
        static int clear = -1;
        if (clear < 0) {
            const char* env = getenv("SB200_LINV_MEMSET");
            clear = (env && env[0] == '1') ? 1 : 0;
        }
        if (clear) SB_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * batch * strideL, stream));
        c.Dinv = Linv; c.ldd = lda; c.strideD = strideL; c.blockD = (long)NB * lda + NB;
    } else {
        c.Dinv = w.take<double>((size_t)batch * cdiv(n, NB) * NB * NB);
        c.ldd = NB; c.strideD = (long)cdiv(n, NB) * NB * NB; c.blockD = (long)NB * NB;
    }
    SB_TRY(potrf_batched(c, stream));
    nll_finalize_kernel<<<batch, 256, 0, stream>>>(n, d, A, lda, strideA, hyp, regmean, regstd, s0, nll, sig2hat, nullptr);
    SB_LAUNCH_CHECK();
    if (!want_grad) return 0;
    SB_TRY(trtri_batched(batch, n, A, lda, strideA, Linv, lda, strideL, tmp, (long)trtri_tmp_elems(n), stream));
    // alpha = Linv' z streams the factor once (HBM bound) while Rinv = Linv' Linv is DMMA bound and reads the same
    // factor: the two run side by side, the matrix-vector product on a per-device side stream forked from `stream` and
    // joined before the contraction (SB200_TRMV_SIDE=0: one after the other on `stream`, A/B experiments)
    SideLane* lane = nullptr;
    SB_TRY(side_lane(&lane));
    if (lane) {
        std::lock_guard<std::mutex> lk(lane->mu);              // fork ... join is enqueued as one unit per device
        SB_CUDA(cudaEventRecord(lane->fork, stream));
        SB_CUDA(cudaStreamWaitEvent(lane->s, lane->fork, 0));
        SB_TRY(trmv_lower_t_batched(batch, n, Linv, lda, strideL, nullptr, A + (long)n * lda, strideA, alpha, n, lane->s));
        SB_CUDA(cudaEventRecord(lane->join, lane->s));
        SB_TRY(lauum_batched(batch, n, Linv, lda, strideL, A, lda, strideA, stream)); // Rinv (lower) overwrites L
        SB_CUDA(cudaStreamWaitEvent(stream, lane->join, 0));
    } else {
        SB_TRY(trmv_lower_t_batched(batch, n, Linv, lda, strideL, nullptr, A + (long)n * lda, strideA, alpha, n, stream));
        SB_TRY(lauum_batched(batch, n, Linv, lda, strideL, A, lda, strideA, stream)); // Rinv (lower) overwrites L
    }
    SB_TRY(gp_grad_contract(batch, n, d, theta, strideTheta, gvar, strideGvar, hyp, regmean, regstd, A, lda, strideA,
                            alpha, n, sig2hat, s0, partial, grad, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
size_t gp_factorize_workspace(int batch, int n) {
    Workspace w(nullptr, 0);
    long lda = lda_for(n);
    w.take<double>((size_t)batch * (n + 1) * lda);
    w.take<double>((size_t)batch * trtri_tmp_elems(n));
    w.take<int>((size_t)batch);
    w.take<int>(potrf_dag_scratch_ints(batch, n + 1));
    return w.off;
}

int gp_factorize(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                 const double* gvar, long strideGvar, const double* hyp, double s0, double* Linv, long ldl,
                 long strideL, double* pw, double* sig2, double* logdet_half, int* info, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream) {
    SB_CHECK_ARG(batch > 0 && n > 0, "gp_factorize: batch=%d n=%d", batch, n);
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "gp_factorize: d=%d outside [1,%d]", d, MAXD);
    SB_CHECK_ARG(ldl >= n && strideL >= (long)n * ldl, "gp_factorize: bad Linv layout");
    SB_CHECK_ARG(workspace_bytes >= gp_factorize_workspace(batch, n), "gp_factorize: workspace too small");
    Workspace w(workspace, workspace_bytes);
    const long lda = lda_for(n);
    const long strideA = (long)(n + 1) * lda;
    double* A = w.take<double>((size_t)batch * strideA);
    double* tmp = w.take<double>((size_t)batch * trtri_tmp_elems(n));
    int* counters = w.take<int>((size_t)batch);
    int* dag_scratch = w.take<int>(potrf_dag_scratch_ints(batch, n + 1));
    SB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int) * batch, stream));
    SB_CUDA(cudaMemsetAsync(info, 0, sizeof(int) * batch, stream));
    SB_CUDA(cudaMemsetAsync(Linv, 0, sizeof(double) * ((size_t)(batch - 1) * strideL + (size_t)n * ldl), stream));
    SB_TRY(gp_assemble(batch, n, d, theta, strideTheta, g, strideG, gvar, strideGvar, hyp, A, lda, strideA, stream));
    CholBatch c{};
    c.batch = batch; c.n = n; c.mrows = n + 1; c.A = A; c.lda = lda; c.strideA = strideA; c.info = info;
    c.counters = counters; c.dag_scratch = dag_scratch;
    c.Dinv = Linv; c.ldd = ldl; c.strideD = strideL; c.blockD = (long)NB * ldl + NB;
    SB_TRY(potrf_batched(c, stream));
    nll_finalize_kernel<<<batch, 256, 0, stream>>>(n, d, A, lda, strideA, hyp, nullptr, nullptr, s0, nullptr, sig2,
                                                   logdet_half);
    SB_LAUNCH_CHECK();
    SB_TRY(trtri_batched(batch, n, A, lda, strideA, Linv, ldl, strideL, tmp, (long)trtri_tmp_elems(n), stream));
    SB_TRY(trmv_lower_t_batched(batch, n, Linv, ldl, strideL, nullptr, A + (long)n * lda, strideA, pw, n, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// PCs that share a factor (same hyper-parameters and gvar) still have their own scores g_j:
//   y_j = Linv g_j,  pw_j = Linv' y_j = R^-1 g_j,  sig2_j = (|y_j|^2 + s0)/(n + s0)     (PCGPwM.py:823-846)
__global__ void __launch_bounds__(256) sig2_kernel(int n, const double* __restrict__ y, double s0,
                                                   double* __restrict__ sig2) {
    __shared__ double red[32];
    const int z = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v = y[(long)z * n + i];
        acc += v * v;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) sig2[z] = (acc + s0) / (n + s0);
}

int gp_weights(int K, int n, const double* Linv, long ldl, long strideL, const int* fidx, const double* g, double s0,
               double* pw, double* sig2, double* ytmp, cudaStream_t stream) {
    SB_CHECK_ARG(K > 0 && n > 0, "gp_weights: empty problem");
    SB_TRY(trmv_lower_batched(K, n, Linv, ldl, strideL, fidx, g, n, ytmp, n, stream));
    sig2_kernel<<<K, 256, 0, stream>>>(n, ytmp, s0, sig2);
    SB_LAUNCH_CHECK();
    SB_TRY(trmv_lower_t_batched(K, n, Linv, ldl, strideL, fidx, ytmp, n, pw, n, stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// multiple of 16: the row-contiguous (n x B) operands are then describable by a TMA tensor map for EVERY population size,
// so the summation order (and with it every bit of the result) does not depend on how a population is chunked
static long ldb_for(int B) { return round_up(B, 16); }

size_t gp_predict_workspace(int K, int nf, int nu, int n, int d, int B, int want_grad) {
    Workspace w(nullptr, 0);
    long ldb = ldb_for(B);
    int Q = 2 + (want_grad ? 2 * d : 0);
    w.take<double>((size_t)nu * n * ldb);
    w.take<double>((size_t)nf * n * ldb);
    if (want_grad) w.take<double>((size_t)nf * n * ldb);
    w.take<double>((size_t)K * cdiv(n, JC) * Q * ldb);
    w.take<int>((size_t)nf + 1 + K);                 // members of every factor
    return w.off;
}

int gp_predict(int K, int nf, int nu, int n, int d, int B, const double* theta, const double* thetanew,
               const double* hypcov, const int* fu, const int* fidx, const double* nug, const double* sig2,
               const double* Linv, long ldl, long strideL, const double* pw, int want_grad, long out_ld,
               double* predvecs, double* predvars, double* dpv, double* dpvar, void* workspace, size_t workspace_bytes,
               cudaStream_t stream) {
    SB_CHECK_ARG(K > 0 && nf > 0 && nu > 0 && n > 0 && B > 0, "gp_predict: empty problem");
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "gp_predict: d=%d outside [1,%d]", d, MAXD);
    SB_CHECK_ARG(workspace_bytes >= gp_predict_workspace(K, nf, nu, n, d, B, want_grad), "gp_predict: workspace too small");
    SB_CHECK_ARG(!want_grad || (dpv && dpvar), "gp_predict: gradient outputs are NULL");
    if (out_ld <= 0) out_ld = K;
    SB_CHECK_ARG(out_ld >= K, "gp_predict: out_ld=%ld < k=%d", out_ld, K);
    // want_grad == 2: the caller's WHOLE population is one theta with d == 1, where the reference applies (1-nug) to
    // predvecs_gradtheta once instead of twice (PCGPwM.py:236, 262-268).  Decided by the caller from the full
    // population, never from the size of a chunk or of a rank's slice.
    const int single_quirk = (want_grad == 2) ? 1 : 0;
    want_grad = want_grad ? 1 : 0;
    Workspace w(workspace, workspace_bytes);
    const long ldb = ldb_for(B);
    const int Q = 2 + (want_grad ? 2 * d : 0);
    const int nchunks = cdiv(n, JC);
    const long strideT = (long)n * ldb;
    double* rT = w.take<double>((size_t)nu * strideT);
    double* T1 = w.take<double>((size_t)nf * strideT);
    double* T2 = want_grad ? w.take<double>((size_t)nf * strideT) : nullptr;
    double* partial = w.take<double>((size_t)K * nchunks * Q * ldb);
    int* mem_off = w.take<int>((size_t)nf + 1 + K);
    int* mem_list = mem_off + nf + 1;

    SB_TRY(cross_cov_t(nu, n, d, B, theta, thetanew, hypcov, rT, ldb, strideT, stream));
    {                                           // T1 = Linv r'            (n x B per factor)
        GemmArgs g{};
        g.M = n; g.N = B; g.K = n;
        g.A = Linv; g.lda = ldl; g.strideA = strideL; g.a_kc = true;
        g.B = rT; g.ldb = ldb; g.strideB = strideT; g.idxB = fu; g.b_kc = false;
        g.C = T1; g.ldc = ldb; g.strideC = strideT;
        g.alpha = 1.0; g.beta = 0.0; g.flags = GEMM_A_LOWER;
        SB_TRY(launch_dgemm(g, nf, -1, stream));
    }
    if (want_grad) {                            // T2 = Linv' T1 = Rinv r'
        GemmArgs g{};
        g.M = n; g.N = B; g.K = n;
        g.A = Linv; g.lda = ldl; g.strideA = strideL; g.a_kc = false;
        g.B = T1; g.ldb = ldb; g.strideB = strideT; g.b_kc = false;
        g.C = T2; g.ldc = ldb; g.strideC = strideT;
        g.alpha = 1.0; g.beta = 0.0; g.flags = GEMM_A_UPPER;
        SB_TRY(launch_dgemm(g, nf, -1, stream));
    }
    {
        dim3 grid(cdiv(B, 32), nchunks, nf);
#define SB_PREDICT_PARTIAL(G, DTV, MPV)                                                                             \
    predict_partial_factor_kernel<G, DTV, MPV><<<grid, 256, 0, stream>>>(n, d, B, K, theta, thetanew, hypcov, mem_off, mem_list, fu, pw, \
                                                                         rT, ldb, strideT, T1, T2, strideT, partial,    \
                                                                         nchunks, Q)
        factor_members_kernel<<<1, 256, 0, stream>>>(K, nf, fidx, mem_off, mem_list);
        SB_LAUNCH_CHECK();
        if (want_grad) {
            if (d <= 4) SB_PREDICT_PARTIAL(true, 4, 3);
            else if (d <= 8) SB_PREDICT_PARTIAL(true, 8, 2);
            else if (d <= 12) SB_PREDICT_PARTIAL(true, 12, 1);
            else SB_PREDICT_PARTIAL(true, 16, 1);
        } else {
            SB_PREDICT_PARTIAL(false, 1, 4);
        }
#undef SB_PREDICT_PARTIAL
        SB_LAUNCH_CHECK();
        dim3 grid2(cdiv(B, 128), K);
        predict_reduce_kernel<<<grid2, 128, 0, stream>>>(K, d, B, ldb, nchunks, Q, partial, nug, sig2, want_grad,
                                                         single_quirk, out_ld, predvecs, predvars, dpv, dpvar);
        SB_LAUNCH_CHECK();
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
int pc_backproject(int m, int K, int B, int d, const double* phi, const double* offset, const double* extravar,
                   const double* predvecs, const double* predvars, const double* dpv, const double* dpvar,
                   double* mean, double* var, double* covxhalf, double* mean_grad, double* covxhalf_grad,
                   cudaStream_t stream) {
    SB_CHECK_ARG(m > 0 && K > 0 && B > 0, "pc_backproject: empty problem");
    SB_CHECK_ARG(m <= 65535, "pc_backproject: m=%d exceeds grid.y", m);
    {
        dim3 grid(cdiv(B, 128), m);
        backproject_meanvar_kernel<<<grid, 128, 0, stream>>>(m, K, B, phi, offset, extravar, predvecs, predvars, mean, var);
        SB_LAUNCH_CHECK();
    }
    if (covxhalf) {
        long total = (long)m * B * K;
        backproject_covxhalf_kernel<<<cdiv(total, 256), 256, 0, stream>>>(m, K, B, phi, predvars, covxhalf);
        SB_LAUNCH_CHECK();
    }
    if (mean_grad) {
        SB_CHECK_ARG(dpv != nullptr, "pc_backproject: dpv is NULL");
        long total = (long)m * B * d;
        backproject_meangrad_kernel<<<cdiv(total, 256), 256, 0, stream>>>(m, K, B, d, phi, dpv, mean_grad);
        SB_LAUNCH_CHECK();
    }
    if (covxhalf_grad) {
        SB_CHECK_ARG(dpvar != nullptr, "pc_backproject: dpvar is NULL");
        long total = (long)m * B * K * d;
        backproject_covxgrad_kernel<<<cdiv(total, 256), 256, 0, stream>>>(m, K, B, d, phi, predvars, dpvar, covxhalf_grad);
        SB_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace sb
