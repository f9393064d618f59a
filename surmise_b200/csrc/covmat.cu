// Matern-3/2 pair kernels.  All four kernels tile the (row, column) pair space 32 x 32 with 256
// threads, stage the length-scale-divided coordinates of the tile's rows and columns in shared
// memory (x / exp(gamma), the reference's own order of operations, matern_covmat.pyx:29-32) and
// write with the fastest thread index along the contiguous output dimension.
// Standalone covmat is HBM-write bound: 8 n1 n2 (+ 8 D n1 n2 with gradients) bytes.
#include "covmat.cuh"
#include <cstdlib>
#include "covmat_dev.cuh"

namespace sb {
namespace {

constexpr int TS = 32;            // tile side
constexpr int XLD = MAXD + 1;     // padded row length of the staged coordinates

using gpdev::PairConsts;
using gpdev::pair_consts;
using gpdev::GpConsts;
using gpdev::load_gp_consts;

// R_pre = c prod(1+S_k) exp(-sum S_k)                       (matern_covmat.pyx:34-48)
__device__ __forceinline__ double rpre_pair(const double* a, const double* b, int d, double c) {
    double prod = c, ssum = 0.0;
    for (int k = 0; k < d; k++) {
        double S = fabs(a[k] - b[k]);
        prod *= (1.0 + S);
        ssum -= S;
    }
    return prod * exp(ssum);
}

__device__ __forceinline__ void stage_scaled(double* dst, const double* __restrict__ X, int row0, int nrows, int d,
                                             const double* ell) {
    for (int idx = threadIdx.x; idx < TS * d; idx += blockDim.x) {
        int r = idx / d, k = idx % d;
        int gr = row0 + r;
        dst[r * XLD + k] = (gr < nrows) ? X[(long)gr * d + k] / ell[k] : 0.0;
    }
}

// rows of X divided by the length-scales, every row padded with zeros up to DT dimensions: the pair loops below then
// run over all DT dimensions unguarded (a padded dimension multiplies the product by exactly 1 and adds exactly 0)
template <int DT>
__device__ __forceinline__ void stage_scaled_n(double* dst, const double* __restrict__ X, int row0, int nrows, int d,
                                               const double* ell, int rows) {
    for (int idx = threadIdx.x; idx < rows * DT; idx += blockDim.x) {
        int r = idx / DT, k = idx % DT;
        int gr = row0 + r;
        dst[r * XLD + k] = (gr < nrows && k < d) ? X[(long)gr * d + k] / ell[k] : 0.0;
    }
}

// rpre_pair over zero-padded coordinates, fully unrolled: same operations in the same order as the runtime-d loop
template <int DT>
__device__ __forceinline__ double rpre_pair_t(const double* a, const double* b, double c) {
    double prod = c, ssum = 0.0;
#pragma unroll
    for (int k = 0; k < DT; k++) {
        double S = fabs(a[k] - b[k]);
        prod *= (1.0 + S);
        ssum -= S;
    }
    return prod * exp(ssum);
}

template <int GRAD>
__global__ void __launch_bounds__(256) matern_covmat_kernel(const double* __restrict__ x1, int n1,
                                                            const double* __restrict__ x2, int n2, int d,
                                                            const double* __restrict__ gammav, double* __restrict__ R,
                                                            long ldr, double* __restrict__ dR) {
    __shared__ double xa[TS * XLD], xb[TS * XLD], ell[MAXD], ginv[MAXD];
    if (threadIdx.x < d) {
        ell[threadIdx.x] = exp(gammav[threadIdx.x]);
        ginv[threadIdx.x] = exp(-gammav[threadIdx.x]);
    }
    __syncthreads();
    const int i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
    stage_scaled(xa, x1, i0, n1, d, ell);
    stage_scaled(xb, x2, j0, n2, d, ell);
    __syncthreads();
    const PairConsts pc = pair_consts(gammav[d]);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = j0 + tx;
    if (j >= n2) return;
    const long plane = (long)n1 * n2;
    for (int r = ty; r < TS; r += 8) {
        int i = i0 + r;
        if (i >= n1) break;
        const double* a = xa + r * XLD;
        const double* b = xb + tx * XLD;
        double rp = rpre_pair(a, b, d, pc.c);
        R[(long)i * ldr + j] = rp + pc.onemc;
        if (GRAD == 1) {                                           // d/dgamma  (pyx:49-57)
            for (int k = 0; k < d; k++) {
                double S = fabs(a[k] - b[k]);
                dR[k * plane + (long)i * n2 + j] = (S * S) / (1.0 + S) * rp;
            }
            dR[d * plane + (long)i * n2 + j] = pc.onemc * (pc.c - rp);
        } else if (GRAD == 2) {                                    // d/dx1     (pyx:13-24, 58-59)
            for (int k = 0; k < d; k++) {
                double s = a[k] - b[k];
                dR[k * plane + (long)i * n2 + j] = (-ginv[k] * s / (1.0 + fabs(s))) * rp;
            }
        }
    }
}

template <int DT>
__global__ void __launch_bounds__(256) gp_assemble_kernel(int n, int d, const double* __restrict__ thetab,
                                                          long strideTheta, const double* __restrict__ gb, long strideG,
                                                          const double* __restrict__ gvarb, long strideGvar,
                                                          const double* __restrict__ hypb, double* __restrict__ Ab,
                                                          long lda, long strideA) {
    const int z = blockIdx.z;
    const int i0 = blockIdx.y * TS, j0 = blockIdx.x * TS;
    if (j0 > i0 + TS - 1) return;                                  // strictly above the diagonal
    __shared__ double xa[TS * XLD], xb[TS * XLD];
    __shared__ GpConsts gc;
    const double* theta = thetab + (long)z * strideTheta;
    const double* hyp = hypb + (long)z * (d + 3);
    double* A = Ab + (long)z * strideA;
    load_gp_consts(&gc, hyp, d);
    __syncthreads();
    stage_scaled_n<DT>(xa, theta, i0, n, d, gc.ell, TS);
    stage_scaled_n<DT>(xb, theta, j0, n, d, gc.ell, TS);
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = j0 + tx;
    if (j >= n) return;
    double bj[DT];
#pragma unroll
    for (int k = 0; k < DT; k++) bj[k] = xb[tx * XLD + k];
    // the thread's four pairs are independent straight-line code (rows past n are staged as zeros: harmless)
    double r0[TS / 8];
#pragma unroll
    for (int u = 0; u < TS / 8; u++) r0[u] = rpre_pair_t<DT>(xa + (ty + 8 * u) * XLD, bj, gc.c) + gc.onemc;
#pragma unroll
    for (int u = 0; u < TS / 8; u++) {
        const int i = i0 + ty + 8 * u;
        if (i > n) continue;
        if (i == n) {                                              // augmented row: g'
            A[(long)n * lda + j] = gb[(long)z * strideG + j];
            continue;
        }
        if (j > i) continue;
        double v = (1.0 - gc.nu) * r0[u];                          // PCGPwM.py:852-857
        if (i == j) {
            v += gc.nu;
            if (gvarb) v += gc.ev * gvarb[(long)z * strideGvar + i];
        }
        A[(long)i * lda + j] = v;
    }
}

// one CTA per 64x64 tile of the lower triangle (16 pairs per thread); partial[z][tile][0..p)
constexpr int GT = 64;

// DT: compile-time bound on d (the per-dimension loops are unrolled into registers up to DT, not MAXD)
template <int DT, int OCC>
__global__ void __launch_bounds__(256, OCC) gp_grad_contract_kernel(int n, int d, const double* __restrict__ thetab,
                                                               long strideTheta, const double* __restrict__ gvarb,
                                                               long strideGvar, const double* __restrict__ hypb,
                                                               const double* __restrict__ Rinvb, long ldr, long strideR,
                                                               const double* __restrict__ alphab, long strideAlpha,
                                                               const double* __restrict__ sig2hat, double s0,
                                                               double* __restrict__ partial, int ntiles) {
    const int z = blockIdx.y;
    // decode linear lower-triangular tile index -> (ti >= tj)
    int t = blockIdx.x;
    int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
    while ((long)(ti + 1) * (ti + 2) / 2 <= t) ti++;
    while ((long)ti * (ti + 1) / 2 > t) ti--;
    const int tj = t - ti * (ti + 1) / 2;
    const int i0 = ti * GT, j0 = tj * GT;

    __shared__ double xa[GT * XLD], xb[GT * XLD], red[MAXP][8];
    __shared__ GpConsts gc;
    const double* theta = thetab + (long)z * strideTheta;
    const double* hyp = hypb + (long)z * (d + 3);
    const double* Rinv = Rinvb + (long)z * strideR;
    const double* alpha = alphab + (long)z * strideAlpha;
    load_gp_consts(&gc, hyp, d);
    __syncthreads();
    stage_scaled_n<DT>(xa, theta, i0, n, d, gc.ell, GT);
    stage_scaled_n<DT>(xb, theta, j0, n, d, gc.ell, GT);
    __syncthreads();
    const double cq = 0.5 * n / ((n + s0) * sig2hat[z]);           // PCGPwM.py:898-902
    const double omn = 1.0 - gc.nu;
    gpdev::GradAcc<DT> G;                                           // length-scales | gamma_d | h_v | h_nu
    G.clear();

    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int j = j0 + tx;
    if (j < n) {
        const double aj = alpha[j];
        const double* b = xb + tx * XLD;
        // four rows per trip: their W entries are loaded together, ahead of the arithmetic of the first
        constexpr int PF = 4;
#pragma unroll 1
        for (int rb = ty; rb < GT; rb += 4 * PF) {
            double wl[PF];
#pragma unroll
            for (int u2 = 0; u2 < PF; u2++) {
                const int i = i0 + rb + 4 * u2;
                wl[u2] = (i < n && j <= i) ? (0.5 * Rinv[(long)i * ldr + j] - cq * alpha[i] * aj) : 0.0;
            }
#pragma unroll
            for (int u2 = 0; u2 < PF; u2++) {
                const int r = rb + 4 * u2, i = i0 + r;
                if (i >= n || j > i) continue;
                double w = wl[u2];
                if (i == j) {                                          // only d/dh_v lives on the diagonal
                    if (gvarb) G.v += w * gc.ev * gvarb[(long)z * strideGvar + i];
                    continue;
                }
                gpdev::contract_offdiag<DT>(G, w, xa + r * XLD, b, gc, omn);
            }
        }
    }
    // warp shuffles, then one pass through shared memory (fixed order)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < DT; k++)
        if (k < d) {
            double sv = warp_sum(G.ls[k]);
            if (lane == 0) red[k][warp] = sv;
        }
    {
        double sg = warp_sum(G.g), sv = warp_sum(G.v), sn = warp_sum(G.nn);
        if (lane == 0) {
            red[d][warp] = sg;
            red[d + 1][warp] = sv;
            red[d + 2][warp] = sn;
        }
    }
    __syncthreads();
    if (threadIdx.x < d + 3) {
        double sv = 0.0;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) sv += red[threadIdx.x][wv];
        partial[((long)z * ntiles + t) * MAXP + threadIdx.x] = sv;
    }
}

// one warp per hyper-parameter: fixed-order strided sum over the tile partials, then the prior term
__global__ void gp_grad_finalize_kernel(int d, int ntiles, const double* __restrict__ partial,
                                        const double* __restrict__ hypb, const double* __restrict__ regmean,
                                        const double* __restrict__ regstd, double* __restrict__ grad) {
    const int z = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31, p = d + 3;
    if (k >= p) return;
    double s = 0.0;
    for (int t = lane; t < ntiles; t += 32) s += partial[((long)z * ntiles + t) * MAXP + k];
    s = warp_sum(s);
    if (lane == 0) {
        double h = hypb[(long)z * p + k];
        s += (1e-8 + h - regmean[k]) / (regstd[k] * regstd[k]);    // PCGPwM.py:905-906
        grad[(long)z * p + k] = s;
    }
}

template <int DT>
__global__ void __launch_bounds__(256) cross_cov_t_kernel(int n, int d, int B, const double* __restrict__ theta,
                                                          const double* __restrict__ thetanew,
                                                          const double* __restrict__ hypcovb, double* __restrict__ rTb,
                                                          long ldb, long strideU) {
    const int u = blockIdx.z;
    const double* hc = hypcovb + (long)u * (d + 1);
    __shared__ double xa[TS * XLD], xb[TS * XLD], ell[MAXD];
    if (threadIdx.x < d) ell[threadIdx.x] = exp(hc[threadIdx.x]);
    __syncthreads();
    const int j0 = blockIdx.y * TS, b0 = blockIdx.x * TS;
    stage_scaled_n<DT>(xa, theta, j0, n, d, ell, TS);
    stage_scaled_n<DT>(xb, thetanew, b0, B, d, ell, TS);
    __syncthreads();
    const PairConsts pc = pair_consts(hc[d]);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int b = b0 + tx;
    if (b >= B) return;
    double* rT = rTb + (long)u * strideU;
    double bn[DT];
#pragma unroll
    for (int k = 0; k < DT; k++) bn[k] = xb[tx * XLD + k];
    double v[TS / 8];
    // argument order (new, fit) as in covmat(theta_new, theta_fit): |a-b| is symmetric
#pragma unroll
    for (int q = 0; q < TS / 8; q++) v[q] = rpre_pair_t<DT>(bn, xa + (ty + 8 * q) * XLD, pc.c) + pc.onemc;
#pragma unroll
    for (int q = 0; q < TS / 8; q++) {
        const int j = j0 + ty + 8 * q;
        if (j < n) rT[(long)j * ldb + b] = v[q];
    }
}

}  // namespace

int grad_contract_tiles(int n) {
    long t = cdiv(n, GT);
    return (int)(t * (t + 1) / 2);
}

int matern_covmat(const double* x1, int n1, const double* x2, int n2, int d, const double* gammav, double* R,
                  long ldr, double* dR, int grad_mode, cudaStream_t stream) {
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "covmat: d=%d outside [1,%d]", d, MAXD);
    SB_CHECK_ARG(grad_mode >= 0 && grad_mode <= 2, "covmat: grad_mode %d", grad_mode);
    SB_CHECK_ARG(grad_mode == 0 || dR != nullptr, "covmat: dR is NULL");
    if (n1 <= 0 || n2 <= 0) return 0;
    dim3 grid(cdiv(n2, TS), cdiv(n1, TS));
    if (grad_mode == 0) matern_covmat_kernel<0><<<grid, 256, 0, stream>>>(x1, n1, x2, n2, d, gammav, R, ldr, dR);
    else if (grad_mode == 1) matern_covmat_kernel<1><<<grid, 256, 0, stream>>>(x1, n1, x2, n2, d, gammav, R, ldr, dR);
    else matern_covmat_kernel<2><<<grid, 256, 0, stream>>>(x1, n1, x2, n2, d, gammav, R, ldr, dR);
    SB_LAUNCH_CHECK();
    return 0;
}

int gp_assemble(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                const double* gvar, long strideGvar, const double* hyp, double* A, long lda, long strideA,
                cudaStream_t stream) {
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "gp_assemble: d=%d outside [1,%d]", d, MAXD);
    dim3 grid(cdiv(n, TS), cdiv(n + 1, TS), batch);
#define SB_ASSEMBLE(DT) \
    gp_assemble_kernel<DT><<<grid, 256, 0, stream>>>(n, d, theta, strideTheta, g, strideG, gvar, strideGvar, hyp, A, lda, strideA)
    if (d <= 4) SB_ASSEMBLE(4);
    else if (d <= 6) SB_ASSEMBLE(6);
    else if (d <= 8) SB_ASSEMBLE(8);
    else if (d <= 10) SB_ASSEMBLE(10);
    else if (d <= 12) SB_ASSEMBLE(12);
    else SB_ASSEMBLE(16);
#undef SB_ASSEMBLE
    SB_LAUNCH_CHECK();
    return 0;
}

int gp_grad_contract(int batch, int n, int d, const double* theta, long strideTheta, const double* gvar,
                     long strideGvar, const double* hyp, const double* regmean, const double* regstd,
                     const double* Rinv, long ldr, long strideR, const double* alpha, long strideAlpha,
                     const double* sig2hat, double s0, double* partial, double* grad, cudaStream_t stream) {
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "gp_grad_contract: d=%d outside [1,%d]", d, MAXD);
    int ntiles = grad_contract_tiles(n);
    dim3 grid(ntiles, batch);
    static int occ3 = -1;                     // SB200_CONTRACT_OCC=2|3: CTAs per SM for d <= 8 (A/B experiments)
    if (occ3 < 0) {
        const char* env = getenv("SB200_CONTRACT_OCC");
        occ3 = (env && env[0] == '2') ? 0 : 1;
    }
#define SB_CONTRACT(DT, OCC)                                                                                            \
    gp_grad_contract_kernel<DT, OCC><<<grid, 256, 0, stream>>>(n, d, theta, strideTheta, gvar, strideGvar, hyp, Rinv, ldr, \
                                                               strideR, alpha, strideAlpha, sig2hat, s0, partial, ntiles)
    if (d <= 4) SB_CONTRACT(4, 3);
    else if (d <= 6) SB_CONTRACT(6, 3);
    else if (d <= 8 && occ3) SB_CONTRACT(8, 3);
    else if (d <= 8) SB_CONTRACT(8, 2);
    else if (d <= 10) SB_CONTRACT(10, 2);
    else if (d <= 12) SB_CONTRACT(12, 2);
    else SB_CONTRACT(16, 2);
#undef SB_CONTRACT
    SB_LAUNCH_CHECK();
    gp_grad_finalize_kernel<<<batch, 32 * (d + 3), 0, stream>>>(d, ntiles, partial, hyp, regmean, regstd, grad);
    SB_LAUNCH_CHECK();
    return 0;
}

int gp_grad_finalize(int batch, int d, int ntiles, const double* partial, const double* hyp, const double* regmean,
                     const double* regstd, double* grad, cudaStream_t stream) {
    gp_grad_finalize_kernel<<<batch, 32 * (d + 3), 0, stream>>>(d, ntiles, partial, hyp, regmean, regstd, grad);
    SB_LAUNCH_CHECK();
    return 0;
}

int cross_cov_t(int nu, int n, int d, int B, const double* theta, const double* thetanew, const double* hypcov,
                double* rT, long ldb, long strideU, cudaStream_t stream) {
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "cross_cov_t: d=%d outside [1,%d]", d, MAXD);
    dim3 grid(cdiv(B, TS), cdiv(n, TS), nu);
#define SB_CROSS(DT) cross_cov_t_kernel<DT><<<grid, 256, 0, stream>>>(n, d, B, theta, thetanew, hypcov, rT, ldb, strideU)
    if (d <= 4) SB_CROSS(4);
    else if (d <= 6) SB_CROSS(6);
    else if (d <= 8) SB_CROSS(8);
    else if (d <= 10) SB_CROSS(10);
    else if (d <= 12) SB_CROSS(12);
    else SB_CROSS(16);
#undef SB_CROSS
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace sb
