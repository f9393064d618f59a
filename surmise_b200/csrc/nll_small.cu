// Single-launch NLL (+ gradient) for the optimiser's regime: n <= 207 points (the reference trains hyper-
// parameters on min(20 d, n) points, PCGPwM.py:744), one CTA per problem, nothing leaves the SM.
//
// Algorithm: symmetric sweep operator on the augmented matrix M = [[R, g], [g', 0]].  Sweeping the n pivots of R
// in order leaves  M_RR = -R^-1,  M_gR = (R^-1 g)' = alpha',  M_gg = -g' R^-1 g,  and the pivots d_k are the
// LDL' pivots of R, so 1/2 log|R| = 1/2 sum log d_k.  That is everything PCGPwM.__negloglik / __negloglikgrad
// (PCGPwM.py:850-907) need, for n^3/2 FMA in one pass (the blocked path spends n^3/3 x 3 in many launches).
// The lower triangle lives in REGISTERS: 16 x 16 threads, thread (ti, tc) owns elements (ti + 16 ii, tc + 16 cc),
// cc <= ii; each pivot step broadcasts the pivot column through shared memory (double buffered: one barrier per
// step) and is a register-resident rank-1 update.  Covariance assembly and the gradient contraction run as
// ordinary pair loops through a packed lower-triangular copy in shared memory (<= 174 KB).
#include "gp.cuh"

namespace sb {
namespace {

constexpr int T16 = 16;
constexpr int XLD = MAXD + 1;

struct SmallArgs {
    int n, d;
    const double* theta; long strideTheta;
    const double* g; long strideG;
    const double* gvar; long strideGvar;
    const double* hyp;
    const double* regmean; const double* regstd;
    double s0;
    int want_grad;
    double* nll; double* grad; int* info;
    long long* stamps;      // optional (instrumentation): clock64() at phase boundaries of CTA 0, 8 slots
};

struct SmallConsts {
    double ell[MAXD];
    double c, onemc, nu, ev;
};

__device__ __forceinline__ long tri(int i, int c) { return (long)i * (i + 1) / 2 + c; }

template <int R>
__global__ void __launch_bounds__(256, 1) gp_nll_small_kernel(SmallArgs p) {
    constexpr int NP = T16 * R;                       // padded order (>= n + 1)
    extern __shared__ double sm[];
    double* P = sm;                                   // packed lower triangle, NP (NP + 1) / 2
    double* ths = P + (long)NP * (NP + 1) / 2;        // [n][XLD] coordinates / length-scale
    double* vbuf = ths + (long)p.n * XLD;             // [2][NP]
    double* red = vbuf + 2 * NP;                      // [MAXP][8]
    double* piv = red + MAXP * 8;                     // [NP] pivots d_k
    __shared__ SmallConsts sc;
    __shared__ int s_bad;

    const int z = blockIdx.x, tid = threadIdx.x, n = p.n, d = p.d;
#define SB_STAMP(slot) do { if (p.stamps && z == 0 && tid == 0) p.stamps[slot] = clock64(); } while (0)
    SB_STAMP(0);
    const double* theta = p.theta + (long)z * p.strideTheta;
    const double* gz = p.g + (long)z * p.strideG;
    const double* gv = p.gvar ? p.gvar + (long)z * p.strideGvar : nullptr;
    const double* hyp = p.hyp + (long)z * (d + 3);

    if (tid < d) sc.ell[tid] = exp(hyp[tid]);
    if (tid == 32) {
        double e = exp(hyp[d]);
        sc.c = 1.0 / (1.0 + e);
        sc.onemc = e / (1.0 + e);
        double en = exp(hyp[d + 2]);
        sc.nu = en / (1.0 + en);                      // PCGPwM.py:853
        sc.ev = exp(hyp[d + 1]);
        s_bad = 0;
    }
    __syncthreads();
    for (int idx = tid; idx < n * d; idx += 256) {
        int r = idx / d, k = idx % d;
        ths[r * XLD + k] = theta[(long)r * d + k] / sc.ell[k];    // matern_covmat.pyx:29-32
    }
    __syncthreads();

    SB_STAMP(1);
    // ---- assemble M (packed lower): one warp per row, lanes along the row
    const int warp = tid >> 5, lane = tid & 31;
    const double omn = 1.0 - sc.nu;
    for (int i = warp; i < NP; i += 8) {
        if (i < n) {
            const double* xi = ths + i * XLD;
            for (int c0 = lane; c0 <= i; c0 += 64) {
                const int c1 = c0 + 32;
                const bool on1 = c1 <= i;
                const double* x0 = ths + c0 * XLD;
                const double* x1 = ths + (on1 ? c1 : c0) * XLD;
                double p0 = sc.c, p1 = sc.c, s0 = 0.0, s1 = 0.0;
                for (int k = 0; k < d; k++) {
                    double S0 = fabs(xi[k] - x0[k]), S1 = fabs(xi[k] - x1[k]);
                    p0 *= (1.0 + S0);
                    p1 *= (1.0 + S1);
                    s0 -= S0;
                    s1 -= S1;
                }
                double v0 = omn * (p0 * exp(s0) + sc.onemc);       // PCGPwM.py:852-857
                double v1 = omn * (p1 * exp(s1) + sc.onemc);
                if (i == c0) v0 += sc.nu + (gv ? sc.ev * gv[i] : 0.0);
                if (i == c1) v1 += sc.nu + (gv ? sc.ev * gv[i] : 0.0);
                P[tri(i, c0)] = v0;
                if (on1) P[tri(i, c1)] = v1;
            }
        } else {
            for (int c = lane; c <= i; c += 32)
                P[tri(i, c)] = (i == n) ? ((c < n) ? gz[c] : 0.0) : ((i == c) ? 1.0 : 0.0);   // g row | identity padding
        }
    }
    __syncthreads();

    SB_STAMP(2);
    // ---- registers <- packed triangle
    const int ti = tid >> 4, tc = tid & 15;
    double a[R][R];
#pragma unroll
    for (int ii = 0; ii < R; ii++)
#pragma unroll
        for (int cc = 0; cc < R; cc++)
            if (cc <= ii) {
                int i = ti + T16 * ii, c = tc + T16 * cc;
                a[ii][cc] = (c <= i) ? P[tri(i, c)] : 0.0;
            }

    SB_STAMP(3);
    // ---- sweep the n pivots of R
    int bad_at = 0;
#pragma unroll
    for (int kb = 0; kb < R; kb++) {
#pragma unroll 1
        for (int kr = 0; kr < T16; kr++) {
            const int k = kb * T16 + kr;
            if (k >= n) break;
            double* v = vbuf + (k & 1) * NP;
            if (tc == kr) {
#pragma unroll
                for (int ii = 0; ii < R; ii++)
                    if (ii >= kb && (ii > kb || ti >= kr)) v[ti + T16 * ii] = a[ii][kb];
            }
            if (ti == kr) {
#pragma unroll
                for (int cc = 0; cc < R; cc++)
                    if (cc <= kb && (cc < kb || tc <= kr)) v[tc + T16 * cc] = a[kb][cc];
            }
            __syncthreads();
            double dk = v[k];
            bool bad = !(dk > 0.0);
            if (bad && bad_at == 0) bad_at = k + 1;
            if (bad) dk = 1.0;
            if (tid == 0) piv[k] = dk;
            const double dinv = 1.0 / dk;
            double wi[R], vc[R];
#pragma unroll
            for (int ii = 0; ii < R; ii++) wi[ii] = v[ti + T16 * ii] * dinv;
#pragma unroll
            for (int cc = 0; cc < R; cc++) vc[cc] = v[tc + T16 * cc];
#pragma unroll
            for (int ii = 0; ii < R; ii++)
#pragma unroll
                for (int cc = 0; cc < R; cc++)
                    if (cc <= ii) a[ii][cc] -= wi[ii] * vc[cc];
            if (tc == kr) {
#pragma unroll
                for (int ii = 0; ii < R; ii++)
                    if (ii >= kb) {
                        int i = ti + T16 * ii;
                        if (i > k) a[ii][kb] = wi[ii];
                        else if (i == k) a[ii][kb] = -dinv;
                    }
            }
            if (ti == kr) {
#pragma unroll
                for (int cc = 0; cc < R; cc++)
                    if (cc <= kb) {
                        int c = tc + T16 * cc;
                        if (c < k) a[kb][cc] = vc[cc] * dinv;
                    }
            }
        }
    }
    SB_STAMP(4);
    if (bad_at != 0 && tid == 0) s_bad = bad_at;

    // ---- packed triangle <- registers:  -R^-1 | alpha' | -g'R^-1 g
#pragma unroll
    for (int ii = 0; ii < R; ii++)
#pragma unroll
        for (int cc = 0; cc < R; cc++)
            if (cc <= ii) {
                int i = ti + T16 * ii, c = tc + T16 * cc;
                if (c <= i) P[tri(i, c)] = a[ii][cc];
            }
    __syncthreads();

    double ld = 0.0;                                               // sum log d_k = log|R|
    for (int k = tid; k < n; k += 256) ld += log(piv[k]);
    ld = warp_sum(ld);
    if (lane == 0) red[warp] = ld;
    __syncthreads();
    double logdet = 0.0;
#pragma unroll
    for (int wv = 0; wv < 8; wv++) logdet += red[wv];
    __syncthreads();
    const double qf = -P[tri(n, n)];                               // g' R^-1 g
    const double sig2hat = (qf + p.s0) / (n + p.s0);               // PCGPwM.py:863
    const int np = d + 3;
    if (tid == 0) {
        double pr = 0.0;
        for (int k = 0; k < np; k++) {
            double t = (1e-8 + hyp[k] - p.regmean[k]) / p.regstd[k];
            pr += t * t;
        }
        p.nll[z] = 0.5 * logdet + 0.5 * n * log(sig2hat) + 0.5 * pr;   // PCGPwM.py:864-867
        p.info[z] = s_bad;
    }
    SB_STAMP(5);
    if (!p.want_grad) return;

    // ---- gradient: sum_ab W_ab dR_j(a,b),  W = 1/2 R^-1 - cq alpha alpha'   (PCGPwM.py:874-907)
    const double cq = 0.5 * n / ((n + p.s0) * sig2hat);
    const double* alpha = P + tri(n, 0);
    double acc[MAXD], acc_g = 0.0, acc_v = 0.0, acc_n = 0.0;
#pragma unroll
    for (int k = 0; k < MAXD; k++) acc[k] = 0.0;
    // two columns per lane per iteration (independent chains hide the FP64 latencies at 2 warps / scheduler);
    // 1/(1+S_k) = prod_{l != k}(1+S_l) / prod_l(1+S_l): one division per pair instead of d
    for (int i = warp; i < n; i += 8) {
        const double ai = alpha[i];
        const double* xi = ths + i * XLD;
        for (int c0 = lane; c0 <= i; c0 += 64) {
            const int c1 = c0 + 32;
            const bool on1 = c1 <= i;
            const int c1s = on1 ? c1 : c0;
            double w0 = -0.5 * P[tri(i, c0)] - cq * ai * alpha[c0];
            double w1 = on1 ? -0.5 * P[tri(i, c1s)] - cq * ai * alpha[c1s] : 0.0;
            if (i == c0) {
                if (gv) acc_v += w0 * sc.ev * gv[i];
                w0 = 0.0;
            }
            if (on1 && i == c1) {
                if (gv) acc_v += w1 * sc.ev * gv[i];
                w1 = 0.0;
            }
            w0 *= 2.0;
            w1 *= 2.0;
            const double* x0 = ths + c0 * XLD;
            const double* x1 = ths + c1s * XLD;
            double S0[MAXD], S1[MAXD], pre0[MAXD], pre1[MAXD];
            double p0 = 1.0, p1 = 1.0, s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int k = 0; k < MAXD; k++)
                if (k < d) {
                    S0[k] = fabs(xi[k] - x0[k]);
                    S1[k] = fabs(xi[k] - x1[k]);
                    pre0[k] = p0;
                    pre1[k] = p1;
                    p0 *= (1.0 + S0[k]);
                    p1 *= (1.0 + S1[k]);
                    s0 -= S0[k];
                    s1 -= S1[k];
                }
            const double rp0 = sc.c * p0 * exp(s0), rp1 = sc.c * p1 * exp(s1);
            const double wr0 = w0 * omn * rp0 / p0, wr1 = w1 * omn * rp1 / p1;     // includes 1 / prod(1+S)
            double suf0 = 1.0, suf1 = 1.0;
#pragma unroll
            for (int k = MAXD - 1; k >= 0; k--)
                if (k < d) {
                    acc[k] += wr0 * (S0[k] * S0[k]) * (pre0[k] * suf0) + wr1 * (S1[k] * S1[k]) * (pre1[k] * suf1);
                    suf0 *= (1.0 + S0[k]);
                    suf1 *= (1.0 + S1[k]);
                }
            acc_g += omn * sc.onemc * (w0 * (sc.c - rp0) + w1 * (sc.c - rp1));
            acc_n -= sc.nu * omn * (w0 * (rp0 + sc.onemc) + w1 * (rp1 + sc.onemc));
        }
    }
#pragma unroll
    for (int k = 0; k < MAXD; k++)
        if (k < d) {
            double sv = warp_sum(acc[k]);
            if (lane == 0) red[k * 8 + warp] = sv;
        }
    {
        double sg = warp_sum(acc_g), sv = warp_sum(acc_v), sn = warp_sum(acc_n);
        if (lane == 0) {
            red[d * 8 + warp] = sg;
            red[(d + 1) * 8 + warp] = sv;
            red[(d + 2) * 8 + warp] = sn;
        }
    }
    __syncthreads();
    if (tid < np) {
        double sv = 0.0;
#pragma unroll
        for (int wv = 0; wv < 8; wv++) sv += red[tid * 8 + wv];
        sv += (1e-8 + hyp[tid] - p.regmean[tid]) / (p.regstd[tid] * p.regstd[tid]);   // PCGPwM.py:905-906
        p.grad[(long)z * np + tid] = sv;
    }
    SB_STAMP(6);
#undef SB_STAMP
}

template <int R>
int launch_small(const SmallArgs& a, int batch, cudaStream_t stream) {
    constexpr int NP = T16 * R;
    size_t smem = ((size_t)NP * (NP + 1) / 2 + (size_t)a.n * XLD + 3 * NP + MAXP * 8) * sizeof(double);
    constexpr size_t kMaxDyn = 227 * 1024 - 1024;      // the kernel also has a few hundred bytes of static shared memory
    SB_CHECK_ARG(smem <= kMaxDyn, "gp_nll_small: n=%d needs %zu B of shared memory", a.n, smem);
    static bool configured[64] = {false};
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        SB_CUDA(cudaFuncSetAttribute(gp_nll_small_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDyn));
        configured[dev] = true;
    }
    gp_nll_small_kernel<R><<<batch, 256, smem, stream>>>(a);
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int gp_nll_small_max_n() { return T16 * 13 - 1; }

int gp_nll_grad_small(int batch, int n, int d, const double* theta, long strideTheta, const double* g, long strideG,
                      const double* gvar, long strideGvar, const double* hyp, const double* regmean,
                      const double* regstd, double s0, int want_grad, double* nll, double* grad, int* info,
                      cudaStream_t stream, long long* stamps) {
    SB_CHECK_ARG(n >= 1 && n <= gp_nll_small_max_n(), "gp_nll_small: n=%d outside [1,%d]", n, gp_nll_small_max_n());
    SB_CHECK_ARG(d >= 1 && d <= MAXD, "gp_nll_small: d=%d outside [1,%d]", d, MAXD);
    SmallArgs a{n, d, theta, strideTheta, g, strideG, gvar, strideGvar, hyp, regmean, regstd, s0, want_grad,
                nll, grad, info, stamps};
    const int R = (n + 1 + T16 - 1) / T16;
    if (R <= 4) return launch_small<4>(a, batch, stream);
    if (R <= 6) return launch_small<6>(a, batch, stream);
    if (R <= 8) return launch_small<8>(a, batch, stream);
    if (R <= 10) return launch_small<10>(a, batch, stream);
    if (R <= 11) return launch_small<11>(a, batch, stream);
    return launch_small<13>(a, batch, stream);
}

}  // namespace sb
