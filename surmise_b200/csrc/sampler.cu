// Device-resident parallel-tempering Langevin Monte Carlo step: everything surmise's PTLMC sampler does between two
// evaluations of the log-posterior (src/surmise/utilitiesmethods/PTLMC.py:203-273) in ONE single-CTA kernel, so that a
// calibration run never leaves the GPU:
//
//   [accept / reject the pending proposal      PTLMC.py:222-237]
//   [temperature-exchange sweep, 5 B proposals PTLMC.py:239-246, 259-273]
//   [step-size adaptation while tuning, or save PTLMC.py:247-257]
//   [Langevin proposal for the next step       PTLMC.py:205-213]
//
// The log-likelihood of the proposals comes from sb200_gp_predict + sb200_woodbury_loglik between two such launches;
// the prior is the uniform box the caller declares (lo, hi): proposals outside it are never accepted and take no part
// in the batch-wide variance test (row_valid).  The chain is a few kilobytes of state, so one CTA holds it in shared
// memory; the exchange sweep is sequential by construction (every swap changes what the next proposal sees) and runs
// in one thread out of shared memory (~40 cycles per proposal).  Random numbers: counter-based Philox4x32-10 keyed by
// the caller's seed, counter = (step, stream, index) -- identical on every rank of a replicated run.
#include "gp.cuh"

namespace sb {
namespace {

constexpr int ST = 256;

struct Philox {
    unsigned k0, k1;
    __device__ Philox(unsigned long long seed) : k0((unsigned)seed), k1((unsigned)(seed >> 32)) {}
    __device__ uint4 operator()(unsigned c0, unsigned c1, unsigned c2, unsigned c3) const {
        unsigned a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            const unsigned hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const unsigned hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            const unsigned n0 = hi1 ^ c1 ^ a, n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// two uniforms in (0,1) from 128 random bits (53 bits each)
__device__ __forceinline__ void uniforms(const uint4& r, double& u0, double& u1) {
    const unsigned long long a = ((unsigned long long)r.x << 32) | r.y, b = ((unsigned long long)r.z << 32) | r.w;
    u0 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    u1 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

enum { STREAM_Z = 0, STREAM_ACCEPT = 1, STREAM_SWAP = 2 };

// scal: [0] tau  [1] accepted fraction accumulated since the last adaptation  [2] total accepted moves
//       [3] total accepted swaps  [4] rho (current base step size)  [5] proposals outside the box (total)
__global__ void __launch_bounds__(ST) ptlmc_step_kernel(
    int B, int d, int numtemps, int tune, int nsave, int mode, unsigned long long seed, double taracc,
    double* __restrict__ thetac, double* __restrict__ fval, double* __restrict__ dfval, double* __restrict__ thetap,
    double* __restrict__ z, double* __restrict__ adjrho, const double* __restrict__ ll, const double* __restrict__ dll,
    const double* __restrict__ temps, const double* __restrict__ lo, const double* __restrict__ hi,
    const double* __restrict__ hc, const double* __restrict__ cov0, double* __restrict__ scal,
    long long* __restrict__ kstep, double* __restrict__ save, int* __restrict__ row_valid) {
    extern __shared__ double sm[];
    double* s_hc = sm;                       // d*d
    double* s_cov = s_hc + d * d;            // d*d
    double* s_lp = s_cov + d * d;            // B : unnormalised log-posterior in position order
    double* s_rho = s_lp + B;                // B : 1/T[i-1] - 1/T[i]
    double* s_logu = s_rho + B;              // 5B
    double* s_th = s_logu + 5 * B;           // B*d staging for the permutation
    double* s_df = s_th + (size_t)B * d;     // B*d
    double* s_red = s_df + (size_t)B * d;    // 32 for block_sum + 2 scalars
    int* s_ord = (int*)(s_red + 34);         // B
    int* s_rt = s_ord + B;                   // 5B
    const int tid = threadIdx.x;
    const Philox rng(seed);
    const long long k = *kstep;              // index of the step whose proposal is pending
    const unsigned klo = (unsigned)k, khi = (unsigned)(k >> 32);

    for (int i = tid; i < d * d; i += ST) {
        s_hc[i] = hc[i];
        s_cov[i] = cov0[i];
    }
    __syncthreads();

    if (mode & 1) {
        // ---- accept / reject (PTLMC.py:214-237) ----
        double nacc = 0.0, nout = 0.0;
        for (int b = tid; b < B; b += ST) {
            const double T = temps[b];
            const bool inside = row_valid[b] != 0;
            const double fp = inside ? ll[b] / T : -INFINITY;
            // qadj = -(2 sum(t1 t2) + sum(t2^2)),  t1 = z / sqrt 2,  t2 = (adjrho / 2) ((dfval + dfvalp) @ hc)
            // the chain's vectors are read from global memory once (registers, d <= MAXD), hc comes from shared memory
            double gs[MAXD], zz[MAXD], dn[MAXD];
#pragma unroll
            for (int i = 0; i < MAXD; i++)
                if (i < d) {
                    dn[i] = dll[(long)b * d + i] / T;
                    gs[i] = dfval[(long)b * d + i] + dn[i];
                    zz[i] = z[(long)b * d + i];
                }
            double q1 = 0.0, q2 = 0.0;
            const double ar = adjrho[b];
            for (int j = 0; j < d; j++) {
                double t2 = 0.0;
#pragma unroll
                for (int i = 0; i < MAXD; i++)
                    if (i < d) t2 += gs[i] * s_hc[i * d + j];
                t2 *= 0.5 * ar;
                double zj = 0.0;
#pragma unroll
                for (int i = 0; i < MAXD; i++)
                    if (i == j) zj = zz[i];
                q1 += zj * 0.70710678118654752440 * t2;
                q2 += t2 * t2;
            }
            const double qadj = -(2.0 * q1 + q2);
            double u0, u1;
            uniforms(rng(klo, khi, STREAM_ACCEPT, (unsigned)b), u0, u1);
            if (!inside) nout += 1.0;
            if (inside && log(u0) < (fp - fval[b]) + qadj) {
                nacc += 1.0;
                fval[b] = fp;
#pragma unroll
                for (int j = 0; j < MAXD; j++)
                    if (j < d) {
                        thetac[(long)b * d + j] = thetap[(long)b * d + j];
                        dfval[(long)b * d + j] = dn[j];
                    }
            }
        }
        nacc = block_sum(nacc, s_red);
        nout = block_sum(nout, s_red);
        if (tid == 0) {
            scal[1] += nacc / B;
            scal[2] += nacc;
            scal[5] += nout;
        }
        // ---- temperature exchange, 5 passes of B adjacent-pair proposals (PTLMC.py:239-246, 259-273) ----
        for (int b = tid; b < B; b += ST) {
            s_lp[b] = fval[b] * temps[b];
            s_ord[b] = b;
            s_rho[b] = b > 0 ? (1.0 / temps[b - 1] - 1.0 / temps[b]) : 0.0;
        }
        for (int j = tid; j < 5 * B; j += ST) {
            double u0, u1;
            const uint4 r = rng(klo, khi, STREAM_SWAP, (unsigned)j);
            uniforms(r, u0, u1);
            s_logu[j] = log(u0);
            s_rt[j] = 1 + (int)(((unsigned long long)r.z * (unsigned)(B - 1)) >> 32);     // uniform on 1 .. B-1
        }
        __syncthreads();
        if (tid == 0 && B > 1) {
            double nsw = 0.0;
            // only the two log-posteriors depend on the previous proposal: its index, threshold and ladder gap are
            // fetched one proposal ahead so that they are off the serial chain
            int rt = s_rt[0];
            double lu = s_logu[0], rh = s_rho[rt];
            for (int j = 0; j < 5 * B; j++) {
                const int rt_n = s_rt[min(j + 1, 5 * B - 1)];
                const double lu_n = s_logu[min(j + 1, 5 * B - 1)], rh_n = s_rho[rt_n];
                const double a = s_lp[rt], b = s_lp[rt - 1];
                if ((a - b) * rh > lu) {
                    s_lp[rt] = b;
                    s_lp[rt - 1] = a;
                    const int t = s_ord[rt - 1];
                    s_ord[rt - 1] = s_ord[rt];
                    s_ord[rt] = t;
                    nsw += 1.0;
                }
                rt = rt_n;
                lu = lu_n;
                rh = rh_n;
            }
            scal[3] += nsw;
        }
        __syncthreads();
        for (int idx = tid; idx < B * d; idx += ST) {
            const int b = idx / d, j = idx % d, o = s_ord[b];
            s_th[idx] = thetac[(long)o * d + j];
            s_df[idx] = temps[o] * dfval[(long)o * d + j] / temps[b];
        }
        __syncthreads();
        for (int idx = tid; idx < B * d; idx += ST) {
            thetac[idx] = s_th[idx];
            dfval[idx] = s_df[idx];
        }
        for (int b = tid; b < B; b += ST) fval[b] = s_lp[b] / temps[b];
        __syncthreads();
        // ---- adaptation while tuning / save afterwards (PTLMC.py:247-257) ----
        if (k < tune && k % 10 == 0) {
            if (tid == 0) {
                const double tau = scal[0] + (scal[1] / 10.0 - taracc) / sqrt(1.0 + (double)k / 10.0);
                scal[0] = tau;
                scal[1] = 0.0;
                s_red[32] = 2.0 * (1.0 + tanh(tau));
                scal[4] = s_red[32];
            }
            __syncthreads();
            for (int b = tid; b < B; b += ST) adjrho[b] = s_red[32] * cbrt(temps[b]);
        } else if (k >= tune && k - tune < nsave) {
            const int nc = B - numtemps;
            for (int idx = tid; idx < nc * d; idx += ST) {
                const int c = idx / d, j = idx % d;
                save[((long)c * nsave + (k - tune)) * d + j] = thetac[(long)(numtemps + c) * d + j];
            }
        }
        __syncthreads();
    }

    if (mode & 2) {
        // ---- Langevin proposal of step kn (PTLMC.py:205-213) ----
        const long long kn = (mode & 1) ? k + 1 : k;
        const unsigned nlo = (unsigned)kn, nhi = (unsigned)(kn >> 32);
        const int npair = (d + 1) / 2;
        for (int idx = tid; idx < B * npair; idx += ST) {            // Box-Muller: two normals per Philox call
            const int b = idx / npair, pr = idx % npair;
            double u0, u1;
            uniforms(rng(nlo, nhi, STREAM_Z, (unsigned)idx), u0, u1);
            const double rad = sqrt(-2.0 * log(u0));
            double sn, cs;
            sincospi(2.0 * u1, &sn, &cs);
            z[(long)b * d + 2 * pr] = rad * cs;
            if (2 * pr + 1 < d) z[(long)b * d + 2 * pr + 1] = rad * sn;
        }
        __syncthreads();
        for (int b = tid; b < B; b += ST) {
            const double ar = adjrho[b];
            bool inside = true;
            double zz[MAXD], df[MAXD];
#pragma unroll
            for (int i = 0; i < MAXD; i++)
                if (i < d) {
                    zz[i] = z[(long)b * d + i];
                    df[i] = dfval[(long)b * d + i];
                }
            for (int j = 0; j < d; j++) {
                double zh = 0.0, dc = 0.0;
#pragma unroll
                for (int i = 0; i < MAXD; i++)
                    if (i < d) {
                        zh += zz[i] * s_hc[i * d + j];
                        dc += df[i] * s_cov[i * d + j];
                    }
                const double tp = thetac[(long)b * d + j] + 1.41421356237309504880 * ar * zh + ar * ar * dc;
                thetap[(long)b * d + j] = tp;
                inside = inside && (tp >= lo[j]) && (tp <= hi[j]);       // NaN -> outside
            }
            row_valid[b] = inside ? 1 : 0;
        }
        if (tid == 0) *kstep = kn;
    } else if (tid == 0 && (mode & 1)) {
        *kstep = k + 1;
    }
}

}  // namespace

size_t ptlmc_step_smem(int B, int d) {
    return (size_t)(2 * d * d + B + B + 5 * B + 2 * (size_t)B * d + 34) * sizeof(double) + (size_t)(B + 5 * B) * sizeof(int);
}

int ptlmc_step(int B, int d, int numtemps, int tune, int nsave, int mode, unsigned long long seed, double taracc,
               double* thetac, double* fval, double* dfval, double* thetap, double* z, double* adjrho, const double* ll,
               const double* dll, const double* temps, const double* lo, const double* hi, const double* hc,
               const double* cov0, double* scal, long long* kstep, double* save, int* row_valid, cudaStream_t stream) {
    SB_CHECK_ARG(B >= 2 && d >= 1 && d <= MAXD, "ptlmc_step: B=%d d=%d outside the supported range", B, d);
    SB_CHECK_ARG(numtemps >= 0 && numtemps < B && tune >= 0 && nsave >= 0, "ptlmc_step: bad ladder / schedule");
    const size_t smem = ptlmc_step_smem(B, d);
    SB_CHECK_ARG(smem <= 200 * 1024, "ptlmc_step: a population of %d chains (d=%d) needs %zu B of shared memory", B, d, smem);
    static DeviceOnce once;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(ptlmc_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        return 0;
    }));
    ptlmc_step_kernel<<<1, ST, smem, stream>>>(B, d, numtemps, tune, nsave, mode, seed, taracc, thetac, fval, dfval, thetap,
                                               z, adjrho, ll, dll, temps, lo, hi, hc, cov0, scal, kstep, save, row_valid);
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace sb
