// Woodbury-form Gaussian log-likelihood (+ d/dtheta) for a whole batch of theta proposals in one
// launch: the PC-space restatement of directbayeswoodbury.loglik / loglik_grad
// (src/surmise/calibrationmethods/directbayeswoodbury.py:261-383).
//
//   v = predvars[b], pv = predvecs[b], D = diag(sqrt v), G = Phi' Sigma^-1 Phi (theta independent)
//   r = (y - offset) - Phi pv   (kept in m-space: expanding r' Sigma^-1 r in k-space cancels badly)
//   q = Phi' Sigma^-1 r,  A = I + D G D,  l = -1/2 log|A| - 1/2 (r' Sigma^-1 r - (Dq)' A^-1 (Dq))
// One CTA per theta; work per theta O(m k + k^3 + d k^2); nothing of size m x B x k x d exists.
// The batch-wide "extravar" trigger (dbw.py:277-284) is a pre-pass that ORs a device flag; the
// main kernel then selects Sigma = yvar or yvar + |extravar| (both Gram matrices are precomputed).
#include "gp.cuh"
#include <algorithm>
#include <atomic>

namespace sb {
namespace {

constexpr int WT = 128;
constexpr size_t WOODBURY_SCRATCH_BYTES = (size_t)1 << 30;      // cap of the global-scratch path's workspace
std::atomic<bool> g_woodbury_force_global{false};      // testing switch: the global-scratch path at any k

// Component a of proposal b in a (possibly) segmented (B, K) table: components arrive in groups of `segk` (one group
// per rank of a PC-sharded prediction, each a dense (B, segk) block, `segstride` doubles apart -- the layout an
// all-gather of the ranks' blocks produces).  segk == K, segstride == 0 is the plain dense table.
__device__ __forceinline__ long seg_at(int b, int a, int segk, long segstride) {
    return (long)(a / segk) * segstride + (long)b * segk + (a % segk);
}
// With a position table (seg_pos[a] = where component a sits among the padded, rank-ordered components) the padded
// slots are never visited: the k x k work stays that of the real components, however many ranks share them.
__device__ __forceinline__ int seg_idx(const int* __restrict__ seg_pos, int a) { return seg_pos ? seg_pos[a] : a; }

// Entry i of the d-vector of component a of proposal b in a segmented (B, K, d) table: the segment offset counts
// doubles (it is the same block stride as for the scalar tables), only the position inside a block scales with d.
__device__ __forceinline__ long seg_at_d(int b, int a, int i, int d, int segk, long segstride) {
    return (long)(a / segk) * segstride + ((long)b * segk + (a % segk)) * d + i;
}

__global__ void __launch_bounds__(WT) woodbury_trigger_kernel(int B, int K, int m, const double* __restrict__ predvars,
                                                              const double* __restrict__ phiT,
                                                              const double* __restrict__ extravar,
                                                              const int* __restrict__ row_valid, int segk,
                                                              long segstride, const int* __restrict__ seg_pos, int* fired) {
    const int b = blockIdx.x;
    if (row_valid && !row_valid[b]) return;          // rows the caller will discard (prior = -inf) take no part in the test
    extern __shared__ double sh[];
    double* v = sh;
    for (int a = threadIdx.x; a < K; a += WT) v[a] = predvars[seg_at(b, seg_idx(seg_pos, a), segk, segstride)];
    __syncthreads();
    bool hit = false;
    for (int x = threadIdx.x; x < m; x += WT) {
        double s = 0.0;
        for (int a = 0; a < K; a++) {
            double ph = phiT[(long)a * m + x];
            s += v[a] * ph * ph;                                   // sum_k covxhalf^2
        }
        double var = extravar[x] + s;                               // emulator variance at (x, theta_b)
        if (fabs(var / (1e-4 + (1.0 + 1e-4) * s)) > 1.0) hit = true;
    }
    if (__syncthreads_or(hit) && threadIdx.x == 0) atomicOr(fired, 1);
}

__global__ void __launch_bounds__(WT) woodbury_loglik_kernel(int B, int K, int m, int d,
                                                             const double* __restrict__ predvecs,
                                                             const double* __restrict__ predvars,
                                                             const double* __restrict__ dpv,
                                                             const double* __restrict__ dpvar,
                                                             const double* __restrict__ phiT,
                                                             const double* __restrict__ resid0,
                                                             const double* __restrict__ sinvb,
                                                             const double* __restrict__ Gb, int* fired,
                                                             int forced_variant, int want_grad, int segk, long segstride,
                                                             const int* __restrict__ seg_pos, double* __restrict__ ll,
                                                             double* __restrict__ dll, double* gws, int b0) {
    extern __shared__ double sh[];
    // The two k x k matrices live in shared memory when they fit (k <= WOODBURY_SMEM_K); beyond that each CTA works on
    // its own slice of a global scratch buffer (L2 resident), with the same code -- block-level barriers order both.
    double* Lm = gws ? gws + (size_t)blockIdx.x * 2 * K * K : sh;      // K*K : A, then L, then Ainv
    double* Xm = Lm + K * K;                                           // K*K : L^-1
    double* w = gws ? sh : Xm + K * K;                                 // m   : Sigma^-1 r
    double* D = w + m;                  // K each below
    double* pvv = D + K;
    double* q = pvv + K;
    double* Dq = q + K;
    double* J3 = Dq + K;
    double* Nv = J3 + K;
    double* MJ3 = Nv + K;
    double* red = MJ3 + K;              // 32
    const int b = blockIdx.x + b0, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int variant = forced_variant >= 0 ? forced_variant : ((*fired) ? 1 : 0);
    if (forced_variant >= 0 && b == 0 && threadIdx.x == 0) *fired = forced_variant;
    const double* sinv = sinvb + (long)variant * m;
    const double* G = Gb + (long)variant * K * K;

    for (int a = tid; a < K; a += WT) {
        D[a] = sqrt(predvars[seg_at(b, seg_idx(seg_pos, a), segk, segstride)]);
        pvv[a] = predvecs[seg_at(b, seg_idx(seg_pos, a), segk, segstride)];
    }
    __syncthreads();
    double t1 = 0.0;
    for (int x = tid; x < m; x += WT) {
        double r = resid0[x];
        for (int a = 0; a < K; a++) r -= phiT[(long)a * m + x] * pvv[a];
        double wx = sinv[x] * r;
        w[x] = wx;
        t1 += r * wx;
    }
    const double term1 = block_sum(t1, red);                       // r' Sigma^-1 r   (also publishes w)
    for (int a = warp; a < K; a += WT / 32) {
        double s = 0.0;
        for (int x = lane; x < m; x += 32) s += phiT[(long)a * m + x] * w[x];
        s = warp_sum(s);
        if (lane == 0) {
            q[a] = s;
            Dq[a] = D[a] * s;
        }
    }
    for (int idx = tid; idx < K * K; idx += WT) {
        int i = idx / K, j = idx % K;
        Lm[idx] = D[i] * G[idx] * D[j] + (i == j ? 1.0 : 0.0);
        Xm[idx] = 0.0;
    }
    // Cholesky of A (k x k), lower, in shared memory
    for (int j = 0; j < K; j++) {
        __syncthreads();
        double ljj = sqrt(Lm[j * K + j]);
        __syncthreads();
        for (int i = j + tid; i < K; i += WT) Lm[i * K + j] = (i == j) ? ljj : Lm[i * K + j] / ljj;
        __syncthreads();
        int rem = K - j - 1;
        for (int idx = tid; idx < rem * rem; idx += WT) {
            int i = j + 1 + idx / rem, c = j + 1 + idx % rem;
            if (c <= i) Lm[i * K + c] -= Lm[i * K + j] * Lm[c * K + j];
        }
    }
    __syncthreads();
    double ldh = 0.0;
    for (int a = tid; a < K; a += WT) ldh += log(Lm[a * K + a]);
    const double logdet_half = block_sum(ldh, red);
    // X = L^-1: each thread owns one column (no cross-thread dependence)
    for (int c = tid; c < K; c += WT) {
        Xm[c * K + c] = 1.0 / Lm[c * K + c];
        for (int i = c + 1; i < K; i++) {
            double s = 0.0;
            for (int j = c; j < i; j++) s += Lm[i * K + j] * Xm[j * K + c];
            Xm[i * K + c] = -s / Lm[i * K + i];
        }
    }
    __syncthreads();
    // Ainv = X' X  (full symmetric) into Lm
    for (int idx = tid; idx < K * K; idx += WT) {
        int a = idx / K, c = idx % K;
        int lo = a > c ? a : c;
        double s = 0.0;
        for (int i = lo; i < K; i++) s += Xm[i * K + a] * Xm[i * K + c];
        Lm[idx] = s;
    }
    __syncthreads();
    double qd = 0.0;
    for (int a = tid; a < K; a += WT) {
        double s = 0.0;
        for (int c = 0; c < K; c++) s += Lm[a * K + c] * Dq[c];
        J3[a] = s;
        qd += Dq[a] * s;
    }
    const double quad = block_sum(qd, red);                        // (Dq)' A^-1 (Dq)
    if (tid == 0) ll[b] = -logdet_half - 0.5 * (term1 - quad);     // dbw.py:303-304
    if (!want_grad) return;

    for (int a = tid; a < K; a += WT) {
        double sN = 0.0, sM = 0.0;
        for (int c = 0; c < K; c++) {
            double mac = G[a * K + c] * D[c];                      // (G D)_ac
            sN += mac * Lm[c * K + a];
            sM += mac * J3[c];
        }
        Nv[a] = sN;
        MJ3[a] = sM;
    }
    __syncthreads();
    for (int i = 0; i < d; i++) {
        double val = 0.0;
        for (int a = tid; a < K; a += WT) {
            double dpa = dpv[seg_at_d(b, seg_idx(seg_pos, a), i, d, segk, segstride)];
            double dD = 0.5 * dpvar[seg_at_d(b, seg_idx(seg_pos, a), i, d, segk, segstride)] / D[a];
            double gd = 0.0;
            for (int c = 0; c < K; c++) gd += G[a * K + c] * dpv[seg_at_d(b, seg_idx(seg_pos, c), i, d, segk, segstride)];
            double dJ2 = -D[a] * gd + dD * q[a];
            // -1/2 dterm3 - 1/2 (dterm1 - dterm2), term by term (dbw.py:344-381 in k-space)
            val += -dD * Nv[a] + q[a] * dpa + dJ2 * J3[a] - dD * J3[a] * MJ3[a];
        }
        double s = block_sum(val, red);
        if (tid == 0) dll[(long)b * d + i] = s;
    }
}


// ---- generic form: any emulator's predictive mean / covxhalf (directbayeswoodbury.py:261-306 as written) -------------
// For an emulator outside the PCGPwM family there is no PC-space structure to exploit: the caller evaluates the
// emulator (its own predict) and hands over mean (B, m), var (B, m) and covxhalf (B, m, K) at the proposals.
//   r = y - mean_b,  J = Sigma^-1/2 C_b,  l = -1/2 (r' Sigma^-1 r - J2' (I + J'J)^-1 J2) - 1/2 log|I + J'J|,  J2 = J' Sigma^-1/2 r
// One CTA per proposal; I + J'J (K x K) and its Cholesky factor in shared memory.
__global__ void __launch_bounds__(WT) mspace_trigger_kernel(int B, int K, int m, const double* __restrict__ varT,
                                                            const double* __restrict__ cxhT, int* fired) {
    const int b = blockIdx.x;
    bool hit = false;
    for (int x = threadIdx.x; x < m; x += WT) {
        const double* c = cxhT + ((long)b * m + x) * K;
        double s = 0.0;
        for (int a = 0; a < K; a++) s += c[a] * c[a];
        if (fabs(varT[(long)b * m + x] / (1e-4 + (1.0 + 1e-4) * s)) > 1.0) hit = true;      // dbw.py:277-279
    }
    if (__syncthreads_or(hit) && threadIdx.x == 0) atomicOr(fired, 1);
}

__global__ void __launch_bounds__(WT) mspace_loglik_kernel(int B, int K, int m, const double* __restrict__ meanT,
                                                           const double* __restrict__ cxhT, const double* __restrict__ y,
                                                           const double* __restrict__ sinv, double* __restrict__ ll) {
    extern __shared__ double sh[];
    double* Am = sh;                    // K*K : I + J'J, then its Cholesky factor (lower)
    double* w = Am + K * K;             // m   : Sigma^-1 r
    double* J2 = w + m;                 // K   : C' Sigma^-1 r, then L^-1 of it
    double* red = J2 + K;               // 32
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* mu = meanT + (long)b * m;
    const double* C = cxhT + (long)b * m * K;
    double t1 = 0.0;
    for (int x = tid; x < m; x += WT) {
        const double r = y[x] - mu[x];
        const double wx = sinv[x] * r;
        w[x] = wx;
        t1 += r * wx;
    }
    const double term1 = block_sum(t1, red);                       // r' Sigma^-1 r   (also publishes w)
    for (int a = warp; a < K; a += WT / 32) {
        double s = 0.0;
        for (int x = lane; x < m; x += 32) s += C[(long)x * K + a] * w[x];
        s = warp_sum(s);
        if (lane == 0) J2[a] = s;
    }
    for (int idx = tid; idx < K * K; idx += WT) {
        const int i = idx / K, j = idx % K;
        double s = 0.0;
        if (j <= i)
            for (int x = 0; x < m; x++) s += C[(long)x * K + i] * sinv[x] * C[(long)x * K + j];
        Am[idx] = (j <= i) ? s + (i == j ? 1.0 : 0.0) : 0.0;
    }
    for (int j = 0; j < K; j++) {                                  // Cholesky, lower, in shared memory
        __syncthreads();
        const double ljj = sqrt(Am[j * K + j]);
        __syncthreads();
        for (int i = j + tid; i < K; i += WT) Am[i * K + j] = (i == j) ? ljj : Am[i * K + j] / ljj;
        __syncthreads();
        const int rem = K - j - 1;
        for (int idx = tid; idx < rem * rem; idx += WT) {
            const int i = j + 1 + idx / rem, c = j + 1 + idx % rem;
            if (c <= i) Am[i * K + c] -= Am[i * K + j] * Am[c * K + j];
        }
    }
    __syncthreads();
    double ldh = 0.0;
    for (int a = tid; a < K; a += WT) ldh += log(Am[a * K + a]);
    const double logdet_half = block_sum(ldh, red);                // 1/2 log|I + J'J|
    for (int j = 0; j < K; j++) {                                  // z = L^-1 J2 by forward substitution, in place
        __syncthreads();
        const double zj = J2[j] / Am[j * K + j];
        __syncthreads();
        if (tid == 0) J2[j] = zj;
        for (int i = j + 1 + tid; i < K; i += WT) J2[i] -= Am[i * K + j] * zj;
    }
    __syncthreads();
    double q = 0.0;
    for (int a = tid; a < K; a += WT) q += J2[a] * J2[a];
    const double term2 = block_sum(q, red);                        // J2' (I + J'J)^-1 J2
    if (tid == 0) ll[b] = -0.5 * (term1 - term2) - logdet_half;    // dbw.py:303-304
}

}  // namespace

// k x k matrices in shared memory?  (m-vector + 7 k-vectors + the two matrices within 227 KB: k <= 112 at m <= 3000)
static bool woodbury_fits_smem(int K, int m) {
    return (size_t)(2 * (size_t)K * K + m + 7 * K + 32) * sizeof(double) <= 227 * 1024 && !g_woodbury_force_global;
}

size_t woodbury_workspace(int B, int K, int m) {
    if (B <= 0 || K <= 0 || woodbury_fits_smem(K, m)) return 0;
    const size_t kk = (size_t)2 * K * K * sizeof(double);
    return kk * std::min<size_t>((size_t)B, std::max<size_t>(1, WOODBURY_SCRATCH_BYTES / kk));
}

int woodbury_loglik(int B, int K, int m, int d, const double* predvecs, const double* predvars, const double* dpv,
                    const double* dpvar, const double* phiT, const double* resid0, const double* extravar,
                    const double* sinv, const double* G, int want_grad, int variant, double* ll, double* dll,
                    int* fired, const int* row_valid, int segk, long segstride, const int* seg_pos, void* ws,
                    size_t ws_bytes, cudaStream_t stream) {
    SB_CHECK_ARG(B > 0 && m > 0, "woodbury: empty batch");
    if (segk <= 0 || (segk >= K && !seg_pos)) {
        segk = K;
        segstride = 0;
        seg_pos = nullptr;
    }
    SB_CHECK_ARG(seg_pos || K % segk == 0, "woodbury: k=%d is not a multiple of the segment width %d", K, segk);
    SB_CHECK_ARG(K >= 1, "woodbury: k=%d", K);
    SB_CHECK_ARG(!want_grad || (dpv && dpvar && dll), "woodbury: gradient buffers are NULL");
    const size_t smem_vec = (size_t)(m + 7 * K + 32) * sizeof(double), kk = (size_t)2 * K * K * sizeof(double);
    const bool in_smem = woodbury_fits_smem(K, m);
    const size_t smem = in_smem ? smem_vec + kk : smem_vec;
    SB_CHECK_ARG(smem <= 227 * 1024, "woodbury: k=%d, m=%d need %zu B of shared memory for the m- and k-vectors", K, m, smem);
    SB_CHECK_ARG(in_smem || (ws && ws_bytes >= kk),
                 "woodbury: k=%d, m=%d keep their k x k systems in global scratch: pass a workspace of "
                 "sb200_woodbury_workspace(B, k, m) bytes (sb200_woodbury_loglik_ws)", K, m);
    static DeviceOnce once;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(woodbury_loglik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        return 0;
    }));
    if (variant < 0) {
        SB_CUDA(cudaMemsetAsync(fired, 0, sizeof(int), stream));
        woodbury_trigger_kernel<<<B, WT, K * sizeof(double), stream>>>(B, K, m, predvars, phiT, extravar, row_valid, segk,
                                                                       segstride, seg_pos, fired);
        SB_LAUNCH_CHECK();
    }
    // in shared memory: one launch; in global scratch: the proposals in chunks that fit the caller's workspace
    const int chunk = in_smem ? B : (int)std::min<size_t>((size_t)B, ws_bytes / kk);
    for (int b0 = 0; b0 < B; b0 += chunk) {
        woodbury_loglik_kernel<<<std::min(chunk, B - b0), WT, smem, stream>>>(
            B, K, m, d, predvecs, predvars, dpv, dpvar, phiT, resid0, sinv, G, fired, variant > 0 ? 1 : variant, want_grad,
            segk, segstride, seg_pos, ll, dll, in_smem ? nullptr : static_cast<double*>(ws), b0);
        SB_LAUNCH_CHECK();
    }
    return 0;
}

int woodbury_loglik_mspace(int B, int K, int m, const double* meanT, const double* varT, const double* cxhT,
                           const double* y, const double* sinv, int want_trigger, double* ll, int* fired,
                           cudaStream_t stream) {
    SB_CHECK_ARG(B > 0 && m > 0 && K >= 1, "woodbury_mspace: B=%d m=%d k=%d", B, m, K);
    SB_CHECK_ARG(B <= 65535 * 1024, "woodbury_mspace: batch too large");
    if (want_trigger) {                       // only the batch-wide unmodelled-variance test (dbw.py:277-279)
        SB_CHECK_ARG(varT && cxhT && fired, "woodbury_mspace: NULL pointer");
        SB_CUDA(cudaMemsetAsync(fired, 0, sizeof(int), stream));
        mspace_trigger_kernel<<<B, WT, 0, stream>>>(B, K, m, varT, cxhT, fired);
        SB_LAUNCH_CHECK();
        return 0;
    }
    SB_CHECK_ARG(meanT && cxhT && y && sinv && ll, "woodbury_mspace: NULL pointer");
    const size_t smem = ((size_t)K * K + m + K + 32) * sizeof(double);
    SB_CHECK_ARG(smem <= 227 * 1024, "woodbury_mspace: k=%d, m=%d need %zu B of shared memory", K, m, smem);
    static DeviceOnce once;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(mspace_loglik_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        return 0;
    }));
    mspace_loglik_kernel<<<B, WT, smem, stream>>>(B, K, m, meanT, cxhT, y, sinv, ll);
    SB_LAUNCH_CHECK();
    return 0;
}

void woodbury_force_global(int on) { g_woodbury_force_global.store(on != 0); }

int woodbury_trigger(int B, int K, int m, const double* predvars, const double* phiT, const double* extravar,
                     int* fired, const int* row_valid, int segk, long segstride, const int* seg_pos, cudaStream_t stream) {
    SB_CHECK_ARG(B > 0 && m > 0 && K >= 1, "woodbury_trigger: empty batch");
    if (segk <= 0 || (segk >= K && !seg_pos)) {
        segk = K;
        segstride = 0;
        seg_pos = nullptr;
    }
    SB_CHECK_ARG(seg_pos || K % segk == 0, "woodbury_trigger: k=%d is not a multiple of the segment width %d", K, segk);
    SB_CUDA(cudaMemsetAsync(fired, 0, sizeof(int), stream));
    woodbury_trigger_kernel<<<B, WT, K * sizeof(double), stream>>>(B, K, m, predvars, phiT, extravar, row_valid, segk,
                                                                   segstride, seg_pos, fired);
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace sb
