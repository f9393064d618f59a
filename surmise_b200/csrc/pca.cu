// Standardisation statistics and the small symmetric eigensolver of the PCA stage of PCGPwM
// (emulationmethods/PCGPwM.py:533-605 __standardizef, 608-654 __PCs).
//
//   * column statistics of the (n x m) output matrix f with missing (non-finite) cells skipped: offset = nanmean,
//     scale = nanstd / sqrt(1 - missing fraction) (PCGPwM.py:553-557), two passes (mean, then centred squares) --
//     HBM-bound, 2 x 8 n m bytes, threads along the contiguous dimension;
//   * fs = (f - offset) / scale with the reference's alternating +2 / -2 seed in the missing cells (PCGPwM.py:565-571);
//   * a batched cyclic Jacobi eigensolver for symmetric p x p matrices, p <= 112, one CTA per matrix, matrix and
//     eigenvectors resident in shared memory: it is the Rayleigh-Ritz / orthonormalisation core of the block subspace
//     iteration that replaces the reference's full SVD (only the components above epsilonPC are ever used).
#include "pca.cuh"

namespace sb {
namespace {

constexpr int CS_T = 128;       // columns per CTA of the statistics kernels
constexpr int CS_MAXSLABS = 64;

__device__ __forceinline__ bool finite_d(double v) { return (__double2hiint(v) & 0x7ff00000) != 0x7ff00000; }

// partial column sums over one slab of rows; CENTERED: sum of (x - mean)^2, else sum of x and the count
template <bool CENTERED>
__global__ void __launch_bounds__(CS_T) col_partial_kernel(int n, int m, const double* __restrict__ f, long ldf,
                                                           const double* __restrict__ mean, int rows_per,
                                                           double* __restrict__ psum, int* __restrict__ pcnt) {
    const int col = blockIdx.x * CS_T + threadIdx.x;
    if (col >= m) return;
    const int r0 = blockIdx.y * rows_per, r1 = min(n, r0 + rows_per);
    const double mu = CENTERED ? mean[col] : 0.0;
    double s0 = 0.0, s1 = 0.0;
    int c = 0;
    int r = r0;
    for (; r + 1 < r1; r += 2) {                    // two independent chains
        const double v0 = f[(long)r * ldf + col], v1 = f[(long)(r + 1) * ldf + col];
        if (finite_d(v0)) {
            const double d0 = v0 - mu;
            s0 += CENTERED ? d0 * d0 : d0;
            c++;
        }
        if (finite_d(v1)) {
            const double d1 = v1 - mu;
            s1 += CENTERED ? d1 * d1 : d1;
            c++;
        }
    }
    if (r < r1) {
        const double v0 = f[(long)r * ldf + col];
        if (finite_d(v0)) {
            const double d0 = v0 - mu;
            s0 += CENTERED ? d0 * d0 : d0;
            c++;
        }
    }
    psum[(long)blockIdx.y * m + col] = s0 + s1;
    if (!CENTERED) pcnt[(long)blockIdx.y * m + col] = c;
}

__global__ void col_mean_kernel(int m, int slabs, const double* __restrict__ psum, const int* __restrict__ pcnt,
                                double* __restrict__ mean, int* __restrict__ cnt) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m) return;
    double s = 0.0;
    int c = 0;
    for (int k = 0; k < slabs; k++) {               // fixed order: deterministic
        s += psum[(long)k * m + col];
        c += pcnt[(long)k * m + col];
    }
    mean[col] = s / (double)c;                      // c == 0 -> NaN, like numpy's nanmean of an empty column
    cnt[col] = c;
}

__global__ void col_scale_kernel(int n, int m, int slabs, const double* __restrict__ pssq, const int* __restrict__ cnt,
                                 double* __restrict__ scale, int* __restrict__ nmiss) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m) return;
    double s = 0.0;
    for (int k = 0; k < slabs; k++) s += pssq[(long)k * m + col];
    const int c = cnt[col];
    const double nanstd = sqrt(s / (double)c);
    const double missfrac = (double)(n - c) / (double)n;
    scale[col] = nanstd / sqrt(1.0 - missfrac);     // PCGPwM.py:556
    nmiss[col] = n - c;
}

// fs = (f - offset) / scale; missing cells get 0 here and their seed in fill_missing_kernel; miss (may be NULL) = mask
__global__ void __launch_bounds__(256) standardize_kernel(int n, int m, const double* __restrict__ f, long ldf,
                                                          const double* __restrict__ offset,
                                                          const double* __restrict__ scale, double* __restrict__ fs,
                                                          long ldfs, unsigned char* __restrict__ miss) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    const int r0 = blockIdx.y * 16;
    if (col >= m) return;
    const double off = offset[col], sc = scale[col];
    for (int r = r0; r < min(n, r0 + 16); r++) {
        const double v = f[(long)r * ldf + col];
        const bool ok = finite_d(v);
        fs[(long)r * ldfs + col] = ok ? (v - off) / sc : 0.0;
        if (miss) miss[(long)r * m + col] = ok ? 0 : 1;
    }
}

// the k-th missing cell of a column (in row order) is seeded with +2, -2, +2, ... minus the mean of that pattern
__global__ void fill_missing_kernel(int n, int m, double* __restrict__ fs, long ldfs,
                                    const unsigned char* __restrict__ miss, const int* __restrict__ nmiss) {
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= m) return;
    const int c = nmiss[col];
    if (c == 0) return;
    const double mu = (c & 1) ? 2.0 / (double)c : 0.0;
    int k = 0;
    for (int r = 0; r < n; r++)
        if (miss[(long)r * m + col]) {
            fs[(long)r * ldfs + col] = ((k & 1) ? -2.0 : 2.0) - mu;
            k++;
        }
}

// sum of squares, two deterministic stages
__global__ void __launch_bounds__(256) sumsq_partial_kernel(long count, const double* __restrict__ x,
                                                            double* __restrict__ part) {
    __shared__ double red[32];
    double s = 0.0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < count; i += (long)gridDim.x * 256) {
        const double v = x[i];
        s += v * v;
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) part[blockIdx.x] = s;
}
__global__ void __launch_bounds__(256) sumsq_final_kernel(int parts, const double* __restrict__ part,
                                                          double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < parts; i += 256) s += part[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = s;
}

// ---- cyclic Jacobi, round-robin ordering: P/2 disjoint rotations per round, P - 1 rounds per sweep -------------
constexpr int JT = 1024;

__global__ void __launch_bounds__(JT) syevj_kernel(int p, const double* __restrict__ Hb, long ldh, long strideH,
                                                   double* __restrict__ Wb, long strideW, double* __restrict__ Zb,
                                                   long ldz, long strideZ, int max_sweeps, int* __restrict__ sweeps_out) {
    extern __shared__ double sm[];
    const int P = (p + 1) & ~1;          // an even number of players; index p (if any) is an idle padding row
    const int ld = P | 1;                // odd leading dimension: column walks are bank-conflict free
    double* A = sm;                      // [P][ld]
    double* VT = A + (long)P * ld;       // [P][ld]  VT[i][r] = component r of eigenvector i
    double* cs = VT + (long)P * ld;      // [P/2][2]
    int* pi = (int*)(cs + P);            // [P/2]
    int* pj = pi + P / 2;                // [P/2]
    __shared__ int s_rot;
    __shared__ double s_scale;
    const int tid = threadIdx.x;
    const double* H = Hb + (long)blockIdx.x * strideH;
    for (int idx = tid; idx < P * P; idx += JT) {
        const int i = idx / P, j = idx % P;
        A[i * ld + j] = (i < p && j < p) ? 0.5 * (H[(long)i * ldh + j] + H[(long)j * ldh + i]) : 0.0;
        VT[i * ld + j] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double mx = 0.0;
        for (int i = 0; i < p; i++) mx = fmax(mx, fabs(A[i * ld + i]));
        s_scale = mx;
    }
    __syncthreads();
    // an off-diagonal element is rotated away while it exceeds eps sqrt(a_ii a_jj) (relative accuracy for the small
    // eigenvalues of a positive definite matrix) and an absolute floor far below the matrix scale
    const double EPS = 1.1102230246251565e-16, floor_abs = 1e-26 * s_scale;
    // One warp per rotation: it derives (c, s) of its pair from the three elements it is about to change (no other
    // warp touches columns i, j in this phase), rotates the two columns, and leaves (c, s) in shared memory for the
    // row phase: two barriers per round.
    const int warp = tid >> 5, lane = tid & 31, nwarps = JT / 32;
    int sweep = 0;
    for (; sweep < max_sweeps; sweep++) {
        __syncthreads();
        if (tid == 0) s_rot = 0;
        for (int r = 0; r < P - 1; r++) {
            __syncthreads();
            for (int k = warp; k < P / 2; k += nwarps) {                    // A <- A J  (columns i, j of every row)
                int i, j;
                if (k == 0) {
                    i = P - 1;
                    j = r;
                } else {
                    i = r + k;
                    if (i >= P - 1) i -= P - 1;
                    j = r - k;
                    if (j < 0) j += P - 1;
                }
                if (i > j) {
                    const int t = i;
                    i = j;
                    j = t;
                }
                const double aij = A[i * ld + j], aii = A[i * ld + i], ajj = A[j * ld + j];
                double c = 1.0, sn = 0.0;
                if (fabs(aij) > EPS * sqrt(fabs(aii) * fabs(ajj)) && fabs(aij) > floor_abs) {
                    const double tau = (ajj - aii) / (2.0 * aij);
                    const double t = copysign(1.0, tau) / (fabs(tau) + sqrt(1.0 + tau * tau));
                    c = rsqrt(1.0 + t * t);
                    sn = t * c;
                }
                __syncwarp();                                                // everyone has read the pair's elements
                if (lane == 0) {
                    cs[2 * k] = c;
                    cs[2 * k + 1] = sn;
                    pi[k] = i;
                    pj[k] = j;
                    if (sn != 0.0) s_rot = 1;
                }
                if (sn != 0.0)
                    for (int row = lane; row < P; row += 32) {
                        const double x = A[row * ld + i], y = A[row * ld + j];
                        A[row * ld + i] = c * x - sn * y;
                        A[row * ld + j] = sn * x + c * y;
                    }
            }
            __syncthreads();
            for (int k = warp; k < P / 2; k += nwarps) {                    // A <- J' A, V <- V J (rows i, j)
                const double sn = cs[2 * k + 1];
                if (sn == 0.0) continue;
                const double c = cs[2 * k];
                const int i = pi[k], j = pj[k];
                for (int col = lane; col < P; col += 32) {
                    double x = A[i * ld + col], y = A[j * ld + col];
                    double xi = c * x - sn * y, yj = sn * x + c * y;
                    if (col == j) xi = 0.0;                                  // the annihilated pair, exactly
                    if (col == i) yj = 0.0;
                    A[i * ld + col] = xi;
                    A[j * ld + col] = yj;
                    x = VT[i * ld + col];
                    y = VT[j * ld + col];
                    VT[i * ld + col] = c * x - sn * y;
                    VT[j * ld + col] = sn * x + c * y;
                }
            }
        }
        __syncthreads();
        if (!s_rot) break;
    }
    __syncthreads();
    // descending order: rank of eigenvalue i among the p of them (ties broken by index)
    double* W = Wb + (long)blockIdx.x * strideW;
    double* Z = Zb + (long)blockIdx.x * strideZ;
    for (int i = tid; i < p; i += JT) {
        const double wi = A[i * ld + i];
        int rank = 0;
        for (int j = 0; j < p; j++) {
            const double wj = A[j * ld + j];
            rank += (wj > wi || (wj == wi && j < i)) ? 1 : 0;
        }
        pi[i] = rank;                    // p <= P, and the pair tables are free now (P/2 + P/2 ints)
        W[rank] = wi;
    }
    __syncthreads();
    for (int idx = tid; idx < p * p; idx += JT) {
        const int i = idx / p, r = idx % p;
        Z[(long)r * ldz + pi[i]] = VT[i * ld + r];
    }
    if (tid == 0 && sweeps_out) sweeps_out[blockIdx.x] = sweep;
}

size_t syevj_smem(int p) {
    const int P = (p + 1) & ~1, ld = P | 1;
    return (size_t)(2 * (long)P * ld + P) * sizeof(double) + (size_t)P * sizeof(int);
}

}  // namespace

int pca_col_stats_slabs(int n) { return (int)std::min<long>(CS_MAXSLABS, std::max<long>(1, cdiv(n, 64))); }

size_t pca_col_stats_workspace(int n, int m) {
    Workspace w(nullptr, 0);
    const int slabs = pca_col_stats_slabs(n);
    w.take<double>((size_t)slabs * m);
    w.take<int>((size_t)slabs * m);
    w.take<int>((size_t)m);
    return w.off;
}

int pca_col_stats(int n, int m, const double* f, long ldf, double* offset, double* scale, int* nmiss, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream) {
    SB_CHECK_ARG(n > 0 && m > 0 && ldf >= m, "col_stats: n=%d m=%d ldf=%ld", n, m, ldf);
    SB_CHECK_ARG(workspace_bytes >= pca_col_stats_workspace(n, m), "col_stats: workspace too small");
    Workspace w(workspace, workspace_bytes);
    const int slabs = pca_col_stats_slabs(n);
    const int rows_per = cdiv(n, slabs);
    double* psum = w.take<double>((size_t)slabs * m);
    int* pcnt = w.take<int>((size_t)slabs * m);
    int* cnt = w.take<int>((size_t)m);
    dim3 grid(cdiv(m, CS_T), slabs);
    col_partial_kernel<false><<<grid, CS_T, 0, stream>>>(n, m, f, ldf, nullptr, rows_per, psum, pcnt);
    SB_LAUNCH_CHECK();
    col_mean_kernel<<<cdiv(m, 128), 128, 0, stream>>>(m, slabs, psum, pcnt, offset, cnt);
    SB_LAUNCH_CHECK();
    col_partial_kernel<true><<<grid, CS_T, 0, stream>>>(n, m, f, ldf, offset, rows_per, psum, nullptr);
    SB_LAUNCH_CHECK();
    col_scale_kernel<<<cdiv(m, 128), 128, 0, stream>>>(n, m, slabs, psum, cnt, scale, nmiss);
    SB_LAUNCH_CHECK();
    return 0;
}

int pca_standardize(int n, int m, const double* f, long ldf, const double* offset, const double* scale, double* fs,
                    long ldfs, unsigned char* miss, const int* nmiss, cudaStream_t stream) {
    SB_CHECK_ARG(n > 0 && m > 0 && ldf >= m && ldfs >= m, "standardize: n=%d m=%d", n, m);
    SB_CHECK_ARG(miss == nullptr || nmiss != nullptr, "standardize: the missing mask needs the per-column counts");
    dim3 grid(cdiv(m, 256), cdiv(n, 16));
    standardize_kernel<<<grid, 256, 0, stream>>>(n, m, f, ldf, offset, scale, fs, ldfs, miss);
    SB_LAUNCH_CHECK();
    if (miss) {
        fill_missing_kernel<<<cdiv(m, 128), 128, 0, stream>>>(n, m, fs, ldfs, miss, nmiss);
        SB_LAUNCH_CHECK();
    }
    return 0;
}

int pca_sumsq(long count, const double* x, double* out, double* scratch, cudaStream_t stream) {
    SB_CHECK_ARG(count > 0, "sumsq: empty input");
    const int parts = (int)std::min<long>(PCA_SUMSQ_PARTS, (count + 255) / 256);
    sumsq_partial_kernel<<<parts, 256, 0, stream>>>(count, x, scratch);
    SB_LAUNCH_CHECK();
    sumsq_final_kernel<<<1, 256, 0, stream>>>(parts, scratch, out);
    SB_LAUNCH_CHECK();
    return 0;
}

int syevj_batched(int batch, int p, const double* H, long ldh, long strideH, double* W, long strideW, double* Z,
                  long ldz, long strideZ, int max_sweeps, int* sweeps, cudaStream_t stream) {
    SB_CHECK_ARG(batch > 0 && p >= 1 && p <= SYEVJ_MAXP, "syevj: batch=%d p=%d (p <= %d)", batch, p, SYEVJ_MAXP);
    SB_CHECK_ARG(ldh >= p && ldz >= p, "syevj: leading dimensions");
    static DeviceOnce once;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(syevj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syevj_smem(SYEVJ_MAXP)));
        return 0;
    }));
    syevj_kernel<<<batch, JT, syevj_smem(p), stream>>>(p, H, ldh, strideH, W, strideW, Z, ldz, strideZ,
                                                      max_sweeps > 0 ? max_sweeps : 40, sweeps);
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace sb
