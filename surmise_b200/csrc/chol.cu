// Batched recursive Cholesky / triangular inverse / L^-T L^-1 on top of the DMMA GEMM.
//
// Column-recursive formulation on a "tall" matrix (mrows >= n): factor the left half (which also
// solves every row below it), one DMMA rank-n1 update of the right half, recurse.  Almost all flops
// land in GEMMs whose K grows with the recursion level, so the trailing updates run at GEMM speed.
// Rows n..mrows-1 are right-hand sides: after the call row n holds (L^-1 g)' when it held g'.
#include "chol.cuh"
#include "dgemm.cuh"

namespace sb {
namespace {

constexpr int NB = CHOL_NB;
constexpr int LDS_ = NB + 1;   // padded shared-memory leading dimension

// One CTA per matrix: unblocked Cholesky of an nb x nb (nb <= 64) diagonal block held in shared
// memory, followed by its explicit inverse.  Only the lower triangle of A is read or written.
__global__ void __launch_bounds__(256) potrf_leaf_kernel(double* __restrict__ Abase, long lda, long strideA,
                                                         double* __restrict__ Dbase, long ldd, long strideD, int nb,
                                                         int j0, int* __restrict__ info) {
    extern __shared__ double sm[];
    double* S = sm;                    // [NB][LDS_]
    double* X = sm + NB * LDS_;        // [NB][LDS_]
    double* diag = X + NB * LDS_;      // [NB]
    const int tid = threadIdx.x, z = blockIdx.x;
    double* A = Abase + (long)z * strideA;
    double* D = Dbase + (long)z * strideD;

    for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        int i = idx / nb, j = idx % nb;
        S[i * LDS_ + j] = (j <= i) ? A[(long)i * lda + j] : 0.0;
        X[i * LDS_ + j] = 0.0;
    }
    int bad_at = 0;
    const int ti = tid >> 4, tc = tid & 15;
    for (int j = 0; j < nb; j++) {
        __syncthreads();
        double ajj = S[j * LDS_ + j];
        bool bad = !(ajj > 0.0);
        if (bad && bad_at == 0) bad_at = j0 + j + 1;
        double ljj = sqrt(bad ? 1.0 : ajj);
        double inv = 1.0 / ljj;
        if (tid == j) diag[j] = ljj;
        if (tid > j && tid < nb) S[tid * LDS_ + j] *= inv;
        __syncthreads();
        for (int i = j + 1 + ti; i < nb; i += 16) {
            double lij = S[i * LDS_ + j];
            for (int c = j + 1 + tc; c <= i; c += 16) S[i * LDS_ + c] -= lij * S[c * LDS_ + j];
        }
    }
    __syncthreads();
    if (tid < nb) S[tid * LDS_ + tid] = diag[tid];
    if (tid == 0 && bad_at != 0 && info[z] == 0) info[z] = bad_at;
    __syncthreads();

    // X = L^-1 by forward substitution, 4 threads per column sharing the dot product
    const int c = tid >> 2, q = tid & 3;
    for (int i = 0; i < nb; i++) {
        double s = 0.0;
        if (c < i && c < nb)
            for (int j = c + q; j < i; j += 4) s += S[i * LDS_ + j] * X[j * LDS_ + c];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0 && c < nb) {
            if (c == i) X[i * LDS_ + c] = 1.0 / diag[i];
            else if (c < i) X[i * LDS_ + c] = -s / diag[i];
        }
        __syncthreads();
    }
    for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
        int i = idx / nb, j = idx % nb;
        if (j <= i) A[(long)i * lda + j] = S[i * LDS_ + j];
        D[(long)i * ldd + j] = X[i * LDS_ + j];
    }
}

int leaf(const CholBatch& c, int j0, int nb, cudaStream_t stream) {
    constexpr size_t smem = (size_t)(2 * NB * LDS_ + NB) * sizeof(double);
    static bool configured[64] = {false};
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !configured[dev]) {
        SB_CUDA(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = true;
    }
    potrf_leaf_kernel<<<c.batch, 256, smem, stream>>>(c.A + (long)j0 * c.lda + j0, c.lda, c.strideA,
                                                     c.Dinv + (long)(j0 / NB) * c.blockD, c.ldd, c.strideD, nb, j0,
                                                     c.info);
    SB_LAUNCH_CHECK();
    return 0;
}

int split_point(int nn) {      // multiple of NB closest to nn/2 from above
    int blocks = (nn + NB - 1) / NB;
    return ((blocks + 1) / 2) * NB;
}

int potrf_rec(const CholBatch& c, int j0, int nn, cudaStream_t stream) {
    if (nn <= NB) {
        SB_TRY(leaf(c, j0, nn, stream));
        int below = c.mrows - (j0 + nn);
        if (below > 0) {                       // panel solve, in place:  P <- P * Dinv'
            GemmArgs g{};
            g.M = below; g.N = nn; g.K = nn;
            g.A = c.A + (long)(j0 + nn) * c.lda + j0; g.lda = c.lda; g.strideA = c.strideA; g.a_kc = true;
            g.B = c.Dinv + (long)(j0 / NB) * c.blockD; g.ldb = c.ldd; g.strideB = c.strideD; g.b_kc = true;
            g.C = c.A + (long)(j0 + nn) * c.lda + j0; g.ldc = c.lda; g.strideC = c.strideA;
            g.alpha = 1.0; g.beta = 0.0; g.flags = GEMM_B_LOWER;
            SB_TRY(launch_dgemm(g, c.batch, 1, stream));
        }
        return 0;
    }
    int n1 = split_point(nn), n2 = nn - n1;
    SB_TRY(potrf_rec(c, j0, n1, stream));
    {                                          // rows [j0+n1, mrows) x cols [j0+n1, j0+nn) -= P Pc'
        GemmArgs g{};
        g.M = c.mrows - (j0 + n1); g.N = n2; g.K = n1;
        g.A = c.A + (long)(j0 + n1) * c.lda + j0; g.lda = c.lda; g.strideA = c.strideA; g.a_kc = true;
        g.B = g.A; g.ldb = c.lda; g.strideB = c.strideA; g.b_kc = true;
        g.C = c.A + (long)(j0 + n1) * c.lda + (j0 + n1); g.ldc = c.lda; g.strideC = c.strideA;
        g.alpha = -1.0; g.beta = 1.0; g.flags = GEMM_C_LOWER;
        SB_TRY(launch_dgemm(g, c.batch, -1, stream));
    }
    return potrf_rec(c, j0 + n1, n2, stream);
}

struct TrtriCtx {
    int batch;
    const double* L; long ldi, strideLi;
    double* X; long ldl, strideL;
    double* tmp; long tmp_stride;
};

int trtri_rec(const TrtriCtx& t, int j0, int nn, cudaStream_t stream) {
    if (nn <= NB) return 0;                    // diagonal blocks of X already hold Dinv
    int n1 = split_point(nn), n2 = nn - n1;
    SB_TRY(trtri_rec(t, j0, n1, stream));
    SB_TRY(trtri_rec(t, j0 + n1, n2, stream));
    long ldt = round_up(n1, 2);
    {                                          // T = L21 * X11          (n2 x n1)
        GemmArgs g{};
        g.M = n2; g.N = n1; g.K = n1;
        g.A = t.L + (long)(j0 + n1) * t.ldi + j0; g.lda = t.ldi; g.strideA = t.strideLi; g.a_kc = true;
        g.B = t.X + (long)j0 * t.ldl + j0; g.ldb = t.ldl; g.strideB = t.strideL; g.b_kc = false;
        g.C = t.tmp; g.ldc = ldt; g.strideC = t.tmp_stride;
        g.alpha = 1.0; g.beta = 0.0; g.flags = GEMM_B_UPPER;     // X11[k][n] == 0 for k < n
        SB_TRY(launch_dgemm(g, t.batch, -1, stream));
    }
    {                                          // X21 = -X22 * T
        GemmArgs g{};
        g.M = n2; g.N = n1; g.K = n2;
        g.A = t.X + (long)(j0 + n1) * t.ldl + (j0 + n1); g.lda = t.ldl; g.strideA = t.strideL; g.a_kc = true;
        g.B = t.tmp; g.ldb = ldt; g.strideB = t.tmp_stride; g.b_kc = false;
        g.C = t.X + (long)(j0 + n1) * t.ldl + j0; g.ldc = t.ldl; g.strideC = t.strideL;
        g.alpha = -1.0; g.beta = 0.0; g.flags = GEMM_A_LOWER;
        SB_TRY(launch_dgemm(g, t.batch, -1, stream));
    }
    return 0;
}

// out[i] = sum_{k >= i} Linv[k][i] y[k]; 32 columns x 8 k-lanes per CTA, fixed-order reduction
__global__ void __launch_bounds__(256) trmv_lower_t_kernel(int n, const double* __restrict__ Lb, long ldl, long strideL,
                                                           const int* __restrict__ fidx,
                                                           const double* __restrict__ yb, long strideY,
                                                           double* __restrict__ outb, long strideOut) {
    __shared__ double part[8][33];
    const int z = blockIdx.y;
    const double* L = Lb + (long)(fidx ? fidx[z] : z) * strideL;
    const double* y = yb + (long)z * strideY;
    const int ci = threadIdx.x & 31, kl = threadIdx.x >> 5;
    const int col0 = blockIdx.x * 32, i = col0 + ci;
    double acc = 0.0;
    if (i < n)
        for (int k = col0 + kl; k < n; k += 8) {
            double v = (k >= i) ? L[(long)k * ldl + i] : 0.0;
            acc += v * y[k];
        }
    part[kl][ci] = acc;
    __syncthreads();
    if (kl == 0 && i < n) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < 8; r++) s += part[r][ci];
        outb[(long)z * strideOut + i] = s;
    }
}

// out[i] = sum_{k <= i} Linv[i][k] x[k]; one warp per row, coalesced along k
__global__ void __launch_bounds__(256) trmv_lower_kernel(int n, const double* __restrict__ Lb, long ldl, long strideL,
                                                         const int* __restrict__ fidx, const double* __restrict__ xb,
                                                         long strideX, double* __restrict__ outb, long strideOut) {
    const int z = blockIdx.y;
    const double* L = Lb + (long)(fidx ? fidx[z] : z) * strideL;
    const double* x = xb + (long)z * strideX;
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    double acc = 0.0;
    for (int k = lane; k <= i; k += 32) acc += L[(long)i * ldl + k] * x[k];
    acc = warp_sum(acc);
    if (lane == 0) outb[(long)z * strideOut + i] = acc;
}

}  // namespace

size_t trtri_tmp_elems(int n) {
    long h = split_point(n);
    return (size_t)(h * round_up(h, 2));
}

int potrf_batched(const CholBatch& c, cudaStream_t stream) {
    SB_CHECK_ARG(c.n > 0 && c.mrows >= c.n && c.batch > 0, "potrf: bad shape n=%d mrows=%d batch=%d", c.n, c.mrows, c.batch);
    SB_CHECK_ARG(c.lda >= c.n, "potrf: lda %ld < n %d", c.lda, c.n);
    return potrf_rec(c, 0, c.n, stream);
}

int trtri_batched(int batch, int n, const double* L, long ldl_in, long strideLin, double* Linv, long ldl,
                  long strideL, double* tmp, long tmp_stride, cudaStream_t stream) {
    SB_CHECK_ARG(n <= NB || (tmp != nullptr && (size_t)tmp_stride >= trtri_tmp_elems(n)), "trtri: temp too small");
    TrtriCtx t{batch, L, ldl_in, strideLin, Linv, ldl, strideL, tmp, tmp_stride};
    return trtri_rec(t, 0, n, stream);
}

int lauum_batched(int batch, int n, const double* Linv, long ldl, long strideL, double* C, long ldc, long strideC,
                  cudaStream_t stream) {
    GemmArgs g{};
    g.M = n; g.N = n; g.K = n;
    g.A = Linv; g.lda = ldl; g.strideA = strideL; g.a_kc = false;    // Aop[m][k] = Linv[k][m], zero for k < m
    g.B = Linv; g.ldb = ldl; g.strideB = strideL; g.b_kc = false;
    g.C = C; g.ldc = ldc; g.strideC = strideC;
    g.alpha = 1.0; g.beta = 0.0; g.flags = GEMM_A_UPPER | GEMM_C_LOWER;
    return launch_dgemm(g, batch, -1, stream);
}

int trmv_lower_t_batched(int batch, int n, const double* Linv, long ldl, long strideL, const int* idx,
                         const double* y, long strideY, double* out, long strideOut, cudaStream_t stream) {
    dim3 grid(cdiv(n, 32), batch);
    trmv_lower_t_kernel<<<grid, 256, 0, stream>>>(n, Linv, ldl, strideL, idx, y, strideY, out, strideOut);
    SB_LAUNCH_CHECK();
    return 0;
}

int trmv_lower_batched(int batch, int n, const double* Linv, long ldl, long strideL, const int* idx,
                       const double* x, long strideX, double* out, long strideOut, cudaStream_t stream) {
    dim3 grid(cdiv(n, 8), batch);
    trmv_lower_kernel<<<grid, 256, 0, stream>>>(n, Linv, ldl, strideL, idx, x, strideX, out, strideOut);
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace sb
