// FP64 tensor-core (DMMA) GEMM used by every dense step of the GP path: Cholesky trailing
// updates, triangular solves through inverted diagonal blocks, triangular inversion, L^-T L^-1
// and the predict products.  One kernel template, four operand-layout instantiations.
#pragma once
#include "common.cuh"

namespace sb {

// C[m][n] = alpha * sum_k Aop[m][k] * Bop[n][k] + beta * C[m][n]      (row-major C, ldc)
//   a_kc : Aop[m][k] = A[m*lda + k]   (k contiguous)      else  A[k*lda + m]   (m contiguous)
//   b_kc : Bop[n][k] = B[n*ldb + k]   (k contiguous)      else  B[k*ldb + n]   (n contiguous)
// Triangular structure only narrows the k-range visited per tile; the operand memory must hold
// real zeros in its "empty" triangle (all such operands are buffers this library zero-fills).
enum GemmFlags : int {
    GEMM_A_LOWER = 1,   // Aop[m][k] == 0 for k > m
    GEMM_A_UPPER = 2,   // Aop[m][k] == 0 for k < m
    GEMM_B_LOWER = 4,   // Bop[n][k] == 0 for k > n
    GEMM_B_UPPER = 8,   // Bop[n][k] == 0 for k < n
    GEMM_C_LOWER = 16   // skip tiles lying strictly above the diagonal of C
};

struct GemmArgs {
    int M, N, K;
    const double* A; long lda; long strideA; const int* idxA;   // batch z uses A + (idxA ? idxA[z] : z) * strideA
    const double* B; long ldb; long strideB; const int* idxB;
    double* C;       long ldc; long strideC;
    double alpha, beta;
    bool a_kc, b_kc;
    int flags;
    // optional second batch level: launch index z = z2 * inner_batch + z1; the strides above (and idxA/idxB)
    // apply to z1, these to z2 (e.g. z1 = matrix of the batch, z2 = diagonal block pair within the matrix)
    int inner_batch;
    long strideA2, strideB2, strideC2;
    int last_M;     // > 0: the last z2 group has this many rows instead of M (0 = all groups alike)
};

// tile: 7 = 64x64 CTA tile, 4 warps, 4 CTAs/SM, cp.async staging; 1 = 128x64 (2 CTAs/SM); 0 = 128x128 (1 CTA/SM);
// 2-4 experimental variants; 8 / 9 = TMA-staged 64x64 / 128x64 (below); -1 = default
int launch_dgemm(const GemmArgs& g, int batch, int tile, cudaStream_t stream);

// tile 8 / 9: the same product with TMA operand staging and an mbarrier pipeline (dgemm_tma.cu; 64x64 / 128x64 tiles);
// returns -2 without launching when the operands cannot be described by tensor maps (launch_dgemm then falls back)
int launch_dgemm_tma(const GemmArgs& g, int batch, int bm, cudaStream_t stream);
double dgemm_algorithmic_flops(const GemmArgs& g);     // 2 M N K with the exact triangular k-ranges
long dgemm_tma_launch_count();

}  // namespace sb
