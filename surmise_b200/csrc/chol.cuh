// Batched blocked Cholesky machinery (all matrices of a batch share n; row-major, lower).
#pragma once
#include "common.cuh"

namespace sb {

constexpr int CHOL_NB = 64;   // leaf block; must not exceed the narrow GEMM tile N (64) -- in-place panel solve

struct CholBatch {
    int batch;        // number of matrices
    int n;            // order of the SPD part
    int mrows;        // rows actually processed (>= n): extra rows below ride along as right-hand sides,
                      //   row n.. holds y' = (L^-1 g)' after the factorisation ("augmented" trick)
    double* A;        // (batch, mrows, lda) lower triangle in, L out (strict upper never touched/read)
    long lda, strideA;
    double* Dinv;     // inverses of the NB x NB diagonal blocks of L: block b of matrix z at
    long ldd;         //   Dinv + z*strideD + b*blockD, leading dimension ldd, zeros above the diagonal
    long strideD, blockD;
    int* info;        // (batch) 0 = ok, j>0 = leading minor j not positive (LAPACK convention)
    int* counters;    // (batch) zero-initialised scratch: arrival counters of the panel kernel (left at zero)
    int* dag_scratch; // potrf_dag_scratch_ints(batch, mrows) ints, or NULL: row-progress counters and the ticket of the
                      //   task-graph kernel (chol_dag.cu); with NULL the recursive multi-launch path is used
};

// task-graph variant (one persistent launch; chol_dag.cu): same contract as potrf_batched
size_t potrf_dag_scratch_ints(int batch, int mrows);
bool potrf_dag_usable(const CholBatch& c);
int potrf_dag(const CholBatch& c, cudaStream_t stream);
void debug_dag_stats(long long* device_ptr);     // instrumentation: 8 clock64 counters per CTA (2 CTAs per SM)

// A <- L (in place), Dinv filled, extra rows <- rows * L^-T.   ~ n^3/3 + (mrows-n) n^2 flop per matrix
int potrf_batched(const CholBatch& c, cudaStream_t stream);

// Linv <- L^-1 (lower, out of place).  Linv's diagonal blocks must already hold Dinv (i.e. the
// factorisation was run with Dinv pointing into Linv: ldd = ldl, blockD = NB*ldl + NB, strideD = strideL)
// and its strict upper triangle must be zero.  tmp: (batch, tmp_stride) doubles, tmp_stride >= (n/2+NB)^2.
// ~ n^3/3 flop per matrix
int trtri_batched(int batch, int n, const double* L, long ldl_in, long strideLin, double* Linv, long ldl,
                  long strideL, double* tmp, long tmp_stride, cudaStream_t stream);

// C_lower <- Linv' Linv  (only tiles touching the lower triangle are written).  ~ n^3/3 flop per matrix
int lauum_batched(int batch, int n, const double* Linv, long ldl, long strideL, double* C, long ldc, long strideC,
                  cudaStream_t stream);

// out[z][i] = sum_{k>=i} Linv[f][k][i] * y[z][k]      (Linv' y, Linv lower triangular), f = idx ? idx[z] : z
int trmv_lower_t_batched(int batch, int n, const double* Linv, long ldl, long strideL, const int* idx,
                         const double* y, long strideY, double* out, long strideOut, cudaStream_t stream);
// out[z][i] = sum_{k<=i} Linv[f][i][k] * x[z][k]       (Linv x)
int trmv_lower_batched(int batch, int n, const double* Linv, long ldl, long strideL, const int* idx,
                       const double* x, long strideX, double* out, long strideOut, cudaStream_t stream);

size_t trtri_tmp_elems(int n);
void debug_panel_stamps(long long* device_ptr);   // instrumentation: 8 clock64 slots written by CTA (0,0) of each panel launch

}  // namespace sb
