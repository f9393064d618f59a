// Standardisation statistics and the small symmetric eigensolver of the PCA stage (PCGPwM.py:533-654).
#pragma once
#include "common.cuh"
#include <algorithm>

namespace sb {

constexpr int SYEVJ_MAXP = 112;        // matrix + eigenvectors of one problem stay in shared memory
constexpr int PCA_SUMSQ_PARTS = 1024;  // scratch doubles of pca_sumsq

// offset = nanmean, scale = nanstd / sqrt(1 - missing fraction) per column of f (n x m, leading dimension ldf),
// nmiss = non-finite cells per column
size_t pca_col_stats_workspace(int n, int m);
int pca_col_stats(int n, int m, const double* f, long ldf, double* offset, double* scale, int* nmiss, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream);
// fs = (f - offset) / scale; with miss != NULL the (n x m) byte mask of the missing cells is written and those cells
// get the reference's alternating seed
int pca_standardize(int n, int m, const double* f, long ldf, const double* offset, const double* scale, double* fs,
                    long ldfs, unsigned char* miss, const int* nmiss, cudaStream_t stream);
// out[0] = sum x_i^2 (deterministic); scratch: PCA_SUMSQ_PARTS doubles
int pca_sumsq(long count, const double* x, double* out, double* scratch, cudaStream_t stream);
// eigen-decomposition of `batch` symmetric p x p matrices (p <= SYEVJ_MAXP): W descending, Z columns = eigenvectors
int syevj_batched(int batch, int p, const double* H, long ldh, long strideH, double* W, long strideW, double* Z,
                  long ldz, long strideZ, int max_sweeps, int* sweeps, cudaStream_t stream);

}  // namespace sb
