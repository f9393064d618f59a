// Batched recursive Cholesky / triangular inverse / L^-T L^-1 on top of the DMMA GEMM.
//
// Column-recursive formulation on a "tall" matrix (mrows >= n): factor the left half (which also
// solves every row below it), one DMMA rank-n1 update of the right half, recurse.  Almost all flops
// land in GEMMs whose K grows with the recursion level, so the trailing updates run at GEMM speed.
// Rows n..mrows-1 are right-hand sides: after the call row n holds (L^-1 g)' when it held g'.
#include "chol.cuh"
#include "dgemm.cuh"
#include <cstdlib>

namespace sb {
namespace {

constexpr int NB = CHOL_NB;

// ------------------------------------------------------------------------------------------------
// Panel kernel: factor one 64-wide diagonal block AND solve the rows below it, in one launch.
//   grid = (row tiles of 128 below the block (at least 1), batch), 256 threads.
// Every CTA redundantly factors the 64x64 block (64^3/3 flop, ~13 us) and inverts it (~3 us), so the panel solve
// needs no second launch and no grid-wide dependency; the CTA that loaded the block last writes L, L^-1 back.
//   phase 1: block -> registers (16x16 thread grid, 4x4 elements each), right-looking Cholesky, two columns per
//            __syncthreads (double-buffered column-pair broadcast through shared memory);
//   phase 2: X = L^-1: the four 16 x 16 diagonal inverses by substitution (one warp each), the rest by two levels
//            of DMMA block merges;
//   phase 3: P <- P X' for this CTA's 128 rows on the FP64 tensor cores (DMMA) out of shared memory.
// FUSE (first block of a two-block leaf): the rank-64 update of the NEXT block column, C -= P Q' with Q = the first
//   64 solved rows, is applied to this CTA's rows in the same pass (every CTA solves those 64 rows redundantly),
//   which removes the smallest trailing-update GEMM launch of the recursion.
constexpr int PT = 128;          // rows per CTA in the panel solve
constexpr int SLD = NB + 4;      // shared leading dimension: (4 g + t) mod 16 distinct -> conflict-free LDS.64
constexpr int WLD = 32 + 4;

__device__ __forceinline__ void cp_async8_plain(double* smem, const double* gmem, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(valid ? 8 : 0) : "memory");
}

// rsqrt for a normal positive double, branch free (MUFU.RSQ64H seed + one third-order step, the sequence the CUDA
// library uses on its fast path) so that two of them interleave; anything else goes through rsqrt() below
__device__ __forceinline__ double rsqrt_normal(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    double e = fma(x, -(y * y), 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}
__device__ __forceinline__ bool normal_positive(double x) {
    return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u;
}

template <bool FUSE>
__global__ void __launch_bounds__(256) potrf_panel_kernel(double* __restrict__ Abase, long lda, long strideA,
                                                          double* __restrict__ Dbase, long ldd, long strideD, int nb,
                                                          int j0, int below, int rows_per_cta, int fuse_nc,
                                                          int* __restrict__ info, int* __restrict__ counters,
                                                          long long* stamps) {
    extern __shared__ double sm[];
    double* S = sm;                       // [NB][SLD]  block, then L
    double* X = S + NB * SLD;             // [NB][SLD]  L^-1
    double* T = X + NB * SLD;             // [PT][SLD]  this CTA's rows of the panel
    double* W = T + PT * SLD;             // [32][WLD]  scratch of the inverse's block merges
    double* invd = W + 32 * WLD;          // [NB]       reciprocals of L's diagonal
    double* Q = invd + NB;                // [NB][SLD]  FUSE only: the first 64 rows below the block, then solved
    const int tid = threadIdx.x, z = blockIdx.y;
#define SB_STAMP(slot) do { if (stamps && z == 0 && blockIdx.x == 0 && tid == 0) stamps[slot] = clock64(); } while (0)
    SB_STAMP(0);
    double* A = Abase + (long)z * strideA;          // points at the diagonal block (j0, j0)
    const int cta_r0 = blockIdx.x * rows_per_cta;   // first panel row of this CTA (relative to the row below the block)
    const int cta_rows = min(rows_per_cta, below - cta_r0);     // <= 0 when there is nothing below

    // asynchronous stage of the first 128 panel rows (overlaps the factorisation)
    auto stage_tile = [&](int r0) {
        const int rows_here = min(PT, cta_r0 + cta_rows - r0);
        const double* P = A + (long)(nb + r0) * lda;
        for (int idx = tid; idx < PT * NB; idx += 256) {
            int r = idx >> 6, c = idx & 63;
            bool ok = (r < rows_here) && (c < nb);
            cp_async8_plain(T + r * SLD + c, ok ? P + (long)r * lda + c : A, ok);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if (FUSE) {
        const double* P = A + (long)nb * lda;
        for (int idx = tid; idx < NB * NB; idx += 256) {
            int r = idx >> 6, c = idx & 63;
            bool ok = (r < below) && (c < nb);
            cp_async8_plain(Q + r * SLD + c, ok ? P + (long)r * lda + c : A, ok);
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    if (cta_rows > 0) stage_tile(cta_r0);

    // ---- phase 1: Cholesky of the (identity padded) 64 x 64 block, register resident, TWO columns per barrier.
    // FP64 results take ~50 cycles on B200, so the serial chain through the pivots is what this phase costs.  A pair
    // of columns (j, j+1) is published together; every thread forms d = a_jj a_j+1,j+1 - a_j+1,j^2 (= p_j p_j+1)
    // and takes rsqrt(p_j) and rsqrt(d) side by side: 1/sqrt(p_j+1) = rsqrt(d) sqrt(p_j), so the second pivot does
    // not wait for the first column's reciprocal square root.
    const int ti = tid >> 4, tc = tid & 15;
    double a[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            int i = ti + 16 * ii, c = tc + 16 * cc;
            double v = 0.0;
            if (i < nb && c <= i) v = A[(long)i * lda + c];
            else if (i >= nb && c == i) v = 1.0;
            a[ii][cc] = v;
        }
    SB_STAMP(1);
    // The block is factored in place by whichever CTA of this matrix loaded it LAST (all CTAs compute the
    // same L), so no CTA can still be reading A's block when it is overwritten.  No spinning: a counter.
    __shared__ int s_writer;
    // the writer CTA below will overwrite the 64 rows under the block: this CTA's copies of them (Q, and T's
    // first rows in CTA 0) must have landed before it counts as arrived
    if (FUSE) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        int old = atomicAdd(&counters[z], 1);
        s_writer = (old == (int)gridDim.x - 1);
        if (s_writer) counters[z] = 0;               // everyone has arrived: re-arm for the next launch
        __threadfence();
    }
    int bad_at = 0;
    // Only the 16-column block index jb is unrolled (it selects registers); the column pair within the block is a
    // real loop, so the code stays small (a fully unrolled version is instruction-fetch bound).
#pragma unroll
    for (int jb = 0; jb < 4; jb++) {
#pragma unroll 1
        for (int jr = 0; jr < 16; jr += 2) {
            const int j = jb * 16 + jr;
            double* cb0 = W + (jr & 2) * NB;          // double-buffered pair of columns (W is free until phase 2)
            double* cb1 = cb0 + NB;
            if (tc == jr) {                           // rows of the 16-blocks above jb are finished: never read again
#pragma unroll
                for (int ii = 0; ii < 4; ii++)
                    if (ii >= jb) cb0[ti + 16 * ii] = a[ii][jb];
            } else if (tc == jr + 1) {
#pragma unroll
                for (int ii = 0; ii < 4; ii++)
                    if (ii >= jb) cb1[ti + 16 * ii] = a[ii][jb];
            }
            __syncthreads();
            double p0 = cb0[j];
            const double q = cb0[j + 1], p1 = cb1[j + 1];
            // this thread's rows / columns of the pair, fetched before the pivot arithmetic (in-order issue: loads
            // placed after the reciprocal square roots would wait for them)
            double x0i[4], x1i[4], x0c[4], x1c[4];
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (e >= jb) {                        // (compile time)
                    x0i[e] = cb0[ti + 16 * e];
                    x1i[e] = cb1[ti + 16 * e];
                    x0c[e] = cb0[tc + 16 * e];
                    x1c[e] = cb1[tc + 16 * e];
                }
            const double det = fma(p0, p1, -(q * q));
            double r0 = rsqrt_normal(p0);
            const double rd = rsqrt_normal(det);
            double s0 = p0 * r0;                      // l(j, j)
            double l10 = q * r0;                      // l(j+1, j)
            double inv1 = rd * s0;                    // 1 / l(j+1, j+1)
            double l11 = det * rd * r0;               // l(j+1, j+1) = sqrt(det) / sqrt(p0)
            if (!(normal_positive(p0) && normal_positive(det))) {
                // block-uniform and rare: a pivot that is not positive (LAPACK-style report, pivot replaced by 1) or
                // a pair whose product left the normal range; one column after the other with the library rsqrt
                if (!(p0 > 0.0)) {
                    if (bad_at == 0) bad_at = j0 + j + 1;
                    p0 = 1.0;
                }
                r0 = rsqrt(p0);
                s0 = p0 * r0;
                l10 = q * r0;
                double p1n = p1 - l10 * l10;
                if (!(p1n > 0.0)) {
                    if (bad_at == 0) bad_at = j0 + j + 2;
                    p1n = 1.0;
                }
                inv1 = rsqrt(p1n);
                l11 = p1n * inv1;
            }
            // Rows / columns in 16-blocks below jb are all still active (no mask), those of block jb are masked by
            // position, those above are finished and skipped -- all decided at compile time (jb is unrolled).
            double ui[4], vi[4], uc[4], vc[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (e < jb) continue;
                ui[e] = x0i[e] * r0;
                vi[e] = (x1i[e] - ui[e] * l10) * inv1;
                uc[e] = x0c[e] * r0;
                vc[e] = (x1c[e] - uc[e] * l10) * inv1;
                if (e == jb) {
                    const int i = ti + 16 * e, c = tc + 16 * e;
                    if (i <= j) ui[e] = 0.0;
                    if (i <= j + 1) vi[e] = 0.0;
                    if (c <= j) uc[e] = 0.0;
                    if (c <= j + 1) vc[e] = 0.0;
                }
            }
#pragma unroll
            for (int ii = 0; ii < 4; ii++)
#pragma unroll
                for (int cc = 0; cc < 4; cc++)
                    if (ii >= jb && cc >= jb && cc <= ii) {
                        a[ii][cc] -= ui[ii] * uc[cc];
                        a[ii][cc] -= vi[ii] * vc[cc];
                    }
            if (tc == jr) {
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    if (ii > jb) a[ii][jb] = ui[ii];
                    else if (ii == jb) {
                        const int i = ti + 16 * ii;
                        if (i > j) a[ii][jb] = ui[ii];
                        else if (i == j) a[ii][jb] = s0;
                    }
                }
            } else if (tc == jr + 1) {
#pragma unroll
                for (int ii = 0; ii < 4; ii++) {
                    if (ii > jb) a[ii][jb] = vi[ii];
                    else if (ii == jb) {
                        const int i = ti + 16 * ii;
                        if (i > j + 1) a[ii][jb] = vi[ii];
                        else if (i == j + 1) a[ii][jb] = l11;
                    }
                }
            }
            if (tid == 0) {
                invd[j] = r0;
                invd[j + 1] = inv1;
            }
        }
    }
    SB_STAMP(2);
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            int i = ti + 16 * ii, c = tc + 16 * cc;
            S[i * SLD + c] = (c <= i) ? a[ii][cc] : 0.0;
            X[i * SLD + c] = 0.0;
        }
    __syncthreads();
    const bool writer = s_writer != 0;
    if (writer && tid == 0 && bad_at != 0 && info[z] == 0) info[z] = bad_at;

    SB_STAMP(3);
    // ---- phase 2: X = L^-1.  (a) warps 0-3 invert one 16 x 16 diagonal sub-block each by forward substitution on the
    // identity, lane (r, h) holding columns 8h .. 8h+7 of row r, rows unscaled until the end so that a step's chain
    // is one shuffle and one FMA; (b) the off-diagonal blocks bottom-up like trtri_levels, X21 = -X22 (L21 X11) for
    // the pairs of 16-blocks and then for the pair of 32-blocks, all on DMMA.
    if (tid < 128) {
        const int lane = tid & 31, k0 = 16 * (tid >> 5), r = lane & 15, h = lane >> 4;
        double b[8], m[16];
#pragma unroll
        for (int j = 0; j < 16; j++) m[j] = (r > j) ? S[(k0 + r) * SLD + k0 + j] * invd[k0 + j] : 0.0;
#pragma unroll
        for (int cc = 0; cc < 8; cc++) b[cc] = (8 * h + cc == r) ? 1.0 : 0.0;
#pragma unroll
        for (int j = 0; j < 15; j++)
#pragma unroll
            for (int cc = 0; cc < 8; cc++) b[cc] -= m[j] * __shfl_sync(0xffffffffu, b[cc], j + 16 * h);
        const double invr = invd[k0 + r];
#pragma unroll
        for (int cc = 0; cc < 8; cc++) X[(k0 + r) * SLD + k0 + 8 * h + cc] = (8 * h + cc <= r) ? b[cc] * invr : 0.0;
    }
    __syncthreads();
    {
        const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
#pragma unroll 1
        for (int sz = 16; sz < NB; sz *= 2) {
            const int tps = sz / 8, per_pair = tps * tps, total = (NB / (2 * sz)) * per_pair;
            for (int tix = warp; tix < total; tix += 8) {             // W = L21 X11
                const int pair = tix / per_pair, rem = tix % per_pair, mi = rem / tps, ni = rem % tps, b0 = pair * 2 * sz;
                double c0 = 0.0, c1 = 0.0;
                for (int k4 = 0; k4 < sz / 4; k4++) {
                    double av = S[(b0 + sz + 8 * mi + g) * SLD + b0 + 4 * k4 + t];
                    double bv = X[(b0 + 4 * k4 + t) * SLD + b0 + 8 * ni + g];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c0), "+d"(c1)
                                 : "d"(av), "d"(bv));
                }
                W[(pair * sz + 8 * mi + g) * WLD + 8 * ni + 2 * t] = c0;
                W[(pair * sz + 8 * mi + g) * WLD + 8 * ni + 2 * t + 1] = c1;
            }
            __syncthreads();
            for (int tix = warp; tix < total; tix += 8) {             // X21 = -X22 W
                const int pair = tix / per_pair, rem = tix % per_pair, mi = rem / tps, ni = rem % tps, b0 = pair * 2 * sz;
                double c0 = 0.0, c1 = 0.0;
                for (int k4 = 0; k4 < sz / 4; k4++) {
                    double av = -X[(b0 + sz + 8 * mi + g) * SLD + b0 + sz + 4 * k4 + t];
                    double bv = W[(pair * sz + 4 * k4 + t) * WLD + 8 * ni + g];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c0), "+d"(c1)
                                 : "d"(av), "d"(bv));
                }
                X[(b0 + sz + 8 * mi + g) * SLD + b0 + 8 * ni + 2 * t] = c0;
                X[(b0 + sz + 8 * mi + g) * SLD + b0 + 8 * ni + 2 * t + 1] = c1;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    SB_STAMP(4);
    if (writer) {
        double* D = Dbase + (long)z * strideD;
        for (int idx = tid; idx < nb * nb; idx += 256) {
            int i = idx / nb, c = idx % nb;
            if (c <= i) A[(long)i * lda + c] = S[i * SLD + c];
            D[(long)i * ldd + c] = X[i * SLD + c];
        }
    }
    SB_STAMP(5);
    if (cta_rows <= 0) return;

    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    if (FUSE) {
        // Q <- Q X' (64 x 64 x 64, warp w owns rows 8w .. 8w+7, in place).  Every CTA gets the same bits; the
        // writer CTA stores them (rows 0..63 below the block), so no CTA can still be loading the unsolved rows.
        double acc[8][2];
#pragma unroll
        for (int ni = 0; ni < 8; ni++) acc[ni][0] = acc[ni][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < NB / 4; k4++) {
            double av = Q[(8 * warp + g) * SLD + 4 * k4 + t];
#pragma unroll
            for (int ni = 0; ni < 8; ni++)
                if (4 * k4 <= 8 * ni + 7) {
                    double bv = X[(8 * ni + g) * SLD + 4 * k4 + t];
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(acc[ni][0]), "+d"(acc[ni][1])
                                 : "d"(av), "d"(bv));
                }
        }
        __syncwarp();
        const int r = 8 * warp + g;
        double* P = A + (long)nb * lda;
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
            int c = 8 * ni + 2 * t;
            Q[r * SLD + c] = acc[ni][0];
            Q[r * SLD + c + 1] = acc[ni][1];
            if (writer && r < below) {
                P[(long)r * lda + c] = acc[ni][0];
                P[(long)r * lda + c + 1] = acc[ni][1];
            }
        }
        // the barrier at the top of the phase-3 loop publishes Q
    }

    // ---- phase 3: P <- P X'  (DMMA from shared memory), 128 rows at a time; warp w owns rows 16w .. 16w+15
    const int m0 = warp * 16;
    for (int r0 = cta_r0; r0 < cta_r0 + cta_rows; r0 += PT) {
        const int rows_here = min(PT, cta_r0 + cta_rows - r0);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        double acc[2][8][2];
#pragma unroll
        for (int mi = 0; mi < 2; mi++)
#pragma unroll
            for (int ni = 0; ni < 8; ni++) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
#pragma unroll
        for (int k4 = 0; k4 < NB / 4; k4++) {
            double av[2];
#pragma unroll
            for (int mi = 0; mi < 2; mi++) av[mi] = T[(m0 + 8 * mi + g) * SLD + 4 * k4 + t];
#pragma unroll
            for (int ni = 0; ni < 8; ni++)
                if (4 * k4 <= 8 * ni + 7) {                      // X[n][k] == 0 for k > n
                    double bv = X[(8 * ni + g) * SLD + 4 * k4 + t];
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                     : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                                     : "d"(av[mi]), "d"(bv));
                }
        }
        double* P = A + (long)(nb + r0) * lda;
        if (!FUSE) {
            __syncthreads();                                         // everyone is done reading T
            if (r0 + PT < cta_r0 + cta_rows) stage_tile(r0 + PT);    // prefetch the next 128 rows while storing
        }
#pragma unroll
        for (int mi = 0; mi < 2; mi++) {
            int r = m0 + 8 * mi + g;
            if (r >= rows_here) continue;
            if (FUSE && r0 + r < NB) continue;                       // those rows are Q: stored by the writer CTA
#pragma unroll
            for (int ni = 0; ni < 8; ni++) {
                int c = 8 * ni + 2 * t;
                if (c < nb) P[(long)r * lda + c] = acc[mi][ni][0];
                if (c + 1 < nb) P[(long)r * lda + c + 1] = acc[mi][ni][1];
            }
        }
        if (FUSE) {
            // C[rows, nb .. nb+fuse_nc) -= P Q'   (the solved rows go back into this warp's own rows of T as operand A)
            __syncwarp();
#pragma unroll
            for (int mi = 0; mi < 2; mi++)
#pragma unroll
                for (int ni = 0; ni < 8; ni++) {
                    int r = m0 + 8 * mi + g, c = 8 * ni + 2 * t;
                    T[r * SLD + c] = -acc[mi][ni][0];
                    T[r * SLD + c + 1] = -acc[mi][ni][1];
                }
            __syncwarp();
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                int r = m0 + 8 * mi + g;
#pragma unroll
                for (int ni = 0; ni < 8; ni++) {
                    int c = 8 * ni + 2 * t;
                    bool ok = r < rows_here;
                    acc[mi][ni][0] = (ok && c < fuse_nc) ? P[(long)r * lda + nb + c] : 0.0;
                    acc[mi][ni][1] = (ok && c + 1 < fuse_nc) ? P[(long)r * lda + nb + c + 1] : 0.0;
                }
            }
#pragma unroll
            for (int k4 = 0; k4 < NB / 4; k4++) {
                double av[2];
#pragma unroll
                for (int mi = 0; mi < 2; mi++) av[mi] = T[(m0 + 8 * mi + g) * SLD + 4 * k4 + t];
#pragma unroll
                for (int ni = 0; ni < 8; ni++) {
                    double bv = Q[(8 * ni + g) * SLD + 4 * k4 + t];
#pragma unroll
                    for (int mi = 0; mi < 2; mi++)
                        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                     : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                                     : "d"(av[mi]), "d"(bv));
                }
            }
            __syncthreads();                                         // everyone is done reading T
            if (r0 + PT < cta_r0 + cta_rows) stage_tile(r0 + PT);
#pragma unroll
            for (int mi = 0; mi < 2; mi++) {
                int r = m0 + 8 * mi + g;
                if (r >= rows_here) continue;
                const int rr = r0 + r;                               // row inside the next block column
#pragma unroll
                for (int ni = 0; ni < 8; ni++) {
                    int c = 8 * ni + 2 * t;                          // strict upper part of the next diagonal block stays
                    if (c < fuse_nc && (rr >= NB || c <= rr)) P[(long)r * lda + nb + c] = acc[mi][ni][0];
                    if (c + 1 < fuse_nc && (rr >= NB || c + 1 <= rr)) P[(long)r * lda + nb + c + 1] = acc[mi][ni][1];
                }
            }
        }
    }
    SB_STAMP(6);
#undef SB_STAMP
}

static std::atomic<long long*> g_panel_stamps{nullptr};      // instrumentation (sb200_debug_panel_stamps)

// fuse_nc > 0 (nb == NB only): also apply this block column's rank-64 update to the next fuse_nc columns
int panel(const CholBatch& c, int j0, int nb, int fuse_nc, cudaStream_t stream) {
    constexpr size_t smem = (size_t)(2 * NB * SLD + PT * SLD + 32 * WLD + NB) * sizeof(double);
    constexpr size_t smem_fuse = smem + (size_t)NB * SLD * sizeof(double);
    static DeviceOnce once;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(potrf_panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SB_CUDA(cudaFuncSetAttribute(potrf_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fuse));
        return 0;
    }));
    int below = c.mrows - (j0 + nb);
    // every CTA repeats the 64x64 factorisation, so keep the grid within one wave of the 148 SMs (1 CTA/SM):
    // each CTA takes as many 128-row sub-tiles as that needs
    int ctas_per_matrix = 148 / c.batch;
    if (ctas_per_matrix < 1) ctas_per_matrix = 1;
    int rows_per_cta = (int)round_up(cdiv(below > 0 ? below : 1, ctas_per_matrix), PT);
    dim3 grid(below > 0 ? cdiv(below, rows_per_cta) : 1, c.batch);
    if (fuse_nc > 0)
        potrf_panel_kernel<true><<<grid, 256, smem_fuse, stream>>>(c.A + (long)j0 * c.lda + j0, c.lda, c.strideA,
                                                                   c.Dinv + (long)(j0 / NB) * c.blockD, c.ldd, c.strideD,
                                                                   nb, j0, below, rows_per_cta, fuse_nc, c.info,
                                                                   c.counters, g_panel_stamps.load());
    else
        potrf_panel_kernel<false><<<grid, 256, smem, stream>>>(c.A + (long)j0 * c.lda + j0, c.lda, c.strideA,
                                                                c.Dinv + (long)(j0 / NB) * c.blockD, c.ldd, c.strideD,
                                                                nb, j0, below, rows_per_cta, 0, c.info, c.counters,
                                                                g_panel_stamps.load());
    SB_LAUNCH_CHECK();
    return 0;
}

bool fuse_leaf_pairs() {       // SB200_PANEL_FUSE=0 keeps the separate rank-64 GEMM (experiments only)
    static int on = -1;
    if (on < 0) {
        const char* env = getenv("SB200_PANEL_FUSE");
        on = (env && env[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

int split_point(int nn) {      // multiple of NB closest to nn/2 from above
    int blocks = (nn + NB - 1) / NB;
    return ((blocks + 1) / 2) * NB;
}

int potrf_rec(const CholBatch& c, int j0, int nn, cudaStream_t stream) {
    if (nn <= NB) return panel(c, j0, nn, 0, stream);   // factor the block and solve every row below it
    int n1 = split_point(nn), n2 = nn - n1;
    if (nn <= 2 * NB && fuse_leaf_pairs()) {            // two-block leaf: the first panel launch carries the update
        SB_TRY(panel(c, j0, NB, n2, stream));
        return panel(c, j0 + NB, n2, 0, stream);
    }
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

// X21 = -X22 (L21 X11) for `pairs` adjacent (s, s2) block pairs starting at block offset j0, spaced 2s apart;
// last_s2 > 0: the last of them is ragged, its lower block has last_s2 < s2 rows
int trtri_merge(const TrtriCtx& t, int j0, int s, int s2, int pairs, int last_s2, cudaStream_t stream) {
    const long step_i = (long)2 * s * (t.ldi + 1), step_x = (long)2 * s * (t.ldl + 1), step_t = (long)s * s;
    {                                          // T = L21 * X11          (s2 x s per pair)
        GemmArgs g{};
        g.M = s2; g.N = s; g.K = s;
        g.A = t.L + (long)(j0 + s) * t.ldi + j0; g.lda = t.ldi; g.strideA = t.strideLi; g.a_kc = true;
        g.B = t.X + (long)j0 * t.ldl + j0; g.ldb = t.ldl; g.strideB = t.strideL; g.b_kc = false;
        g.C = t.tmp; g.ldc = s; g.strideC = t.tmp_stride;
        g.alpha = 1.0; g.beta = 0.0; g.flags = GEMM_B_UPPER;     // X11[k][n] == 0 for k < n
        g.inner_batch = t.batch; g.strideA2 = step_i; g.strideB2 = step_x; g.strideC2 = step_t; g.last_M = last_s2;
        SB_TRY(launch_dgemm(g, t.batch * pairs, -1, stream));
    }
    {                                          // X21 = -X22 * T
        GemmArgs g{};
        g.M = s2; g.N = s; g.K = s2;
        g.A = t.X + (long)(j0 + s) * t.ldl + (j0 + s); g.lda = t.ldl; g.strideA = t.strideL; g.a_kc = true;
        g.B = t.tmp; g.ldb = s; g.strideB = t.tmp_stride; g.b_kc = false;
        g.C = t.X + (long)(j0 + s) * t.ldl + j0; g.ldc = t.ldl; g.strideC = t.strideL;
        g.alpha = -1.0; g.beta = 0.0; g.flags = GEMM_A_LOWER;
        g.inner_batch = t.batch; g.strideA2 = step_x; g.strideB2 = step_t; g.strideC2 = step_x; g.last_M = last_s2;
        SB_TRY(launch_dgemm(g, t.batch * pairs, -1, stream));
    }
    return 0;
}

// Bottom-up: at level s all pairs of adjacent s-blocks are independent, so one (two-level batched) launch pair
// handles every pair of every matrix, the ragged last pair (right block shorter) included as a shorter last group.
int trtri_levels(const TrtriCtx& t, int n, cudaStream_t stream) {
    for (long s = NB; s < n; s *= 2) {
        int full = (int)(n / (2 * s));                 // pairs with two complete s-blocks
        long rest = n - (long)full * 2 * s;            // columns after the last full pair
        const bool ragged = rest > s;
        SB_CHECK_ARG((long)t.batch * (full + 1) <= 65535, "trtri: batch x pairs exceeds grid.z");
        if (full > 0 && ragged) SB_TRY(trtri_merge(t, 0, (int)s, (int)s, full + 1, (int)(rest - s), stream));
        else if (full > 0) SB_TRY(trtri_merge(t, 0, (int)s, (int)s, full, 0, stream));
        else if (ragged) SB_TRY(trtri_merge(t, 0, (int)s, (int)(rest - s), 1, 0, stream));
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

void debug_panel_stamps(long long* p) { g_panel_stamps.store(p); }

size_t trtri_tmp_elems(int n) {
    // level s uses full_pairs * s * s (+ rest * s for a ragged pair) <= n^2 / 4 (+ slack for block rounding)
    long h = (long)n / 2 + NB;
    return (size_t)(h * h);
}

int potrf_batched(const CholBatch& c, cudaStream_t stream) {
    SB_CHECK_ARG(c.n > 0 && c.mrows >= c.n && c.batch > 0, "potrf: bad shape n=%d mrows=%d batch=%d", c.n, c.mrows, c.batch);
    SB_CHECK_ARG(c.lda >= c.n, "potrf: lda %ld < n %d", c.lda, c.n);
    SB_CHECK_ARG(c.counters != nullptr, "potrf: counters is NULL");
    // one persistent task-graph launch unless switched off (SB200_POTRF_DAG=0: A/B experiments) or not applicable
    static int dag = -1;
    if (dag < 0) {
        const char* env = getenv("SB200_POTRF_DAG");
        dag = (env && env[0] == '0') ? 0 : ((env && env[0] == '2') ? 2 : 1);     // 2: always (experiments)
    }
    // measured on B200 (scripts/potrf_ab.py, task graph vs recursion): 1.20 vs 1.46 ms (n=2000, one matrix), 2.54 vs 3.37 ms
    // (n=2000 x 20), 6.83 vs 8.32 ms (n=4000 x 8), 7.3 vs 10.6 ms (n=8000, one matrix), 44.9 vs 47.4 ms (n=8000 x 8,
    // 30.4 TFLOP/s) -- the task graph wins at every shape once its operands come in by TMA
    if (dag && c.n > NB && potrf_dag_usable(c)) return potrf_dag(c, stream);
    return potrf_rec(c, 0, c.n, stream);
}

int trtri_batched(int batch, int n, const double* L, long ldl_in, long strideLin, double* Linv, long ldl,
                  long strideL, double* tmp, long tmp_stride, cudaStream_t stream) {
    SB_CHECK_ARG(n <= NB || (tmp != nullptr && (size_t)tmp_stride >= trtri_tmp_elems(n)), "trtri: temp too small");
    TrtriCtx t{batch, L, ldl_in, strideLin, Linv, ldl, strideL, tmp, tmp_stride};
    return trtri_levels(t, n, stream);
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
