// Task-graph batched Cholesky: ONE persistent launch factors every matrix of the batch.
//
// The recursive formulation in chol.cu spends a launch per 64-wide panel and per trailing update; at the C2 shape
// (20 matrices of order 2000) the 32 serial panel launches and the ramp / tail of ~30 mid-size GEMM launches cost more
// than the arithmetic.  Here the factorisation is a left-looking tile algorithm (64 x 64 tiles) whose tasks are handed
// out by a ticket counter to CTAs that stay resident:
//   task (z, i, j), i >= j:   C = A_ij - sum_{k<j} L_ik L_jk'          (DMMA, cp.async pipeline, full K at once)
//                             i == j:  L_jj = chol(C), X_jj = L_jj^-1    (register-resident 64 x 64 factorisation)
//                             i >  j:  L_ij = C X_jj'                    (DMMA out of shared memory)
// Dependencies are tracked per tile ROW: prog[z][i] = number of finished tiles of row i (a row finishes left to
// right).  A task spins (one thread, acquire loads) until the row counters of its two operand rows cover the next
// 64-wide k-slab, so it accumulates everything that is ready and only waits for what is not.  Tickets are issued in
// an order in which every dependency of a task has a smaller ticket -- column by column, with the diagonal task of
// column c+1 right behind the one tile it needs from column c -- so the CTAs that are running can always make
// progress: no co-residency assumption, no deadlock, and the 64 x 64 factorisations (the serial chain) overlap the
// bulk updates of the previous column.  Rows n..mrows-1 are right-hand sides, as in potrf_batched.
#include "chol.cuh"
#include "dgemm_tile.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cuda.h>           // CUtensorMap (type and enums only; the encoder is fetched through the runtime)

namespace sb {
namespace {

using namespace tile;

constexpr int NB = CHOL_NB;          // 64
constexpr int SLD = NB + 4;
constexpr int WLD = 32 + 4;
constexpr int DT = 256;              // threads per CTA
constexpr int DSTAGES = 4;
constexpr int KT_PER_SLAB = NB / BK;
constexpr int STAGE_ELEMS = 3 * BK * NB;                           // A (128 rows) and B (64 rows) operand tiles of one stage
constexpr int SM_PIPE = DSTAGES * STAGE_ELEMS;                     // doubles
constexpr int SM_FACTOR = 3 * NB * SLD;                            // X, T (128 rows); S, W, invd alias T in diagonal tasks
constexpr int SM_DOUBLES = SM_PIPE > SM_FACTOR ? SM_PIPE : SM_FACTOR;
constexpr int YIELD_SLOTS = 512;                                   // per-SM yield counters (%smid < 512)
constexpr long long SPIN_LIMIT = 4000000000LL;                     // ~2 s of clock64: a lost dependency aborts, not hangs

struct DagArgs {
    double* A; long lda, strideA;
    double* Dinv; long ldd, strideD, blockD;
    int n, mrows, batch, nt, mtp, total;        // nt column tiles (64), mtp row pairs (128 rows)
    int* info; int* prog; int* ticket;      // prog: batch * mtp counters; ticket[0] = next task, ticket[1] = abort flag
    int* xflag;                             // batch * nt: 1 once L_jj and its inverse are in global memory
    int* smflag;                            // per SM: number of 64 x 64 factorisations running there (NULL: no yielding)
    int nowait;                             // experiments only (SB200_DAG_NOWAIT=1): ignore dependencies, results are wrong
    long long* stats;                       // instrumentation (may be NULL): per CTA [total, waiting, diagonal, solve, tasks, factor, chunk, ndiag]
};

__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double rsqrt_normal(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    double e = fma(x, -(y * y), 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}
__device__ __forceinline__ bool normal_positive(double x) {
    return (unsigned)(__double2hiint(x) - 0x00100000) < 0x7fe00000u;
}


// ---- TMA / mbarrier plumbing of the task-graph kernel's operand pipeline ------------------------------------------
// Operand tiles (64 rows x 16 k-contiguous doubles = 64 x 128 B) are fetched by cp.async.bulk.tensor (SASS UTMALDG) with
// the 128-byte swizzle: row r of a tile lies at r * 128 B and its 16-byte chunk c at position c ^ (r & 7).  A warp's
// DMMA fragment loads stay bank-conflict free under that swizzle when lane t of k-step s takes
// k = 8 (t >> 1) + 2 s + (t & 1) instead of 4 s + t -- any assignment of the 16 k's of a tile to (step, lane) is valid
// as long as both operands use the same one.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// a transfer that never completes would otherwise hang the GPU: trap after ~2 s
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > SPIN_LIMIT) __trap();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}

// all threads: returns min(prog_a, prog_b) once it is >= need (uniform); s_avail is a shared scratch word
__device__ __forceinline__ int wait_progress(const int* pa, const int* pb, int need, int* s_avail, int* abort_flag,
                                             int tid, long long& waited) {
    if (tid == 0) {
        int v;
        const long long t0 = clock64();
        for (;;) {
            v = ld_acquire(pa);
            if (pb) v = min(v, ld_acquire(pb));
            if (v >= need) break;
            if (clock64() - t0 > SPIN_LIMIT || ld_acquire(abort_flag)) {
                atomicExch(abort_flag, 1);
                v = 1 << 30;
                break;
            }
            __nanosleep(40);
        }
        waited += clock64() - t0;
        *s_avail = v;
    }
    __syncthreads();
    const int v = *s_avail;
    __syncthreads();
    return v;
}

// Cholesky of the 64 x 64 block held in S (lower part, identity padded by the caller) and its inverse:
// S <- L (zeros above the diagonal), X <- L^-1.  Same scheme as potrf_panel_kernel (chol.cu): register-resident
// right-looking factorisation, two columns per barrier, then 16 x 16 diagonal inverses by substitution and two levels
// of DMMA block merges.  Returns the 1-based index (offset j0) of the first non-positive pivot, or 0.
__device__ int factor_invert_64(double* S, double* X, double* W, double* invd, int j0, int tid) {
    const int ti = tid >> 4, tc = tid & 15;
    double a[4][4];
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            const int i = ti + 16 * ii, c = tc + 16 * cc;
            a[ii][cc] = (c <= i) ? S[i * SLD + c] : 0.0;
        }
    int bad_at = 0;
#pragma unroll
    for (int jb = 0; jb < 4; jb++) {
#pragma unroll 1
        for (int jr = 0; jr < 16; jr += 2) {
            const int j = jb * 16 + jr;
            double* cb0 = W + (jr & 2) * NB;
            double* cb1 = cb0 + NB;
            if (tc == jr) {
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
            double x0i[4], x1i[4], x0c[4], x1c[4];
#pragma unroll
            for (int e = 0; e < 4; e++)
                if (e >= jb) {
                    x0i[e] = cb0[ti + 16 * e];
                    x1i[e] = cb1[ti + 16 * e];
                    x0c[e] = cb0[tc + 16 * e];
                    x1c[e] = cb1[tc + 16 * e];
                }
            const double det = fma(p0, p1, -(q * q));
            double r0 = rsqrt_normal(p0);
            const double rd = rsqrt_normal(det);
            double s0 = p0 * r0;
            double l10 = q * r0;
            double inv1 = rd * s0;
            double l11 = det * rd * r0;
            if (!(normal_positive(p0) && normal_positive(det))) {
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
    __syncthreads();                      // the last column pair has been read by everyone
#pragma unroll
    for (int ii = 0; ii < 4; ii++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            const int i = ti + 16 * ii, c = tc + 16 * cc;
            S[i * SLD + c] = (c <= i) ? a[ii][cc] : 0.0;
            X[i * SLD + c] = 0.0;
        }
    __syncthreads();
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
                    dmma_8x8x4(c0, c1, av, bv);
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
                    dmma_8x8x4(c0, c1, av, bv);
                }
                X[(b0 + sz + 8 * mi + g) * SLD + b0 + 8 * ni + 2 * t] = c0;
                X[(b0 + sz + 8 * mi + g) * SLD + b0 + 8 * ni + 2 * t + 1] = c1;
            }
            __syncthreads();
        }
    }
    return bad_at;
}

// out (64 x 64) = T (64 rows, SLD) X' for the rows [row_lo, row_hi) of a 64-row chunk, all 8 warps (32 x 16 each);
// stored to dst (row 0 of the chunk), columns < nbj
__device__ __forceinline__ void solve_chunk64(const double* T, const double* X, double* dst, long lda, int row_lo,
                                              int row_hi, int nbj, int warp, int g, int t) {
    constexpr int MI = 4, NI = 2;
    const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 16;
    double out[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) out[i][j][0] = out[i][j][1] = 0.0;
#pragma unroll
    for (int k4 = 0; k4 < NB / 4; k4++) {
        if (4 * k4 > wn0 + 15) break;                                 // X[n][k] == 0 for k > n (warp-uniform)
        double a[MI], b[NI];
#pragma unroll
        for (int i = 0; i < MI; i++) a[i] = T[(wm0 + i * 8 + g) * SLD + 4 * k4 + t];
#pragma unroll
        for (int j = 0; j < NI; j++) b[j] = X[(wn0 + j * 8 + g) * SLD + 4 * k4 + t];
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
            for (int j = 0; j < NI; j++) dmma_8x8x4(out[i][j][0], out[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < MI; i++) {
        const int row = wm0 + i * 8 + g;
        if (row < row_lo || row >= row_hi) continue;
#pragma unroll
        for (int j = 0; j < NI; j++) {
            const int col = wn0 + j * 8 + 2 * t;
            double* ptr = dst + (long)row * lda + col;
            if (col + 1 < nbj) *reinterpret_cast<double2*>(ptr) = make_double2(out[i][j][0], out[i][j][1]);
            else if (col < nbj) ptr[0] = out[i][j][0];
        }
    }
}

template <bool TMA>
__global__ void __launch_bounds__(DT, 2) potrf_dag_kernel(DagArgs p, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(16) double sm_raw[];
    // the swizzled TMA tiles need 1024-byte alignment (the launch reserves the slack)
    double* sm = TMA ? reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~(uintptr_t)1023) : sm_raw;
    __shared__ int s_task, s_avail;
    __shared__ __align__(8) unsigned long long full_bar[DSTAGES], empty_bar[DSTAGES];
    unsigned git = 0;                                     // k-tiles this CTA has consumed so far (all tasks): stage = git % S
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    constexpr int WM = 32, WN = 32, MI = WM / 8, NI = WN / 8, WARPS_N = NB / WN;      // 4 x 2 warps on a 128 x 64 tile
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    int* abort_flag = p.ticket + 1;
    unsigned smid;
    asm("mov.u32 %0, %%smid;\n" : "=r"(smid));
    int* my_yield = (p.smflag && smid < (unsigned)YIELD_SLOTS) ? p.smflag + smid : nullptr;
    const int avail0 = p.nowait ? (1 << 30) : 0;
    if (TMA) {
        if (tid == 0) {
            for (int i = 0; i < DSTAGES; i++) {
                mbar_init(&full_bar[i], 1);               // one arrival (the producer's expect_tx) + the bytes
                mbar_init(&empty_bar[i], DT / 32);        // one arrival per warp
            }
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
    }

    // ticket -> task decoding state (tickets of one CTA increase, so the column segment only moves forward)
    int seg = -1, seg_start = 0, seg_size = p.batch;
    long long c_wait = 0, c_diag = 0, c_solve = 0, c_tasks = 0, c_factor = 0, c_chunk = 0, c_ndiag = 0;
    const long long c_begin = clock64();

    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(p.ticket, 1);
        __syncthreads();
        const int ticket = s_task;
        if (ticket >= p.total) break;
        c_tasks++;
        while (ticket >= seg_start + seg_size) {
            seg_start += seg_size;
            seg++;
            const int a = ((seg & 1) && (seg + 1) / 2 < p.mtp) ? 1 : 0, b = (seg + 1 < p.nt) ? 1 : 0;
            seg_size = p.batch * (a + b + (p.mtp - seg / 2 - 1 - a));
        }
        int tI, tj;                                       // row pair, tile column
        const int rr = ticket - seg_start, z = rr % p.batch;
        {
            int sub = rr / p.batch;
            if (seg < 0) {
                tI = tj = 0;
            } else {
                const int a = ((seg & 1) && (seg + 1) / 2 < p.mtp) ? 1 : 0;
                const bool b = seg + 1 < p.nt;
                if (a && sub == 0) {                      // the tile the next diagonal task is waiting for
                    tI = (seg + 1) / 2;
                    tj = seg;
                } else {
                    sub -= a;
                    if (b && sub == 0) {                  // the next column's diagonal task
                        tI = (seg + 1) / 2;
                        tj = seg + 1;
                    } else {
                        if (b) sub--;
                        tI = seg / 2 + 1 + a + sub;
                        tj = seg;
                    }
                }
            }
        }
        const int tJ = tj >> 1;                           // row pair holding tile row tj
        const bool diag = tI == tJ;
        const int diag_off = (tj & 1) ? NB : 0;           // where tile row tj sits inside its pair
        double* A = p.A + (long)z * p.strideA;
        const int r0 = tI * 2 * NB, c0 = tj * NB;
        const int nbj = min(NB, p.n - c0);                // width of this column of tiles
        const int rows_here = min(2 * NB, p.mrows - r0);
        const int* prog_i = p.prog + (long)z * p.mtp + tI;
        const int* prog_j = p.prog + (long)z * p.mtp + tJ;
        const int ktiles = c0 / BK;
        const bool idle_warp = diag && diag_off && warp < 4;   // rows above the diagonal tile: nothing to compute

        double* As = sm;
        double* Bs = sm + DSTAGES * BK * 2 * NB;
        int avail = avail0;                               // leading k-slabs (tiles) of both operand rows known finished

        // accumulators start at -C (C = A_ij - sum L L'  ->  -(sum L L' - C))
        double acc[MI][NI][2];
#pragma unroll
        for (int i = 0; i < MI; i++) {
            const int row = r0 + wm0 + i * 8 + g;
#pragma unroll
            for (int j = 0; j < NI; j++) {
                const int col = c0 + wn0 + j * 8 + 2 * t;
                double v0 = 0.0, v1 = 0.0;
                if (!idle_warp && row < p.mrows && col + 1 < c0 + nbj) {
                    const double2 o = *reinterpret_cast<const double2*>(A + (long)row * p.lda + col);
                    v0 = o.x;
                    v1 = o.y;
                } else if (!idle_warp && row < p.mrows && col < c0 + nbj) {
                    v0 = A[(long)row * p.lda + col];
                }
                acc[i][j][0] = -v0;
                acc[i][j][1] = -v1;
            }
        }

        if constexpr (TMA) {
            // One elected thread (lane 0 of warp 0) produces: it waits until all warps have released the slot and, at
            // every 64-wide k-slab, until the operand rows' progress counters cover it, then posts the byte count and
            // issues the bulk-tensor loads.  Everyone consumes: wait for the slot's bytes, multiply, release the slot.
            // No CTA-wide barrier inside the loop -- warps drift up to DSTAGES - 1 tiles apart.
            const unsigned tx_bytes = (unsigned)((diag ? 2 : 3) * NB * BK * sizeof(double));
            auto produce = [&](int ktl) {
                const unsigned it = git + (unsigned)ktl, slot = it % DSTAGES, ph = (it / DSTAGES) & 1;
                mbar_wait(&empty_bar[slot], ph ^ 1);
                const int slab = ktl / KT_PER_SLAB;
                // The 64 x 64 factorisations are the serial chain of the whole factorisation, and their DFMAs share the
                // FP64 pipe with this CTA's DMMAs (18 us alone, 36 us beside a busy neighbour): while one runs on this
                // SM, an update task stops feeding its pipeline at the next slab boundary.
                if (my_yield && !diag && (ktl % KT_PER_SLAB) == 0 && ld_acquire(my_yield) > 0) {
                    const long long t0 = clock64();
                    while (ld_acquire(my_yield) > 0 && clock64() - t0 < 200000) __nanosleep(200);
                    c_wait += clock64() - t0;
                }
                if (slab >= avail) {
                    const long long t0 = clock64();
                    for (;;) {
                        int v = ld_acquire(prog_i);
                        if (!diag) v = min(v, ld_acquire(prog_j));
                        if (v > slab) {
                            avail = v;
                            break;
                        }
                        if (clock64() - t0 > SPIN_LIMIT || ld_acquire(abort_flag)) {
                            atomicExch(abort_flag, 1);
                            avail = 1 << 30;
                            break;
                        }
                        __nanosleep(40);
                    }
                    c_wait += clock64() - t0;
                    asm volatile("fence.proxy.async;\n" ::: "memory");      // the loads below go through the async proxy
                }
                double* dst = sm + (size_t)slot * STAGE_ELEMS;
                mbar_expect_tx(&full_bar[slot], tx_bytes);
                tma_load_3d(dst, &tmap, ktl * BK, r0, z, &full_bar[slot]);
                tma_load_3d(dst + NB * BK, &tmap, ktl * BK, r0 + NB, z, &full_bar[slot]);
                if (!diag) tma_load_3d(dst + 2 * NB * BK, &tmap, ktl * BK, c0, z, &full_bar[slot]);
            };
            const bool producer = tid == 0;
            if (producer) {
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // the buffers were last touched by ld/st.shared
                for (int s = 0; s < DSTAGES - 1 && s < ktiles; s++) produce(s);
            }
            // swizzled fragment columns of this lane for the four k-steps of a tile (see the helper comment)
            int colk[BK / 4];
#pragma unroll
            for (int k4 = 0; k4 < BK / 4; k4++) colk[k4] = ((((t >> 1) * 4 + k4) ^ g) << 1) | (t & 1);
#pragma unroll 1
            for (int kt = 0; kt < ktiles; kt++) {
                if (producer && kt + DSTAGES - 1 < ktiles) produce(kt + DSTAGES - 1);
                __syncwarp();
                const unsigned it = git + (unsigned)kt, slot = it % DSTAGES, ph = (it / DSTAGES) & 1;
                mbar_wait(&full_bar[slot], ph);
                if (!idle_warp) {
                    const double* as = sm + (size_t)slot * STAGE_ELEMS;
                    const double* bs = diag ? as + diag_off * BK : as + 2 * NB * BK;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 4; k4++) {
                        double a[MI], b[NI];
#pragma unroll
                        for (int i = 0; i < MI; i++) a[i] = as[(wm0 + i * 8 + g) * BK + colk[k4]];
#pragma unroll
                        for (int j = 0; j < NI; j++) b[j] = bs[(wn0 + j * 8 + g) * BK + colk[k4]];
#pragma unroll
                        for (int i = 0; i < MI; i++)
#pragma unroll
                            for (int j = 0; j < NI; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[slot]);
            }
            git += (unsigned)ktiles;
            __syncthreads();                              // every warp is done with the stage buffers: reuse them below
            avail = avail0;                               // (the producer's knowledge is not shared: the solve re-checks)
        } else {
        auto stage_load = [&](int kt) {
            const int slot = kt % DSTAGES;
            load_tile<2 * NB, true, DT>(As + slot * BK * 2 * NB, A, p.lda, r0, p.mrows, kt * BK, c0, true, tid);
            if (!diag) load_tile<NB, true, DT>(Bs + slot * BK * NB, A, p.lda, c0, p.mrows, kt * BK, c0, true, tid);
        };
#pragma unroll 1
        for (int s = 0; s < DSTAGES - 1; s++) {
            if (s < ktiles) {
                if (s / KT_PER_SLAB >= avail)
                    avail = wait_progress(prog_i, diag ? nullptr : prog_j, s / KT_PER_SLAB + 1, &s_avail, abort_flag, tid, c_wait);
                stage_load(s);
            }
            cp_async_commit();
        }
#pragma unroll 1
        for (int kt = 0; kt < ktiles; kt++) {
            cp_async_wait<DSTAGES - 2>();
            __syncthreads();
            {
                const int nx = kt + DSTAGES - 1;
                if (nx < ktiles) {
                    if (nx / KT_PER_SLAB >= avail)
                        avail = wait_progress(prog_i, diag ? nullptr : prog_j, nx / KT_PER_SLAB + 1, &s_avail, abort_flag, tid, c_wait);
                    stage_load(nx);
                }
                cp_async_commit();
            }
            if (idle_warp) continue;
            const double* as = As + (kt % DSTAGES) * BK * 2 * NB;
            // the diagonal task multiplies the pair's rows by the rows of tile row tj, which are part of the same stage
            const double* bs = diag ? as + diag_off * 4 : Bs + (kt % DSTAGES) * BK * NB;
            const int bld = diag ? 2 * NB : NB;
#pragma unroll
            for (int k4 = 0; k4 < BK / 4; k4++) {
                double a[MI], b[NI];
#pragma unroll
                for (int i = 0; i < MI; i++) a[i] = as[((k4 * 2 * NB) + wm0 + i * 8 + g) * 4 + t];
#pragma unroll
                for (int j = 0; j < NI; j++) b[j] = bs[((k4 * bld) + wn0 + j * 8 + g) * 4 + t];
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NI; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();                                  // the pipeline buffers are free: reuse them below
        }

        const long long c_fin = clock64();
        double* X = sm;                                   // [64][SLD]   L_jj^-1
        double* T = sm + NB * SLD;                        // [128][SLD]  updated tile (off-diagonal tasks)
        if (diag) {
            double* S = T;                                // [64][SLD]   the diagonal block, then L_jj; later a 64-row chunk
            double* W = S + NB * SLD;
            double* invd = W + 32 * WLD;
            const int lo = diag_off + nbj;                // rows of the pair below the block: solved after the factorisation
            // park their updated values in place (global memory), the block itself goes to shared memory
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int row = wm0 + i * 8 + g, col = wn0 + j * 8 + 2 * t + e;
                        const double v = -acc[i][j][e];
                        if (row >= diag_off && row < diag_off + NB) {
                            const int br = row - diag_off;
                            S[br * SLD + col] = (br < nbj && col < nbj) ? v : ((br == col) ? 1.0 : 0.0);
                        }
                        if (row >= lo && row < rows_here && col < nbj) A[(long)(r0 + row) * p.lda + c0 + col] = v;
                    }
            __syncthreads();
            const long long c_f0 = clock64();
            if (my_yield && tid == 0) atomicAdd(my_yield, 1);
            const int bad_at = factor_invert_64(S, X, W, invd, c0, tid);
            if (my_yield && tid == 0) atomicSub(my_yield, 1);
            c_factor += clock64() - c_f0;
            c_ndiag++;
            double* D = p.Dinv + (long)z * p.strideD + (long)tj * p.blockD;
            for (int idx = tid; idx < nbj * nbj; idx += DT) {
                const int i = idx / nbj, c = idx % nbj;
                if (c <= i) A[(long)(c0 + i) * p.lda + c0 + c] = S[i * SLD + c];
                D[(long)i * p.ldd + c] = X[i * SLD + c];
            }
            if (tid == 0 && bad_at != 0 && p.info[z] == 0) p.info[z] = bad_at;
            // L_jj^-1 is what every other task of this column waits for: announce it before solving the rows below
            __threadfence();
            __syncthreads();                                          // (also: S is no longer read)
            if (tid == 0) atomicExch(p.xflag + (long)z * p.nt + tj, 1);
            const long long c_c0 = clock64();
            for (int ch = lo; ch < rows_here; ch += NB) {             // rows below the block, 64 at a time: P <- P X'
                const int rows = min(NB, rows_here - ch);
                double* P = A + (long)(r0 + ch) * p.lda + c0;
                for (int idx = tid; idx < NB * NB; idx += DT) {
                    const int r = idx >> 6, c = idx & 63;
                    S[r * SLD + c] = (r < rows && c < nbj) ? P[(long)r * p.lda + c] : 0.0;
                }
                __syncthreads();
                solve_chunk64(S, X, P, p.lda, 0, rows, nbj, warp, g, t);
                __syncthreads();                                      // the chunk buffer is free again
            }
            c_chunk += clock64() - c_c0;
            c_diag += clock64() - c_fin;
        } else {
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) {
                    const int row = wm0 + i * 8 + g, col = wn0 + j * 8 + 2 * t;
                    T[row * SLD + col] = (col < nbj) ? -acc[i][j][0] : 0.0;
                    T[row * SLD + col + 1] = (col + 1 < nbj) ? -acc[i][j][1] : 0.0;
                }
            if (!p.nowait) wait_progress(p.xflag + (long)z * p.nt + tj, nullptr, 1, &s_avail, abort_flag, tid, c_wait);
            const double* D = p.Dinv + (long)z * p.strideD + (long)tj * p.blockD;
            for (int idx = tid; idx < NB * NB; idx += DT) {
                const int i = idx >> 6, c = idx & 63;
                X[i * SLD + c] = (i < nbj && c < nbj) ? __ldcg(D + (long)i * p.ldd + c) : 0.0;
            }
            __syncthreads();
            double out[MI][NI][2];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) out[i][j][0] = out[i][j][1] = 0.0;
#pragma unroll
            for (int k4 = 0; k4 < NB / 4; k4++) {
                if (4 * k4 > wn0 + WN - 1) break;                     // X[n][k] == 0 for k > n (warp-uniform)
                double a[MI], b[NI];
#pragma unroll
                for (int i = 0; i < MI; i++) a[i] = T[(wm0 + i * 8 + g) * SLD + 4 * k4 + t];
#pragma unroll
                for (int j = 0; j < NI; j++) b[j] = X[(wn0 + j * 8 + g) * SLD + 4 * k4 + t];
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NI; j++) dmma_8x8x4(out[i][j][0], out[i][j][1], a[i], b[j]);
            }
#pragma unroll
            for (int i = 0; i < MI; i++) {
                const int row = wm0 + i * 8 + g;
                if (row >= rows_here) continue;
#pragma unroll
                for (int j = 0; j < NI; j++) {
                    const int col = wn0 + j * 8 + 2 * t;
                    double* ptr = A + (long)(r0 + row) * p.lda + c0 + col;
                    if (col + 1 < nbj) *reinterpret_cast<double2*>(ptr) = make_double2(out[i][j][0], out[i][j][1]);
                    else if (col < nbj) ptr[0] = out[i][j][0];
                }
            }
            c_solve += clock64() - c_fin;
        }
        // publish: every thread's stores are fenced, then one thread advances the row counter
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(p.prog + (long)z * p.mtp + tI, 1);
    }
    // a dependency that never arrived (SPIN_LIMIT) aborted the waits: the factors are not valid, say so for every matrix
    if (ld_acquire(abort_flag))
        for (int zz = tid; zz < p.batch; zz += DT) p.info[zz] = 0x7fffffff;
    if (p.stats && tid == 0) {
        long long* st = p.stats + 8 * (long)blockIdx.x;
        st[0] = clock64() - c_begin;
        st[1] = c_wait;
        st[2] = c_diag;
        st[3] = c_solve;
        st[4] = c_tasks;
        st[5] = c_factor;
        st[6] = c_chunk;
        st[7] = c_ndiag;
    }
}

std::atomic<long long*> g_dag_stats{nullptr};

}  // namespace

void debug_dag_stats(long long* p) { g_dag_stats.store(p); }

size_t potrf_dag_scratch_ints(int batch, int mrows) {      // row-pair counters, ticket + abort flag, diagonal flags
    return (size_t)batch * cdiv(mrows, 2 * NB) + 2 + (size_t)batch * cdiv(mrows, NB) + YIELD_SLOTS;
}

bool potrf_dag_usable(const CholBatch& c) {
    return (c.lda % 2 == 0) && (c.strideA % 2 == 0) && (((uintptr_t)c.A & 15) == 0) && c.dag_scratch != nullptr &&
           (long)c.batch * cdiv(c.mrows, NB) * (cdiv(c.mrows, NB) + 1) / 2 < (1L << 30);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static std::atomic<void*> fn{nullptr};
    void* f = fn.load();
    if (!f) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        fn.store(f);
    }
    return (EncodeTiledFn)f;
}

// 0 = cp.async staging, 1 = TMA + mbarrier pipeline (default); SB200_DAG_TMA=0 selects the former (A/B experiments)
static int dag_use_tma() {
    static int on = -1;
    if (on < 0) {
        const char* env = getenv("SB200_DAG_TMA");
        on = (env && env[0] == '0') ? 0 : 1;
    }
    return on;
}

int potrf_dag(const CholBatch& c, cudaStream_t stream) {
    constexpr size_t smem = (size_t)SM_DOUBLES * sizeof(double);
    constexpr size_t smem_tma = smem + 1024;
    static DeviceOnce once;
    static int sm_count[64];
    int dev = 0;
    SB_CUDA(cudaGetDevice(&dev));
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(potrf_dag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SB_CUDA(cudaFuncSetAttribute(potrf_dag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma));
        int n = 0;
        SB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
        if (dev >= 0 && dev < 64) sm_count[dev] = n;
        return 0;
    }));
    int sms = (dev >= 0 && dev < 64) ? sm_count[dev] : 0;
    if (sms <= 0) SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DagArgs a{};
    a.A = c.A; a.lda = c.lda; a.strideA = c.strideA;
    a.Dinv = c.Dinv; a.ldd = c.ldd; a.strideD = c.strideD; a.blockD = c.blockD;
    a.n = c.n; a.mrows = c.mrows; a.batch = c.batch;
    a.nt = cdiv(c.n, NB); a.mtp = cdiv(c.mrows, 2 * NB);
    long total = 0;
    for (int j = 0; j < a.nt; j++) total += (long)(a.mtp - j / 2);
    total *= c.batch;
    a.total = (int)total;
    a.info = c.info;
    a.prog = c.dag_scratch;
    a.ticket = c.dag_scratch + (size_t)c.batch * a.mtp;
    a.xflag = a.ticket + 2;
    {
        static int yield = -1;                     // SB200_DAG_YIELD=0: no yielding to the diagonal tasks (A/B experiments)
        if (yield < 0) {
            const char* env = getenv("SB200_DAG_YIELD");
            yield = (env && env[0] == '0') ? 0 : 1;
        }
        a.smflag = yield ? a.xflag + (size_t)c.batch * a.nt : nullptr;
    }
    a.stats = g_dag_stats.load();
    {
        const char* env = getenv("SB200_DAG_NOWAIT");
        a.nowait = (env && env[0] == '1') ? 1 : 0;
    }
    SB_CUDA(cudaMemsetAsync(c.dag_scratch, 0, sizeof(int) * potrf_dag_scratch_ints(c.batch, c.mrows), stream));
    const int grid = (int)std::min<long>(total, 2L * sms);

    // operand tiles as a 3-d tensor (k, row, matrix): 64 rows x 16 doubles per box, 128-byte swizzle, zero fill out of bounds
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    bool tma = dag_use_tma() != 0 && encode_tiled() != nullptr && c.batch >= 1;
    if (tma) {
        const cuuint64_t dims[3] = {(cuuint64_t)c.lda, (cuuint64_t)c.mrows, (cuuint64_t)c.batch};
        const cuuint64_t strides[2] = {(cuuint64_t)c.lda * sizeof(double), (cuuint64_t)c.strideA * sizeof(double)};
        const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)NB, 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = encode_tiled()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)c.A, dims, strides, box, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) tma = false;                // e.g. a stride the encoder refuses: cp.async staging instead
    }
    auto launch = [&]() {
        if (tma) potrf_dag_kernel<true><<<grid, DT, smem_tma, stream>>>(a, tmap);
        else potrf_dag_kernel<false><<<grid, DT, smem, stream>>>(a, tmap);
    };
    if (gemm_profile_enabled()) {                  // bench.py's roofline leg: CUDA events around this launch
        cudaEvent_t e0, e1;
        SB_TRY(profile_event_pair(&e0, &e1));
        SB_CUDA(cudaEventRecord(e0, stream));
        launch();
        SB_CUDA(cudaEventRecord(e1, stream));
        const double nn = (double)c.n;             // n^3/3 for the factor + n^2 per right-hand-side row
        kernel_profile_add(1, e0, e1, (double)c.batch * (nn * nn * nn / 3.0 + (double)(c.mrows - c.n) * nn * nn));
        SB_LAUNCH_CHECK();
        return 0;
    }
    launch();
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace sb
