// DMMA (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) GEMM for sm_100a.
//
// tcgen05.mma has no f64 kind, so on B200 the FP64 tensor path is the warp-level DMMA.  The kernel
// is a classic 3-stage cp.async pipeline: 64x64 CTA tile (default; 128x64 and 128x128 kept as variants), BK = 16,
// 4 (8) warps, each warp owning a 32x32 (64x32) accumulator block held in registers.  Operand tiles are staged in shared
// memory in a fragment-friendly order so that every LDS.64 of a warp is bank-conflict free:
//   k-contiguous operand : [k/4][row][k%4]  -> a warp's 8x4 fragment is 256 contiguous bytes
//   row-contiguous operand: [k][row + pad 4] -> (4t + g) mod 16 distinct per half-warp
// Algorithmic work: 2*M*N*Keff flop where Keff is the (possibly triangular) k-range per tile.
#include "dgemm.cuh"
#include "dgemm_tile.cuh"
#include <algorithm>
#include <cstdlib>
using std::min;
using std::max;

namespace sb {
namespace {

using namespace tile;

template <int BM, int BN, int WM, int WN, bool AKC, bool BKC>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, (BM * BN <= 64 * 64) ? SB_MIN64 : ((BM * BN <= 64 * 128) ? 2 : 1))
    dgemm_dmma_kernel(GemmArgs p) {
    constexpr int A_TILE = BK * (BM + 4);
    constexpr int B_TILE = BK * (BN + 4);
    constexpr int WARPS_N = BN / WN;
    constexpr int WARPS_M = BM / WM;
    constexpr int NTHREADS = WARPS_M * WARPS_N * 32;
    constexpr int MI = WM / 8, NI = WN / 8;

    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_TILE;

    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int z1 = p.inner_batch > 0 ? z % p.inner_batch : z;
    const long z2 = p.inner_batch > 0 ? z / p.inner_batch : 0;
    // two-level batches may end with a shorter group (the ragged last block pair of a trtri level)
    const int M = (p.last_M > 0 && p.inner_batch > 0 && z2 == (long)gridDim.z / p.inner_batch - 1) ? p.last_M : p.M;
    // longest-k tiles first: with a lower-triangular A the k-range grows with m, so walk the row tiles backwards
    // (lower-triangular B / upper-triangular A and B already start with their longest tiles)
    const int tile_m = (p.flags & GEMM_A_LOWER) ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
    const int tile_n = (p.flags & GEMM_B_LOWER) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
    const int n0 = tile_n * BN;
    // lower-only output: row tiles of column block J start ON the diagonal (row n0), so only the first tile of each
    // column block straddles it (25 % of one tile wasted per block instead of 50 % of two)
    const int m0 = (p.flags & GEMM_C_LOWER) ? n0 + tile_m * BM : tile_m * BM;
    if (m0 >= M) return;
    const int m_end = min(M, m0 + BM), n_end = min(p.N, n0 + BN);

    const double* A = p.A + (long)(p.idxA ? p.idxA[z1] : z1) * p.strideA + z2 * p.strideA2;
    const double* B = p.B + (long)(p.idxB ? p.idxB[z1] : z1) * p.strideB + z2 * p.strideB2;
    double* C = p.C + (long)z1 * p.strideC + z2 * p.strideC2;

    int k_lo = 0, k_hi = p.K;
    if (p.flags & GEMM_A_LOWER) k_hi = min(k_hi, m_end);
    if (p.flags & GEMM_A_UPPER) k_lo = max(k_lo, m0);
    if (p.flags & GEMM_B_LOWER) k_hi = min(k_hi, n_end);
    if (p.flags & GEMM_B_UPPER) k_lo = max(k_lo, n0);
    k_lo &= ~(BK - 1);
    const int ktiles = (k_hi > k_lo) ? (k_hi - k_lo + BK - 1) / BK : 0;

    const bool a16 = ((p.lda & 1) == 0) && (((uintptr_t)A & 15) == 0);
    const bool b16 = ((p.ldb & 1) == 0) && (((uintptr_t)B & 15) == 0);

    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    const bool c16 = ((p.ldc & 1) == 0) && (((uintptr_t)C & 15) == 0);
    const double alpha = p.alpha, beta = p.beta;

    // the cp.async prologue goes out first so that its latency overlaps the (blocking) loads of the C tile below
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < ktiles) {
            load_tile<BM, AKC, NTHREADS>(As + s * A_TILE, A, p.lda, m0, M, k_lo + s * BK, k_hi, a16, tid);
            load_tile<BN, BKC, NTHREADS>(Bs + s * B_TILE, B, p.ldb, n0, p.N, k_lo + s * BK, k_hi, b16, tid);
        }
        cp_async_commit();
    }

    // beta * C is folded into the accumulators up front: alpha (A B' + (beta/alpha) C).  The C loads then overlap
    // the cp.async prologue instead of serialising (load -> fma -> store per fragment) in the epilogue.
    double acc[MI][NI][2];
    const bool fold = (beta != 0.0) && (alpha != 0.0);
    const double cscale = fold ? beta / alpha : 0.0;
#pragma unroll
    for (int i = 0; i < MI; i++) {
        const int row = m0 + wm0 + i * 8 + g;
#pragma unroll
        for (int j = 0; j < NI; j++) {
            const int col = n0 + wn0 + j * 8 + 2 * t;
            double c0 = 0.0, c1 = 0.0;
            if (fold && row < M) {
                const double* ptr = C + (long)row * p.ldc + col;
                if (c16 && col + 1 < p.N) {
                    double2 o = *reinterpret_cast<const double2*>(ptr);
                    c0 = o.x;
                    c1 = o.y;
                } else {
                    if (col < p.N) c0 = ptr[0];
                    if (col + 1 < p.N) c1 = ptr[1];
                }
            }
            acc[i][j][0] = cscale * c0;
            acc[i][j][1] = cscale * c1;
        }
    }

    // warp-level k-range: inside the CTA's k-range a warp whose rows (columns) lie deeper in the triangle only
    // meets stored zeros for part of it, and a warp block entirely above the diagonal of a lower-only C is never
    // written -- those k-slabs are staged (other warps need them) but their DMMAs are skipped
    int wkt_lo = 0, wkt_hi = ktiles;
    {
        const int wr0 = m0 + wm0, wc0 = n0 + wn0;
        int wk_lo = k_lo, wk_hi = k_hi;
        if (p.flags & GEMM_A_LOWER) wk_hi = min(wk_hi, wr0 + WM);
        if (p.flags & GEMM_A_UPPER) wk_lo = max(wk_lo, wr0);
        if (p.flags & GEMM_B_LOWER) wk_hi = min(wk_hi, wc0 + WN);
        if (p.flags & GEMM_B_UPPER) wk_lo = max(wk_lo, wc0);
        if ((p.flags & GEMM_C_LOWER) && wc0 > wr0 + WM - 1) wk_hi = wk_lo;
        if (wr0 >= M || wc0 >= p.N) wk_hi = wk_lo;
        wkt_lo = (wk_lo - k_lo) / BK;
        wkt_hi = wk_hi > wk_lo ? (wk_hi - k_lo + BK - 1) / BK : 0;
    }

    for (int kt = 0; kt < ktiles; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = kt + STAGES - 1;
            if (nx < ktiles) {
                int slot = nx % STAGES;
                load_tile<BM, AKC, NTHREADS>(As + slot * A_TILE, A, p.lda, m0, M, k_lo + nx * BK, k_hi, a16, tid);
                load_tile<BN, BKC, NTHREADS>(Bs + slot * B_TILE, B, p.ldb, n0, p.N, k_lo + nx * BK, k_hi, b16, tid);
            }
            cp_async_commit();
        }
        if (kt < wkt_lo || kt >= wkt_hi) continue;
        const double* as = As + (kt % STAGES) * A_TILE;
        const double* bs = Bs + (kt % STAGES) * B_TILE;
#pragma unroll
        for (int k4 = 0; k4 < BK / 4; k4++) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; i++)
                a[i] = AKC ? as[((k4 * BM) + wm0 + i * 8 + g) * 4 + t] : as[(k4 * 4 + t) * (BM + 4) + wm0 + i * 8 + g];
#pragma unroll
            for (int j = 0; j < NI; j++)
                b[j] = BKC ? bs[((k4 * BN) + wn0 + j * 8 + g) * 4 + t] : bs[(k4 * 4 + t) * (BN + 4) + wn0 + j * 8 + g];
#pragma unroll
            for (int i = 0; i < MI; i++)
#pragma unroll
                for (int j = 0; j < NI; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    const bool lower_only = (p.flags & GEMM_C_LOWER) != 0;
    const double beta_ep = fold ? 0.0 : beta;          // only the degenerate alpha == 0 case still reads C here
#pragma unroll
    for (int i = 0; i < MI; i++) {
        int row = m0 + wm0 + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < NI; j++) {
            int col = n0 + wn0 + j * 8 + 2 * t;
            double* ptr = C + (long)row * p.ldc + col;
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
            // lower-only output: never touch an element above the diagonal of C
            const bool w0 = (col < p.N) && !(lower_only && col > row);
            const bool w1 = (col + 1 < p.N) && !(lower_only && col + 1 > row);
            if (w0 && w1 && c16) {
                if (beta_ep != 0.0) {
                    double2 o = *reinterpret_cast<const double2*>(ptr);
                    v0 += beta_ep * o.x;
                    v1 += beta_ep * o.y;
                }
                *reinterpret_cast<double2*>(ptr) = make_double2(v0, v1);
            } else {
                if (w0) {
                    if (beta_ep != 0.0) v0 += beta_ep * ptr[0];
                    ptr[0] = v0;
                }
                if (w1) {
                    if (beta_ep != 0.0) v1 += beta_ep * ptr[1];
                    ptr[1] = v1;
                }
            }
        }
    }
}

// 2*M*N*K with the exact element-wise triangular k-ranges (tile independent): the algorithmic flop of one problem
double algorithmic_flops(const GemmArgs& g, int, int) {
    double fl = 0.0;
    for (int m = 0; m < g.M; m++) {
        // count (n, k) pairs with nonzero products for this row
        // k-range from A: [alo, ahi); from B (depends on n): handled per case below
        long alo = (g.flags & GEMM_A_UPPER) ? m : 0;
        long ahi = (g.flags & GEMM_A_LOWER) ? (long)min(g.K, m + 1) : g.K;
        long nmax = (g.flags & GEMM_C_LOWER) ? (long)min(g.N, m + 1) : g.N;
        if (ahi <= alo || nmax <= 0) continue;
        if (!(g.flags & (GEMM_B_LOWER | GEMM_B_UPPER))) {
            fl += 2.0 * (double)(ahi - alo) * (double)nmax;
        } else {
            for (long n = 0; n < nmax; n++) {
                long lo = alo, hi = ahi;
                if (g.flags & GEMM_B_UPPER) lo = lo > n ? lo : n;
                if (g.flags & GEMM_B_LOWER) hi = hi < n + 1 ? hi : n + 1;
                if (hi > lo) fl += 2.0 * (double)(hi - lo);
            }
        }
    }
    return fl;
}

template <int BM, int BN, int WM, int WN, bool AKC, bool BKC>
int launch_one(const GemmArgs& g, int batch, cudaStream_t stream) {
    constexpr size_t smem = (size_t)STAGES * (BK * (BM + 4) + BK * (BN + 4)) * sizeof(double);
    auto kern = dgemm_dmma_kernel<BM, BN, WM, WN, AKC, BKC>;
    static DeviceOnce once;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return 0;
    }));
    constexpr int NTHREADS = (BM / WM) * (BN / WN) * 32;
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), batch);
    if (gemm_profile_enabled()) {
        cudaEvent_t e0, e1;
        SB_TRY(profile_event_pair(&e0, &e1));
        SB_CUDA(cudaEventRecord(e0, stream));
        kern<<<grid, NTHREADS, smem, stream>>>(g);
        SB_CUDA(cudaEventRecord(e1, stream));
        double fl = algorithmic_flops(g, BM, BN) * batch;
        if (g.last_M > 0 && g.inner_batch > 0) {             // the last group of a two-level batch is shorter
            GemmArgs gl = g;
            gl.M = g.last_M;
            fl += (algorithmic_flops(gl, BM, BN) - algorithmic_flops(g, BM, BN)) * g.inner_batch;
        }
        gemm_profile_add(e0, e1, fl);
        SB_LAUNCH_CHECK();
        return 0;
    }
    kern<<<grid, NTHREADS, smem, stream>>>(g);
    SB_LAUNCH_CHECK();
    return 0;
}

template <int BM, int BN, int WM, int WN>
int launch_layout(const GemmArgs& g, int batch, cudaStream_t stream) {
    if (g.a_kc) {
        if (g.b_kc) return launch_one<BM, BN, WM, WN, true, true>(g, batch, stream);
        return launch_one<BM, BN, WM, WN, true, false>(g, batch, stream);
    }
    if (g.b_kc) return launch_one<BM, BN, WM, WN, false, true>(g, batch, stream);
    return launch_one<BM, BN, WM, WN, false, false>(g, batch, stream);
}

}  // namespace

double dgemm_algorithmic_flops(const GemmArgs& g) { return algorithmic_flops(g, 0, 0); }

static std::atomic<long> g_tma_launches{0};
long dgemm_tma_launch_count() { return g_tma_launches.load(); }

int launch_dgemm(const GemmArgs& g, int batch, int tile, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return 0;
    SB_CHECK_ARG(batch <= 65535, "dgemm: batch %d exceeds grid.z limit", batch);
    SB_CHECK_ARG(cdiv(g.M, 128) <= 65535, "dgemm: M %d too large", g.M);
    // measured on B200 (scripts/gemm_launch_table.py, gemm_variants.py): 64x64 tiles with 4 warps at 4 CTAs/SM match the
    // 128x64 tile at 2 CTAs/SM on a large square product (32.7 vs 32.3 TFLOP/s at 4096^3) and beat it on every
    // triangular / mid-size batched shape of this path (less tile-granularity waste, better wave balance, more CTAs to
    // hide each other's prologue / epilogue): -8 % GEMM time over a C2 NLL+grad step.  The 128x64 tile itself had
    // replaced 128x128 at 1 CTA/SM for the same reasons (32.6 vs 31.1 TFLOP/s at 4096^3, 23.2 vs 20.7 on the batched
    // trailing update, 19.4 vs 15.2 on the predict product).
    if (tile < 0) {
        const char* env = getenv("SB200_GEMM_TILE");              // experiments only
        tile = env ? atoi(env) : 7;
        // default: TMA-staged 64x64 tiles (scripts/gemm_tma_ab.py on B200: 33.7 vs 32.8 TFLOP/s at 4096^3, 1-3 % faster on
        // the batched mid-size products of the trtri / lauum / predict steps); the smallest products (K < 128, a few
        // k-tiles per CTA) do not amortise the barrier set-up and keep the cp.async kernel
        if (!env && g.K >= 128) tile = 8;
    }
    if (tile == 8 || tile == 9) {                      // TMA + mbarrier pipeline (dgemm_tma.cu), 64x64 / 128x64 tiles
        SB_CHECK_ARG(batch <= 65535, "dgemm: batch %d exceeds grid.z limit", batch);
        const int rc = launch_dgemm_tma(g, batch, tile == 8 ? 64 : 128, stream);
        if (rc != -2) {
            if (rc == 0) g_tma_launches.fetch_add(1, std::memory_order_relaxed);
            return rc;
        }
        tile = 7;                                      // operands a tensor map cannot describe: cp.async staging
    }
    if (tile == 7) return launch_layout<64, 64, 32, 32>(g, batch, stream);
    if (tile == 1) return launch_layout<128, 64, 32, 32>(g, batch, stream);
    if (tile == 2) return launch_layout<128, 128, 32, 32>(g, batch, stream);     // 16 warps
    if (tile == 3) return launch_layout<128, 64, 32, 16>(g, batch, stream);      // 16 warps, narrow
    if (tile == 4) return launch_layout<64, 128, 32, 32>(g, batch, stream);      // 2 CTAs / SM
    return launch_layout<128, 128, 64, 32>(g, batch, stream);
}

}  // namespace sb
