// DMMA GEMM with TMA operand staging (cp.async.bulk.tensor -> SASS UTMALDG) and an mbarrier full/empty pipeline.
//
// Same contract as dgemm_dmma_kernel (dgemm.cu): C = alpha Aop Bop' + beta C with optional triangular k-ranges, batched
// (two levels, gather indices), four operand-layout combinations.  What changes is the tile movement:
//   * one elected thread issues bulk-tensor loads of 64-row operand boxes into 128-byte-swizzled shared memory and
//     posts the byte count on the stage's "full" mbarrier; every warp waits on it, multiplies, and releases the stage
//     on its "empty" mbarrier.  There is no CTA-wide barrier in the k loop and no per-thread address arithmetic;
//   * k-contiguous operand  : tensor (k, row, z1, z2), box 16 x 64       -> shared [row][16 k];
//     row-contiguous operand: tensor (row % 16, k, row / 16, z1, z2), box 16 x 16 x 4 -> shared [row / 16][k][16 rows];
//     both are 128-byte rows whose 16-byte chunks the hardware XORs with (row index & 7);
//   * lane t of k-step s takes k = 8 (t >> 1) + 2 ((t + s) & 3) + (t & 1) of the 16-wide tile (instead of 4 s + t): with
//     that assignment the fragment loads of BOTH layouts are bank-conflict free under the swizzle.
// Out-of-range rows / k are zero-filled by the tensor map bounds.  Shapes the tensor maps cannot describe (odd leading
// dimensions, row-contiguous operands whose extent is not a multiple of 16 and not padded) return -2: the caller
// falls back to the cp.async kernel.
#include "dgemm.cuh"
#include "dgemm_tile.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
using std::min;
using std::max;

namespace sb {
namespace {

using tile::dmma_8x8x4;
constexpr int BK = 16, BN = 64;
constexpr long long SPIN_LIMIT = 4000000000LL;

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
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity))
        if (clock64() - t0 > SPIN_LIMIT) __trap();          // a transfer that never lands must not hang the GPU
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], "
        "[%7];\n" ::"r"(smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}

// 64 operand rows [mn0, mn0 + 64) x 16 k [k0, k0 + 16) of batch element (z1, z2) -> 8 KB of shared memory
template <bool KC>
__device__ __forceinline__ void load_box(double* dst, const CUtensorMap* map, int mn0, int k0, int z1, int z2,
                                         unsigned long long* bar) {
    if (KC) tma_load_4d(dst, map, k0, mn0, z1, z2, bar);
    else tma_load_5d(dst, map, 0, k0, mn0 >> 4, z1, z2, bar);
}

template <int BM, bool AKC, bool BKC, int STAGES>
__global__ void __launch_bounds__(BM * 2, BM == 128 ? 2 : 4)
    dgemm_tma_kernel(GemmArgs p, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA_last,
                     const __grid_constant__ CUtensorMap mapB) {
    constexpr int WM = 32, WN = 32, MI = 4, NI = 4, WARPS_N = BN / WN, NWARPS = (BM / WM) * WARPS_N;
    constexpr int A_ELEMS = BM * BK, STAGE_ELEMS = (BM + BN) * BK;
    extern __shared__ __align__(16) double smem_raw[];
    double* sm = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES];

    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int z1 = p.inner_batch > 0 ? z % p.inner_batch : z;
    const int z2 = p.inner_batch > 0 ? z / p.inner_batch : 0;
    // the last group of a two-level batch may have fewer rows: its own tensor map bounds the rows, so the loads never
    // touch memory below the shorter block (rows >= M of any other tile are zero-filled by the map as well)
    const bool last_group = p.last_M > 0 && p.inner_batch > 0 && z2 == (int)gridDim.z / p.inner_batch - 1;
    const int M = last_group ? p.last_M : p.M;
    const CUtensorMap* mA = last_group ? &mapA_last : &mapA;
    const int tile_m = (p.flags & GEMM_A_LOWER) ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
    const int tile_n = (p.flags & GEMM_B_LOWER) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x;
    const int n0 = tile_n * BN;
    const int m0 = (p.flags & GEMM_C_LOWER) ? n0 + tile_m * BM : tile_m * BM;
    if (m0 >= M) return;
    const int m_end = min(M, m0 + BM), n_end = min(p.N, n0 + BN);

    // batch coordinates of the operands inside their tensor maps (a shared operand has extent 1)
    const int za = p.strideA == 0 ? 0 : (p.idxA ? p.idxA[z1] : z1), za2 = p.strideA2 == 0 ? 0 : z2;
    const int zb = p.strideB == 0 ? 0 : (p.idxB ? p.idxB[z1] : z1), zb2 = p.strideB2 == 0 ? 0 : z2;
    double* C = p.C + (long)z1 * p.strideC + (long)z2 * p.strideC2;

    int k_lo = 0, k_hi = p.K;
    if (p.flags & GEMM_A_LOWER) k_hi = min(k_hi, m_end);
    if (p.flags & GEMM_A_UPPER) k_lo = max(k_lo, m0);
    if (p.flags & GEMM_B_LOWER) k_hi = min(k_hi, n_end);
    if (p.flags & GEMM_B_UPPER) k_lo = max(k_lo, n0);
    k_lo &= ~(BK - 1);
    const int ktiles = (k_hi > k_lo) ? (k_hi - k_lo + BK - 1) / BK : 0;

    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
    const bool c16 = ((p.ldc & 1) == 0) && (((uintptr_t)C & 15) == 0);
    const double alpha = p.alpha, beta = p.beta;

    if (tid == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], NWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto produce = [&](int kt) {
        const int slot = kt % STAGES, ph = (kt / STAGES) & 1;
        mbar_wait(&empty_bar[slot], ph ^ 1);
        double* dst = sm + (size_t)slot * STAGE_ELEMS;
        const int k0 = k_lo + kt * BK;
        mbar_expect_tx(&full_bar[slot], (unsigned)(STAGE_ELEMS * sizeof(double)));
#pragma unroll
        for (int h = 0; h < BM / 64; h++) load_box<AKC>(dst + h * 64 * BK, mA, m0 + 64 * h, k0, za, za2, &full_bar[slot]);
        load_box<BKC>(dst + A_ELEMS, &mapB, n0, k0, zb, zb2, &full_bar[slot]);
    };
    const bool producer = tid == 0;
    if (producer)
        for (int s = 0; s < STAGES - 1 && s < ktiles; s++) produce(s);

    // beta * C folded into the accumulators up front (overlaps the first transfers)
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

    // warp-level k-range (see dgemm.cu): slabs a warp only meets stored zeros in are staged but not multiplied
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

    // per-lane fragment offsets under the 128-byte swizzle (doubles, relative to a 64-row box; see the header comment)
    //   k-contiguous  : row * 16 + (((k >> 1) ^ (row & 7)) << 1) + (k & 1),           row & 7 == g
    //   row-contiguous: ((row >> 4) * 16 + k) * 16 + ((((row & 15) >> 1) ^ (k & 7)) << 1) + (row & 1)
    int kk[4];
#pragma unroll
    for (int s = 0; s < 4; s++) kk[s] = 8 * (t >> 1) + 2 * ((t + s) & 3) + (t & 1);
    int offA[4][AKC ? 1 : 2], offB[4][BKC ? 1 : 2];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int k = kk[s];
        if (AKC) offA[s][0] = g * BK + ((((k >> 1) ^ g) << 1) | (k & 1));
        else {
            offA[s][0] = k * BK + ((((g >> 1)) ^ (k & 7)) << 1) + (g & 1);            // rows 0..7 of a 16-row group
            offA[s][AKC ? 0 : 1] = k * BK + (((4 + (g >> 1)) ^ (k & 7)) << 1) + (g & 1);   // rows 8..15
        }
        if (BKC) offB[s][0] = g * BK + ((((k >> 1) ^ g) << 1) | (k & 1));
        else {
            offB[s][0] = k * BK + ((((g >> 1)) ^ (k & 7)) << 1) + (g & 1);
            offB[s][BKC ? 0 : 1] = k * BK + (((4 + (g >> 1)) ^ (k & 7)) << 1) + (g & 1);
        }
    }

#pragma unroll 1
    for (int kt = 0; kt < ktiles; kt++) {
        if (producer && kt + STAGES - 1 < ktiles) produce(kt + STAGES - 1);
        __syncwarp();
        const int slot = kt % STAGES, ph = (kt / STAGES) & 1;
        mbar_wait(&full_bar[slot], ph);
        if (kt >= wkt_lo && kt < wkt_hi) {
            // warp-relative bases: a k-contiguous box is [row][16], a row-contiguous one [row / 16][16 k][16]
            const double* as = sm + (size_t)slot * STAGE_ELEMS + (AKC ? wm0 * BK : (wm0 >> 4) * 16 * BK);
            const double* bs = sm + (size_t)slot * STAGE_ELEMS + A_ELEMS + (BKC ? wn0 * BK : (wn0 >> 4) * 16 * BK);
            // The tensor maps zero-fill beyond K, but a triangular k-range can end inside a tile (k_hi < K): what lies
            // between k_hi and the end of the tile is real memory (padding, the next row) and must not be multiplied.
            const int kvalid = k_hi - (k_lo + kt * BK);                // >= 16 except in a ragged last tile
#pragma unroll
            for (int s = 0; s < 4; s++) {
                double a[MI], b[NI];
#pragma unroll
                for (int i = 0; i < MI; i++)
                    a[i] = AKC ? as[i * 8 * BK + offA[s][0]] : as[(i >> 1) * 16 * BK + offA[s][AKC ? 0 : (i & 1)]];
#pragma unroll
                for (int j = 0; j < NI; j++)
                    b[j] = BKC ? bs[j * 8 * BK + offB[s][0]] : bs[(j >> 1) * 16 * BK + offB[s][BKC ? 0 : (j & 1)]];
                if (kk[s] >= kvalid) {
#pragma unroll
                    for (int i = 0; i < MI; i++) a[i] = 0.0;
#pragma unroll
                    for (int j = 0; j < NI; j++) b[j] = 0.0;
                }
#pragma unroll
                for (int i = 0; i < MI; i++)
#pragma unroll
                    for (int j = 0; j < NI; j++) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[slot]);
    }

    const bool lower_only = (p.flags & GEMM_C_LOWER) != 0;
    const double beta_ep = fold ? 0.0 : beta;
#pragma unroll
    for (int i = 0; i < MI; i++) {
        int row = m0 + wm0 + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < NI; j++) {
            int col = n0 + wn0 + j * 8 + 2 * t;
            double* ptr = C + (long)row * p.ldc + col;
            double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
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

// Tensor map of one operand: `rows` logical rows (M or N), K columns; kc = k-contiguous.  false: not describable.
bool make_map(CUtensorMap* map, const double* base, bool kc, long ld, int rows, int K, long stride1, long stride2,
              int batch1, int batch2) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (ld & 1) || (stride1 & 1) || (stride2 & 1) || ld <= 0) return false;
    const cuuint64_t e1 = stride1 == 0 ? 1 : (cuuint64_t)max(batch1, 1), e2 = stride2 == 0 ? 1 : (cuuint64_t)max(batch2, 1);
    // a unit-extent dimension still needs a legal (multiple of 16 bytes, non-zero) stride
    const cuuint64_t s1 = (cuuint64_t)(stride1 == 0 ? 16 : stride1) * sizeof(double);
    const cuuint64_t s2 = (cuuint64_t)(stride2 == 0 ? 16 : stride2) * sizeof(double);
    memset(map, 0, sizeof(*map));
    CUresult r;
    if (kc) {
        const cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, e1, e2};
        const cuuint64_t strides[3] = {(cuuint64_t)ld * sizeof(double), s1, s2};
        const cuuint32_t box[4] = {16, 64, 1, 1}, es[4] = {1, 1, 1, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // rows in groups of 16 contiguous doubles: the last group must not run past the row
        if ((rows & 15) != 0 && ld < round_up(rows, 16)) return false;
        const cuuint64_t dims[5] = {16, (cuuint64_t)K, (cuuint64_t)cdiv(rows, 16), e1, e2};
        const cuuint64_t strides[4] = {(cuuint64_t)ld * sizeof(double), 128, s1, s2};
        const cuuint32_t box[5] = {16, 16, 4, 1, 1}, es[5] = {1, 1, 1, 1, 1};
        r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    return r == CUDA_SUCCESS;
}

template <int BM, bool AKC, bool BKC>
int launch_tma(const GemmArgs& g, int batch, const CUtensorMap& mapA, const CUtensorMap& mapA_last, const CUtensorMap& mapB,
               cudaStream_t stream) {
    constexpr int STAGES = BM == 128 ? 4 : 3;
    constexpr size_t smem = (size_t)STAGES * (BM + BN) * BK * sizeof(double) + 1024;
    auto kern = dgemm_tma_kernel<BM, AKC, BKC, STAGES>;
    static DeviceOnce once;
    SB_TRY(once.run([&]() -> int {
        SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        return 0;
    }));
    dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), batch);
    if (gemm_profile_enabled()) {
        cudaEvent_t e0, e1;
        SB_TRY(profile_event_pair(&e0, &e1));
        SB_CUDA(cudaEventRecord(e0, stream));
        kern<<<grid, BM * 2, smem, stream>>>(g, mapA, mapA_last, mapB);
        SB_CUDA(cudaEventRecord(e1, stream));
        double fl = dgemm_algorithmic_flops(g) * batch;
        if (g.last_M > 0 && g.inner_batch > 0) {
            GemmArgs gl = g;
            gl.M = g.last_M;
            fl += (dgemm_algorithmic_flops(gl) - dgemm_algorithmic_flops(g)) * g.inner_batch;
        }
        gemm_profile_add(e0, e1, fl);
        SB_LAUNCH_CHECK();
        return 0;
    }
    kern<<<grid, BM * 2, smem, stream>>>(g, mapA, mapA_last, mapB);
    SB_LAUNCH_CHECK();
    return 0;
}

template <int BM>
int launch_layout(const GemmArgs& g, int batch, const CUtensorMap& a, const CUtensorMap& al, const CUtensorMap& b,
                  cudaStream_t stream) {
    if (g.a_kc) {
        if (g.b_kc) return launch_tma<BM, true, true>(g, batch, a, al, b, stream);
        return launch_tma<BM, true, false>(g, batch, a, al, b, stream);
    }
    if (g.b_kc) return launch_tma<BM, false, true>(g, batch, a, al, b, stream);
    return launch_tma<BM, false, false>(g, batch, a, al, b, stream);
}

}  // namespace

// bm: 64 or 128.  Returns -2 (nothing launched) when the operands cannot be described by tensor maps.
int launch_dgemm_tma(const GemmArgs& g, int batch, int bm, cudaStream_t stream) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return 0;
    if (g.K <= 0) return -2;                          // beta-only update: the plain kernel handles it
    const int inner = g.inner_batch > 0 ? g.inner_batch : batch;
    const int outer = g.inner_batch > 0 ? batch / g.inner_batch : 1;
    // with gather indices the extent of the gathered dimension is not known here: any index the caller passes is legal
    const int extA = g.idxA ? 65535 : inner, extB = g.idxB ? 65535 : inner;
    CUtensorMap mapA, mapA_last, mapB;
    if (!make_map(&mapA, g.A, g.a_kc, g.lda, g.M, g.K, g.strideA, g.strideA2, extA, outer)) return -2;
    if (g.last_M > 0 && g.inner_batch > 0) {
        if (!make_map(&mapA_last, g.A, g.a_kc, g.lda, g.last_M, g.K, g.strideA, g.strideA2, extA, outer)) return -2;
    } else {
        mapA_last = mapA;
    }
    if (!make_map(&mapB, g.B, g.b_kc, g.ldb, g.N, g.K, g.strideB, g.strideB2, extB, outer)) return -2;
    if (bm == 128) return launch_layout<128>(g, batch, mapA, mapA_last, mapB, stream);
    return launch_layout<64>(g, batch, mapA, mapA_last, mapB, stream);
}

}  // namespace sb
