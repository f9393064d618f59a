// Operand staging of the DMMA GEMM kernels (cp.async pipeline, fragment-ordered shared-memory tiles) -- shared by
// dgemm.cu and the task-graph Cholesky kernel (chol_dag.cu).
#pragma once
#include "common.cuh"

namespace sb {
namespace tile {

#ifndef SB_BK
#define SB_BK 16
#endif
#ifndef SB_STAGES
#define SB_STAGES 3
#endif
#ifndef SB_MIN64
#define SB_MIN64 4
#endif
constexpr int BK = SB_BK;
constexpr int STAGES = SB_STAGES;

__device__ __forceinline__ void cp_async16(double* smem, const double* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Stage one BMN x BK operand tile.  Elements outside [0,MN) x [kbase,k_hi) are zero-filled by
// cp.async's src-size mechanism, so ragged M/N/K and triangular k-ranges need no special code.
template <int BMN, bool KC, int NTHREADS>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ G, long ld, int mn0, int MN,
                                          int kbase, int k_hi, bool vec16, int tid) {
    const double* dummy = (const double*)((uintptr_t)G & ~(uintptr_t)15);
    if (KC) {
        // chunk order: (16-byte half of a k-quad, row, k-quad).  A warp then writes 16 rows x 32 contiguous bytes of
        // one k-quad = 512 contiguous bytes of shared memory (conflict-free; walking along k first put the four
        // k-quads of a row 4 KB apart, a 4-way bank conflict on every staging write) and still reads whole 32-byte
        // sectors from global memory
        constexpr int CHUNKS = BMN * (BK / 2);
#pragma unroll
        for (int c = tid; c < CHUNKS; c += NTHREADS) {
            int mn = (c >> 1) % BMN, kl = 4 * (c / (2 * BMN)) + 2 * (c & 1);
            int gk = kbase + kl, gmn = mn0 + mn;
            bool rok = gmn < MN;
            bool v0 = rok && gk < k_hi, v1 = rok && (gk + 1) < k_hi;
            double* dst = s + ((kl >> 2) * BMN + mn) * 4 + (kl & 3);
            const double* src = G + (long)gmn * ld + gk;
            if (vec16) {
                cp_async16(dst, v0 ? src : dummy, v1 ? 16 : (v0 ? 8 : 0));
            } else {
                cp_async8(dst, v0 ? src : dummy, v0 ? 8 : 0);
                cp_async8(dst + 1, v1 ? src + 1 : dummy, v1 ? 8 : 0);
            }
        }
    } else {
        constexpr int CHUNKS = BK * (BMN / 2);
#pragma unroll
        for (int c = tid; c < CHUNKS; c += NTHREADS) {
            int k = c / (BMN / 2), ml = 2 * (c % (BMN / 2));
            int gk = kbase + k, gmn = mn0 + ml;
            bool kok = gk < k_hi;
            bool v0 = kok && gmn < MN, v1 = kok && (gmn + 1) < MN;
            double* dst = s + k * (BMN + 4) + ml;
            const double* src = G + (long)gk * ld + gmn;
            if (vec16) {
                cp_async16(dst, v0 ? src : dummy, v1 ? 16 : (v0 ? 8 : 0));
            } else {
                cp_async8(dst, v0 ? src : dummy, v0 ? 8 : 0);
                cp_async8(dst + 1, v1 ? src + 1 : dummy, v1 ? 8 : 0);
            }
        }
    }
}

}  // namespace tile
}  // namespace sb
