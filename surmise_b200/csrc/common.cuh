// Shared helpers for the surmise_b200 CUDA library (sm_100a only, fp64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstddef>
#include <atomic>
#include <mutex>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "surmise_b200 kernels are written for sm_100a (B200) only"
#endif

namespace sb {

// ---- error plumbing: every C-ABI entry point returns 0, <0 (bad argument) or >0 (cudaError_t) ----
void set_last_error(const char* fmt, ...);
void count_launch();

// optional per-launch timing of the DMMA GEMM (bench.py's roofline leg): when enabled, launch_dgemm brackets
// each launch with CUDA events on its stream and books the launch's algorithmic flop count
bool gemm_profile_enabled();
int profile_event_pair(cudaEvent_t* a, cudaEvent_t* b);      // timing events from a pool (no creation cost once warm)
void gemm_profile_add(cudaEvent_t start, cudaEvent_t stop, double flops);
void kernel_profile_add(int kind, cudaEvent_t start, cudaEvent_t stop, double flops);   // 0 = DMMA GEMM, 1 = task-graph Cholesky

#define SB_CHECK_ARG(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            sb::set_last_error(__VA_ARGS__);            \
            return -1;                                  \
        }                                               \
    } while (0)

#define SB_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            sb::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                               __FILE__, __LINE__);                                     \
            return (int)e__;                                                            \
        }                                                                               \
    } while (0)

// every kernel launch in the library is followed by this: counts launches (sb200_launch_count)
#define SB_LAUNCH_CHECK()              \
    do {                               \
        sb::count_launch();            \
        SB_CUDA(cudaGetLastError());   \
    } while (0)

#define SB_TRY(expr)            \
    do {                        \
        int rc__ = (expr);      \
        if (rc__ != 0) return rc__; \
    } while (0)


// One-time, per-device kernel attribute set-up (cudaFuncSetAttribute for > 48 KB of dynamic shared memory) that is
// safe when several host threads make their first call on a fresh device at the same moment: double-checked under
// a mutex; devices beyond the table are simply configured on every call (the attribute call is idempotent).
struct DeviceOnce {
    std::atomic<int> done[64];
    std::mutex mu;
    DeviceOnce() {
        for (auto& d : done) d.store(0, std::memory_order_relaxed);
    }
    template <class F>
    int run(F&& configure) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) {
            set_last_error("cudaGetDevice failed: %s", cudaGetErrorString(e));
            return (int)e;
        }
        if (dev < 0 || dev >= 64) return configure();
        if (done[dev].load(std::memory_order_acquire)) return 0;
        std::lock_guard<std::mutex> lk(mu);
        if (done[dev].load(std::memory_order_relaxed)) return 0;
        int rc = configure();
        if (rc == 0) done[dev].store(1, std::memory_order_release);
        return rc;
    }
};

static inline long round_up(long v, long a) { return (v + a - 1) / a * a; }
static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// bump allocator over a caller-provided workspace (the library never allocates device memory)
struct Workspace {
    char* base;
    size_t cap;
    size_t off;
    Workspace(void* p, size_t bytes) : base((char*)p), cap(bytes), off(0) {}
    template <typename T>
    T* take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
        if (base == nullptr) {           // sizing pass
            off += bytes;
            return nullptr;
        }
        if (off + bytes > cap) return nullptr;
        T* r = (T*)(base + off);
        off += bytes;
        return r;
    }
};

constexpr int MAXD = 16;          // max parameter dimension d supported by the register-resident loops
constexpr int MAXP = MAXD + 3;    // hyper-parameters per GP: d length-scales + scale + log var-const + logit nugget

// ---- device helpers ----
#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 1 / x for x known to be a normal number >= 1 (here always 1 + |s|): MUFU.RCP64H seed and two Newton steps, branch
// free.  __drcp_rn adds a range check and a slow path (BSSY / BRA / BSYNC around every call) that these loops, which
// take one reciprocal per pair and dimension, do not need; the result is within 1 ulp of the rounded quotient.
__device__ __forceinline__ double rcp_ge1(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// block-wide sum; `red` must hold >= 32 doubles; result valid in all threads
__device__ __forceinline__ double block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
    if (w == 0) {
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    return red[0];
}
#endif

}  // namespace sb
