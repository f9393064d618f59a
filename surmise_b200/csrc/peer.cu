// All-gather of the PC-sharded prediction blocks over NVLink peer memory, without a collective library call.
//
// The device-resident PT-LMC step of a PC-sharded calibration needs, every step, each rank's (B, kmax, 2 + 2d) block
// of predictions on every rank before the (replicated) Woodbury kernel runs.  The blocks live in SYMMETRIC memory
// (one allocation per rank, mapped into every peer: torch.distributed._symmetric_memory supplies the mapping).  After
// its predict kernels a rank
//   post : stores its block straight into its slot of every peer's buffer (16-byte st.global over NVLink), fences at
//          system scope and raises its flag on the peer to the new epoch (the last CTA per peer does that);
//   wait : one CTA spins (acquire loads at system scope on its OWN flags) until every peer's flag has reached the
//          epoch, then advances the epoch.
// Both kernels take everything that changes from step to step from device memory (the epoch), so a captured CUDA graph
// replays them unchanged.  Buffers are double-buffered by the caller (parity of the step): a slot is overwritten two
// steps later, after its reader has posted the step in between -- which in stream order is after its Woodbury kernel
// read it.  A flag that does not arrive within ~2 s sets the error word instead of hanging.
#include "gp.cuh"
#include <algorithm>

namespace sb {
namespace {

constexpr long long PEER_SPIN_LIMIT = 4000000000LL;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

// grid (ctas_per_peer, world): blockIdx.y = destination rank (the own rank returns at once)
__global__ void __launch_bounds__(256) peer_post_kernel(const double* __restrict__ block, long count,
                                                        const unsigned long long* __restrict__ peer_bufs, long slot_off,
                                                        long flag_off, int rank, int world,
                                                        const unsigned long long* __restrict__ epoch, int* __restrict__ arrive) {
    const int dst = blockIdx.y;
    if (dst == rank) return;
    double* out = reinterpret_cast<double*>(peer_bufs[dst]) + slot_off;
    const long n2 = count >> 1;                                        // 16-byte pieces (count is even: ld padded to 16)
    const double2* src2 = reinterpret_cast<const double2*>(block);
    double2* out2 = reinterpret_cast<double2*>(out);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (long)gridDim.x * blockDim.x) out2[i] = src2[i];
    if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[count - 1] = block[count - 1];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(&arrive[dst], 1);
        if (old == (int)gridDim.x - 1) {                               // every CTA of this destination has fenced its stores
            arrive[dst] = 0;
            __threadfence_system();
            unsigned long long* flag = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(peer_bufs[dst]) + flag_off) + rank;
            st_release_sys(flag, epoch[0] + 1);
        }
    }
}

__global__ void __launch_bounds__(64) peer_wait_kernel(const unsigned long long* __restrict__ my_flags, int rank, int world,
                                                       unsigned long long* __restrict__ epoch, int* __restrict__ error) {
    const unsigned long long target = epoch[0] + 1;
    const long long t0 = clock64();
    const bool failed = *reinterpret_cast<volatile int*>(error) != 0;      // one missed epoch: do not wait for the rest
    for (int src = threadIdx.x; src < world && !failed; src += blockDim.x) {
        if (src == rank) continue;
        while (ld_acquire_sys(my_flags + src) < target) {
            if (clock64() - t0 > PEER_SPIN_LIMIT) {
                atomicExch(error, 1);
                break;
            }
            __nanosleep(100);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) epoch[0] = target;
}

}  // namespace

int peer_allgather_post(const double* block, long count, const unsigned long long* peer_bufs, long slot_off, long flag_off,
                        int rank, int world, const unsigned long long* epoch, int* arrive, cudaStream_t stream) {
    SB_CHECK_ARG(world >= 2 && rank >= 0 && rank < world && count > 0, "peer_allgather_post: rank %d of %d, %ld doubles", rank,
                 world, count);
    SB_CHECK_ARG((((uintptr_t)block) & 15) == 0 && (slot_off & 1) == 0, "peer_allgather_post: blocks must be 16-byte aligned");
    const int ctas = (int)std::min<long>(8, std::max<long>(1, count / 4096));
    dim3 grid(ctas, world);
    peer_post_kernel<<<grid, 256, 0, stream>>>(block, count, peer_bufs, slot_off, flag_off, rank, world, epoch, arrive);
    SB_LAUNCH_CHECK();
    return 0;
}

int peer_allgather_wait(const unsigned long long* my_flags, int rank, int world, unsigned long long* epoch, int* error,
                        cudaStream_t stream) {
    SB_CHECK_ARG(world >= 2 && rank >= 0 && rank < world, "peer_allgather_wait: rank %d of %d", rank, world);
    peer_wait_kernel<<<1, 64, 0, stream>>>(my_flags, rank, world, epoch, error);
    SB_LAUNCH_CHECK();
    return 0;
}

}  // namespace sb
