// FP64 pipe micro-benchmark: DFMA (CUDA-core FP64) vs DMMA.8x8x4 issue rate per SM on B200.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters, long long* cyc) {
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void dmma_kernel(double* out, int iters, long long* cyc) {
    double c[8][2];
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = threadIdx.x + j;
    double a = 1.0000001, b = 0.999999;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// mixed: even warps issue DFMA, odd warps DMMA -- do the two share one pipe?
__global__ void mixed_kernel(double* out, int iters, long long* cyc, int mode) {
    // mode 0: even warps DFMA, odd DMMA; 1: even idle, odd DMMA; 2: all DMMA; 3: even DFMA, odd idle
    const bool odd = ((threadIdx.x >> 7) & 1) != 0;      // warps 4-7, 12-15, ...: both kinds on every SM sub-partition
    const bool tensor = mode == 2 || (odd && mode != 3);
    const bool idle = (mode == 1 && !odd) || (mode == 3 && odd);
    double c[8][2];
    for (int j = 0; j < 8; j++) c[j][0] = c[j][1] = threadIdx.x + j;
    double a = 1.0000001, b = 0.999999, cc = 1e-9;
    __syncthreads();
    long long t0 = clock64();
    if (tensor) {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
        }
    } else if (!idle) {
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int j = 0; j < 8; j++) c[j][0] = fma(c[j][0], a, cc);
        }
    }
    __syncthreads();
    long long t1 = clock64();
    double s = 0; for (int j = 0; j < 8; j++) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int which = 0; which < 2; which++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (which == 0) dfma_kernel<<<148, threads>>>(out, iters, cyc); else dmma_kernel<<<148, threads>>>(out, iters, cyc);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double fma_per_thread = (double)iters * 8 * (which ? 8.0 : 1.0);   // DMMA: 256 FMA / 32 lanes = 8 per lane
            double total = fma_per_thread * threads * 148;
            printf("%s threads/SM=%4d  cycles=%lld  FMA/clk/SM=%.1f  TFLOP/s=%.2f\n", which ? "DMMA" : "DFMA", threads, h,
                   fma_per_thread * threads / (double)h, 2 * total / (ms * 1e-3) / 1e12);
        }
    }
    for (int mode = 0; mode < 4; mode++)
        for (int threads : {256, 1024}) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            mixed_kernel<<<148, threads>>>(out, iters, cyc, mode);
            cudaEventRecord(e0);
            mixed_kernel<<<148, threads>>>(out, iters, cyc, mode);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long h = (long long)(ms * 1e-3 * 1.965e9);      // whole-kernel time in SM clocks at 1965 MHz
            int wt = mode == 2 ? threads : (mode == 3 ? 0 : threads / 2), wf = (mode == 0 || mode == 3) ? threads / 2 : 0;
            double dfma = (double)iters * 8 * wf, dmma = (double)iters * 8 * 8 * wt;
            printf("MIXED mode %d threads/SM=%4d  cycles=%lld  DFMA %.1f + DMMA %.1f FMA/clk/SM\n", mode, threads, h, dfma / h, dmma / h);
        }
    return 0;
}
