// Dependent-issue latencies on B200 (one warp, one CTA): DFMA, DMUL, MUFU.RSQ64H + refinement, LDS, STS+BAR+LDS, SHFL.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double rsqrt_normal(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x));
    double e = fma(x, -(y * y), 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}
__global__ void lat(double* out, long long* cyc, double x0, int threads_active) {
    __shared__ double sm[256];
    double x = x0 + threadIdx.x * 1e-9;
    long long t0, t1;
    const int N = 512;
    // 1. dependent DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = fma(x, 0.999999, 1e-9);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // 2. dependent DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = x * 1.0000001;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // 3. dependent rsqrt_normal (MUFU + 4 deep)
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N / 4; i++) x = rsqrt_normal(x) + 1.0;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // 4. STS + BAR + LDS round trip
    t0 = clock64();
    for (int i = 0; i < N / 4; i++) {
        sm[threadIdx.x] = x;
        __syncthreads();
        x += sm[(threadIdx.x + 1) & (blockDim.x - 1)];
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // 5. dependent double shuffle
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // 6. dependent DADD+DSETP select
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) x = (x > 0.5) ? x + 1e-9 : x - 1e-9;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // 7. dependent MUFU.RSQ64H alone (seed only)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(x)); x = y; }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // 8. dependent FFMA for reference
    float f = (float)x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, 0.999f, 1e-3f);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    out[threadIdx.x] = x + f;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
    const char* names[8] = {"DFMA", "DMUL", "rsqrt_normal+DADD (MUFU+4 deep+1)", "STS+BAR+LDS(+DADD)", "SHFL.64", "DSETP+select+DADD", "MUFU.RSQ64H", "FFMA"};
    int div[8] = {512, 512, 128, 128, 512, 512, 512, 512};
    for (int threads : {32, 256}) {
        for (int rep = 0; rep < 2; rep++) lat<<<1, threads>>>(out, cyc, 1.2345, threads);
        cudaDeviceSynchronize();
        long long h[8]; cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
        printf("threads=%d\n", threads);
        for (int i = 0; i < 8; i++) printf("  %-36s %7.1f cycles per dependent step\n", names[i], (double)h[i] / div[i]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
