// FP64 peak micro-benchmark for sm_100a: vector DFMA vs tensor DMMA shapes.
// Used to establish the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 figure).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1684(double* out, int iters, double a, double b) {
    double c[NACC][4]; double av[2] = {a, a + 1}, bv[1] = {b};
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma1684(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double a, double b) {
    double c[NACC][4]; double av[4] = {a, a + 1, a + 2, a + 3}, bv[2] = {b, b + 1};
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma1688(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a, double b) {
    double c[NACC][4]; double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a + i;
#pragma unroll
    for (int i = 0; i < 4; i++) bv[i] = b + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma16816(c[i], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int nsm = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    const int iters = 20000;
    for (int wps = 4; wps <= 16; wps *= 2) {   // warps per SM (one CTA per SM)
        int threads = wps * 32; int blocks = nsm;
        double nthreads = (double)threads * blocks;
        float ms = time_it([&] { k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DFMA            %8.2f TFLOP/s\n", wps, nthreads * iters * 16 * 2 / ms / 1e9);
        double nwarps = nthreads / 32;
        ms = time_it([&] { k_dmma884<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m8n8k4 x8   %8.2f TFLOP/s\n", wps, nwarps * iters * 8 * (8.0 * 8 * 4 * 2) / ms / 1e9);
        ms = time_it([&] { k_dmma884<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m8n8k4 x16  %8.2f TFLOP/s\n", wps, nwarps * iters * 16 * (8.0 * 8 * 4 * 2) / ms / 1e9);
        ms = time_it([&] { k_dmma1684<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m16n8k4 x8  %8.2f TFLOP/s\n", wps, nwarps * iters * 8 * (16.0 * 8 * 4 * 2) / ms / 1e9);
        ms = time_it([&] { k_dmma1688<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m16n8k8 x8  %8.2f TFLOP/s\n", wps, nwarps * iters * 8 * (16.0 * 8 * 8 * 2) / ms / 1e9);
        ms = time_it([&] { k_dmma16816<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("warps/SM %2d  DMMA m16n8k16 x8 %8.2f TFLOP/s\n", wps, nwarps * iters * 8 * (16.0 * 8 * 16 * 2) / ms / 1e9);
    }
    // dependent-chain latency of m8n8k4 (1 warp, 1 accumulator)
    {
        float ms = time_it([&] { k_dmma884<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA m8n8k4 dependent issue: %.1f ns per mma\n", ms * 1e6 / iters);
        ms = time_it([&] { k_dmma16816<1><<<1, 32>>>(out, iters, 1.0000001, 1e-9); });
        printf("DMMA m16n8k16 dependent issue: %.1f ns per mma\n", ms * 1e6 / iters);
    }
    return 0;
}
