// ubench.cu -- micro-benchmarks behind the round-2 kernel decisions (sm_100a):
//   1. DMMA.8x8x4 issue rate as a function of warps per SM sub-partition and of the number of independent accumulators
//      (= distance between dependent DMMAs in one warp's stream) -> the latency a single warp's chain must cover
//   2. DFMA with a uniform-register operand fed by LDCU from constant memory vs DFMA on vector registers
//   3. LDCU stream (16 x LDCU.64 per "matrix", dependent DFMAs) at different constant working sets
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_chain(double* out, int iters, double a, double b)
{
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < iters; it++)
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// One child contraction of the codon pruning kernel as a stream: per k-step 4 x 128-bit B-fragment loads feed 8 DMMAs.
// MODE 0: B fragments by ld.global.nc (L1-resident 32 KB tile), 1: by ld.shared, 2: B kept in registers (no loads)
template <int MODE>
__global__ void dmma_stream(const double2* __restrict__ gB, double* out, int iters)
{
    extern __shared__ double2 sB[];
    const int lane = threadIdx.x & 31;
    if (MODE == 1) {
        for (int e = threadIdx.x; e < 16 * 4 * 32; e += blockDim.x) sB[e] = gB[e];
        __syncthreads();
    }
    double a[16], c[8][2];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = 1.0 + 1e-9 * (k + lane);
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = lane; }
    double2 breg[4];
#pragma unroll
    for (int pr = 0; pr < 4; pr++) breg[pr] = gB[pr * 32 + lane];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int ks = 0; ks < 16; ks++) {
            double b[8];
#pragma unroll
            for (int pr = 0; pr < 4; pr++) {
                double2 v;
                if (MODE == 0) v = __ldg(gB + (ks * 4 + pr) * 32 + lane);
                else if (MODE == 1) v = sB[(ks * 4 + pr) * 32 + lane];
                else v = breg[pr];
                b[2 * pr] = v.x; b[2 * pr + 1] = v.y;
            }
#pragma unroll
            for (int nt = 0; nt < 8; nt++) dmma884(c[nt][0], c[nt][1], a[ks], b[nt]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// P-stationary formulation: the P slab of a warp (one m-tile of 8 output states x 64 input states = 16 A fragments) stays in
// registers, the child vectors of a chunk of 64 patterns are streamed from shared memory as B fragments (one LDS.64 per DMMA,
// row stride 68 doubles), 8 accumulators (n-tiles) per warp.  MT2: two m-tiles per warp (each B fragment feeds two DMMAs).
template <int MT>
__global__ void dmma_pstat(double* out, int iters)
{
    extern __shared__ double sV[];      // [64 patterns][68]
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3;
    for (int e = threadIdx.x; e < 64 * 68; e += blockDim.x) sV[e] = 1e-3 * (e % 7);
    __syncthreads();
    double a[MT][16], c[MT][8][2];
#pragma unroll
    for (int m = 0; m < MT; m++) {
#pragma unroll
        for (int k = 0; k < 16; k++) a[m][k] = 1.0 + 1e-9 * (k + lane + m);
#pragma unroll
        for (int i = 0; i < 8; i++) { c[m][i][0] = i; c[m][i][1] = lane; }
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int ks = 0; ks < 16; ks++) {
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                const double b = sV[(8 * nt + g) * 68 + 4 * ks + q];
#pragma unroll
                for (int m = 0; m < MT; m++) dmma884(c[m][nt][0], c[m][nt][1], a[m][ks], b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int m = 0; m < MT; m++)
#pragma unroll
        for (int i = 0; i < 8; i++) s += c[m][i][0] + c[m][i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__constant__ double cM[7680];
__constant__ int cIdx[1024];

// per "op": one 4x4 matrix from constant memory (uniform index), applied to PPT vectors per thread
template <int PPT>
__global__ void dfma_const(double* out, int nops, int iters)
{
    double v[PPT][4];
#pragma unroll
    for (int u = 0; u < PPT; u++) { v[u][0] = 1.0 + threadIdx.x * 1e-9; v[u][1] = 0.5; v[u][2] = 0.25 + u; v[u][3] = 0.125; }
    for (int it = 0; it < iters; it++)
#pragma unroll 1
        for (int op = 0; op < nops; op++) {
            const double* P = cM + cIdx[op] * 16;
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                double t[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double s = P[4 * j] * v[u][0];
                    s = fma(P[4 * j + 1], v[u][1], s); s = fma(P[4 * j + 2], v[u][2], s); t[j] = fma(P[4 * j + 3], v[u][3], s);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) v[u][j] = t[j];
            }
        }
    double s = 0;
#pragma unroll
    for (int u = 0; u < PPT; u++) s += v[u][0] + v[u][1] + v[u][2] + v[u][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same with the matrices in shared memory (broadcast LDS.128), as prune4v2 does without PCONST
template <int PPT>
__global__ void dfma_smem(double* out, const double* gM, int nmat, int nops, int iters)
{
    extern __shared__ double sM[];
    for (int e = threadIdx.x; e < nmat * 16; e += blockDim.x) sM[e] = gM[e];
    __syncthreads();
    double v[PPT][4];
#pragma unroll
    for (int u = 0; u < PPT; u++) { v[u][0] = 1.0 + threadIdx.x * 1e-9; v[u][1] = 0.5; v[u][2] = 0.25 + u; v[u][3] = 0.125; }
    for (int it = 0; it < iters; it++)
#pragma unroll 1
        for (int op = 0; op < nops; op++) {
            const double2* P2 = reinterpret_cast<const double2*>(sM + cIdx[op] * 16);
            double P[16];
#pragma unroll
            for (int e = 0; e < 8; e++) { double2 x = P2[e]; P[2 * e] = x.x; P[2 * e + 1] = x.y; }
#pragma unroll
            for (int u = 0; u < PPT; u++) {
                double t[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    double s = P[4 * j] * v[u][0];
                    s = fma(P[4 * j + 1], v[u][1], s); s = fma(P[4 * j + 2], v[u][2], s); t[j] = fma(P[4 * j + 3], v[u][3], s);
                }
#pragma unroll
                for (int j = 0; j < 4; j++) v[u][j] = t[j];
            }
        }
    double s = 0;
#pragma unroll
    for (int u = 0; u < PPT; u++) s += v[u][0] + v[u][1] + v[u][2] + v[u][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float timeit(void (*launch)(void*), void* arg)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(arg); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); launch(arg); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

struct DA { int nacc, threads, iters; double* out; int nsm; };
static void run_dmma(void* p)
{
    DA* a = (DA*)p;
    switch (a->nacc) {
    case 1: dmma_chain<1><<<a->nsm, a->threads>>>(a->out, a->iters * 8, 1.0000001, 1e-9); break;
    case 2: dmma_chain<2><<<a->nsm, a->threads>>>(a->out, a->iters * 4, 1.0000001, 1e-9); break;
    case 4: dmma_chain<4><<<a->nsm, a->threads>>>(a->out, a->iters * 2, 1.0000001, 1e-9); break;
    default: dmma_chain<8><<<a->nsm, a->threads>>>(a->out, a->iters, 1.0000001, 1e-9); break;
    }
}
struct FA { int ppt, threads, blocks, nops, iters, smem, nmat; double* out; const double* gM; };
static void run_const(void* p)
{
    FA* a = (FA*)p;
    if (a->ppt == 1) dfma_const<1><<<a->blocks, a->threads>>>(a->out, a->nops, a->iters);
    else if (a->ppt == 2) dfma_const<2><<<a->blocks, a->threads>>>(a->out, a->nops, a->iters);
    else if (a->ppt == 4) dfma_const<4><<<a->blocks, a->threads>>>(a->out, a->nops, a->iters);
    else dfma_const<8><<<a->blocks, a->threads>>>(a->out, a->nops, a->iters);
}
static void run_smem(void* p)
{
    FA* a = (FA*)p;
    size_t sm = (size_t)a->nmat * 128;
    if (a->ppt == 1) dfma_smem<1><<<a->blocks, a->threads, sm>>>(a->out, a->gM, a->nmat, a->nops, a->iters);
    else if (a->ppt == 2) dfma_smem<2><<<a->blocks, a->threads, sm>>>(a->out, a->gM, a->nmat, a->nops, a->iters);
    else if (a->ppt == 4) dfma_smem<4><<<a->blocks, a->threads, sm>>>(a->out, a->gM, a->nmat, a->nops, a->iters);
    else dfma_smem<8><<<a->blocks, a->threads, sm>>>(a->out, a->gM, a->nmat, a->nops, a->iters);
}

int main()
{
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    double* out; cudaMalloc(&out, (size_t)nsm * 8 * 1024 * 8);
    printf("device %s, %d SMs\n", prop.name, nsm);
    printf("== DMMA.8x8x4: TFLOP/s by warps per SM and independent accumulators per warp (peak = 1 DMMA / 16 clk / sub-partition)\n");
    for (int threads : {128, 256, 512, 1024})
        for (int nacc : {1, 2, 4, 8}) {
            DA a{nacc, threads, 20000, out, nsm};
            float ms = timeit(run_dmma, &a);
            double flop = (double)nsm * (threads / 32) * 20000.0 * 8 * 512.0;
            printf("  %2d warps/SM (%d per sub-partition), %d accumulators: %7.2f TFLOP/s\n", threads / 32, threads / 128, nacc, flop / ms / 1e9);
        }
    {
        double2* gB; cudaMalloc(&gB, 16 * 4 * 32 * sizeof(double2));
        cudaMemset(gB, 0, 16 * 4 * 32 * sizeof(double2));
        cudaFuncSetAttribute(dmma_stream<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
        printf("== one child contraction as a stream (128 DMMA + 64 B-fragment loads of 128 bits per 8 patterns), TFLOP/s\n");
        for (int bps : {1, 2, 4})
            for (int threads : {128, 256, 512}) {
                if (bps * threads > 1024) continue;
                const int iters = 400;
                double flop = (double)nsm * bps * (threads / 32) * iters * 128.0 * 512.0;
                float ms[3];
                for (int mode = 0; mode < 3; mode++) {
                    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                    float best = 1e30f;
                    for (int r = 0; r < 4; r++) {
                        cudaEventRecord(e0);
                        if (mode == 0) dmma_stream<0><<<nsm * bps, threads>>>(gB, out, iters);
                        else if (mode == 1) dmma_stream<1><<<nsm * bps, threads, 32768>>>(gB, out, iters);
                        else dmma_stream<2><<<nsm * bps, threads>>>(gB, out, iters);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float t; cudaEventElapsedTime(&t, e0, e1);
                        if (r && t < best) best = t;
                    }
                    ms[mode] = best;
                }
                printf("  %d CTAs/SM x %2d warps: B by LDG %6.2f | by LDS %6.2f | in registers %6.2f\n", bps, threads / 32, flop / ms[0] / 1e9,
                       flop / ms[1] / 1e9, flop / ms[2] / 1e9);
            }
    }
    {
        cudaFuncSetAttribute(dmma_pstat<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 68 * 8);
        cudaFuncSetAttribute(dmma_pstat<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 68 * 8);
        printf("== P-stationary stream (A = P slab in registers, B = child vectors by LDS.64 from a 64-pattern chunk), TFLOP/s\n");
        for (int mt : {1, 2})
            for (int bps : {1, 2})
                for (int threads : {128, 256}) {
                    const int iters = 400;
                    double flop = (double)nsm * bps * (threads / 32) * iters * 128.0 * mt * 512.0;
                    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                    float best = 1e30f;
                    for (int r = 0; r < 4; r++) {
                        cudaEventRecord(e0);
                        if (mt == 1) dmma_pstat<1><<<nsm * bps, threads, 64 * 68 * 8>>>(out, iters);
                        else dmma_pstat<2><<<nsm * bps, threads, 64 * 68 * 8>>>(out, iters);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float t; cudaEventElapsedTime(&t, e0, e1);
                        if (r && t < best) best = t;
                    }
                    printf("  %d m-tile(s) per warp, %d CTAs/SM x %d warps: %6.2f\n", mt, bps, threads / 32, flop / best / 1e9);
                }
    }
    return 0;
    // constant matrices
    double hM[7680]; int hI[1024];
    for (int i = 0; i < 7680; i++) hM[i] = (i % 16) % 5 == 0 ? 0.97 : 0.01;
    cudaMemcpyToSymbol(cM, hM, sizeof(hM));
    double* gM; cudaMalloc(&gM, sizeof(hM)); cudaMemcpy(gM, hM, sizeof(hM), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(dfma_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(dfma_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(dfma_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(dfma_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    printf("== 4x4 matrix-vector ops (16 DP instr per pattern): matrices by LDCU (uniform operands) vs broadcast LDS; TFLOP/s counts 32 flop per matvec\n");
    for (int nmat : {1, 16, 125, 480}) {
        for (int i = 0; i < 1024; i++) hI[i] = i % nmat;
        cudaMemcpyToSymbol(cIdx, hI, sizeof(hI));
        for (int ppt : {1, 2, 4, 8})
            for (int threads : {128, 256}) {
                for (int bps : {1, 2, 4}) {
                    FA a{ppt, threads, nsm * bps, 1000, 20, 0, nmat, out, gM};
                    float m1 = timeit(run_const, &a), m2 = timeit(run_smem, &a);
                    double flop = (double)a.blocks * threads * ppt * 1000.0 * 20 * 32.0;
                    printf("  %3d matrices (%5.1f KB), ppt %d, %3d thr x %d blocks/SM: LDCU %6.2f TF | LDS %6.2f TF\n", nmat, nmat * 0.125, ppt, threads, bps,
                           flop / m1 / 1e9, flop / m2 / 1e9);
                }
            }
    }
    return 0;
}
