// prune4.cuh -- the 4-state (nucleotide) pruning kernel: a register-resident tree walk on the FP64 vector pipe.
//
// With n = 4 a contraction is 16 FMAs per pattern and branch: a DMMA tile (8 x 8 x 4) would be three quarters
// padding, and the work is far too thin to amortise fragment traffic.  Here ONE THREAD owns PPT site patterns of
// one rate class and walks the whole tree for them (same planner program as the DMMA kernel, src/baseml.c:1517-1570):
// the node being built lives in 4 registers per pattern, a node that must wait for a sibling subtree is parked in
// shared memory ([slot][state][thread], conflict-free), and the 4x4 P(t) of every branch of the class sits in shared
// memory, read as warp-wide broadcasts.  Conditional vectors never touch HBM: per class the kernel reads the tip
// codes (1 byte per tip and pattern) and writes one double per pattern.
#pragma once
#include "kernels.cuh"

namespace pb200 {

struct Prune4Params {
    const int* prog;
    int prog_in_const;
    int prog_base;              // offset of the program in c_prog
    const unsigned char* z;     // [ns][zstride] compact codes
    long long zstride, npatt;
    const double* P;            // [nclass][nnode][pstride]: internal branch P[j][k] row-major; tip branch PT[code][j]
    int pstride;
    int nnode, nclass, nslots, semantics, nscale;
    int p_in_smem;              // 1: the class's matrices are staged in shared memory
    int ns, ncodes;
    const double* pi;           // [>=4]
    const unsigned char* amb_mask;   // [ncodes]: bit k set = state k compatible with the code (codes 0..3: 1<<code)
    double* fhK;                // [nclass][npatt]
};

template <int PPT, int THREADS, bool PSM, bool CP>
__global__ void __launch_bounds__(THREADS) prune4_kernel(const Prune4Params p)
{
    extern __shared__ __align__(16) double sm4[];
    const int tid = threadIdx.x;
    const int cls = blockIdx.y;
    const long long N = p.npatt;
    double* Psm = sm4;                                                   // [nnode][16] (if p_in_smem)
    double* slots = sm4 + (PSM ? (size_t)p.nnode * p.pstride : 0);      // [nslots][PPT][4][THREADS]
    const double* Pg = p.P + (size_t)cls * p.nnode * p.pstride;
    if (PSM) {
        for (int e = tid; e < p.nnode * p.pstride; e += THREADS) Psm[e] = Pg[e];
        __syncthreads();
    }
    // program words: constant memory (uniform loads) when the program fits there, else read-only global loads;
    // an explicit branch per access keeps the two address spaces apart (a selected pointer would become generic)
    auto PW = [&](int i) -> int { if (CP) return c_prog[i]; else return __ldg(p.prog + i); };      // absolute positions: see prune_kernel
    const int cb = CP ? p.prog_base : 0, hdr0 = cb + PROG_HDR;
    const int nops = PW(cb), ntip = PW(cb + 2);
    const int ops0 = cb + PW(cb + 3);
    // per-warp ring of tip codes (cp.async, 4 bytes per lane): [slot][PPT][32], two halves of HALF4 tips; one half is
    // consumed (no synchronisation per tip) while the other is being refilled, one cp.async group per half
    constexpr int HALF4 = 8, RING4 = 2 * HALF4;
    const int warp = tid >> 5, lane = tid & 31;
    unsigned char* ring = reinterpret_cast<unsigned char*>(slots + (size_t)p.nslots * PPT * 4 * THREADS) + (size_t)warp * RING4 * PPT * 32;
    const double pi0 = p.pi[0], pi1 = p.pi[1], pi2 = p.pi[2], pi3 = p.pi[3];

    for (long long base = (long long)blockIdx.x * THREADS * PPT; base < N; base += (long long)gridDim.x * THREADS * PPT) {
        long long h[PPT];
        bool live[PPT];
#pragma unroll
        for (int u = 0; u < PPT; u++) { h[u] = base + (long long)u * THREADS + tid; live[u] = h[u] < N; if (!live[u]) h[u] = N - 1; }
        double acc[PPT][4], lnsc[PPT];
#pragma unroll
        for (int u = 0; u < PPT; u++) lnsc[u] = 0.0;
        // fill one half of the ring with the codes of tips [tip0, tip0 + HALF4) of the walk: ONE cp.async group
        auto half_fill = [&](int half, int tip0) {
            if (lane < PPT * 8) {
                const int u = lane >> 3, l4 = (lane & 7) * 4;
                const long long hh = base + (long long)u * THREADS + warp * 32 + l4;
#pragma unroll
                for (int i = 0; i < HALF4; i++) {
                    const int tipi = tip0 + i;
                    if (tipi < ntip) {
                        unsigned char* dst = ring + ((half * HALF4 + i) * PPT + u) * 32 + l4;
                        if (hh + 4 <= p.zstride) cp_async4(dst, p.z + (size_t)PW(hdr0 + tipi) * p.zstride + hh);
                        else *reinterpret_cast<unsigned int*>(dst) = 0u;      // past the padded row: dead patterns read state 0
                    }
                }
            }
            cp_async_commit();
        };
        half_fill(0, 0);
        half_fill(1, HALF4);
        int ti = 0;
        int pc = ops0;
        for (int op = 0; op < nops; op++) {
            const int nchild = PW(pc), out_slot = PW(pc + 1), sidx = PW(pc + 2);
            pc += OP_HDR;
            bool started = false;
            for (int c = 0; c < nchild; c++) {
                const int kind = PW(pc), idx = PW(pc + 1), node = PW(pc + 2);
                pc += CH_REC;
                if (kind == CH_PARTIAL) {
#pragma unroll
                    for (int u = 0; u < PPT; u++)
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            double v = slots[(((size_t)idx * PPT + u) * 4 + j) * THREADS + tid];
                            acc[u][j] = started ? acc[u][j] * v : v;
                        }
                    started = true;
                    continue;
                }
                const double* Pm = (PSM ? Psm : Pg) + (size_t)node * p.pstride;
                if (kind == CH_TIP) {
                    // tip (src/baseml.c:1547-1558): row `code` of PT[code][j] = sum of the columns of P the code allows
                    if (ti % HALF4 == 0) {                    // entering a half: its group (all but the newest) has landed
                        cp_async_wait<1>();
                        __syncwarp();
                    }
                    int codes[PPT];
#pragma unroll
                    for (int u = 0; u < PPT; u++) codes[u] = min((int)ring[((ti % RING4) * PPT + u) * 32 + lane], p.ncodes - 1);
                    ti++;
                    if (ti % HALF4 == 0) {                    // half consumed: refill it with the tips two halves ahead
                        __syncwarp();
                        half_fill(((ti / HALF4) - 1) & 1, ti + HALF4);
                    }
#pragma unroll
                    for (int u = 0; u < PPT; u++) {
                        const int code = codes[u];
                        double cj[4];
                        {
                            const double2 a = *reinterpret_cast<const double2*>(Pm + code * 4);
                            const double2 b = *reinterpret_cast<const double2*>(Pm + code * 4 + 2);
                            cj[0] = a.x; cj[1] = a.y; cj[2] = b.x; cj[3] = b.y;
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) acc[u][j] = started ? acc[u][j] * cj[j] : cj[j];
                    }
                    started = true;
                }
                else {
                    // internal child (src/baseml.c:1559-1566): 4x4 matrix-vector product; the child's vector comes from
                    // this thread's registers (CH_DIRECT: finished by the previous op) or from its parked slot
                    double Pr[16];
#pragma unroll
                    for (int e = 0; e < 16; e += 2) { double2 t = *reinterpret_cast<const double2*>(Pm + e); Pr[e] = t.x; Pr[e + 1] = t.y; }
#pragma unroll
                    for (int u = 0; u < PPT; u++) {
                        double v[4];
                        if (kind == CH_DIRECT) {
#pragma unroll
                            for (int k = 0; k < 4; k++) v[k] = acc[u][k];
                        }
                        else {
#pragma unroll
                            for (int k = 0; k < 4; k++) v[k] = slots[(((size_t)idx * PPT + u) * 4 + k) * THREADS + tid];
                        }
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            double t = Pr[j * 4] * v[0];
                            t = fma(Pr[j * 4 + 1], v[1], t);
                            t = fma(Pr[j * 4 + 2], v[2], t);
                            t = fma(Pr[j * 4 + 3], v[3], t);
                            acc[u][j] = (started && kind != CH_DIRECT) ? acc[u][j] * t : t;
                        }
                    }
                    started = true;
                }
            }
            if (sidx >= 0) {
                // NodeScale (src/treesub.c:7213-7227)
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double mx = fmax(fmax(acc[u][0], acc[u][1]), fmax(acc[u][2], acc[u][3]));
                    if (mx < 1e-300) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 1.0; lnsc[u] += -800.0; }
                    else { acc[u][0] /= mx; acc[u][1] /= mx; acc[u][2] /= mx; acc[u][3] /= mx; lnsc[u] += log(mx); }
                }
            }
            if (out_slot == OUT_REG) {
            }
            else if (out_slot >= 0) {
#pragma unroll
                for (int u = 0; u < PPT; u++)
#pragma unroll
                    for (int j = 0; j < 4; j++) slots[(((size_t)out_slot * PPT + u) * 4 + j) * THREADS + tid] = acc[u][j];
            }
            else {
                // root (src/treesub.c:7780-7798, 7726-7750)
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double fh = pi0 * acc[u][0];
                    fh = fma(pi1, acc[u][1], fh); fh = fma(pi2, acc[u][2], fh); fh = fma(pi3, acc[u][3], fh);
                    double v;
                    if (p.semantics == SEM_LFUN) { if (fh <= 0) fh = 1e-80; v = log(fh) + lnsc[u]; }
                    else { if (fh <= 0) fh = 1e-300; v = p.nscale ? log(fh) + lnsc[u] : fh; }
                    if (live[u]) p.fhK[(size_t)cls * N + h[u]] = v;
                }
            }
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

}  // namespace pb200
