// prune4v2.cuh -- the 4-state (nucleotide) pruning kernel, second generation: a register-resident tree walk on the
// FP64 vector pipe driven by a MACRO-OP program (plan.h: Plan4), built so that the interpretive overhead per FP64
// instruction stays below one.
//
// One thread owns PPT site patterns of one rate class and carries ONE running vector per pattern (4 doubles)
// through the whole tree (src/baseml.c:1517-1570, ConditionalPNode).  Every binary internal node is one macro-op
// whose last arithmetic instruction writes the running vector (no register copies); partial products that must
// wait for a sibling subtree are parked in shared memory as 16-byte halves ([slot][pattern][half][thread],
// conflict-free 128-bit accesses).  The class's 4x4 P(t) (internal branches) and PT[code][4] rows (tip branches,
// ambiguity sums pre-added by pt_kernel) are staged once per block in shared memory and read as warp broadcasts /
// per-code gathers.  Tip codes come from a WALK-ORDERED packing of com.z built once per (data, tree): word g of
// pattern h holds the codes of tips 4g..4g+3 in the order the program consumes them, so a thread streams its
// codes with coalesced 32-bit loads issued DEPTH words ahead and extracts them with one byte-permute per tip.
// Conditional vectors never touch HBM: per class the kernel reads 1 byte per tip and pattern and writes one double.
#pragma once
#include "kernels.cuh"

namespace pb200 {

struct Prune4v2Params {
    const int* ops;             // macro-ops (2 ints each) in global memory; constant memory copy when ops_in_const
    int nops;
    const unsigned int* zw;     // [ngroups][zwstride] walk-ordered tip codes, 4 per word
    long long zwstride, npatt;
    int ngroups;
    const double* P;            // [nclass][nnode][pstride]: internal branch P[j][k] row-major; tip branch PT[code][j]
    int pstride;
    int nnode, nclass, nslots, semantics, nscale;
    const double* pi;           // [>=4]
    double* fhK;                // [nclass][npatt]
};

// walk-ordered packing: zw[g][h] = codes of tips tip_order[4g..4g+3] at pattern h (missing tips: 0)
__global__ void pack_tips4_kernel(const unsigned char* __restrict__ z, long long zstride, const int* __restrict__ tip_order,
                                  int ntip, int ngroups, unsigned int* __restrict__ zw)
{
    const long long total = (long long)ngroups * zstride;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(e / zstride);
        const long long h = e - (long long)g * zstride;
        unsigned int w = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int t = 4 * g + i;
            if (t < ntip) w |= (unsigned int)z[(size_t)tip_order[t] * zstride + h] << (8 * i);
        }
        zw[e] = w;
    }
}

template <int PPT, int THREADS, int MINB, bool PSM, bool CP>
__global__ void __launch_bounds__(THREADS, MINB) prune4v2_kernel(const Prune4v2Params p)
{
    extern __shared__ __align__(16) double sm4v2[];
    constexpr int DEPTH = 3;                        // code words in flight ahead of the one being consumed
    const int tid = threadIdx.x;
    const int cls = blockIdx.y;
    const long long N = p.npatt;
    const int pstride = p.pstride;
    double* Psm = sm4v2;                                                              // [nnode][pstride] (if PSM)
    double2* slots = reinterpret_cast<double2*>(sm4v2 + (PSM ? (size_t)p.nnode * pstride : 0));   // [nslots][PPT][2][THREADS]
    const double* Pg = p.P + (size_t)cls * p.nnode * pstride;
    if (PSM) {
        const double2* src = reinterpret_cast<const double2*>(Pg);
        double2* dst = reinterpret_cast<double2*>(Psm);
        for (int e = tid; e < p.nnode * pstride / 2; e += THREADS) dst[e] = src[e];
        __syncthreads();
    }
    const double* Pbase = PSM ? Psm : Pg;
    const double pi0 = p.pi[0], pi1 = p.pi[1], pi2 = p.pi[2], pi3 = p.pi[3];
    const int nops = p.nops, ngroups = p.ngroups;
    double2* myslot = slots + tid;

    for (long long base = (long long)blockIdx.x * THREADS * PPT; base < N; base += (long long)gridDim.x * THREADS * PPT) {
        long long h[PPT];
        const unsigned int* zp[PPT];
        unsigned int cw[PPT], nw[PPT][DEPTH];
#pragma unroll
        for (int u = 0; u < PPT; u++) {
            h[u] = base + (long long)u * THREADS + tid;
            const long long hc = h[u] < N ? h[u] : N - 1;       // dead lanes walk a valid pattern, nothing is written
            zp[u] = p.zw + hc;
            cw[u] = __ldg(zp[u]);
#pragma unroll
            for (int d = 0; d < DEPTH; d++) nw[u][d] = (d + 1 < ngroups) ? __ldg(zp[u] + (size_t)(d + 1) * p.zwstride) : 0u;
            zp[u] += (size_t)(DEPTH + 1) * p.zwstride;           // next word to fetch
        }
        int ti = 0;
        double acc[PPT][4], lnsc[PPT];
#pragma unroll
        for (int u = 0; u < PPT; u++) { lnsc[u] = 0.0; acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 1.0; }

        // next tip code of every pattern of this thread (offset of its PT row in doubles = 4*code)
        auto take = [&](int (&off)[PPT]) {
            const unsigned int sel = 0x4440u | (unsigned int)(ti & 3);
#pragma unroll
            for (int u = 0; u < PPT; u++) off[u] = (int)__byte_perm(cw[u], 0u, sel) * 4;
            ti++;
            if ((ti & 3) == 0) {
                const bool more = (ti >> 2) + DEPTH < ngroups;
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    cw[u] = nw[u][0];
#pragma unroll
                    for (int d = 0; d + 1 < DEPTH; d++) nw[u][d] = nw[u][d + 1];
                    if (more) { nw[u][DEPTH - 1] = __ldg(zp[u]); zp[u] += p.zwstride; }
                }
            }
        };
        auto row = [&](const double* Pm, int off, double (&c)[4]) {
            const double2 a = *reinterpret_cast<const double2*>(Pm + off);
            const double2 b = *reinterpret_cast<const double2*>(Pm + off + 2);
            c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
        };
        auto load16 = [&](const double* Pm, double (&Pr)[16]) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) { const double2 t = *reinterpret_cast<const double2*>(Pm + e); Pr[e] = t.x; Pr[e + 1] = t.y; }
        };
        auto matvec = [&](const double (&Pr)[16], const double (&v)[4], double (&t)[4]) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double s = Pr[j * 4] * v[0];
                s = fma(Pr[j * 4 + 1], v[1], s);
                s = fma(Pr[j * 4 + 2], v[2], s);
                t[j] = fma(Pr[j * 4 + 3], v[3], s);
            }
        };

        for (int op = 0; op < nops; op++) {
            int w0, w1;
            if (CP) { w0 = c_prog[2 * op]; w1 = c_prog[2 * op + 1]; }
            else { const int2 w = __ldg(reinterpret_cast<const int2*>(p.ops) + op); w0 = w.x; w1 = w.y; }
            const int opc = w0 & 15;
            const int a = w1 & 0xffff, b = (w1 >> 16) & 0xffff;
            double2* sl = myslot + (size_t)((w0 >> 8) & 255) * (PPT * 2 * THREADS);
            switch (opc) {
            case M4_CHERRY: {
                // two tip children (src/baseml.c:1547-1558): rows of PT[code][j] = sum of the columns of P the code allows
                int oa[PPT], ob[PPT];
                take(oa); take(ob);
                const double *Pa = Pbase + (size_t)a * pstride, *Pb = Pbase + (size_t)b * pstride;
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double ra[4], rb[4];
                    row(Pa, oa[u], ra); row(Pb, ob[u], rb);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[u][j] = ra[j] * rb[j];
                }
                break;
            }
            case M4_MV_TIP: {
                // internal child still in registers (src/baseml.c:1559-1566), then a tip child
                double Pr[16];
                load16(Pbase + (size_t)a * pstride, Pr);
                int ob[PPT];
                take(ob);
                const double* Pb = Pbase + (size_t)b * pstride;
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double t[4], rb[4];
                    row(Pb, ob[u], rb);
                    matvec(Pr, acc[u], t);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[u][j] = t[j] * rb[j];
                }
                break;
            }
            case M4_MV_PUSH: {
                double Pr[16];
                load16(Pbase + (size_t)a * pstride, Pr);
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double t[4];
                    matvec(Pr, acc[u], t);
                    sl[(u * 2 + 0) * THREADS] = make_double2(t[0], t[1]);
                    sl[(u * 2 + 1) * THREADS] = make_double2(t[2], t[3]);
                }
                break;
            }
            case M4_MV_POP: {
                double Pr[16];
                load16(Pbase + (size_t)a * pstride, Pr);
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double t[4];
                    const double2 s0 = sl[(u * 2 + 0) * THREADS], s1 = sl[(u * 2 + 1) * THREADS];
                    matvec(Pr, acc[u], t);
                    acc[u][0] = t[0] * s0.x; acc[u][1] = t[1] * s0.y; acc[u][2] = t[2] * s1.x; acc[u][3] = t[3] * s1.y;
                }
                break;
            }
            case M4_PUSH: {
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    sl[(u * 2 + 0) * THREADS] = make_double2(acc[u][0], acc[u][1]);
                    sl[(u * 2 + 1) * THREADS] = make_double2(acc[u][2], acc[u][3]);
                }
                break;
            }
            case M4_TIP_MUL: {
                int oa[PPT];
                take(oa);
                const double* Pa = Pbase + (size_t)a * pstride;
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double ra[4];
                    row(Pa, oa[u], ra);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[u][j] *= ra[j];
                }
                break;
            }
            default: {
                // root (src/treesub.c:7780-7798, 7726-7750)
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double fh = pi0 * acc[u][0];
                    fh = fma(pi1, acc[u][1], fh); fh = fma(pi2, acc[u][2], fh); fh = fma(pi3, acc[u][3], fh);
                    double v;
                    if (p.semantics == SEM_LFUN) { if (fh <= 0) fh = 1e-80; v = log(fh) + lnsc[u]; }
                    else { if (fh <= 0) fh = 1e-300; v = p.nscale ? log(fh) + lnsc[u] : fh; }
                    if (h[u] < N) p.fhK[(size_t)cls * N + h[u]] = v;
                }
                break;
            }
            }
            if (w0 & M4_SCALE_FLAG) {
                // NodeScale (src/treesub.c:7213-7227)
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double mx = fmax(fmax(acc[u][0], acc[u][1]), fmax(acc[u][2], acc[u][3]));
                    if (mx < 1e-300) { acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 1.0; lnsc[u] += -800.0; }
                    else { acc[u][0] /= mx; acc[u][1] /= mx; acc[u][2] /= mx; acc[u][3] /= mx; lnsc[u] += log(mx); }
                }
            }
        }
    }
}

}  // namespace pb200
