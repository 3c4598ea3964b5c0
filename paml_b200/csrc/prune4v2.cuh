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
// codes with coalesced 32-bit loads issued four words (16 tips) ahead and extracts them with one byte-permute per tip;
// where each tip's code sits is static, so it is encoded in the op (no run-time bookkeeping).
// Conditional vectors never touch HBM: per class the kernel reads 1 byte per tip and pattern and writes one double.
#pragma once
#include "kernels.cuh"

namespace pb200 {

struct Prune4v2Params {
    const int* ops;             // macro-ops (2 ints each) in global memory; constant memory copy when ops_in_const
    int nops;
    int prog_base;              // offset of the op words in c_prog
    const unsigned int* zw;     // [ngroups][zwstride] walk-ordered tip codes, 4 per word
    long long zwstride, npatt;
    int ngroups;
    const double* P;            // [nclass][nnode][pstride]: internal branch P[j][k] row-major; tip branch PT[code][j]
    int pstride;
    int nnode, nclass, nslots, semantics, nscale;
    const double* pi;           // [>=4]
    double* fhK;                // [nclass][npatt]
    // PCONST: the internal branches' matrices of classes cls0 .. cls0 + gridDim.y - 1 sit in c_P4 as dense 4x4 blocks (16
    // doubles each): class blockIdx.y, node v (>= ns) at cp_base + blockIdx.y * cp_class_stride + 16 v  (cp_base = -16 ns);
    // shared memory then holds only the tip matrices
    int ns, cls0, cp_class_stride, cp_base;
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

__device__ __forceinline__ double2 lds_f64x2(unsigned int addr)
{
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_f64x2_volatile(unsigned int addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64x2(unsigned int addr, double a, double b)
{
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(addr), "d"(a), "d"(b) : "memory");
}

template <int PPT, int THREADS, int MINB, bool PSM, bool CP, bool PCONST>
__global__ void __launch_bounds__(THREADS, MINB) prune4v2_kernel(const Prune4v2Params p)
{
    extern __shared__ __align__(16) double sm4v2[];
    const int tid = threadIdx.x;
    const int cls = p.cls0 + blockIdx.y;
    const long long N = p.npatt;
    const int pstride = p.pstride;
    const unsigned int pstride_b = (unsigned int)pstride * 8u;
    const double* Pg = p.P + (size_t)cls * p.nnode * pstride;
    // matrices staged in shared memory: all of the class's (PSM) or, when the internal branches' are in c_P4, the tips' only
    // (tips are nodes 0 .. ns-1: the first ns matrices)
    const int nstaged = PCONST ? p.ns : p.nnode;
    if (PSM) {
        const double2* src = reinterpret_cast<const double2*>(Pg);
        double2* dst = reinterpret_cast<double2*>(sm4v2);
        for (int e = tid; e < nstaged * pstride / 2; e += THREADS) dst[e] = src[e];
        __syncthreads();
    }
    // shared-memory byte addresses: the staged matrices [nstaged][pstride] (if PSM), then the parked slots
    // [nslots][PPT][2][THREADS] x 16 bytes
    const unsigned int smem0 = (unsigned int)__cvta_generic_to_shared(sm4v2);
    const unsigned int slot0 = smem0 + (PSM ? (unsigned int)nstaged * pstride_b : 0u) + (unsigned int)tid * 16u;
    constexpr unsigned int SLOT_BYTES = PPT * 2 * THREADS * 16;
    const double pi0 = p.pi[0], pi1 = p.pi[1], pi2 = p.pi[2], pi3 = p.pi[3];
    const int nops = p.nops, ngroups = p.ngroups;
    const size_t zws = (size_t)p.zwstride;

    // 16 bytes of matrix `node` of this class at byte offset `off` (a warp-wide broadcast or a per-code gather)
    auto ldP = [&](unsigned int node_off, unsigned int off) -> double2 {
        if (PSM) return lds_f64x2(smem0 + node_off + off);
        return __ldg(reinterpret_cast<const double2*>(reinterpret_cast<const char*>(Pg) + node_off + off));
    };

    for (long long base = (long long)blockIdx.x * THREADS * PPT; base < N; base += (long long)gridDim.x * THREADS * PPT) {
        long long h[PPT];
        const unsigned int* zp[PPT];
        unsigned int cw[PPT], r[PPT][4];                 // current code word; queue of the next four (r[0] is the next one)
#pragma unroll
        for (int u = 0; u < PPT; u++) {
            h[u] = base + (long long)u * THREADS + tid;
            const long long hc = h[u] < N ? h[u] : N - 1;       // dead lanes walk a valid pattern, nothing is written
            zp[u] = p.zw + hc;
            cw[u] = __ldg(zp[u]);
#pragma unroll
            for (int d = 1; d <= 4; d++) r[u][d - 1] = (d < ngroups) ? __ldg(zp[u] + (size_t)d * zws) : 0u;
            zp[u] += 5 * zws;                                   // next word to fetch
        }
        double acc[PPT][4], lnsc[PPT];
#pragma unroll
        for (int u = 0; u < PPT; u++) { lnsc[u] = 0.0; acc[u][0] = acc[u][1] = acc[u][2] = acc[u][3] = 1.0; }

        // byte offset of the PT row of the next tip, for every pattern of this thread; `f` is the op's static tip field
        auto take = [&](int f, unsigned int (&off)[PPT]) {
            const int adv = (f >> 2) & 3;
            if (adv) {
                // shift the queue (three register moves per pattern every fourth tip: cheaper than indexing it, which made
                // the compiler clone the op bodies per queue position) and fetch the word four ahead
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    cw[u] = r[u][0]; r[u][0] = r[u][1]; r[u][1] = r[u][2]; r[u][2] = r[u][3];
                    if (adv == 1) r[u][3] = __ldg(zp[u]);
                    zp[u] += zws;
                }
            }
            const unsigned int sel = 0x4440u | (unsigned int)(f & 3);
#pragma unroll
            for (int u = 0; u < PPT; u++) off[u] = __byte_perm(cw[u], 0u, sel) * 32u;
        };
        auto row = [&](unsigned int node_off, unsigned int off, double (&c)[4]) {
            const double2 a = ldP(node_off, off), b = ldP(node_off, off + 16u);
            c[0] = a.x; c[1] = a.y; c[2] = b.x; c[3] = b.y;
        };
        // the 4x4 matrix of internal branch `node` (warp-uniform): uniform constant loads when PCONST, else 8 broadcast loads
        auto load16 = [&](int node, int op, double (&Pr)[16]) {
            if (PCONST) {
                // the node index is re-read from the constant-memory op word with the (uniform) op counter, so that the
                // compiler can see the address is warp-uniform and use the uniform datapath (LDCU -> UR operands of DFMA)
                const int ci = (c_prog[p.prog_base + 2 * op + 1] & 0xffff) * 16 + p.cp_base + (int)blockIdx.y * p.cp_class_stride;
#pragma unroll
                for (int e = 0; e < 16; e++) Pr[e] = c_P4[ci + e];
            }
            else {
                const unsigned int node_off = (unsigned int)node * pstride_b;
#pragma unroll
                for (int e = 0; e < 16; e += 2) { const double2 t = ldP(node_off, 8u * e); Pr[e] = t.x; Pr[e + 1] = t.y; }
            }
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
        auto op_words = [&](int op, int& w0, int& w1) {
            if (CP) { w0 = c_prog[p.prog_base + 2 * op]; w1 = c_prog[p.prog_base + 2 * op + 1]; }
            else { const int2 w = __ldg(reinterpret_cast<const int2*>(p.ops) + op); w0 = w.x; w1 = w.y; }
        };

        int nw0, nw1;
        op_words(0, nw0, nw1);
#pragma unroll 1
        for (int op = 0; op < nops; op++) {
            const int w0 = nw0, w1 = nw1;
            op_words(op + 1 < nops ? op + 1 : op, nw0, nw1);            // the next op's words are in flight during this op
            const int opc = w0 & 15;
            const int a_node = w1 & 0xffff;
            const unsigned int a_off = (unsigned int)a_node * pstride_b, b_off = ((unsigned int)w1 >> 16) * pstride_b;
            const unsigned int sl = slot0 + (((unsigned int)w0 >> M4_SLOT_SHIFT) & 255u) * SLOT_BYTES;
            switch (opc) {
            case M4_CHERRY: {
                // two tip children (src/baseml.c:1547-1558): rows of PT[code][j] = sum of the columns of P the code allows
                unsigned int oa[PPT], ob[PPT];
                take(w0 >> M4_TIPF0_SHIFT, oa); take(w0 >> M4_TIPF1_SHIFT, ob);
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double ra[4], rb[4];
                    row(a_off, oa[u], ra); row(b_off, ob[u], rb);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[u][j] = ra[j] * rb[j];
                }
                break;
            }
            case M4_MV_TIP: {
                // internal child still in registers (src/baseml.c:1559-1566), then a tip child
                double Pr[16];
                load16(a_node, op, Pr);
                unsigned int ob[PPT];
                take(w0 >> M4_TIPF0_SHIFT, ob);
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double t[4], rb[4];
                    row(b_off, ob[u], rb);
                    matvec(Pr, acc[u], t);
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[u][j] = t[j] * rb[j];
                }
                break;
            }
            case M4_MV_PUSH: {
                double Pr[16];
                load16(a_node, op, Pr);
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double t[4];
                    matvec(Pr, acc[u], t);
                    sts_f64x2(sl + (u * 2 + 0) * THREADS * 16, t[0], t[1]);
                    sts_f64x2(sl + (u * 2 + 1) * THREADS * 16, t[2], t[3]);
                }
                break;
            }
            case M4_MV_POP: {
                double Pr[16];
                load16(a_node, op, Pr);
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double t[4];
                    const double2 s0 = lds_f64x2_volatile(sl + (u * 2 + 0) * THREADS * 16), s1 = lds_f64x2_volatile(sl + (u * 2 + 1) * THREADS * 16);
                    matvec(Pr, acc[u], t);
                    acc[u][0] = t[0] * s0.x; acc[u][1] = t[1] * s0.y; acc[u][2] = t[2] * s1.x; acc[u][3] = t[3] * s1.y;
                }
                break;
            }
            case M4_PUSH: {
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    sts_f64x2(sl + (u * 2 + 0) * THREADS * 16, acc[u][0], acc[u][1]);
                    sts_f64x2(sl + (u * 2 + 1) * THREADS * 16, acc[u][2], acc[u][3]);
                }
                break;
            }
            case M4_TIP_MUL: {
                unsigned int oa[PPT];
                take(w0 >> M4_TIPF0_SHIFT, oa);
#pragma unroll
                for (int u = 0; u < PPT; u++) {
                    double ra[4];
                    row(a_off, oa[u], ra);
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
